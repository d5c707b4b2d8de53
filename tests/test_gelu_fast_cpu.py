"""`gelu_fast` (csrc/common.cuh): nn.GELU() (erf form, what the reference's SimpleConvNet and time MLPs use) as
v * Phi(v) with the normal tail written as 2^P(|v|).  The polynomial coefficients are read out of the CUDA source and the
function is re-evaluated here in fp32 Horner arithmetic against scipy's erf: |error| <= 2e-7 on the whole line -- three
orders of magnitude below the fp16 rounding of the stored activation (tests/micro/gelu_fit.py is the fit itself)."""
import os
import re

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _coefficients():
    src = open(os.path.join(ROOT, "dyffusion_b200", "csrc", "common.cuh")).read()
    body = src[src.index("float gelu_fast(float v)"):]
    body = body[:body.index("return v *")]
    lead = float(re.search(r"float p = ([-+0-9.e]+)f;", body).group(1))
    rest = [float(m) for m in re.findall(r"p = fmaf\(p, a, ([-+0-9.e]+)f\);", body)]
    clamp = float(re.search(r"fminf\(fabsf\(v\), ([0-9.]+)f\)", body).group(1))
    return [lead] + rest, clamp


def test_gelu_fast_matches_the_erf_form():
    coef, clamp = _coefficients()
    assert len(coef) == 8 and clamp == 6.0
    v = np.concatenate([np.linspace(-12, 12, 400001), np.array([0.0, -0.0, 1e-8, -1e-8, 30.0, -30.0])]).astype(np.float32)
    a = np.minimum(np.abs(v), np.float32(clamp))
    p = np.full_like(a, np.float32(coef[0]))
    for c in coef[1:]:
        p = p * a + np.float32(c)                       # fp32 Horner, like the fmaf chain (up to the fused rounding)
    e = np.exp2(p.astype(np.float64))
    got = v.astype(np.float64) * np.where(v > 0, 1.0 - e, e)
    want = 0.5 * v.astype(np.float64) * (1.0 + erf(v.astype(np.float64) / np.sqrt(2.0)))
    assert np.abs(got - want).max() <= 2e-7
    assert abs(got[np.argmax(v == 30.0)] - 30.0) < 1e-6 and abs(got[np.argmax(v == -30.0)]) < 1e-6  # clamped tail
