"""Developer tool (CPU): the fit behind `gelu_fast` (csrc/common.cuh): 0.5 erfc(a / sqrt 2) = 2^P(a) on a in [0, 6], P a
weighted Chebyshev least-squares polynomial, checked in emulated fp32 Horner arithmetic against scipy's erf/erfc."""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from numpy.polynomial import polynomial as Pn
from scipy.special import erf, erfc

A = 6.0
a = np.cos(np.linspace(0, np.pi, 4001)) * A / 2 + A / 2
tail = 0.5 * erfc(a / np.sqrt(2))
w = tail * np.maximum(a, 0.25)  # GELU error = a * tail * ln2 * dP
for deg in (6, 7, 8):
    p = Ch.cheb2poly(Ch.chebfit(2 * a / A - 1, np.log2(tail), deg, w=w))
    comp = np.zeros(1)
    for k, ck in enumerate(p):
        comp = Pn.polyadd(comp, ck * Pn.polypow(np.array([-1.0, 2 / A]), k))
    xs = np.linspace(0, 8, 200001)
    ac = np.minimum(xs, A).astype(np.float32)
    acc = np.float32(comp[-1]) * np.ones_like(ac)
    for ck in comp[-2::-1]:
        acc = acc * ac + np.float32(ck)
    e = np.exp2(acc.astype(np.float64))
    err_pos = np.abs(xs * ((1 - e) - 0.5 * (1 + erf(xs / np.sqrt(2)))))
    err_neg = np.abs(xs * (e - 0.5 * erfc(xs / np.sqrt(2))))
    print(deg, "max |GELU err|", max(err_pos.max(), err_neg.max()), "max |Phi err|", np.abs((1 - e) - 0.5 * (1 + erf(xs / np.sqrt(2)))).max())
    print("   coefficients (ascending):", [float(np.float32(c)) for c in comp])
