"""`FastDiv` (csrc/common.cuh): the multiply-high + shift constants the persistent kernels use to decode their work items
(n / d for run-time d without a hardware division).  The host constructor is compiled with nvcc (no GPU needed to run it) and
its constants are checked here against exact integer division for every n the kernels can pass (n, d < 2^31)."""
import os
import random
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SRC = r'''
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
int main(int argc, char** argv) {
  for (int i = 1; i < argc; ++i) {
    dyf::FastDiv f((uint32_t)strtoul(argv[i], nullptr, 10));
    printf("%u %u %u\n", f.d, f.mul, f.shr);
  }
  return 0;
}
'''


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not installed")
def test_fastdiv_constants_divide_exactly(tmp_path):
    src = tmp_path / "fastdiv_probe.cu"
    src.write_text(SRC)
    exe = tmp_path / "fastdiv_probe"
    subprocess.run([NVCC, "-std=c++17", "-I", os.path.join(ROOT, "dyffusion_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    rng = random.Random(7)
    divisors = sorted(set(list(range(1, 300)) + [2 ** k for k in range(31)] + [2 ** k - 1 for k in range(2, 32)] +
                          [2 ** k + 1 for k in range(1, 30)] + [rng.randrange(1, 2 ** 31) for _ in range(400)]))
    out = subprocess.run([str(exe)] + [str(d) for d in divisors], check=True, capture_output=True, text=True).stdout.split("\n")
    rows = [tuple(int(v) for v in line.split()) for line in out if line.strip()]
    assert [r[0] for r in rows] == divisors
    for d, mul, shr in rows:
        assert 0 < mul < 2 ** 32 and shr < 32
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - d] + [rng.randrange(0, 2 ** 31) for _ in range(64)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                q = ((((n * mul) >> 32) + n) & 0xFFFFFFFF) >> shr   # = (__umulhi(n, mul) + n) >> shr in 32-bit arithmetic
                assert q == n // d, (n, d, mul, shr)
