"""The drop-in boundary is a C ABI: include/dyffusion_b200.h must compile as plain C99 and be callable from a C program
linked against libdyffusion_b200.so (tests/c_abi_probe.c).  Runs without a GPU: handle bookkeeping works on the host, every
compute entry point must answer DYF_ERR_CUDA ("no CPU fallback") rather than compute."""
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "dyffusion_b200")

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")


def test_header_is_plain_c99(tmp_path):
    src = tmp_path / "hdr.c"
    src.write_text('#include "dyffusion_b200.h"\nint main(void) { return sizeof(dyf_net_desc) == 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{INC}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_program_links_and_gets_the_documented_errors(tmp_path):
    import dyffusion_b200.engine  # noqa: F401  (builds nothing; fails loudly if the library is missing)
    exe = str(tmp_path / "c_abi_probe")
    flags = [] if torch.cuda.is_available() else ["-DEXPECT_NO_GPU"]
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", *flags, f"-I{INC}", os.path.join(ROOT, "tests", "c_abi_probe.c"),
                        f"-L{LIBDIR}", "-ldyffusion_b200", f"-Wl,-rpath,{LIBDIR}", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and "c_abi_probe: ok" in run.stdout, run.stdout + run.stderr
