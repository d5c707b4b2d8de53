"""Builds the engine's C-ABI shared library in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdyffusion_b200.so")
SOURCES = ["abi.cu", "net.cu", "sampler.cu", "conv_mma.cu", "conv_umma.cu", "conv_up.cu", "conv_flat.cu", "aux_kernels.cu", "attn_kernels.cu", "attn_fused.cu", "metrics_kernels.cu",
           "dataset_kernels.cu", "optim_kernels.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC,-O2,-fvisibility=hidden", "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dyffusion_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, bf16: bool = False) -> str:
    """bf16=True (or DYF_ACT_BF16=1 in the environment) builds the bf16-storage variant as libdyffusion_b200_bf16.so
    (load it with DYF_LIB=<path>; A/B measurements of the 16-bit storage type)."""
    bf16 = bf16 or os.environ.get("DYF_ACT_BF16") == "1"
    if bf16:
        return _build_variant("bf16", ["-DDYF_ACT_BF16"], verbose)
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(HERE, "build", "nvcc.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


def _build_variant(tag: str, defines, verbose: bool) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    lib = os.path.join(HERE, f"libdyffusion_b200_{tag}.so")
    bdir = os.path.join(HERE, "build", tag)
    os.makedirs(bdir, exist_ok=True)
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({tag}):\n{out}")
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed ({tag}):\n{r.stdout}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
