"""Builds libdmc_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

    python build_native.py [--force]

(kept outside the package so that building never imports the package, which loads the library)

Flags that matter for bit parity: -fmad=false (no FMA contraction) and no --use_fast_math (IEEE division and
square root).  -lineinfo keeps the ncu source page mapped to these files.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "depthmapcompression_b200")
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdmc_b200.so")
SOURCES = ["dmc_kernels_8u.cu", "dmc_bwrf8u_h2.cu", "dmc_bwrf8u_c3_h2.cu", "dmc_joint_bwrf.cu", "dmc_front8u.cu", "dmc_kernels_32f.cu", "dmc_bwrf32f_tiled.cu", "dmc_jpeg.cu", "dmc_hostlink.cu", "dmc_render.cu", "dmc_capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2", "--shared", "-Xptxas", "-v", "--threads", "0"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "dmc_c.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    cmd = [_nvcc()] + NVCC_FLAGS + (["-ccbin", ccbin] if ccbin else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-4000:])
    if verbose:
        print(log)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("built", LIB)
