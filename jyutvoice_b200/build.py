"""Builds the in-tree CUDA library (sm_100a only) with nvcc.  `python -m jyutvoice_b200.build`."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libjyutvoice_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-diag-suppress", "177",
]


def sources():
    out = []
    for root, _, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(os.path.dirname(HERE), "include", "jyutvoice_b200.h"))
    return out


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("JYUTVOICE_B200_NVCC_FLAGS", "").split()  # experiments, e.g. -DJV_ATTN_POLY_EVERY=4
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB, os.path.join(CSRC, "lib.cu")]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
