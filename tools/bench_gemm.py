"""Micro-benchmark of the tcgen05 GEMM over the estimator / HiFT shapes (CUDA events, no profiler)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: F401,E402  (initialises the CUDA context the same way the product does)
from jyutvoice_b200 import _lib  # noqa: E402

torch.zeros(1, device="cuda")
L = _lib.lib()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 38656
only = sys.argv[2] if len(sys.argv) > 2 else None
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cases = [
    ("QKV      N=1536 K=256  bf16 out", M, 1536, 256, 1, 16),
    ("FF1      N=1024 K=256  GELU bf16", M, 1024, 256, 1, 16 | 2),
    ("out-proj N=256  K=512  resid f32 + LN2", M, 256, 512, 1, 1 | 8),
    ("out-proj N=256  K=512  resid f32", M, 256, 512, 1, 1),
    ("xb16 out-proj N=256 K=512 resid + LN2", M, 256, 512, 1, 1 | 8 | 32),
    ("xb16 FF2 N=256 K=1024 resid + LN2", M, 256, 1024, 1, 1 | 8 | 32),
    ("xb16 conv N=256 K=3x256 LN1+Mish+resid+LN2", M, 256, 256, 3, 1 | 4 | 8 | 32),
    ("FF2      N=256  K=1024 resid f32 + LN2", M, 256, 1024, 1, 1 | 8),
    ("conv     N=256  K=3x256 plain bf16", M, 256, 256, 3, 16),
    ("conv     N=256  K=3x256 LN1+Mish bf16", M, 256, 256, 3, 16 | 4),
    ("conv     N=256  K=3x256 LN1+Mish+resid f32+LN2", M, 256, 256, 3, 1 | 4 | 8),
    ("plain    N=256  K=256  bf16", M, 256, 256, 1, 16),
    ("hift s2  N=64   K=11x64 bf16 (M x 8)", M * 8, 64, 64, 11, 16),
    ("hift s0  N=256  K=7x256 bf16", M // 4, 256, 256, 7, 16),
]
for name, m, n, k, taps, mode in cases:
    if only and not (name == only[1:] if only.startswith('=') else name.startswith(only)):
        continue
    ms = ctypes.c_double()
    rc = L.jv_bench_gemm(m, n, k, taps, mode, iters, ctypes.byref(ms))
    if rc != 0:
        print(name, "rc", rc, L.jv_last_error().decode())
        continue
    fl = 2.0 * m * n * k * taps
    print(f"{name:50s} M={m:7d}: {ms.value*1e3:9.1f} us  {fl/ms.value/1e9:8.1f} TFLOP/s", flush=True)
