"""[needs a tracing build: JYUTVOICE_B200_NVCC_FLAGS=-DJV_TRACE python -m jyutvoice_b200.build --force]
Where the MMA thread and epilogue warp 0 of gemm_taps_tc_kernel wait, for the launches with one epilogue mask
(clock64 sums per CTA, last such launch of one estimator forward).
usage: python tools/gemm_trace.py <epi mask, e.g. 54 = out-proj, 55 = conv2, 8 = QKV / FF1> [batch=64] [frames=300]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic, _lib  # noqa: E402
epi = int(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 300
dev = torch.device("cuda:0")
cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="bf16"))
cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)
cfm = cfm.to(dev)
g = torch.Generator().manual_seed(0)
mu = torch.randn(B, 80, T, generator=g).to(dev)
spks = torch.randn(B, 80, generator=g).to(dev)
cfm(mu, None, 1, 1.0, spks, None, lengths=[T] * B)
L = ctypes.CDLL(_lib.LIB_PATH)
buf = torch.zeros(256 * 8, dtype=torch.int64, device=dev)
L.jv_debug_gemm_trace(ctypes.c_void_p(buf.data_ptr()), epi)
cfm(mu, None, 1, 1.0, spks, None, lengths=[T] * B)
torch.cuda.synchronize()
L.jv_debug_gemm_trace(ctypes.c_void_p(0), -1)
t = buf.cpu().view(-1, 8).double()
mma = t[t[:, 0] > 0]
epi_rows = t[t[:, 4] > 0]
print(f"epilogue mask {epi}: {len(mma)} MMA threads, {len(epi_rows)} epilogue warps reported")
for k, n in enumerate(["MMA thread total", "  wait accumulator free (tempty)", "  wait operands (full)", "  units"]):
    x = mma[:, k]
    print(f"{n:36s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  max {x.max():9.0f}")
for k, n in enumerate(["epilogue warp 0 total", "  wait accumulator (tfull)", "  wait residual tile", "  tiles"]):
    x = epi_rows[:, 4 + k]
    print(f"{n:36s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  max {x.max():9.0f}")
