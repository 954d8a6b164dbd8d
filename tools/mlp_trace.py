"""[needs a tracing build: JYUTVOICE_B200_NVCC_FLAGS=-DJV_TRACE python -m jyutvoice_b200.build --force]
Where the MMA thread of mlp_fused_pair_kernel waits (clock64 sums per CTA pair, last launch of one estimator forward).
usage: python tools/mlp_trace.py [batch=64] [frames=300]"""
import ctypes, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic, _lib  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dev = torch.device("cuda:0")
cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="bf16"))
cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)
cfm = cfm.to(dev)
g = torch.Generator().manual_seed(0)
mu = torch.randn(B, 80, T, generator=g).to(dev)
spks = torch.randn(B, 80, generator=g).to(dev)
cfm(mu, None, 1, 1.0, spks, None, lengths=[T] * B)
L = ctypes.CDLL(_lib.LIB_PATH)
buf = torch.zeros(4096 * 8, dtype=torch.int64, device=dev)
os.environ["JYUTVOICE_B200_ATTN"] = os.environ.get("JYUTVOICE_B200_ATTN", "1")
L.jv_debug_attention_trace(ctypes.c_void_p(buf.data_ptr()))
cfm(mu, None, 1, 1.0, spks, None, lengths=[T] * B)
torch.cuda.synchronize()
L.jv_debug_attention_trace(ctypes.c_void_p(0))
t = buf.cpu().view(-1, 8)[:74].double()
names = ["MMA thread total", "wait LNX tile (a_full)", "wait weights (w_full)", "wait acc1 drained", "wait H written (h_full)", "wait acc2 drained", "tiles", "FF1 issue (16 MMAs x chunks, waits excluded)"]
for k, n in enumerate(names):
    x = t[:, k]
    print(f"{n:28s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  max {x.max():9.0f}")
issue = t[:, 0] - t[:, 1:6].sum(1)
print(f"issue time total             mean {issue.mean():9.0f}   FF1 part {t[:, 7].mean():9.0f}   FF2 part (+ loop) {(issue - t[:, 7]).mean():9.0f}")
print(f"per MMA: FF1 (N=128) {(t[:, 7] / (t[:, 6] * 128)).mean():6.1f} clk   FF2 (N=256) {((issue - t[:, 7]) / (t[:, 6] * 64)).mean():6.1f} clk")
