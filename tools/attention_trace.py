"""[needs a tracing build: JYUTVOICE_B200_NVCC_FLAGS=-DJV_TRACE python -m jyutvoice_b200.build --force]
Per-CTA timeline of attention_tc2_kernel (clock64 stamps written by thread 0 of every CTA of ONE launch).
usage: python tools/attention_trace.py [batch=64] [frames=300]"""
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
lens = [T] * B
mu = torch.randn(B, 80, T, generator=g).to(dev)
spks = torch.randn(B, 80, generator=g).to(dev)
cfm(mu, None, 1, 1.0, spks, None, lengths=lens)
L = ctypes.CDLL(_lib.LIB_PATH)
nq = (T + 127) // 128
n_cta = nq * 8 * 2 * B
buf = torch.zeros(max(n_cta * 8, 296 * 16), dtype=torch.int64, device=dev)
L.jv_debug_attention_trace(ctypes.c_void_p(buf.data_ptr()))
cfm(mu, None, 1, 1.0, spks, None, lengths=lens)   # every attention launch overwrites the buffer: the last one stays
torch.cuda.synchronize()
L.jv_debug_attention_trace(ctypes.c_void_p(0))
if os.environ.get("JYUTVOICE_B200_ATTN", "2") == "3":
    t = buf.cpu().double()[: 296 * 16].view(296, 16)
    t = t[t[:, 8] > 0]
    names = ["setup", "item decode (global loads)", "wait first S of item", "softmax loops", "  waiting S (j>=1)", "  waiting PV (late)",
             "final PV wait", "O read-out + store", "CTA total", "items", "tiles"]
    print(f"persistent kernel: {len(t)} CTAs; per-CTA sums over its items (clock cycles)")
    for k, nme in enumerate(names):
        x = t[:, k]
        print(f"{nme:30s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  p90 {x.quantile(0.9):9.0f}  max {x.max():9.0f}")
    sys.exit(0)
t = buf.cpu().view(n_cta, 8).double()
t = t[t[:, 5] > 0]
full = t.view(-1, 8)
def stat(name, x):
    print(f"{name:34s} mean {x.mean():9.0f}  p50 {x.median():9.0f}  p90 {x.quantile(0.9):9.0f}  max {x.max():9.0f}  clk")
qt = torch.arange(n_cta)[: len(t)] % nq
for sel, label in ((qt < nq - 1, "full query tiles"), (qt == nq - 1, "last query tile")):
    s = full[sel]
    print(f"--- {label}: {len(s)} CTAs")
    stat("setup (entry -> after pdl_wait)", s[:, 1] - s[:, 0])
    stat("first S ready (loads + S MMA)", s[:, 2] - s[:, 1])
    stat("softmax loop (tile 0 -> last P)", s[:, 3] - s[:, 2])
    stat("  of which waiting for S (j>=1)", s[:, 6])
    stat("  of which waiting for PV (late)", s[:, 7])
    stat("final PV wait", s[:, 4] - s[:, 3])
    stat("O read-out + store", s[:, 5] - s[:, 4])
    stat("CTA total", s[:, 5] - s[:, 0])
span = full[:, 5].max() - full[:, 0].min()
print(f"kernel span (first entry -> last exit) ~ {span:.0f} clk (SM clocks are not synchronised across SMs: indicative only)")
