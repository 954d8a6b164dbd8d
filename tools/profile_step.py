"""One short pass of the hot path for ncu (launch list / --set full captures).
usage: python tools/profile_step.py [batch] [frames] [nfe] [precision] [what=all|cfm|hift]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, HiFTGenerator, synthetic  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 300
nfe = int(sys.argv[3]) if len(sys.argv) > 3 else 1
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
what = sys.argv[5] if len(sys.argv) > 5 else "all"

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
lens = [T] * B
mu = torch.randn(B, 80, T, generator=g).to(dev)
spks = torch.randn(B, 80, generator=g).to(dev)
mel = (torch.randn(B, 80, T, generator=g) * 2 - 5).to(dev)
if what in ("all", "cfm"):
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=prec))
    cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)
    cfm = cfm.to(dev)
    mel, _ = cfm(mu, None, nfe, 1.0, spks, None, lengths=lens)
if what in ("all", "hift"):
    hift = HiFTGenerator(precision=prec)
    hift.load_state_dict(synthetic.make_hift_state_dict(), strict=True)
    hift = hift.to(dev)
    wav, _ = hift.inference(mel, lengths=lens)
torch.cuda.synchronize()
print("ok", float(mel.abs().mean()))
