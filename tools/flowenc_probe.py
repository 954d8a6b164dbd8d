import sys, time; sys.path.insert(0, '.')
import torch, numpy as np
from jyutvoice_b200 import FlowEncoder, synthetic
from oracle import flow_encoder as ofe
sd = synthetic.make_flow_encoder_state_dict()
enc = FlowEncoder(); enc.load_state_dict(sd, strict=True); enc = enc.cuda()
g = torch.Generator().manual_seed(7)
for T, st in ((75, False), (150, True)):
    token = torch.randint(0, 6561, (1, T), generator=g)
    h, _ = enc(token.cuda(), torch.tensor([T]).cuda(), streaming=st)
    with torch.no_grad(): ref, _ = ofe.flow_encoder_forward(sd, token, st)
    print("T", T, st, "max err", (h.cpu() - ref).abs().max().item(), "ref max", ref.abs().max().item())
for B, T in ((1, 75), (1, 250), (64, 150)):
    token = torch.randint(0, 6561, (B, T), generator=g).cuda()
    lens = torch.full((B,), T).cuda()
    for _ in range(3): enc(token, lens)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): enc(token, lens)
    torch.cuda.synchronize(); print("B", B, "T", T, "ms per call", (time.perf_counter() - t0) / 5 * 1e3)
