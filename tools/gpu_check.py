"""Diagnostic run on the GPU box: prints per-kernel errors instead of pass/fail (development aid)."""
import ctypes
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from jyutvoice_b200 import _lib  # noqa: E402


def section(name):
    print(f"\n===== {name} =====", flush=True)


def run(fn):
    try:
        fn()
    except Exception:
        traceback.print_exc()
        print("FAILED", flush=True)


def gemm_case(prec, M, N, K, bias=True):
    L = _lib.lib()
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    C = torch.full((M, N), float("nan"), device="cuda")
    p = lambda z: ctypes.c_void_p(0 if z is None else z.data_ptr())
    torch.cuda.synchronize()
    t0 = time.time()
    rc = L.jv_test_gemm(_lib.PREC[prec], M, N, K, p(A), p(W), p(b), p(C), None)
    torch.cuda.synchronize()
    dt = time.time() - t0
    if rc != 0:
        print(f"gemm {prec} M={M} N={N} K={K}: rc={rc} {L.jv_last_error().decode()}")
        return
    if prec == "bf16":
        Ar, Wr = A.bfloat16().double(), W.bfloat16().double()
    else:
        Ar, Wr = A.double(), W.double()
    ref = Ar @ Wr.T + (b.double() if bias else 0)
    err = (C.double() - ref).abs().max().item()
    print(f"gemm {prec} M={M} N={N} K={K}: max_abs_err={err:.3e} ref_max={ref.abs().max().item():.2f} nan={torch.isnan(C).sum().item()} ({dt*1e3:.1f} ms)", flush=True)


def gemms():
    for prec in ("fp32", "bf16"):
        for (M, N, K) in [(128, 256, 64), (128, 256, 256), (300, 256, 256), (1000, 1536, 256), (777, 80, 256), (4096, 1024, 256),
                          (2048, 256, 1024), (515, 64, 128), (40000, 256, 768)]:
            gemm_case(prec, M, N, K)


def estimator_checks():
    from oracle import weights, estimator as oe
    from oracle.make_golden import est_inputs, cfm_inputs
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder
    sd = weights.make_estimator_state_dict()
    nb = weights.noise_bank()
    GOLD = os.path.join(ROOT, "tests", "golden")
    for prec in ("fp32", "bf16"):
        section(f"estimator {prec}")
        est = CausalConditionalDecoder(precision=prec)
        cfm = CausalConditionalCFM(estimator=est)
        cfm.load_state_dict(sd, strict=True)
        cfm = cfm.cuda()
        g = np.load(os.path.join(GOLD, "estimator_fwd.npz"))
        x, mask, mu, t, spks, cond = est_inputs(int(g["seed"]), int(g["R"]), int(g["T"]), list(g["lens"]))
        t0 = time.time()
        v = est(x.cuda(), mask.cuda(), mu.cuda(), t.cuda(), spks.cuda(), cond.cuda()).cpu()
        print(f"first forward {time.time()-t0:.2f}s")
        ref = torch.from_numpy(g["out"])
        print(f"estimator_fwd vs golden: max_abs={(v-ref).abs().max().item():.3e} ref_max={ref.abs().max().item():.3f} "
              f"rel_rms={((v-ref).pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item():.3e} nan={torch.isnan(v).sum().item()}", flush=True)
        for name in ("cfm_T33_n4", "cfm_T50_n10"):
            g = np.load(os.path.join(GOLD, name + ".npz"))
            T, n = int(g["T"]), int(g["n_timesteps"])
            mu, spks = cfm_inputs(int(g["seed"]), T)
            torch.cuda.synchronize(); t0 = time.time()
            mel, _ = cfm(mu.cuda(), torch.ones(1, 1, T).cuda(), n, 1.0, spks.cuda(), torch.zeros(1, 80, T).cuda())
            torch.cuda.synchronize(); dt = time.time() - t0
            mel = mel.cpu()
            ref = torch.from_numpy(g["out"])
            print(f"{name}: max_abs={(mel-ref).abs().max().item():.3e} ref_max={ref.abs().max().item():.3f} "
                  f"rel_rms={((mel-ref).pow(2).mean().sqrt()/ref.pow(2).mean().sqrt()).item():.3e} ({dt*1e3:.1f} ms)", flush=True)
        # ragged batch vs per-utterance oracle
        B, lens = 3, [41, 17, 30]
        gg = torch.Generator().manual_seed(77)
        mu = torch.randn(B, 80, max(lens), generator=gg)
        spks = torch.randn(B, 80, generator=gg)
        mask = torch.zeros(B, 1, max(lens))
        for i, l in enumerate(lens):
            mask[i, 0, :l] = 1
        mel, _ = cfm(mu.cuda(), mask.cuda(), 3, 1.0, spks.cuda(), None)
        mel = mel.cpu()
        with torch.no_grad():
            ref = oe.cfm_forward_batch(sd, nb, mu, lens, 3, 1.0, spks, None)
        print(f"ragged batch n=3: max_abs={(mel-ref).abs().max().item():.3e} pad_max={max(mel[1,:,17:].abs().max().item(), mel[2,:,30:].abs().max().item()):.1e}", flush=True)
        # timing at bench shape
        if prec == "bf16":
            B, T = 16, 300
            mu = torch.randn(B, 80, T).cuda(); spks = torch.randn(B, 80).cuda(); mask = torch.ones(B, 1, T).cuda()
            cfm(mu, mask, 2, 1.0, spks, None)
            torch.cuda.synchronize(); t0 = time.time()
            cfm(mu, mask, 10, 1.0, spks, None)
            torch.cuda.synchronize(); dt = time.time() - t0
            print(f"B=16 T=300 n=10 bf16: {dt*1e3:.1f} ms -> {B*T/50/dt:.1f} audio-s/s (CFM only)", flush=True)
        del cfm, est


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.__version__)
    section("gemm")
    run(gemms)
    run(estimator_checks)
    print("launches", _lib.lib().jv_launch_count())
