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
    from oracle import estimator as oe
    from jyutvoice_b200 import synthetic as weights
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


def hift_checks():
    from oracle import hift as oh
    from jyutvoice_b200 import synthetic as weights
    from oracle.make_golden import hift_mel
    from jyutvoice_b200 import HiFTGenerator
    from conftest import snr_db
    GOLD = os.path.join(ROOT, "tests", "golden")
    for prec in ("fp32", "bf16"):
        for tag, f0b in (("unvoiced", None), ("voiced", 200.0)):
            section(f"hift {prec} {tag}")
            sd = weights.make_hift_state_dict(f0_bias=f0b)
            hift = HiFTGenerator(precision=prec)
            hift.load_state_dict(sd, strict=True)
            hift = hift.cuda()
            g = np.load(os.path.join(GOLD, f"hift_{tag}.npz"))
            B, T = int(g["B"]), int(g["T"])
            mel = hift_mel(int(g["seed"]), B, T)
            f0 = hift.predict_f0(mel.cuda()).cpu()
            f0r = torch.from_numpy(g["f0"])
            print(f"f0: max_abs={(f0-f0r).abs().max().item():.3e} ref_max={f0r.abs().max().item():.3f}", flush=True)
            rng = oh.draw_source_rng(B, T * 480, torch.Generator().manual_seed(int(g["rng_seed"])))
            s = hift.source(f0r.cuda(), rng).cpu()
            sr = torch.from_numpy(g["s"])
            print(f"source (golden f0): max_abs={(s-sr).abs().max().item():.3e}", flush=True)
            wav = hift.decode(mel.cuda(), sr.cuda()).cpu()
            wr = torch.from_numpy(g["wav_decode"])
            print(f"decode(x, s_golden): snr={snr_db(wr, wav):.1f} dB max_abs={(wav-wr).abs().max().item():.3e} nan={torch.isnan(wav).sum().item()}", flush=True)
            wav_i, s_i = hift.inference(mel.cuda(), rng=rng)
            wi = torch.from_numpy(g["wav_inference"])
            print(f"inference(rng shared): snr={snr_db(wi, wav_i.cpu()):.1f} dB s_max_abs={(s_i.cpu()-sr).abs().max().item():.3e}", flush=True)
            # ragged batch: each utterance must equal its own unpadded oracle call
            lens = [30, 19]
            melr = mel.clone()
            melr[1, :, 19:] = 0
            sr2 = sr.clone()
            wav_r = hift.decode(melr.cuda(), sr2.cuda(), lengths=lens).cpu()
            with torch.no_grad():
                ref1 = oh.decode(sd, mel[1:2, :, :19], sr[1:2, :, :19 * 480])
            print(f"ragged decode utt1 (T=19 of 30): snr={snr_db(ref1, wav_r[1:2, :19*480]):.1f} dB tail_max={wav_r[1, 19*480:].abs().max().item():.1e} "
                  f"utt0 snr={snr_db(wr[0:1], wav_r[0:1]):.1f}", flush=True)
            if prec == "bf16" and tag == "unvoiced":
                B, T = 16, 300
                melb = (torch.randn(B, 80, T) * 2 - 5).cuda()
                hift.inference(melb)
                torch.cuda.synchronize(); t0 = time.time()
                hift.inference(melb)
                torch.cuda.synchronize(); dt = time.time() - t0
                print(f"B=16 T=300 hift bf16: {dt*1e3:.1f} ms -> {B*T/50/dt:.1f} audio-s/s (HiFT only)", flush=True)
            del hift


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "estimator", "hift"]
    print(torch.cuda.get_device_name(0), torch.__version__)
    if "gemm" in which:
        section("gemm")
        run(gemms)
    if "estimator" in which:
        run(estimator_checks)
    if "hift" in which:
        run(hift_checks)
    print("launches", _lib.lib().jv_launch_count())
