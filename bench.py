#!/usr/bin/env python
"""Headline benchmark: audio-seconds generated per second (1/RTF) of the CFM (10 NFE, CFG) + HiFT hot
path on synthetic utterances of BASELINE.json's throughput config (batch 64 per GPU, ~6 s each, bf16).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the oracle port on host cores

One "step" = one pass of the hot path over one batch: CausalConditionalCFM.forward -> HiFTGenerator.inference.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
FRAMES_PER_SEC = 50.0  # 24000 / 480


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the newest committed
    `ncu --set full` summary (profiles/*_gemm_traffic.json); None if there is none."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_gemm_traffic.json")))
    if not files:
        return None, None
    try:
        with open(files[-1]) as f:
            d = json.load(f)
        return float(d["avg_dram_bytes_per_launch"]), os.path.relpath(files[-1], ROOT)
    except Exception:
        return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained"), "bf16_burst": d.get("bf16_tflops"), "hbm": d.get("hbm_gbs"),
                "src": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "src": "fallback"}


def make_workload(batch, frames, seed):
    """Synthetic utterances of the named shape: lengths U{0.9 T .. 1.1 T} (config 5: Tx in U{45..55} tokens)."""
    g = torch.Generator().manual_seed(seed)
    lo, hi = int(frames * 0.9), int(frames * 1.1)
    lens = torch.randint(lo, hi + 1, (batch,), generator=g).tolist()
    Tmax = hi
    mu = torch.randn(batch, 80, Tmax, generator=g)
    spks = torch.randn(batch, 80, generator=g)
    for i, l in enumerate(lens):
        mu[i, :, l:] = 0
    return lens, Tmax, mu, spks


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def cpu_reference_pass(est_sd, hift_sd, noise_bank, frames, nfe, seed):
    """One utterance through the oracle port (B=1, the reference's only mode); returns audio seconds."""
    from oracle import estimator as oe, hift as oh
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(1, 80, frames, generator=g)
    spks = torch.randn(1, 80, generator=g)
    with torch.no_grad():
        mel = oe.cfm_forward(est_sd, noise_bank, mu, torch.ones(1, 1, frames), nfe, 1.0, spks, torch.zeros(1, 80, frames))
        rng = oh.draw_source_rng(1, 480 * frames, g)
        oh.inference(hift_sd, mel, rng)
    return frames / FRAMES_PER_SEC


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from jyutvoice_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    est_sd = synthetic.make_estimator_state_dict()
    hift_sd = synthetic.make_hift_state_dict()
    nb = synthetic.noise_bank()
    for w in range(args.warmup):
        cpu_reference_pass(est_sd, hift_sd, nb, min(args.frames, 100), args.nfe, w)
    t0 = time.perf_counter()
    audio = 0.0
    for k in range(args.steps):
        audio += cpu_reference_pass(est_sd, hift_sd, nb, args.frames, args.nfe, 100 + k)
    dt = time.perf_counter() - t0
    value = audio / dt
    sample = f"{args.steps} utterance(s) of {args.frames} frames, batch 1 loop (the reference's only mode), fp32, {args.nfe} NFE + HiFT"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(1, args.steps), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": f"BASELINE configs[4] per-GPU slice: batch {args.batch} utterances x ~{args.frames / FRAMES_PER_SEC:.0f} s "
                        f"({int(args.frames * 0.9)}..{int(args.frames * 1.1)} mel frames), n_timesteps={args.nfe}, CFG 0.7, "
                        f"CFM solve + HiFT vocoder",
            "batch_per_gpu": args.batch, "frames": args.frames, "n_timesteps": args.nfe, "precision": args.precision,
            "l2": "no explicit flush: per-step working set (185 MB bf16 weights + >1 GB activations) exceeds the 126 MB L2",
            "weights": "random-init (jyutvoice_b200.synthetic, PyTorch-default statistics)"}


def run_ours(args):
    import torch.distributed as dist
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, HiFTGenerator, synthetic, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a B200; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    est_sd = synthetic.make_estimator_state_dict()
    hift_sd = synthetic.make_hift_state_dict()
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=args.precision))
    cfm.load_state_dict(est_sd, strict=True)
    cfm = cfm.to(dev)
    hift = HiFTGenerator(precision=args.precision)
    hift.load_state_dict(hift_sd, strict=True)
    hift = hift.to(dev)

    lens, Tmax, mu_h, spks_h = make_workload(args.batch, args.frames, 1000 + rank)
    audio_s = sum(lens) / FRAMES_PER_SEC
    mu_pin, spks_pin = mu_h.pin_memory(), spks_h.pin_memory()
    mu_d, spks_d = mu_pin.to(dev), spks_pin.to(dev)
    wav_pin = torch.empty((args.batch, 480 * Tmax), dtype=torch.float32).pin_memory()
    gather_buf = [torch.empty((args.batch, 480 * Tmax), dtype=torch.float32, device=dev) for _ in range(world)] if world > 1 else None

    def step(mu, spks):
        mel, _ = cfm(mu, None, args.nfe, 1.0, spks, None, lengths=lens)
        wav, _ = hift.inference(mel, lengths=lens)
        if world > 1:  # the only collective of the path: gather the waveforms
            dist.all_gather(gather_buf, wav)
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(mu_d, spks_d)
    barrier()

    # ---- timed region 1: device-resident inputs (value)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.jv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(mu_d, spks_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.jv_launch_count() - launches0

    # ---- timed region 2: end to end through the public API with host buffers (e2e)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        mu = mu_pin.to(dev, non_blocking=True)
        spks = spks_pin.to(dev, non_blocking=True)
        wav = step(mu, spks)
        wav_pin.copy_(wav, non_blocking=True)
        torch.cuda.synchronize()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel: events around every tcgen05 GEMM launch, same steps
    kms, kfl, kn = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    roof = None
    if args.precision == "bf16":
        _lib.check(L.jv_profile_begin())
        for _ in range(max(1, min(args.steps, 2))):
            step(mu_d, spks_d)
        _lib.check(L.jv_profile_end(ctypes.byref(kms), ctypes.byref(kfl), ctypes.byref(kn)))
        peaks = load_peaks()
        achieved = kfl.value / (kms.value * 1e-3) / 1e12 if kms.value > 0 else 0.0
        nprof = max(1, min(args.steps, 2))
        traffic, traffic_src = load_traffic()
        roof = {"bound": "tensor", "kernel": "gemm_taps_tc_kernel (tcgen05 bf16, all conv / linear contractions)",
                "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_sustained"], "traffic": traffic, "traffic_unit": "DRAM bytes per launch",
                "traffic_source": traffic_src, "peak_source": peaks["src"] + " sustained bf16",
                "launches_per_step": kn.value / nprof, "kernel_ms_per_step": kms.value / nprof,
                "algo_tflop_per_step": kfl.value / nprof / 1e12, "share_of_step": (kms.value / nprof) / (ms / args.steps)}

    # max over ranks, sum of audio
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        a = torch.tensor([audio_s], dtype=torch.float64, device=dev)
        dist.all_reduce(a, op=dist.ReduceOp.SUM)
        audio_total = float(a[0])
    else:
        audio_total = audio_s

    if rank == 0:
        value = audio_total * args.steps / (ms * 1e-3)
        e2e_value = audio_total * args.steps / (ms_e2e * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": mu_pin.numel() * 4 + spks_pin.numel() * 4,
                    "d2h_bytes_per_step": wav_pin.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "audio_seconds_per_step": audio_total,
        }
        if roof is not None:
            line["roofline"] = roof
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            nb = synthetic.noise_bank()
            cpu_reference_pass(est_sd, hift_sd, nb, 100, args.nfe, 0)  # warm-up
            t0 = time.perf_counter()
            n_utts = 2
            audio = sum(cpu_reference_pass(est_sd, hift_sd, nb, args.frames, args.nfe, 10 + i) for i in range(n_utts))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": audio / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{n_utts} of the {args.batch} utterances ({args.frames} frames each), oracle port, "
                                              f"batch-1 loop, fp32, torch CPU with {cores} threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--nfe", type=int, default=10)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: W >= 3
        run_ours(args)


if __name__ == "__main__":
    main()
