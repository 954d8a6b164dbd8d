#!/usr/bin/env python
"""Headline benchmark: audio-seconds generated per second (1/RTF) of the CFM (10 NFE, CFG) + HiFT hot
path on synthetic utterances of BASELINE.json's throughput config (batch 64 per GPU, ~6 s each, bf16).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (torchrun for N > 1)
  python bench.py --impl reference [--steps K] [--warmup W]      # CPU arm: the reference's own modules (oracle/_ref) on host cores

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


class CpuArm:
    """The reference's CPU implementation of the path, batch-1 loop (its only mode), fp32, torch CPU threads.
    kind "reference": the reference's own modules imported from oracle/_ref (byte-compiled from /root/reference by
    oracle/build_ref.py; the source tree itself when it is present) through oracle/ref_shims.py.
    kind "port": the oracle restatement (oracle/estimator.py, oracle/hift.py), used only when oracle/_ref is absent."""

    def __init__(self, est_sd, hift_sd):
        from jyutvoice_b200 import synthetic
        self.kind = "port"
        self.est_sd, self.hift_sd = est_sd, hift_sd
        self.nb = synthetic.noise_bank()
        try:
            from oracle import ref_shims
            if ref_shims.reference_available():
                self.cfm = ref_shims.build_reference_cfm()
                self.cfm.load_state_dict(est_sd, strict=True)
                self.hift = ref_shims.build_reference_hift()
                self.hift.load_state_dict(hift_sd, strict=True)
                self.kind = "reference"
        except Exception as e:  # fall back to the port, and say why
            print(f"bench.py: reference arm falls back to the oracle port ({type(e).__name__}: {e})", file=sys.stderr)
            self.kind = "port"

    def one_utterance(self, frames, nfe, seed):
        """CFM solve + HiFT inference of one synthetic utterance; returns its audio seconds."""
        g = torch.Generator().manual_seed(seed)
        mu = torch.randn(1, 80, frames, generator=g)
        spks = torch.randn(1, 80, generator=g)
        if self.kind == "reference":
            with torch.inference_mode():
                mel, _ = self.cfm(mu, torch.ones(1, 1, frames), nfe, 1.0, spks, torch.zeros(1, 80, frames))
                self.hift.inference(speech_feat=mel)
        else:
            from oracle import estimator as oe, hift as oh
            with torch.no_grad():
                mel = oe.cfm_forward(self.est_sd, self.nb, mu, torch.ones(1, 1, frames), nfe, 1.0, spks, torch.zeros(1, 80, frames))
                oh.inference(self.hift_sd, mel, oh.draw_source_rng(1, 480 * frames, g))
        return frames / FRAMES_PER_SEC

    def describe(self, n_utts, frames, nfe, cores):
        what = "the reference's own modules (oracle/_ref)" if self.kind == "reference" else "oracle port"
        return (f"{n_utts} utterance(s) of {frames} frames, {what}, batch-1 loop (the reference's only mode), fp32, "
                f"{nfe} NFE + HiFT, torch CPU with {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from jyutvoice_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    arm = CpuArm(synthetic.make_estimator_state_dict(), synthetic.make_hift_state_dict())
    for w in range(args.warmup):
        arm.one_utterance(min(args.frames, 100), args.nfe, w)
    times = []
    audio = 0.0
    for k in range(args.steps):
        t0 = time.perf_counter()
        audio += arm.one_utterance(args.frames, args.nfe, 100 + k)
        times.append(time.perf_counter() - t0)
    dt = sum(times)
    value = audio / dt
    sample = arm.describe(args.steps, args.frames, args.nfe, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / max(1, args.steps), "p50_ms": sorted(times)[len(times) // 2] * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "fp32"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, precision=None):
    return {"workload": f"BASELINE configs[4] per-GPU slice: batch {args.batch} utterances x ~{args.frames / FRAMES_PER_SEC:.0f} s "
                        f"({int(args.frames * 0.9)}..{int(args.frames * 1.1)} mel frames), n_timesteps={args.nfe}, CFG 0.7, "
                        f"CFM solve + HiFT vocoder",
            "batch_per_gpu": args.batch, "frames": args.frames, "n_timesteps": args.nfe, "precision": precision or args.precision,
            "l2": "no explicit flush: per-step working set (185 MB bf16 weights + >1 GB activations) exceeds the 126 MB L2",
            "weights": "random-init (jyutvoice_b200.synthetic, PyTorch-default statistics)"}


def global_workload(world, batch, frames):
    """BASELINE config 5: a fixed set of 512 synthetic utterances, 64 per GPU.  Slice r of the set is
    make_workload(batch, frames, 1000 + r); a run on `world` GPUs takes the first `world` slices and shards them with
    sharding.shard_utterances (at world = 1 that is exactly slice 0 in its own order)."""
    lens, mus, spk = [], [], []
    Tmax = int(frames * 1.1)
    for r in range(world):
        l, _, mu, sp = make_workload(batch, frames, 1000 + r)
        lens += l
        mus.append(mu)
        spk.append(sp)
    return lens, Tmax, torch.cat(mus), torch.cat(spk)


def p50(xs):
    s = sorted(xs)
    return s[len(s) // 2]


def small_latency(precision, est_sd, hift_sd, dev, frames=99, nfe=10, reps=7):
    """BASELINE config 1 (one ~2 s utterance, batch 1, 10 NFE): p50 wall latency of CFM + HiFT through the Python API,
    host tensors in, waveform back on the host."""
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, HiFTGenerator
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=precision))
    cfm.load_state_dict(est_sd, strict=True)
    cfm = cfm.to(dev)
    hift = HiFTGenerator(precision=precision)
    hift.load_state_dict(hift_sd, strict=True)
    hift = hift.to(dev)
    g = torch.Generator().manual_seed(0)
    mu = torch.randn(1, 80, frames, generator=g).pin_memory()
    spks = torch.randn(1, 80, generator=g).pin_memory()
    out = torch.empty((1, 480 * frames)).pin_memory()
    ts = []
    for i in range(reps + 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel, _ = cfm(mu.to(dev, non_blocking=True), None, nfe, 1.0, spks.to(dev, non_blocking=True), None, lengths=[frames])
        wav, _ = hift.inference(mel, lengths=[frames])
        out.copy_(wav, non_blocking=True)
        torch.cuda.synchronize()
        if i >= 2:
            ts.append((time.perf_counter() - t0) * 1e3)
    return p50(ts)


def run_ours(args):
    import torch.distributed as dist
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, HiFTGenerator, synthetic, sharding, _lib

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

    # the sweep's utterance set, sharded across the ranks (weak scaling: 64 utterances per GPU)
    all_lens, Tmax, all_mu, all_spks = global_workload(world, args.batch, args.frames)
    total = len(all_lens)
    plan = sharding.shard_utterances(all_lens, world)
    mine = plan[rank]
    lens = [all_lens[i] for i in mine]
    audio_s = sum(lens) / FRAMES_PER_SEC
    mu_pin, spks_pin = all_mu[mine].contiguous().pin_memory(), all_spks[mine].contiguous().pin_memory()
    del all_mu, all_spks
    mu_d, spks_d = mu_pin.to(dev), spks_pin.to(dev)
    n_mine = len(mine)
    wav_pin = torch.empty((n_mine, 480 * Tmax), dtype=torch.float32).pin_memory()
    wav_lens = torch.tensor([480 * l for l in lens], dtype=torch.int64, device=dev)

    def step(mu, spks):
        mel, _ = cfm(mu, None, args.nfe, 1.0, spks, None, lengths=lens)
        wav, _ = hift.inference(mel, lengths=lens)
        if world > 1:  # the only collective of the path: every rank ends up with all waveforms, in the set's order
            sharding.gather_waveforms(wav, wav_lens, mine, total, 480 * Tmax)
        return wav

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(mu_d, spks_d)
    barrier()

    # ---- timed region 1: device-resident inputs (value); one event per step boundary gives the per-step p50
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.jv_launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    evs[0].record()
    for k in range(args.steps):
        step(mu_d, spks_d)
        evs[k + 1].record()
    barrier()
    ms = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    launches = L.jv_launch_count() - launches0

    # ---- timed region 2: end to end through the public API with host buffers (e2e)
    barrier()
    e2e_ms = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        mu = mu_pin.to(dev, non_blocking=True)
        spks = spks_pin.to(dev, non_blocking=True)
        wav = step(mu, spks)
        wav_pin.copy_(wav, non_blocking=True)
        torch.cuda.synchronize()
        e2e_ms.append((time.perf_counter() - t0) * 1e3)
    barrier()
    ms_e2e = (time.perf_counter() - t_all) * 1e3
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
        # whole-step view: algorithmic FLOPs of the step (BASELINE.md section 4 formulas) over the step time
        step_tflop = (args.nfe * 2 * sum(132161536.0 * t + 114688.0 * t * t for t in lens) + 612304320.0 * sum(lens)) / 1e12
        roof["step_algo_tflop"] = step_tflop
        roof["step_frac"] = step_tflop / (ms / args.steps * 1e-3) / peaks["bf16_sustained"]

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
            "ms_per_step": ms / args.steps, "p50_ms": p50(step_ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": mu_pin.numel() * 4 + spks_pin.numel() * 4,
                    "d2h_bytes_per_step": wav_pin.numel() * 4, "ms_per_step": ms_e2e / args.steps, "p50_ms": p50(e2e_ms)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "audio_seconds_per_step": audio_total,
            "utterances_per_step": total,
            "sharding": "jyutvoice_b200.sharding.shard_utterances over the first 64 x n_gpus of the 512-utterance set; "
                        "gather_waveforms (NCCL all_gather) inside the timed step" if world > 1 else "single GPU: slice 0 of the set",
        }
        if roof is not None:
            line["roofline"] = roof
        if world == 1 and not args.no_extra:
            # (rank 0, single-GPU runs only: the multi-GPU launches stay short)
            extra = {}
            del cfm, hift
            torch.cuda.empty_cache()
            extra["config1_latency_p50_ms"] = {"workload": "BASELINE configs[0]: one utterance of 99 frames (~2 s), batch 1, 10 NFE, "
                                                           "CFM + HiFT, host in / host out",
                                               "bf16": small_latency("bf16", est_sd, hift_sd, dev),
                                               "fp32": small_latency("fp32", est_sd, hift_sd, dev)}
            if args.precision == "bf16":  # the <= 1e-3 / >= 60 dB mode on the headline workload: one timed step
                cfm32 = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="fp32"))
                cfm32.load_state_dict(est_sd, strict=True)
                cfm32 = cfm32.to(dev)
                hift32 = HiFTGenerator(precision="fp32")
                hift32.load_state_dict(hift_sd, strict=True)
                hift32 = hift32.to(dev)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for i in range(2):
                    if i == 1:
                        f0.record()
                    mel, _ = cfm32(mu_d, None, args.nfe, 1.0, spks_d, None, lengths=lens)
                    hift32.inference(mel, lengths=lens)
                f1.record()
                torch.cuda.synchronize()
                ms32 = f0.elapsed_time(f1)
                extra["fp32_mode"] = {"value": audio_s / (ms32 * 1e-3), "unit": UNIT, "ms_per_step": ms32, "steps": 1,
                                      "note": "same workload, precision fp32 (the mel <= 1e-3 / wav >= 60 dB mode)"}
                del cfm32, hift32
            line["extra"] = extra
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            arm = CpuArm(est_sd, hift_sd)
            arm.one_utterance(100, args.nfe, 0)  # warm-up
            t0 = time.perf_counter()
            n_utts = 2
            audio = sum(arm.one_utterance(args.frames, args.nfe, 10 + i) for i in range(n_utts))
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": audio / dt, "unit": UNIT, "cores": cores, "kind": arm.kind,
                                    "sample": arm.describe(n_utts, args.frames, args.nfe, cores)}
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
    ap.add_argument("--no-extra", action="store_true", help="skip the config-1 latency and fp32-mode lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3  # timing rule: W >= 3
        run_ours(args)


if __name__ == "__main__":
    main()
