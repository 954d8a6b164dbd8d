"""GPU (-m gpu): the CUDA path, called through the C-ABI via the Python mirrors, against
 (1) golden vectors made by the unmodified reference, (2) the oracle on the same seeded inputs,
 (3) size-independent properties at BASELINE.json's sizes.
Tolerances (north_star): fp32 mode mel max-abs <= 1e-3, waveform SNR >= 60 dB; integer indexing exact.
bf16 mode ("stated looser bound", BASELINE.md section 5, frozen from B200 measurements): mel max-abs <= 1e-1 and
rel-RMS <= 2e-2; waveform SNR >= 45 dB at decode(x, s)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, snr_db

pytestmark = pytest.mark.gpu

FP32_MEL_TOL = 1e-3
BF16_MEL_TOL = 1e-1
BF16_MEL_RELRMS = 2e-2
FP32_SNR = 60.0
BF16_SNR = 45.0


def rel_rms(x, ref):
    return float((x - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())


@pytest.fixture(scope="module")
def cfms(est_sd):
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder
    out = {}
    for prec in ("fp32", "bf16"):
        cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=prec))
        cfm.load_state_dict(est_sd, strict=True)
        out[prec] = cfm.cuda()
    return out


@pytest.fixture(scope="module")
def hifts(hift_sd, hift_sd_voiced):
    from jyutvoice_b200 import HiFTGenerator
    out = {}
    for prec in ("fp32", "bf16"):
        for tag, sd in (("unvoiced", hift_sd), ("voiced", hift_sd_voiced)):
            h = HiFTGenerator(precision=prec)
            h.load_state_dict(sd, strict=True)
            out[(prec, tag)] = h.cuda()
    return out


# ------------------------------------------------------------------------------ GEMM engines
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("shape", [(128, 256, 64), (300, 256, 256), (1000, 1536, 256), (777, 80, 256), (2048, 256, 1024), (515, 64, 128)])
def test_gemm_engine(prec, shape):
    from jyutvoice_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    C = torch.full((M, N), float("nan"), device="cuda")
    p = lambda z: ctypes.c_void_p(z.data_ptr())
    _lib.check(_lib.lib().jv_test_gemm(_lib.PREC[prec], M, N, K, p(A), p(W), p(b), p(C), None))
    if prec == "bf16":
        ref = A.bfloat16().double() @ W.bfloat16().double().T + b.double()
    else:
        ref = A.double() @ W.double().T + b.double()
    assert (C.double() - ref).abs().max().item() <= 5e-5


# ------------------------------------------------------------------------------ estimator / CFM
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_estimator_forward_golden(cfms, prec):
    from oracle.make_golden import est_inputs
    g = np.load(os.path.join(GOLDEN, "estimator_fwd.npz"))
    x, mask, mu, t, spks, cond = est_inputs(int(g["seed"]), int(g["R"]), int(g["T"]), list(g["lens"]))
    v = cfms[prec].estimator(x.cuda(), mask.cuda(), mu.cuda(), t.cuda(), spks.cuda(), cond.cuda()).cpu()
    ref = torch.from_numpy(g["out"])
    err = (v - ref).abs().max().item()
    if prec == "fp32":
        assert err <= 1e-4
    else:
        assert err <= BF16_MEL_TOL and rel_rms(v, ref) <= BF16_MEL_RELRMS
    assert float(v[1, :, 20:].abs().max()) == 0.0  # padded frames exactly zero, as in the reference


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_streaming_golden(cfms, prec):
    """streaming=True (static chunk mask, chunk 50) against the reference's own streaming outputs, then back to full
    context on the same handle (the chunk is state of the handle)."""
    from oracle.make_golden import est_inputs, cfm_inputs
    g = np.load(os.path.join(GOLDEN, "estimator_fwd_stream.npz"))
    x, mask, mu, t, spks, cond = est_inputs(int(g["seed"]), int(g["R"]), int(g["T"]), list(g["lens"]))
    args = (x.cuda(), mask.cuda(), mu.cuda(), t.cuda(), spks.cuda(), cond.cuda())
    v = cfms[prec].estimator(*args, streaming=True).cpu()
    ref = torch.from_numpy(g["out"])
    err = (v - ref).abs().max().item()
    if prec == "fp32":
        assert err <= 1e-4
    else:
        assert err <= BF16_MEL_TOL and rel_rms(v, ref) <= BF16_MEL_RELRMS
    v_full = cfms[prec].estimator(*args, streaming=False).cpu()
    assert (v_full - ref).abs().max().item() > 1e-2  # the mask is really applied, and really switched off again
    g = np.load(os.path.join(GOLDEN, "cfm_T130_n3_stream.npz"))
    T, n = int(g["T"]), int(g["n_timesteps"])
    mu, spks = cfm_inputs(int(g["seed"]), T)
    mel, _ = cfms[prec](mu.cuda(), torch.ones(1, 1, T).cuda(), n, 1.0, spks.cuda(), torch.zeros(1, 80, T).cuda(), streaming=True)
    ref = torch.from_numpy(g["out"])
    err = (mel.cpu() - ref).abs().max().item()
    if prec == "fp32":
        assert err <= FP32_MEL_TOL
    else:
        assert err <= BF16_MEL_TOL and rel_rms(mel.cpu(), ref) <= BF16_MEL_RELRMS


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["cfm_T33_n4", "cfm_T50_n10", "cfm_T40_n5_cond"])
def test_cfm_golden(cfms, prec, name):
    from oracle.make_golden import cfm_inputs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    T, n = int(g["T"]), int(g["n_timesteps"])
    mu, spks = cfm_inputs(int(g["seed"]), T)
    cond = torch.zeros(1, 80, T)
    temperature = 1.0
    if name.endswith("cond"):
        cond[:, :, :12] = torch.randn(1, 80, 12, generator=torch.Generator().manual_seed(90))
        temperature = float(g["temperature"])
    mel, none = cfms[prec](mu.cuda(), torch.ones(1, 1, T).cuda(), n, temperature, spks.cuda(), cond.cuda())
    assert none is None and mel.dtype == torch.float32
    ref = torch.from_numpy(g["out"])
    err = (mel.cpu() - ref).abs().max().item()
    if prec == "fp32":
        assert err <= FP32_MEL_TOL
    else:
        assert err <= BF16_MEL_TOL and rel_rms(mel.cpu(), ref) <= BF16_MEL_RELRMS


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cfm_ragged_batch_equals_batch1_oracle(cfms, prec, est_sd, noise_bank):
    """The reference is batch-1 only; a batch must give, per utterance, the unpadded B=1 answer."""
    from oracle import estimator as oe
    lens = [41, 1, 17, 30]
    B = len(lens)
    g = torch.Generator().manual_seed(77)
    mu = torch.randn(B, 80, max(lens), generator=g)
    spks = torch.randn(B, 80, generator=g)
    mask = torch.zeros(B, 1, max(lens))
    for i, l in enumerate(lens):
        mask[i, 0, :l] = 1
    mel, _ = cfms[prec](mu.cuda(), mask.cuda(), 3, 1.0, spks.cuda(), None)
    mel = mel.cpu()
    with torch.no_grad():
        ref = oe.cfm_forward_batch(est_sd, noise_bank, mu, lens, 3, 1.0, spks, None)
    tol = FP32_MEL_TOL if prec == "fp32" else BF16_MEL_TOL
    assert (mel - ref).abs().max().item() <= tol
    for i, l in enumerate(lens):
        if l < max(lens):
            assert float(mel[i, :, l:].abs().max()) == 0.0


def test_cfm_batch_invariance_full_size(cfms):
    """BASELINE config 5 shape (64 x ~300 frames, 10 NFE, bf16): an utterance's mel must not depend on
    what else is in the batch (rows are independent), and the run is deterministic."""
    cfm = cfms["bf16"]
    g = torch.Generator().manual_seed(5)
    B, T = 64, 330
    lens = torch.randint(270, 331, (B,), generator=g).tolist()
    mu = torch.randn(B, 80, T, generator=g).cuda()
    spks = torch.randn(B, 80, generator=g).cuda()
    mel, _ = cfm(mu, None, 10, 1.0, spks, None, lengths=lens)
    mel2, _ = cfm(mu, None, 10, 1.0, spks, None, lengths=lens)
    assert torch.equal(mel, mel2)
    assert torch.isfinite(mel).all()
    for i in (0, 37, 63):
        one, _ = cfm(mu[i:i + 1, :, : lens[i]].contiguous(), None, 10, 1.0, spks[i:i + 1], None, lengths=[lens[i]])
        assert torch.equal(one[0], mel[i, :, : lens[i]])


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cfm_long_utterance_config4_shape(cfms, prec, est_sd, noise_bank):
    """BASELINE config 4 shape: a 30 s utterance (T = 1500, 24 key tiles, attention = 57 % of the flops), against the
    batch-1 oracle; one Euler step keeps the CPU side to seconds."""
    from oracle import estimator as oe
    T = 1500
    g = torch.Generator().manual_seed(1500)
    mu = torch.randn(1, 80, T, generator=g)
    spks = torch.randn(1, 80, generator=g)
    mel, _ = cfms[prec](mu.cuda(), None, 1, 1.0, spks.cuda(), None, lengths=[T])
    with torch.no_grad():
        ref = oe.cfm_forward(est_sd, noise_bank, mu, torch.ones(1, 1, T), 1, 1.0, spks, torch.zeros(1, 80, T))
    err = (mel.cpu() - ref).abs().max().item()
    if prec == "fp32":
        assert err <= FP32_MEL_TOL
    else:
        assert err <= BF16_MEL_TOL and rel_rms(mel.cpu(), ref) <= BF16_MEL_RELRMS


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_streaming_ragged_vs_oracle(cfms, prec, est_sd):
    """streaming=True on a ragged batch whose lengths straddle chunk (50) and key-tile (64) boundaries."""
    from oracle import estimator as oe
    from oracle.make_golden import est_inputs
    lens = [330, 50, 129, 64, 51, 200]
    x, mask, mu, t, spks, cond = est_inputs(33, len(lens), max(lens), lens)
    v = cfms[prec].estimator(x.cuda(), mask.cuda(), mu.cuda(), t.cuda(), spks.cuda(), cond.cuda(), streaming=True).cpu()
    with torch.no_grad():
        ref = torch.zeros_like(v)
        for i, l in enumerate(lens):  # the oracle on each utterance alone, unpadded
            ref[i:i + 1, :, :l] = oe.estimator_forward(est_sd, x[i:i + 1, :, :l], mask[i:i + 1, :, :l], mu[i:i + 1, :, :l],
                                                       t[i:i + 1], spks[i:i + 1], cond[i:i + 1, :, :l], chunk=50)
    err = (v - ref).abs().max().item()
    if prec == "fp32":
        assert err <= 1e-4
    else:
        assert err <= BF16_MEL_TOL and rel_rms(v, ref) <= BF16_MEL_RELRMS


def test_split_solve_experimental_path(est_sd, noise_bank):
    """JYUTVOICE_B200_SPLIT=1 (opt-in): a large solve as two concurrent half-batches on two streams / graph branches must
    give bit-identical mels to the default single chain (same kernels, same per-utterance arithmetic)."""
    import subprocess
    import sys
    code = (
        "import torch\n"
        "from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic\n"
        "cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision='bf16'))\n"
        "cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)\n"
        "cfm = cfm.cuda()\n"
        "g = torch.Generator().manual_seed(9)\n"
        "lens = [int(v) for v in torch.randint(270, 331, (64,), generator=g)]\n"
        "mu = torch.randn(64, 80, 330, generator=g).cuda(); spks = torch.randn(64, 80, generator=g).cuda()\n"
        "mel, _ = cfm(mu, None, 4, 1.0, spks, None, lengths=lens)\n"
        "print('SUM', float(mel.double().sum()), float(mel.double().abs().sum()))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for split in ("0", "1"):
        env = dict(os.environ, JYUTVOICE_B200_SPLIT=split, PYTHONPATH=root)
        out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append(out.stdout.split("SUM")[1].split()[:2])
    assert outs[0] == outs[1]


def test_fused_mlp_experimental_path():
    """JYUTVOICE_B200_MLP=1 (opt-in): FF1 + GELU + FF2 + residual + norm in one tcgen05 kernel (csrc/mlp_tc.cuh).  The
    switch is read once per process, hence the subprocess; same golden, same bf16 gate."""
    import subprocess
    import sys
    code = (
        "import os, numpy as np, torch\n"
        "from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic\n"
        "from oracle.make_golden import est_inputs\n"
        "g = np.load(os.path.join('tests', 'golden', 'estimator_fwd.npz'))\n"
        "cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision='bf16'))\n"
        "cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)\n"
        "cfm = cfm.cuda()\n"
        "a = [z.cuda() for z in est_inputs(int(g['seed']), int(g['R']), int(g['T']), list(g['lens']))]\n"
        "v = cfm.estimator(*a).cpu()\n"
        "ref = torch.from_numpy(g['out'])\n"
        "print('ERR', float((v - ref).abs().max()), float((v - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()))\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, JYUTVOICE_B200_MLP="1", JYUTVOICE_B200_FORCE_MODES="1", PYTHONPATH=root)  # (small input: force the fused kernel)
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    err, rel = [float(z) for z in out.stdout.split("ERR")[1].split()[:2]]
    assert err <= BF16_MEL_TOL and rel <= BF16_MEL_RELRMS


def test_mlp_tail_split_experimental_path(tmp_path):
    """JYUTVOICE_B200_MLP_TAIL=1 (opt-in): the pair-tiles that do not fill the last wave of the fused feed-forward are
    split by hidden chunk over all SMs and summed by a row kernel.  Not bit-identical to the default (another summation
    order for those rows), so: the path is taken (more launches), the result is finite and within bf16 noise of the default."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch\n"
        "from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic, _lib\n"
        "cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision='bf16'))\n"
        "cfm.load_state_dict(synthetic.make_estimator_state_dict(), strict=True)\n"
        "cfm = cfm.cuda()\n"
        "g = torch.Generator().manual_seed(9)\n"
        "lens = [int(v) for v in torch.randint(270, 331, (64,), generator=g)]\n"
        "mu = torch.randn(64, 80, 330, generator=g).cuda(); spks = torch.randn(64, 80, generator=g).cuda()\n"
        "n0 = _lib.lib().jv_launch_count()\n"
        "mel, _ = cfm(mu, None, 2, 1.0, spks, None, lengths=lens)\n"
        "torch.cuda.synchronize()\n"
        "np.save(sys.argv[1], mel.cpu().numpy())\n"
        "print('LAUNCHES', _lib.lib().jv_launch_count() - n0)\n"
    )
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mels, launches = [], []
    for tail in ("0", "1"):
        env = dict(os.environ, JYUTVOICE_B200_MLP_TAIL=tail, JYUTVOICE_B200_GRAPH="0", PYTHONPATH=root)
        f = str(tmp_path / f"mel{tail}.npy")
        out = subprocess.run([sys.executable, "-c", code, f], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        mels.append(torch.from_numpy(np.load(f)))
        launches.append(int(out.stdout.split("LAUNCHES")[1].split()[0]))
    assert launches[1] == launches[0] + 2 * 2 * 56      # two extra launches per fused feed-forward, 56 blocks, 2 NFE
    assert torch.isfinite(mels[1]).all()
    assert (mels[0] - mels[1]).abs().max().item() <= BF16_MEL_TOL and rel_rms(mels[1], mels[0]) <= BF16_MEL_RELRMS


def test_cfm_solve_replays_a_graph_and_matches_eager_counts(cfms):
    """Steps 1 .. n-1 of a solve replay one captured Euler step; the kernel-launch counter stays what eager gives."""
    from jyutvoice_b200 import _lib
    L = _lib.lib()
    cfm = cfms["bf16"]
    mu = torch.randn(2, 80, 40, generator=torch.Generator().manual_seed(4)).cuda()
    spks = torch.zeros(2, 80).cuda()
    cfm(mu, None, 2, spks=spks, lengths=[40, 33])  # n < 3: eager
    g0, l0 = L.jv_graph_launch_count(), L.jv_launch_count()
    cfm(mu, None, 2, spks=spks, lengths=[40, 33])
    per2 = L.jv_launch_count() - l0
    assert L.jv_graph_launch_count() == g0
    l1 = L.jv_launch_count()
    cfm(mu, None, 6, spks=spks, lengths=[40, 33])
    per6 = L.jv_launch_count() - l1
    if os.environ.get("JYUTVOICE_B200_GRAPH", "1") != "0":
        assert L.jv_graph_launch_count() == g0 + 5
    per_step = (per6 - per2) // 4
    assert per_step > 250 and (per6 - per2) % 4 == 0  # four more Euler steps, each the same ~275 launches (fused feed-forward)


def test_c_abi_error_codes(cfms):
    """Raw C-ABI calls: status codes and messages instead of exceptions / crashes (include/jyutvoice_b200.h)."""
    from jyutvoice_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda:0")
    h = cfms["fp32"].estimator.handle(dev)
    T = 12
    lens = _lib.i32_array([T])
    mu = torch.zeros(1, 80, T, device=dev)
    spks = torch.zeros(1, 80, device=dev)
    noise = torch.zeros(80, 64, device=dev)
    out = torch.empty(1, 80, T, device=dev)
    tspan = _lib.f32_array([0.0, 0.5, 1.0])
    p = lambda z: ctypes.c_void_p(z.data_ptr())
    need = L.jv_cfm_solve_workspace_bytes(h, 1, lens)
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device=dev)

    def solve(hh, lens_, ws_bytes, n=2, noise_stride=64):
        return L.jv_cfm_solve(hh, 1, T, lens_, p(mu), p(spks), None, p(noise), noise_stride, 1.0, n, tspan, 0.7, p(out), p(ws),
                              ws_bytes, None)

    assert solve(h, lens, need) == _lib.JV_OK
    assert solve(h, lens, 1024) == _lib.JV_ERR_STATE and b"workspace too small" in L.jv_last_error()
    assert solve(h, _lib.i32_array([T + 1]), need) == _lib.JV_ERR_INVALID and b"lens[0]" in L.jv_last_error()
    assert solve(h, _lib.i32_array([0]), need) == _lib.JV_ERR_INVALID
    assert solve(h, lens, need, n=0) == _lib.JV_ERR_INVALID
    assert solve(h, lens, need, noise_stride=4) == _lib.JV_ERR_INVALID and b"noise bank" in L.jv_last_error()
    assert solve(None, lens, need) == _lib.JV_ERR_STATE
    assert L.jv_cfm_solve_workspace_bytes(None, 1, lens) == 0
    assert L.jv_estimator_set_chunk(h, -1) == _lib.JV_ERR_INVALID
    # a fresh handle with no weights: finalize names the first missing key, forward refuses to run
    h2 = ctypes.c_void_p()
    assert L.jv_estimator_create(0, _lib.PREC["fp32"], ctypes.byref(h2)) == _lib.JV_OK
    try:
        assert L.jv_estimator_finalize(h2) == _lib.JV_ERR_STATE and b"time_mlp" in L.jv_last_error() or b"missing" in L.jv_last_error()
        assert solve(h2, lens, need) == _lib.JV_ERR_STATE
        shape = (ctypes.c_int64 * 2)(3, 0)
        w = torch.zeros(3, 3)
        assert L.jv_estimator_set_weight(h2, b"x", ctypes.c_void_p(w.data_ptr()), shape, 2) == _lib.JV_ERR_INVALID  # empty dim
        assert L.jv_estimator_set_weight(h2, None, ctypes.c_void_p(w.data_ptr()), shape, 2) == _lib.JV_ERR_INVALID
    finally:
        L.jv_estimator_destroy(h2)
    assert solve(h, lens, need) == _lib.JV_OK  # the good handle is unaffected


def test_finalize_is_strict_about_keys(est_sd):
    """load_state_dict(strict=True) semantics at the C level: an extra key fails finalize and is named."""
    from jyutvoice_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.jv_estimator_create(0, _lib.PREC["fp32"], ctypes.byref(h)) == _lib.JV_OK
    try:
        _lib.set_weights(h, L.jv_estimator_set_weight, [(k[len("estimator."):], v) for k, v in est_sd.items()])
        extra = torch.zeros(4, 4)
        shape = (ctypes.c_int64 * 2)(4, 4)
        assert L.jv_estimator_set_weight(h, b"mid_blocks.0.1.0.attn1.to_q.lora", ctypes.c_void_p(extra.data_ptr()), shape, 2) == _lib.JV_OK
        assert L.jv_estimator_finalize(h) != _lib.JV_OK
        assert b"to_q.lora" in L.jv_last_error()
    finally:
        L.jv_estimator_destroy(h)


def test_cfm_errors(cfms):
    cfm = cfms["fp32"]
    mu = torch.zeros(2, 80, 10).cuda()
    with pytest.raises(ValueError):
        cfm(mu, None, 2, spks=torch.zeros(2, 80).cuda(), lengths=[10, 11])
    with pytest.raises(ValueError):
        cfm(mu, None, 2, spks=torch.zeros(2, 80).cuda(), lengths=[10, 0])
    # T <= chunk: the streaming chunk mask is a no-op (one chunk covers every key)
    a, _ = cfm(mu, torch.ones(2, 1, 10).cuda(), 2, spks=torch.zeros(2, 80).cuda(), streaming=True)
    b, _ = cfm(mu, torch.ones(2, 1, 10).cuda(), 2, spks=torch.zeros(2, 80).cuda(), streaming=False)
    assert torch.equal(a, b)


# ------------------------------------------------------------------------------ HiFT
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["unvoiced", "voiced"])
def test_hift_golden(hifts, prec, tag):
    from oracle import hift as oh
    from oracle.make_golden import hift_mel
    hift = hifts[(prec, tag)]
    g = np.load(os.path.join(GOLDEN, f"hift_{tag}.npz"))
    B, T = int(g["B"]), int(g["T"])
    mel = hift_mel(int(g["seed"]), B, T)
    f0 = hift.predict_f0(mel.cuda()).cpu()
    f0_ref = torch.from_numpy(g["f0"])
    assert (f0 - f0_ref).abs().max().item() <= (1e-3 if prec == "fp32" else 5e-2) * max(1.0, float(f0_ref.max()) / 100)
    rng = oh.draw_source_rng(B, T * 480, torch.Generator().manual_seed(int(g["rng_seed"])))
    s_ref = torch.from_numpy(g["s"])
    s = hift.source(f0_ref.cuda(), rng).cpu()
    assert (s - s_ref).abs().max().item() <= 1e-5  # fp64 phase prefix == torch's CPU cumsum
    wav = hift.decode(mel.cuda(), s_ref.cuda()).cpu()
    assert wav.shape == (B, 480 * T)
    assert snr_db(torch.from_numpy(g["wav_decode"]), wav) >= (FP32_SNR if prec == "fp32" else BF16_SNR)
    if prec == "fp32":
        wav_i, s_i = hift.inference(mel.cuda(), rng=rng)
        assert s_i.shape == (B, 1, 480 * T)
        assert snr_db(torch.from_numpy(g["wav_inference"]), wav_i.cpu()) >= FP32_SNR


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_hift_ragged_equals_unpadded(hifts, prec, hift_sd):
    """Naive zero padding leaks ~41 dB backwards through the non-causal convs;
    every utterance of a ragged batch must equal its own unpadded decode."""
    from oracle import hift as oh
    hift = hifts[(prec, "unvoiced")]
    g = torch.Generator().manual_seed(8)
    lens = [23, 9, 1, 16]
    B, T = len(lens), max(lens)
    mel = torch.randn(B, 80, T, generator=g) * 2 - 5
    s = torch.randn(B, 1, 480 * T, generator=g) * 0.05
    wav = hift.decode(mel.cuda(), s.cuda(), lengths=lens).cpu()
    for i, l in enumerate(lens):
        with torch.no_grad():
            ref = oh.decode(hift_sd, mel[i:i + 1, :, :l], s[i:i + 1, :, : 480 * l])
        assert snr_db(ref, wav[i:i + 1, : 480 * l]) >= (FP32_SNR if prec == "fp32" else BF16_SNR)
        assert float(wav[i, 480 * l:].abs().sum()) == 0.0


def test_hift_full_size_properties(hifts):
    """BASELINE config 3 shape (32 mels of 80 x 500): finite, clamped, deterministic, batch-invariant."""
    hift = hifts[("bf16", "unvoiced")]
    g = torch.Generator().manual_seed(3)
    mel = (torch.randn(32, 80, 500, generator=g) * 2 - 5).cuda()
    rng = {"phase": torch.zeros(32, 9, 1), "noise": torch.randn(32, 9, 480 * 500, generator=g)}
    wav, s = hift.inference(mel, rng=rng)
    wav2, _ = hift.inference(mel, rng=rng)
    assert torch.equal(wav, wav2)
    assert torch.isfinite(wav).all() and float(wav.abs().max()) <= 0.99 + 1e-6
    one, _ = hift.inference(mel[5:6], rng={"phase": rng["phase"][5:6], "noise": rng["noise"][5:6]})
    # not bit-equal by design: large batches run the convs in slab mode (K blocks outer, taps inner), small ones tap-major,
    # so the fp32 accumulation order differs; the utterance itself must not depend on its neighbours beyond that
    assert snr_db(one[0], wav[5]) >= 55.0


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_padded_tensors_longer_than_every_utterance(cfms, hifts, prec, hift_sd, est_sd, noise_bank):
    """The caller's tensors may be padded beyond the longest utterance (T = 48 > max(lens) = 41): the workspace queries
    see only the lengths, so nothing sized by them may use the tensor stride (regression: 8-GPU bench, ranks whose random
    lengths did not reach the padded width)."""
    from oracle import estimator as oe, hift as oh
    lens = [41, 30]
    T = 48
    g = torch.Generator().manual_seed(48)
    mu = torch.randn(2, 80, T, generator=g)
    spks = torch.randn(2, 80, generator=g)
    mel, _ = cfms[prec](mu.cuda(), None, 2, 1.0, spks.cuda(), None, lengths=lens)
    with torch.no_grad():
        ref = oe.cfm_forward_batch(est_sd, noise_bank, mu, lens, 2, 1.0, spks, None)
    assert (mel.cpu() - ref).abs().max().item() <= (FP32_MEL_TOL if prec == "fp32" else BF16_MEL_TOL)
    assert float(mel[:, :, 41:].abs().max()) == 0.0
    hift = hifts[(prec, "voiced")]
    x = (torch.randn(2, 80, T, generator=g) * 2 - 5)
    wav, s = hift.inference(x.cuda(), lengths=lens)
    assert wav.shape == (2, 480 * T) and s.shape == (2, 1, 480 * T)
    for i, l in enumerate(lens):  # each utterance equals its own unpadded call (same RNG draws are not shared: compare decode)
        one = hift.decode(x[i:i + 1, :, :l].contiguous().cuda(), s[i:i + 1, :, : 480 * l].contiguous())
        both = hift.decode(x.cuda(), s, lengths=lens)
        assert snr_db(one.cpu(), both[i:i + 1, : 480 * l].cpu()) >= (FP32_SNR if prec == "fp32" else 55.0)
        assert float(both[i, 480 * l:].abs().max()) == 0.0


def test_inference_draws_rng_like_the_reference(hifts):
    """Default rng: Uniform on CPU, randn on the device, in the reference's order (generator.py:155-158,171,235)."""
    hift = hifts[("fp32", "unvoiced")]
    mel = (torch.randn(1, 80, 12) * 2 - 5).cuda()
    torch.manual_seed(123)
    w1, s1 = hift.inference(mel)
    torch.manual_seed(123)
    w2, s2 = hift.inference(mel)
    assert torch.equal(w1, w2) and torch.equal(s1, s2)
    torch.manual_seed(124)
    w3, _ = hift.inference(mel)
    assert not torch.equal(w1, w3)


# ------------------------------------------------------------------------------ headline sizes against the oracle
def _oracle_cfm_utts(est_sd, noise_bank, lens, mu, spks, idx, nfe):
    """The oracle (B = 1, unpadded: the reference's only mode) on a few utterances of a batch."""
    from oracle import estimator as oe
    out = {}
    with torch.no_grad():
        for i in idx:
            T = lens[i]
            out[i] = oe.cfm_forward(est_sd, noise_bank, mu[i:i + 1, :, :T], torch.ones(1, 1, T), nfe, 1.0, spks[i:i + 1],
                                    torch.zeros(1, 80, T))
    return out


@pytest.fixture(scope="module")
def headline_oracle(est_sd, noise_bank):
    """bench.py's own batch (BASELINE config 5 per-GPU slice: 64 utterances of 270..330 frames, 10 NFE): the oracle's mel
    for three of them (~5 s of CPU each), shared by the fp32 and bf16 cases."""
    from bench import make_workload
    lens, Tmax, mu, spks = make_workload(64, 300, 1000)
    idx = (0, 31, 63)
    return lens, Tmax, mu, spks, _oracle_cfm_utts(est_sd, noise_bank, lens, mu, spks, idx, 10)


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_headline_config5_vs_oracle(cfms, prec, headline_oracle):
    """The benchmarked configuration itself (B = 64, ~300 frames, 10 NFE, CFG): GPU mel against the B = 1 oracle."""
    lens, Tmax, mu, spks, ref = headline_oracle
    mel, _ = cfms[prec](mu.cuda(), None, 10, 1.0, spks.cuda(), None, lengths=lens)
    mel = mel.cpu()
    assert torch.isfinite(mel).all()
    for i, r in ref.items():
        got = mel[i:i + 1, :, : lens[i]]
        err = (got - r).abs().max().item()
        if prec == "fp32":
            assert err <= FP32_MEL_TOL, (i, err)
        else:
            assert err <= BF16_MEL_TOL and rel_rms(got, r) <= BF16_MEL_RELRMS, (i, err, rel_rms(got, r))
        assert float(mel[i, :, lens[i]:].abs().max()) == 0.0
    if prec == "bf16":
        assert cfms[prec].estimator.saturation_count() == 0


def test_config2_vs_oracle(cfms, est_sd, noise_bank):
    """BASELINE config 2: batch 16 of ~300 frames, n_timesteps = 10, CFG, bf16 on one GPU; two utterances against the oracle."""
    from bench import make_workload
    lens, Tmax, mu, spks = make_workload(16, 300, 2000)
    ref = _oracle_cfm_utts(est_sd, noise_bank, lens, mu, spks, (3, 12), 10)
    mel, _ = cfms["bf16"](mu.cuda(), None, 10, 1.0, spks.cuda(), None, lengths=lens)
    mel = mel.cpu()
    for i, r in ref.items():
        got = mel[i:i + 1, :, : lens[i]]
        assert (got - r).abs().max().item() <= BF16_MEL_TOL and rel_rms(got, r) <= BF16_MEL_RELRMS


@pytest.mark.parametrize("prec", ["bf16", "fp32"])
def test_config3_hift_vs_oracle(hifts, prec, hift_sd):
    """BASELINE config 3: HiFT alone on 32 mels of 80 x 500; two of them against oracle.decode with the same source."""
    from oracle import hift as oh
    hift = hifts[(prec, "unvoiced")]
    g = torch.Generator().manual_seed(3)
    mel = torch.randn(32, 80, 500, generator=g) * 2 - 5
    rng = oh.draw_source_rng(32, 480 * 500, g)
    wav, s = hift.inference(mel.cuda(), rng=rng)
    wav, s = wav.cpu(), s.cpu()
    for i in (1, 30):
        with torch.no_grad():
            ref = oh.decode(hift_sd, mel[i:i + 1], s[i:i + 1])
        assert snr_db(ref, wav[i:i + 1]) >= (FP32_SNR if prec == "fp32" else BF16_SNR)


# ------------------------------------------------------------------------------ rows a4 / a9 directly
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_time_embedding_vs_oracle(cfms, prec, est_sd):
    """SinusoidalPosEmb -> TimestepEmbedding -> per-resnet Mish -> Linear (decoder.py:15-30, 127-171, 101-103), fp32 in
    both modes: the table the solver builds once per solve."""
    import torch.nn.functional as F
    from oracle import estimator as oe
    t = oe.t_span_cosine(10)[:-1]
    got = cfms[prec].estimator.time_embedding(t.cuda()).cpu()
    with torch.no_grad():
        temb = oe.time_embedding(est_sd, t)
        names = ["down_blocks.0.0"] + [f"mid_blocks.{i}.0" for i in range(12)] + ["up_blocks.0.0"]
        ref = torch.stack([F.linear(F.mish(temb), est_sd[f"estimator.{n}.mlp.1.weight"], est_sd[f"estimator.{n}.mlp.1.bias"])
                           for n in names], dim=1)
    assert got.shape == ref.shape == (10, 14, 256)
    assert (got - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
@pytest.mark.parametrize("tag", ["unvoiced", "voiced"])
def test_stft_golden(hifts, prec, tag):
    """HiFTGenerator._stft (generator.py:371-381) against the reference's own s_stft; a ragged batch equals the unpadded
    call (per-utterance reflect padding at the right edge)."""
    hift = hifts[(prec, tag)]
    g = np.load(os.path.join(GOLDEN, f"hift_{tag}.npz"))
    s = torch.from_numpy(g["s"])
    ref = torch.from_numpy(g["s_stft"])
    re, im = hift._stft(s.squeeze(1).cuda())
    got = torch.cat([re, im], dim=1).cpu()
    assert got.shape == ref.shape
    tol = 1e-5 if prec == "fp32" else 4e-3 * float(ref.abs().max())  # bf16 mode stores s_stft as the bf16 conv operand
    assert (got - ref).abs().max().item() <= tol
    lens = [30, 11]
    re2, im2 = hift._stft(s.squeeze(1).cuda(), lengths=lens)
    one_re, one_im = hift._stft(s[1:2, 0, : 480 * 11].contiguous().cuda())
    assert torch.equal(re2[1:2, :, : 120 * 11 + 1], one_re) and torch.equal(im2[1:2, :, : 120 * 11 + 1], one_im)
    assert float(re2[1, :, 120 * 11 + 1:].abs().max()) == 0.0
    assert torch.equal(re2[0], re[0])


# ------------------------------------------------------------------------------ 16-bit residual stream under harder weights
@pytest.mark.parametrize("scale", [1.0, 4.0])
def test_reference_init_weights_through_bf16_mode(scale, noise_bank):
    """Weights drawn the way the reference's own constructor does (kaiming-normal, std sqrt(2 / fan_in), zero biases:
    decoder.py:414-430): bf16 mode (fp16 residual stream) must stay inside its bound relative to the oracle without
    saturating the stream.  The 4x-scaled stress set does exceed the fp16 range (measured on B200): there the counter
    must see it, and the next call -- which the mirror runs with an fp32 stream -- must be back inside the bf16 bound."""
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic
    from oracle import estimator as oe
    sd = synthetic.make_estimator_state_dict(seed=99, init="reference", scale=scale)
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="bf16"))
    cfm.load_state_dict(sd, strict=True)
    cfm = cfm.cuda()
    lens = [150, 97]
    g = torch.Generator().manual_seed(5)
    mu = torch.randn(2, 80, 150, generator=g)
    spks = torch.randn(2, 80, generator=g)
    mel, _ = cfm(mu.cuda(), None, 4, 1.0, spks.cuda(), None, lengths=lens)
    mel = mel.cpu()
    with torch.no_grad():
        ref = oe.cfm_forward_batch(sd, noise_bank, mu, lens, 4, 1.0, spks, None)
    scale_ref = max(1.0, float(ref.abs().max()) / 6.0)  # the bound is stated for mels of abs-max ~6
    print(f"reference-init x{scale}: ref abs-max {float(ref.abs().max()):.2f}, err {(mel - ref).abs().max().item():.3e}, "
          f"rel-rms {rel_rms(mel, ref):.3e}, saturated stores {cfm.estimator.saturation_count()}")
    assert torch.isfinite(mel).all()
    if scale == 1.0:
        assert cfm.estimator.saturation_count() == 0
        assert (mel - ref).abs().max().item() <= BF16_MEL_TOL * scale_ref and rel_rms(mel, ref) <= BF16_MEL_RELRMS
        # what the 16-bit stream itself costs: the same bf16 contractions with an fp32 stream
        cfm.estimator.set_stream_format("fp32")
        mel32, _ = cfm(mu.cuda(), None, 4, 1.0, spks.cuda(), None, lengths=lens)
        print(f"  fp16 stream vs fp32 stream (both bf16 contractions): rel-rms {rel_rms(mel, mel32.cpu()):.3e}")
        assert rel_rms(mel, mel32.cpu()) <= BF16_MEL_RELRMS
    else:
        n_sat = cfm.estimator.saturation_count()
        assert n_sat > 0
        import warnings
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            mel, _ = cfm(mu.cuda(), None, 4, 1.0, spks.cuda(), None, lengths=lens)
        assert any("fp32" in str(x.message) for x in w) and cfm.estimator.stream_format == "fp32"
        assert cfm.estimator.saturation_count() == n_sat  # nothing is clipped any more
        mel = mel.cpu()
        # No accuracy claim for this set: with 4x weights the attention logits are 16x larger, the soft-max is nearly
        # one-hot and bf16 operand rounding flips its winners -- any bf16 implementation decorrelates from fp32 here
        # (measured rel-RMS 0.87 with an fp32 stream as well).  What is asserted is the mechanism: detection, fallback, finite.
        print(f"  fp32 stream: err {(mel - ref).abs().max().item():.3e}, rel-rms {rel_rms(mel, ref):.3e}")
        assert torch.isfinite(mel).all()


def test_fp16_stream_saturation_is_detected_and_falls_back(noise_bank):
    """Weights scaled until the fp16 stream clips: the counter must see it, the mirror must warn and switch the stream
    to fp32 on the next call, and that call must be finite and un-clipped (close to the oracle in relative terms)."""
    import warnings
    from jyutvoice_b200 import CausalConditionalCFM, CausalConditionalDecoder, synthetic
    from oracle import estimator as oe
    sd = synthetic.make_estimator_state_dict(seed=99, init="reference", scale=1.0)
    for k in list(sd):  # blow up the residual branches only: to_out / ff.net.2 write straight into the stream
        if k.endswith("attn1.to_out.0.weight") or k.endswith("ff.net.2.weight"):
            sd[k] = sd[k] * 30000.0
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="bf16"))
    cfm.load_state_dict(sd, strict=True)
    cfm = cfm.cuda()
    g = torch.Generator().manual_seed(6)
    mu = torch.randn(1, 80, 64, generator=g)
    spks = torch.randn(1, 80, generator=g)
    cfm(mu.cuda(), None, 1, 1.0, spks.cuda(), None, lengths=[64])
    assert cfm.estimator.saturation_count() > 0
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        mel, _ = cfm(mu.cuda(), None, 1, 1.0, spks.cuda(), None, lengths=[64])
    assert any("fp32" in str(x.message) for x in w)
    assert cfm.estimator.stream_format == "fp32"
    with torch.no_grad():
        ref = oe.cfm_forward(sd, noise_bank, mu, torch.ones(1, 1, 64), 1, 1.0, spks, torch.zeros(1, 80, 64))
    assert torch.isfinite(mel).all()
    assert rel_rms(mel.cpu(), ref) <= 5e-2


def test_no_ffma_fallback_in_bf16_mode(cfms, hifts):
    """bf16 mode must run every contraction of both graphs on the tcgen05 kernel: an unsupported shape would be lowered to
    the FFMA engine (~50x slower) and counted."""
    from jyutvoice_b200 import _lib
    L = _lib.lib()
    n0 = L.jv_simt_fallback_count()
    mu = torch.randn(2, 80, 70, generator=torch.Generator().manual_seed(4)).cuda()
    mel, _ = cfms["bf16"](mu, None, 2, spks=torch.zeros(2, 80).cuda(), lengths=[70, 33])
    hifts[("bf16", "voiced")].inference(mel, lengths=[70, 33])
    assert L.jv_simt_fallback_count() == n0


# ------------------------------------------------------------------------------ text front
@pytest.fixture(scope="module")
def text_front():
    from jyutvoice_b200 import TextEncoder, DurationPredictor, synthetic
    params = dict(n_feats=80, n_channels=192, filter_channels=768, filter_channels_dp=256, n_heads=2, n_layers=6, kernel_size=3,
                  p_dropout=0.1, gin_channels=192, prenet=True)
    enc = TextEncoder("RoPE Encoder", params, n_vocab=97, n_lang=4, n_tone=7)
    enc.load_state_dict(synthetic.make_text_encoder_state_dict(), strict=True)
    dp = DurationPredictor(in_channels=576, filter_channels=256, kernel_size=3, p_dropout=0.1, gin_channels=192)
    dp.load_state_dict(synthetic.make_duration_predictor_state_dict(), strict=True)
    return enc.cuda(), dp.cuda()


def test_text_encoder_and_duration_predictor_vs_oracle(text_front):
    """A ragged batch (the reference's synthesise is batch-1) against oracle/text_encoder.py: fp32 FFMA, 1e-4."""
    from jyutvoice_b200 import synthetic
    from oracle import text_encoder as ot
    enc, dp = text_front
    g = torch.Generator().manual_seed(3)
    B, Tx = 4, 57
    ri = lambda hi: torch.randint(0, hi, (B, Tx), generator=g)
    x, lang, tone, wp, sp = torch.randint(1, 97, (B, Tx), generator=g), ri(4), ri(7), ri(4), ri(4)
    spk = torch.randn(B, 192, generator=g)
    xl = torch.tensor([57, 1, 23, 40])
    hx, mu, mask = enc(x.cuda(), xl.cuda(), lang.cuda(), tone.cuda(), wp.cuda(), sp.cuda(), spk.cuda())
    logw = dp(hx, mask, spk.cuda())
    esd, dsd = synthetic.make_text_encoder_state_dict(), synthetic.make_duration_predictor_state_dict()
    with torch.no_grad():
        rx, rmu, rmask = ot.text_encoder_forward(esd, x, xl, lang, tone, wp, sp, spk)
        rlogw = ot.duration_predictor_forward(dsd, rx, rmask, spk)
    assert torch.equal(mask.cpu(), rmask)
    assert (hx.cpu() - rx).abs().max().item() <= 1e-4
    assert (mu.cpu() - rmu).abs().max().item() <= 1e-4
    assert (logw.cpu() - rlogw).abs().max().item() <= 1e-4
    for b, l in enumerate(xl.tolist()):  # exactly zero beyond each utterance, as the reference's `* x_mask`
        if l < Tx:
            assert float(hx[b, :, l:].abs().max()) == 0.0 and float(mu[b, :, l:].abs().max()) == 0.0 and float(logw[b, :, l:].abs().max()) == 0.0


def test_length_regulator_bit_exact():
    """The GPU length regulator against the integer fixtures of the reference's own code (lengths.npz): y_lengths, the
    alignment map and the gathered mu_y, including length_scale != 1, single-token and ragged cases."""
    from jyutvoice_b200 import length_regulate
    from jyutvoice_b200.text import attn_from_frame_token
    g = np.load(os.path.join(GOLDEN, "lengths.npz"))
    for ci in range(int(g["n_cases"])):
        logw = torch.from_numpy(g[f"c{ci}_logw"])
        xl = torch.from_numpy(g[f"c{ci}_x_lengths"])
        B, _, Tx = logw.shape
        x_mask = (torch.arange(Tx)[None, :] < xl[:, None]).unsqueeze(1).float()
        mu_x = torch.randn(B, 80, Tx, generator=torch.Generator().manual_seed(ci)) * x_mask
        mu_y, y_len, ft, _ = length_regulate(logw.cuda(), x_mask.cuda(), mu_x.cuda(), float(g[f"c{ci}_length_scale"]))
        assert np.array_equal(y_len.cpu().numpy(), g[f"c{ci}_y_lengths"])
        attn = attn_from_frame_token(ft, Tx).cpu()
        ref_attn = torch.from_numpy(g[f"c{ci}_attn"]).float()
        assert torch.equal(attn, ref_attn)
        ref_mu_y = torch.matmul(ref_attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
        assert torch.equal(mu_y.cpu(), ref_mu_y)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["synth_c1", "synth_prompt"])
def test_synthesise_from_tokens_golden(name, precision, text_front, est_sd):
    """End to end from token ids (text encoder, duration predictor, length regulator, CFM all on the GPU) against the
    reference's real synthesise(): integer outputs exact, encoder_outputs to fp32 round-off, mel inside the mode's bound."""
    from jyutvoice_b200 import JyutVoiceTTS, CausalConditionalCFM, CausalConditionalDecoder, synthetic
    from oracle.make_golden import synth_inputs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    enc, dp = text_front
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=precision))
    cfm.load_state_dict(est_sd, strict=True)
    tts = JyutVoiceTTS(encoder=enc, decoder=cfm, dp=dp)
    tts.spk_embed_affine_layer.load_state_dict(synthetic.make_spk_affine_state_dict())
    tts = tts.cuda()
    inp = synth_inputs(int(g["seed"]), int(g["Tx"]))
    prompt = int(g["prompt"])
    gg = torch.Generator().manual_seed(int(g["seed"]) + 500)
    pf = torch.randn(1, prompt, 80, generator=gg) if prompt else None
    ph = torch.randn(1, prompt, 80, generator=gg) if prompt else None
    cu = lambda z: None if z is None else z.cuda()
    out = tts.synthesise(*[cu(z) for z in inp], prompt_feat=cu(pf), prompt_h=cu(ph), n_timesteps=int(g["n_timesteps"]),
                         length_scale=float(g["length_scale"]))
    assert np.array_equal(out["mel_lengths"].cpu().numpy(), g["mel_lengths"])
    assert np.array_equal(out["attn"].cpu().numpy().astype(np.uint8), g["attn"])
    assert (out["encoder_outputs"].cpu() - torch.from_numpy(g["encoder_outputs"])).abs().max().item() <= 1e-4
    ref = torch.from_numpy(g["decoder_outputs"])
    err = (out["decoder_outputs"].cpu() - ref).abs().max().item()
    assert err <= (FP32_MEL_TOL if precision == "fp32" else BF16_MEL_TOL), err


# ------------------------------------------------------------------------------ speech-token encoder
@pytest.fixture(scope="module")
def flow_encoder():
    from jyutvoice_b200 import FlowEncoder, synthetic
    enc = FlowEncoder()
    enc.load_state_dict(synthetic.make_flow_encoder_state_dict(), strict=True)
    return enc.cuda()


FLOWENC_TOL = 2e-4   # fp32 (3xTF32 contractions, 10 layers, |hidden| ~ 5): measured ~2e-5


def test_flow_encoder_golden_gpu(flow_encoder):
    """Token ids -> prompt_h against the reference's own FlowEncoder outputs: full context and the static chunk mask."""
    from oracle.make_golden import flow_encoder_tokens
    g = np.load(os.path.join(GOLDEN, "flow_encoder.npz"))
    for ci in range(int(g["n_cases"])):
        T = int(g[f"c{ci}_T"])
        token = flow_encoder_tokens(int(g[f"c{ci}_seed"]), T).cuda()
        h, masks = flow_encoder(token, torch.tensor([T]).cuda(), streaming=bool(g[f"c{ci}_streaming"]))
        assert h.shape == (1, 2 * T, 80) and np.array_equal(masks.cpu().numpy(), g[f"c{ci}_masks"])
        err = (h.cpu() - torch.from_numpy(g[f"c{ci}_h"])).abs().max().item()
        assert err <= FLOWENC_TOL, (ci, err)


def test_flow_encoder_ragged_batch_vs_oracle(flow_encoder):
    """A ragged batch (the reference's call site is batch 1) against the oracle utterance by utterance, including a 1-token
    prompt, negative (clamped) ids, and exact zeros beyond each utterance; the encoder-only module on features gives the
    same hidden states as the token path."""
    from jyutvoice_b200 import UpsampleConformerEncoder, synthetic
    from oracle import flow_encoder as ofe
    sd = synthetic.make_flow_encoder_state_dict()
    g = torch.Generator().manual_seed(7)
    lens = [75, 1, 40, 66]
    B, T = len(lens), max(lens)
    token = torch.randint(0, 6561, (B, T), generator=g)
    token[2, 3] = -1
    for streaming in (False, True):
        h, masks = flow_encoder(token.cuda(), torch.tensor(lens).cuda(), streaming=streaming)
        h = h.cpu()
        for b, l in enumerate(lens):
            with torch.no_grad():
                ref, _ = ofe.flow_encoder_forward(sd, token[b:b + 1, :l], streaming)
            err = (h[b:b + 1, : 2 * l] - ref).abs().max().item()
            assert err <= FLOWENC_TOL, (streaming, b, err)
            assert float(h[b, 2 * l:].abs().max()) == 0.0 if 2 * l < 2 * T else True
            assert int(masks[b].sum()) == 2 * l
    enc = UpsampleConformerEncoder()
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    enc = enc.cuda()
    xs = sd["input_embedding.weight"][token.clamp(min=0)].cuda()
    hid, masks2 = enc(xs, torch.tensor(lens).cuda())
    for b, l in enumerate(lens):
        with torch.no_grad():
            _, ref_hid = ofe.flow_encoder_forward(sd, token[b:b + 1, :l], False)
        assert (hid[b:b + 1, : 2 * l].cpu() - ref_hid).abs().max().item() <= FLOWENC_TOL
    assert torch.equal(masks2, masks)


def test_flow_encoder_errors(flow_encoder):
    from jyutvoice_b200 import UpsampleConformerEncoder
    with pytest.raises(ValueError):
        flow_encoder(torch.zeros(2, 5, dtype=torch.long).cuda(), torch.tensor([5, 6]).cuda())      # length beyond T
    with pytest.raises(RuntimeError):
        flow_encoder.cpu()(torch.zeros(1, 5, dtype=torch.long), torch.tensor([5]))                 # no CPU path
    flow_encoder.cuda()
    with pytest.raises(ValueError):
        UpsampleConformerEncoder(macaron_style=True)
    enc = UpsampleConformerEncoder().cuda()                                                        # zero weights: still runs
    out, _ = enc(torch.zeros(1, 3, 512).cuda(), torch.tensor([3]).cuda())
    assert torch.isfinite(out).all()


def test_prompt_h_from_tokens_feeds_synthesise(flow_encoder, text_front, est_sd):
    """The voice-cloning call sequence of infer.py:386-432 on the GPU: speech tokens -> FlowEncoder -> prompt_h -> synthesise;
    the generated part has the regulated length and the result is finite."""
    from jyutvoice_b200 import JyutVoiceTTS, CausalConditionalCFM, CausalConditionalDecoder, synthetic
    from oracle.make_golden import synth_inputs
    enc, dp = text_front
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="bf16"))
    cfm.load_state_dict(est_sd, strict=True)
    tts = JyutVoiceTTS(encoder=enc, decoder=cfm, dp=dp)
    tts.spk_embed_affine_layer.load_state_dict(synthetic.make_spk_affine_state_dict())
    tts = tts.cuda()
    Tp = 30
    token = torch.randint(0, 6561, (1, Tp), generator=torch.Generator().manual_seed(2)).cuda()
    prompt_h, _ = flow_encoder(token, torch.tensor([Tp]).cuda())
    prompt_feat = torch.randn(1, 2 * Tp, 80, generator=torch.Generator().manual_seed(3)).cuda()
    inp = [z.cuda() for z in synth_inputs(11, 20)]
    out = tts.synthesise(*inp, prompt_feat=prompt_feat, prompt_h=prompt_h, n_timesteps=4)
    assert out["decoder_outputs"].shape[-1] == int(out["mel_lengths"][0]) and torch.isfinite(out["decoder_outputs"]).all()
