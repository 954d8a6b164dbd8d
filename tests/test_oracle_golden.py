"""CPU: the oracle restatement against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle (the reference's own
tests hold nothing for this path)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, snr_db
from oracle import estimator as oe, hift as oh, lengths as ol
from jyutvoice_b200 import synthetic as weights
from oracle.make_golden import est_inputs, cfm_inputs, hift_mel


def _check_weights(sd, stored):
    s, a = weights.checksum(sd)
    assert abs(s - stored[0]) <= 1e-6 * abs(stored[1]) and abs(a - stored[1]) <= 1e-9 * abs(stored[1]), \
        "weights drawn here differ from the ones the golden was made with"


def test_table_sizes(est_sd, hift_sd):
    assert len(est_sd) == 910
    assert sum(v.numel() for v in est_sd.values()) == 71302480  # reference README.md:171,233
    assert len(hift_sd) == 328
    assert sum(v.numel() for v in hift_sd.values()) == 20821295


def test_estimator_forward_golden(est_sd):
    g = np.load(os.path.join(GOLDEN, "estimator_fwd.npz"))
    _check_weights(est_sd, g["weight_checksum"])
    x, mask, mu, t, spks, cond = est_inputs(int(g["seed"]), int(g["R"]), int(g["T"]), list(g["lens"]))
    with torch.no_grad():
        v = oe.estimator_forward(est_sd, x, mask, mu, t, spks, cond)
    assert (v - torch.from_numpy(g["out"])).abs().max().item() <= 2e-5
    # padded output frames are exactly zero
    assert float(v[1, :, 20:].abs().max()) == 0.0


def test_streaming_golden(est_sd, noise_bank):
    """streaming=True: static chunk mask (chunk 50) in every attention (decoder.py:950-953, mask.py:91-126)."""
    g = np.load(os.path.join(GOLDEN, "estimator_fwd_stream.npz"))
    x, mask, mu, t, spks, cond = est_inputs(int(g["seed"]), int(g["R"]), int(g["T"]), list(g["lens"]))
    with torch.no_grad():
        v = oe.estimator_forward(est_sd, x, mask, mu, t, spks, cond, chunk=50)
        v_full = oe.estimator_forward(est_sd, x, mask, mu, t, spks, cond)
    ref = torch.from_numpy(g["out"])
    assert (v - ref).abs().max().item() <= 2e-5
    assert (v_full - ref).abs().max().item() > 1e-2  # the chunk mask matters at T = 120
    g = np.load(os.path.join(GOLDEN, "cfm_T130_n3_stream.npz"))
    T, n = int(g["T"]), int(g["n_timesteps"])
    mu, spks = cfm_inputs(int(g["seed"]), T)
    with torch.no_grad():
        mel = oe.cfm_forward(est_sd, noise_bank, mu, torch.ones(1, 1, T), n, 1.0, spks, torch.zeros(1, 80, T), chunk=50)
    assert (mel - torch.from_numpy(g["out"])).abs().max().item() <= 1e-4


def test_module_goldens(est_sd, hift_sd):
    """Per-module fixtures from the reference's own submodules: CausalBlock1D, two ResNet blocks,
    a transformer block with a ragged key mask, two HiFT ResBlocks."""
    from oracle.make_golden import module_inputs
    g = np.load(os.path.join(GOLDEN, "modules.npz"))
    lens, mask, x320, x256, temb, xh = module_inputs(int(g["seed"]))
    p = "estimator."
    with torch.no_grad():
        cb = oe.causal_block(est_sd, p + "down_blocks.0.0.block1", x320, mask)
        rn = oe.resnet(est_sd, p + "down_blocks.0.0", x320, mask, temb)
        rm = oe.resnet(est_sd, p + "mid_blocks.3.0", x256, mask, temb)
        tb = oe.tblock(est_sd, p + "mid_blocks.3.1.2", x256.transpose(1, 2).contiguous(), oe.attn_bias(mask, 45))
        r3 = oh.resblock(hift_sd, "resblocks.6", xh, 3)
        r11 = oh.resblock(hift_sd, "resblocks.8", xh, 11)
    for name, got in (("causal_block", cb), ("resnet", rn), ("resnet_mid", rm), ("tblock", tb), ("hift_resblock_s2_k3", r3),
                      ("hift_resblock_s2_k11", r11)):
        ref = torch.from_numpy(g[name])
        assert got.shape == ref.shape, name
        assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), name
    assert float(cb[1, :, 17:].abs().max()) == 0.0  # CausalBlock1D re-applies the mask


@pytest.mark.parametrize("name", ["cfm_T33_n4", "cfm_T50_n10"])
def test_cfm_golden(est_sd, noise_bank, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    T, n = int(g["T"]), int(g["n_timesteps"])
    mu, spks = cfm_inputs(int(g["seed"]), T)
    with torch.no_grad():
        mel = oe.cfm_forward(est_sd, noise_bank, mu, torch.ones(1, 1, T), n, 1.0, spks, torch.zeros(1, 80, T))
    assert (mel - torch.from_numpy(g["out"])).abs().max().item() <= 1e-4


def test_cfm_golden_cond(est_sd, noise_bank):
    g = np.load(os.path.join(GOLDEN, "cfm_T40_n5_cond.npz"))
    mu, spks = cfm_inputs(9, 40)
    gg = torch.Generator().manual_seed(90)
    cond = torch.zeros(1, 80, 40)
    cond[:, :, :12] = torch.randn(1, 80, 12, generator=gg)
    with torch.no_grad():
        mel = oe.cfm_forward(est_sd, noise_bank, mu, torch.ones(1, 1, 40), 5, 0.8, spks, cond)
    assert (mel - torch.from_numpy(g["out"])).abs().max().item() <= 1e-4


@pytest.mark.parametrize("tag", ["unvoiced", "voiced"])
def test_hift_golden(tag, hift_sd, hift_sd_voiced):
    sd = hift_sd if tag == "unvoiced" else hift_sd_voiced
    g = np.load(os.path.join(GOLDEN, f"hift_{tag}.npz"))
    _check_weights(sd, g["weight_checksum"])
    B, T = int(g["B"]), int(g["T"])
    mel = hift_mel(int(g["seed"]), B, T)
    s_ref = torch.from_numpy(g["s"])
    with torch.no_grad():
        f0 = oh.f0_predict(sd, mel)
        assert (f0 - torch.from_numpy(g["f0"])).abs().max().item() <= 1e-3 * max(1.0, float(g["f0"].max()))
        st = oh.stft(s_ref.squeeze(1))
        assert (st - torch.from_numpy(g["s_stft"])).abs().max().item() <= 1e-5
        wav = oh.decode(sd, mel, s_ref)
        assert snr_db(torch.from_numpy(g["wav_decode"]), wav) >= 90.0
        rng = oh.draw_source_rng(B, T * 480, torch.Generator().manual_seed(int(g["rng_seed"])))
        wav_i, s = oh.inference(sd, mel, rng)
        # source: fp32 cumsum phase makes this seam looser when voiced
        assert (s - s_ref).abs().max().item() <= (1e-6 if tag == "unvoiced" else 2e-4)
        assert snr_db(torch.from_numpy(g["wav_inference"]), wav_i) >= (90.0 if tag == "unvoiced" else 60.0)


def test_lengths_golden():
    g = np.load(os.path.join(GOLDEN, "lengths.npz"))
    for ci in range(int(g["n_cases"])):
        logw = g[f"c{ci}_logw"]
        x_lengths = g[f"c{ci}_x_lengths"]
        Tx = logw.shape[-1]
        x_mask = ol.sequence_mask(x_lengths, Tx)[:, None, :].astype(np.float32)
        # exp() is taken by torch in the product; feed the same float32 exp here
        w_ceil, y_lengths = ol.durations(torch.from_numpy(logw).numpy(), x_mask, float(g[f"c{ci}_length_scale"]))
        w_t = (torch.ceil(torch.exp(torch.from_numpy(logw)) * torch.from_numpy(x_mask)) * float(g[f"c{ci}_length_scale"]))
        y_t = torch.clamp_min(w_t.sum([1, 2]), 1).long().numpy()
        assert np.array_equal(y_t, g[f"c{ci}_y_lengths"])
        y_lengths = y_t
        w_ceil = w_t.numpy()
        y_mask = ol.sequence_mask(y_lengths, int(y_lengths.max()))[:, None, :].astype(np.float32)
        assert np.array_equal(y_mask, g[f"c{ci}_y_mask"])
        assert np.array_equal(ol.make_pad_mask(y_lengths), g[f"c{ci}_pad_mask"])
        attn_mask = x_mask[:, 0, :, None] * y_mask[:, 0, None, :]
        path = ol.generate_path(w_ceil[:, 0], attn_mask)
        assert np.array_equal(path[:, None], g[f"c{ci}_attn"])


@pytest.mark.parametrize("name", ["synth_c1", "synth_prompt"])
def test_text_front_golden(name):
    """oracle/text_encoder.py against the reference's real TextEncoder / DurationPredictor outputs (synthetic weights)."""
    from oracle import text_encoder as ot
    from oracle.make_golden import synth_inputs
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    inp = synth_inputs(int(g["seed"]), int(g["Tx"]))
    esd, dsd = weights.make_text_encoder_state_dict(), weights.make_duration_predictor_state_dict()
    with torch.no_grad():
        x, mu, mask = ot.text_encoder_forward(esd, *inp)
        logw = ot.duration_predictor_forward(dsd, x, mask, inp[-1])
        mu_y, y_len, attn = ot.regulate(logw, mask, mu, float(g["length_scale"]))
    assert (x - torch.from_numpy(g["enc_x"])).abs().max().item() <= 1e-5
    assert (mu - torch.from_numpy(g["enc_mu"])).abs().max().item() <= 1e-5
    assert (logw - torch.from_numpy(g["logw"])).abs().max().item() <= 1e-5
    assert np.array_equal(y_len.numpy(), g["mel_lengths"])
    assert np.array_equal(attn.numpy().astype(np.uint8), g["attn"])
    assert (mu_y - torch.from_numpy(g["encoder_outputs"])).abs().max().item() <= 1e-5


def test_flow_encoder_golden():
    """oracle/flow_encoder.py against the reference's speech-token encoder outputs (synthetic weights): full context and
    the static chunk mask."""
    from oracle import flow_encoder as ofe
    from oracle.make_golden import flow_encoder_tokens
    g = np.load(os.path.join(GOLDEN, "flow_encoder.npz"))
    sd = weights.make_flow_encoder_state_dict()
    assert abs(sum(float(v.double().abs().sum()) for v in sd.values()) - float(g["weights_checksum"])) <= 1e-6 * float(g["weights_checksum"])
    for ci in range(int(g["n_cases"])):
        token = flow_encoder_tokens(int(g[f"c{ci}_seed"]), int(g[f"c{ci}_T"]))
        with torch.no_grad():
            h, hid = ofe.flow_encoder_forward(sd, token, bool(g[f"c{ci}_streaming"]))
        assert (hid - torch.from_numpy(g[f"c{ci}_hidden"])).abs().max().item() <= 1e-5
        assert (h - torch.from_numpy(g[f"c{ci}_h"])).abs().max().item() <= 1e-5
