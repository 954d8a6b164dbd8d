"""JyutVoiceTTS.synthesise mirror against golden vectors of the reference's real synthesise()
(real TextEncoder + DurationPredictor; their outputs are stored in the fixture and replayed by stubs, because
those two modules are outside the hot path and their 25 M weights cannot be committed)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import GOLDEN


class ReplayEncoder(nn.Module):
    n_feats = 80

    def __init__(self, g):
        super().__init__()
        self.x, self.mu, self.mask = (torch.from_numpy(g[k]) for k in ("enc_x", "enc_mu", "enc_mask"))

    def forward(self, x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed):
        d = x.device
        return self.x.to(d), self.mu.to(d), self.mask.to(d)


class ReplayDP(nn.Module):
    def __init__(self, g):
        super().__init__()
        self.logw = torch.from_numpy(g["logw"])

    def forward(self, x, x_mask, spk_embed):
        return self.logw.to(x.device)


def build(g, precision, est_sd):
    from jyutvoice_b200 import JyutVoiceTTS, CausalConditionalCFM, CausalConditionalDecoder, synthetic
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision=precision))
    cfm.load_state_dict(est_sd, strict=True)
    tts = JyutVoiceTTS(encoder=ReplayEncoder(g), decoder=cfm, dp=ReplayDP(g))
    tts.spk_embed_affine_layer.load_state_dict(synthetic.make_spk_affine_state_dict())
    return tts


def inputs(g):
    from oracle.make_golden import synth_inputs
    inp = synth_inputs(int(g["seed"]), int(g["Tx"]))
    prompt = int(g["prompt"])
    gg = torch.Generator().manual_seed(int(g["seed"]) + 500)
    pf = torch.randn(1, prompt, 80, generator=gg) if prompt else None
    ph = torch.randn(1, prompt, 80, generator=gg) if prompt else None
    return inp, pf, ph


@pytest.mark.parametrize("name", ["synth_c1", "synth_prompt"])
def test_length_regulation_bit_exact_cpu(name, est_sd):
    """Integer indexing (mel_lengths, attn, gathered encoder_outputs) with the CFM call stubbed out: runs on CPU."""
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    tts = build(g, "fp32", est_sd)
    seen = {}

    def fake_decoder(mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, streaming=False, lengths=None):
        seen["lengths"] = lengths
        seen["cond"] = cond
        return torch.zeros_like(mu), None
    tts.decoder.forward = fake_decoder
    inp, pf, ph = inputs(g)
    out = tts.synthesise(*inp, prompt_feat=pf, prompt_h=ph, n_timesteps=int(g["n_timesteps"]), length_scale=float(g["length_scale"]))
    assert out["mel_lengths"].dtype == torch.int64
    assert np.array_equal(out["mel_lengths"].numpy(), g["mel_lengths"])
    assert np.array_equal(out["attn"].numpy().astype(np.uint8), g["attn"])
    assert np.array_equal(out["encoder_outputs"].numpy(), g["encoder_outputs"])
    assert seen["lengths"] == [int(g["prompt"]) + int(g["mel_lengths"][0])]
    assert set(out) == {"encoder_outputs", "decoder_outputs", "attn", "mel", "mel_lengths", "rtf"}
    assert out["decoder_outputs"].shape[-1] == int(g["mel_lengths"][0])


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["synth_c1", "synth_prompt"])
def test_synthesise_matches_reference(name, precision, est_sd):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    tts = build(g, precision, est_sd).cuda()
    inp, pf, ph = inputs(g)
    cu = lambda z: None if z is None else z.cuda()
    out = tts.synthesise(*[cu(z) for z in inp], prompt_feat=cu(pf), prompt_h=cu(ph), n_timesteps=int(g["n_timesteps"]),
                         length_scale=float(g["length_scale"]))
    assert np.array_equal(out["mel_lengths"].cpu().numpy(), g["mel_lengths"])
    assert np.array_equal(out["attn"].cpu().numpy().astype(np.uint8), g["attn"])
    ref = torch.from_numpy(g["decoder_outputs"])
    err = (out["decoder_outputs"].cpu() - ref).abs().max().item()
    assert err <= (1e-3 if precision == "fp32" else 1e-1), err
    assert out["mel"] is out["decoder_outputs"] and out["rtf"] > 0


@pytest.mark.gpu
def test_synthesise_batched_equals_single(est_sd):
    """The reference raises for batch != 1 (jyutvoice_tts.py:205-211); the batch must equal per-utterance calls."""
    g = np.load(os.path.join(GOLDEN, "synth_c1.npz"))
    from jyutvoice_b200 import JyutVoiceTTS, CausalConditionalCFM, CausalConditionalDecoder

    class Enc(nn.Module):
        n_feats = 80

        def forward(self, x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed):
            gen = torch.Generator().manual_seed(3)
            B, Tx = x.shape
            mask = (torch.arange(Tx)[None, :] < x_lengths.cpu()[:, None]).float().unsqueeze(1)
            h = torch.randn(4, 16, 12, generator=gen)[:B, :, :Tx]
            mu = torch.randn(4, 80, 12, generator=gen)[:B, :, :Tx] * mask
            return h.to(x.device), mu.to(x.device), mask.to(x.device)

    class DP(nn.Module):
        def forward(self, x, x_mask, spk_embed):
            gen = torch.Generator().manual_seed(4)
            return (torch.rand(4, 1, 12, generator=gen)[: x.shape[0], :, : x.shape[2]] * 1.5).to(x.device)

    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder(precision="fp32"))
    cfm.load_state_dict(est_sd, strict=True)
    tts = JyutVoiceTTS(encoder=Enc(), decoder=cfm, dp=DP()).cuda()
    B, Tx = 3, 12
    x = torch.ones(B, Tx, dtype=torch.long).cuda()
    xl = torch.tensor([12, 7, 10]).cuda()
    spk = torch.randn(B, 192, generator=torch.Generator().manual_seed(9)).cuda()
    out = tts.synthesise(x, xl, x, x, x, x, spk, n_timesteps=2)
    # utterance 0 alone (full-length tokens: the stub encoder is deterministic in its first row)
    one = tts.synthesise(x[:1], xl[:1], x[:1], x[:1], x[:1], x[:1], spk[:1], n_timesteps=2)
    L0 = int(out["mel_lengths"][0])
    assert int(one["mel_lengths"][0]) == L0
    assert (out["decoder_outputs"][0, :, :L0] - one["decoder_outputs"][0, :, :L0]).abs().max().item() <= 1e-5
    for b in range(B):
        Lb = int(out["mel_lengths"][b])
        assert float(out["decoder_outputs"][b, :, Lb:].abs().sum()) == 0.0
