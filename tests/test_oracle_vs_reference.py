"""CPU, this container only: oracle vs the unmodified reference imported live from /root/reference.
Skipped where the reference tree is absent (the GPU box)."""
import pytest
import torch

from oracle import ref_shims, estimator as oe, hift as oh
from jyutvoice_b200 import synthetic as weights
from conftest import snr_db

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present")


def test_estimator_state_dict_is_the_references(est_sd):
    cfm = ref_shims.build_reference_cfm()
    res = cfm.load_state_dict(est_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(cfm.rand_noise, weights.noise_bank())
    g = torch.Generator().manual_seed(123)
    R, T = 2, 19
    x = torch.randn(R, 80, T, generator=g)
    mu = torch.randn(R, 80, T, generator=g)
    spks = torch.randn(R, 80, generator=g)
    cond = torch.randn(R, 80, T, generator=g)
    mask = torch.ones(R, 1, T)
    mask[1, 0, 11:] = 0
    t = torch.tensor([0.3, 0.7])
    with torch.no_grad():
        ref = cfm.estimator(x, mask, mu, t, spks, cond)
        mine = oe.estimator_forward(est_sd, x, mask, mu, t, spks, cond)
    assert (ref - mine).abs().max().item() <= 1e-5


def test_hift_matches_reference(hift_sd):
    hift = ref_shims.build_reference_hift()
    res = hift.load_state_dict(hift_sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    g = torch.Generator().manual_seed(4)
    mel = torch.randn(1, 80, 21, generator=g) * 2 - 5
    with torch.no_grad():
        torch.manual_seed(5)
        wav_r, s_r = hift.inference(mel)
        rng = oh.draw_source_rng(1, 21 * 480, torch.Generator().manual_seed(5))
        wav_m, s_m = oh.inference(hift_sd, mel, rng)
    assert (s_r - s_m).abs().max().item() <= 1e-6
    assert snr_db(wav_r, wav_m) >= 90.0


def test_flow_encoder_matches_reference():
    """oracle/flow_encoder.py against the reference's UpsampleConformerEncoder (+ the two layers of infer.py's FlowEncoder):
    the synthetic weights load strict, and outputs agree for a 1-token prompt, full context and the static chunk mask."""
    from oracle import flow_encoder as ofe
    m = ref_shims.build_reference_flow_encoder()
    sd = weights.make_flow_encoder_state_dict()
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    from jyutvoice.utils.mask import make_pad_mask
    g = torch.Generator().manual_seed(1)
    for T, streaming in ((1, False), (29, False), (58, True)):
        token = torch.randint(0, 6561, (1, T), generator=g)
        tl = torch.tensor([T])
        with torch.no_grad():
            x = m.input_embedding(torch.clamp(token, min=0)) * (~make_pad_mask(tl)).float().unsqueeze(-1)
            hid, masks = m.encoder(x, tl, streaming=streaming)
            h = m.encoder_proj(hid)
            mine_h, mine_hid = ofe.flow_encoder_forward(sd, token, streaming)
        assert masks.shape == (1, 1, 2 * T) and bool(masks.all())
        assert (hid - mine_hid).abs().max().item() <= 1e-5
        assert (h - mine_h).abs().max().item() <= 1e-5
