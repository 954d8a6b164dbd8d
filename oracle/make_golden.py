"""Generate tests/golden/*.npz from the UNMODIFIED reference imported from /root/reference.

Run here (the container that has /root/reference):   python -m oracle.make_golden
The reference cannot travel to the GPU box, so its outputs are committed as small fixtures.
Inputs are re-derived in the tests from the seeds stored in each file; weights come from
jyutvoice_b200/synthetic.py (deterministic) and their fingerprint is stored so a silent RNG difference
between machines is detected instead of producing a bogus parity failure.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from jyutvoice_b200 import synthetic as weights  # noqa: E402  (deterministic random-init weights under the reference keys)

OUT = os.path.join(ROOT, "tests", "golden")


def est_inputs(seed, R, T, lens):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(R, 80, T, generator=g)
    mu = torch.randn(R, 80, T, generator=g)
    spks = torch.randn(R, 80, generator=g)
    cond = torch.randn(R, 80, T, generator=g) * 0.3
    t = torch.rand(R, generator=g)
    mask = torch.zeros(R, 1, T)
    for i, l in enumerate(lens):
        mask[i, 0, :l] = 1
    return x, mask, mu, t, spks, cond


def cfm_inputs(seed, T):
    g = torch.Generator().manual_seed(seed)
    mu = torch.randn(1, 80, T, generator=g)
    spks = torch.randn(1, 80, generator=g)
    return mu, spks


def hift_mel(seed, B, T):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 80, T, generator=g) * 2 - 5


def synth_inputs(seed, Tx):
    """~50 token ids, random ids per stream, random speaker embedding."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(1, 97, (1, Tx), generator=g)
    lang = torch.randint(0, 4, (1, Tx), generator=g)
    tone = torch.randint(0, 7, (1, Tx), generator=g)
    word_pos = torch.randint(0, 4, (1, Tx), generator=g)
    syllable_pos = torch.randint(0, 4, (1, Tx), generator=g)
    spk_embed = torch.randn(1, 192, generator=g)
    return x, torch.tensor([Tx]), lang, tone, word_pos, syllable_pos, spk_embed


def synth_golden():
    """Runs the reference's real TextEncoder + DurationPredictor + synthesise with the synthetic weights of
    jyutvoice_b200/synthetic.py.  The encoder / dp outputs are stored too: they pin oracle/text_encoder.py and the GPU
    text front, and let the CPU tests replay them."""
    import types
    ref_shims.install()
    from jyutvoice.models.jyutvoice_tts import JyutVoiceTTS
    from jyutvoice.models.text_encoder import TextEncoder
    from jyutvoice.models.duration_predictor import DurationPredictor
    enc_params = types.SimpleNamespace(n_feats=80, n_channels=192, filter_channels=768, filter_channels_dp=256, n_heads=2,
                                       n_layers=6, kernel_size=3, p_dropout=0.1, gin_channels=192, prenet=True)
    encoder = TextEncoder(encoder_type="RoPE Encoder", encoder_params=enc_params, n_vocab=97, n_lang=4, n_tone=7)
    dp = DurationPredictor(in_channels=576, filter_channels=256, kernel_size=3, p_dropout=0.1, gin_channels=192)
    # deterministic synthetic weights under the reference's keys (strict load = the key tables are the reference's)
    encoder.load_state_dict(weights.make_text_encoder_state_dict(), strict=True)
    dp.load_state_dict(weights.make_duration_predictor_state_dict(), strict=True)
    cfm = ref_shims.build_reference_cfm()
    cfm.load_state_dict(weights.make_estimator_state_dict(), strict=True)
    tts = JyutVoiceTTS(encoder=encoder, decoder=cfm, dp=dp, output_size=80, spk_embed_dim=192)
    aff = weights.make_spk_affine_state_dict()
    tts.spk_embed_affine_layer.load_state_dict(aff)
    tts.eval()
    for name, seed, Tx, ls, n, prompt in (("synth_c1", 0, 51, 1.0, 10, 0), ("synth_prompt", 1, 21, 2.0, 4, 17)):
        inp = synth_inputs(seed, Tx)
        g = torch.Generator().manual_seed(seed + 500)
        pf = torch.randn(1, prompt, 80, generator=g) if prompt else None
        ph = torch.randn(1, prompt, 80, generator=g) if prompt else None
        with torch.no_grad():
            ex, emu, emask = encoder(*inp)
            logw = dp(ex, emask, inp[-1])
            out = tts.synthesise(*inp, prompt_feat=pf, prompt_h=ph, n_timesteps=n, temperature=1.0, length_scale=ls)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, Tx=Tx, length_scale=ls, n_timesteps=n, prompt=prompt,
                            enc_x=ex.numpy(), enc_mu=emu.numpy(), enc_mask=emask.numpy(), logw=logw.numpy(),
                            mel_lengths=out["mel_lengths"].numpy(), attn=out["attn"].numpy().astype(np.uint8),
                            encoder_outputs=out["encoder_outputs"].numpy(), decoder_outputs=out["decoder_outputs"].numpy())
        print(name, "T =", int(out["mel_lengths"][0]))


def flow_encoder_tokens(seed, T):
    return torch.randint(0, 6561, (1, T), generator=torch.Generator().manual_seed(seed))


def flow_encoder_golden():
    """The reference's speech-token encoder (infer.py:35-82 FlowEncoder = input_embedding + UpsampleConformerEncoder +
    encoder_proj) on two prompts: full context, and the encoder's static chunk mask (streaming=True)."""
    m = ref_shims.build_reference_flow_encoder()
    from jyutvoice.utils.mask import make_pad_mask
    sd = weights.make_flow_encoder_state_dict()
    m.load_state_dict(sd, strict=True)
    out = {"weights_checksum": np.float64(sum(float(v.double().abs().sum()) for v in sd.values())), "n_cases": 2}
    for ci, (seed, T, streaming) in enumerate(((41, 37, False), (43, 61, True))):
        token = flow_encoder_tokens(seed, T)
        token_len = torch.tensor([T])
        with torch.no_grad():   # infer.py:77-82 verbatim
            mask = (~make_pad_mask(token_len)).float().unsqueeze(-1)
            x = m.input_embedding(torch.clamp(token, min=0)) * mask
            hid, masks = m.encoder(x, token_len, streaming=streaming)
            h = m.encoder_proj(hid)
        out.update({f"c{ci}_seed": seed, f"c{ci}_T": T, f"c{ci}_streaming": int(streaming), f"c{ci}_h": h.numpy(),
                    f"c{ci}_hidden": hid.numpy(), f"c{ci}_masks": masks.numpy()})
    np.savez_compressed(os.path.join(OUT, "flow_encoder.npz"), **out)


def module_inputs(seed):
    """Inputs of the per-module fixtures: a ragged 3-row batch of 45 frames."""
    g = torch.Generator().manual_seed(seed)
    lens = [45, 17, 32]
    T = 45
    mask = torch.zeros(3, 1, T)
    for i, l in enumerate(lens):
        mask[i, 0, :l] = 1
    x320 = torch.randn(3, 320, T, generator=g)
    x256 = torch.randn(3, 256, T, generator=g)
    temb = torch.randn(3, 1024, generator=g)
    xh = torch.randn(2, 64, 333, generator=g)  # HiFT stage-2 shaped ResBlock input
    return lens, mask, x320, x256, temb, xh


def module_golden():
    """Per-module outputs of the UNMODIFIED reference submodules: CausalBlock1D, CausalResnetBlock1D, BasicTransformerBlock
    with a ragged key mask, and two HiFT ResBlocks (kernel 3 and 11) — they pin the oracle below the estimator / decode level."""
    ref_shims.install()
    from jyutvoice.utils.mask import add_optional_chunk_mask
    from jyutvoice.utils.common import mask_to_bias
    cfm = ref_shims.build_reference_cfm()
    cfm.load_state_dict(weights.make_estimator_state_dict(), strict=True)
    est = cfm.estimator
    hift = ref_shims.build_reference_hift()
    hift.load_state_dict(weights.make_hift_state_dict(), strict=True)
    lens, mask, x320, x256, temb, xh = module_inputs(31)
    out = {"seed": 31}
    with torch.no_grad():
        resnet0 = est.down_blocks[0][0]
        out["causal_block"] = resnet0.block1(x320, mask).numpy()            # down_blocks.0.0.block1 (320 -> 256)
        out["resnet"] = resnet0(x320, mask, temb).numpy()                   # down_blocks.0.0
        mid = est.mid_blocks[3][0]
        out["resnet_mid"] = mid(x256, mask, temb).numpy()                   # mid_blocks.3.0 (256 -> 256)
        tb = est.mid_blocks[3][1][2]                                        # mid_blocks.3.1.2
        h = x256.transpose(1, 2).contiguous()
        am = add_optional_chunk_mask(h, mask.bool(), False, False, 0, 0, -1).repeat(1, h.size(1), 1)
        out["tblock"] = tb(hidden_states=h, attention_mask=mask_to_bias(am, h.dtype), timestep=None).numpy()
        out["hift_resblock_s2_k3"] = hift.resblocks[6](xh).numpy()          # stage 2, kernel 3
        out["hift_resblock_s2_k11"] = hift.resblocks[8](xh).numpy()         # stage 2, kernel 11
    np.savez_compressed(os.path.join(OUT, "modules.npz"), **out)
    print("modules", {k: getattr(v, "shape", v) for k, v in out.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    torch.manual_seed(0)

    # ---------------- estimator / CFM ----------------
    cfm = ref_shims.build_reference_cfm()
    sd = weights.make_estimator_state_dict()
    cfm.load_state_dict(sd, strict=True)
    cs = weights.checksum(sd)
    with torch.no_grad():
        lens = [37, 20, 29]
        x, mask, mu, t, spks, cond = est_inputs(5, 3, 37, lens)
        v = cfm.estimator(x, mask, mu, t, spks, cond)
        np.savez_compressed(os.path.join(OUT, "estimator_fwd.npz"), seed=5, R=3, T=37, lens=np.array(lens),
                            out=v.numpy(), weight_checksum=np.array(cs))
        for name, seed, T, n in (("cfm_T50_n10", 7, 50, 10), ("cfm_T33_n4", 8, 33, 4)):
            mu, spks = cfm_inputs(seed, T)
            mel, _ = cfm(mu, torch.ones(1, 1, T), n, 1.0, spks, torch.zeros(1, 80, T))
            np.savez_compressed(os.path.join(OUT, name + ".npz"), seed=seed, T=T, n_timesteps=n, out=mel.numpy(),
                                weight_checksum=np.array(cs))
        # with a prompt-style non-zero cond
        mu, spks = cfm_inputs(9, 40)
        g = torch.Generator().manual_seed(90)
        cond = torch.zeros(1, 80, 40)
        cond[:, :, :12] = torch.randn(1, 80, 12, generator=g)
        mel, _ = cfm(mu, torch.ones(1, 1, 40), 5, 0.8, spks, cond)
        np.savez_compressed(os.path.join(OUT, "cfm_T40_n5_cond.npz"), seed=9, T=40, n_timesteps=5, temperature=0.8,
                            out=mel.numpy(), weight_checksum=np.array(cs))
        # streaming=True (static chunk mask, chunk 50): decoder.py:950-953
        lens = [120, 70, 101]
        x, mask, mu, t, spks, cond = est_inputs(15, 3, 120, lens)
        v = cfm.estimator(x, mask, mu, t, spks, cond, streaming=True)
        np.savez_compressed(os.path.join(OUT, "estimator_fwd_stream.npz"), seed=15, R=3, T=120, lens=np.array(lens),
                            out=v.numpy(), weight_checksum=np.array(cs))
        mu, spks = cfm_inputs(17, 130)
        mel, _ = cfm(mu, torch.ones(1, 1, 130), 3, 1.0, spks, torch.zeros(1, 80, 130), streaming=True)
        np.savez_compressed(os.path.join(OUT, "cfm_T130_n3_stream.npz"), seed=17, T=130, n_timesteps=3, out=mel.numpy(),
                            weight_checksum=np.array(cs))
    del cfm

    # ---------------- HiFT ----------------
    hift = ref_shims.build_reference_hift()
    for tag, f0b in (("unvoiced", None), ("voiced", 200.0)):
        sdh = weights.make_hift_state_dict(f0_bias=f0b)
        hift.load_state_dict(sdh, strict=True)
        csh = weights.checksum(sdh)
        B, T = 2, 30
        mel = hift_mel(3, B, T)
        with torch.no_grad():
            f0 = hift.f0_predictor(mel)
            torch.manual_seed(11)
            wav, s = hift.inference(mel)
            wav_d = hift.decode(mel, s)
            sr, si = hift._stft(s.squeeze(1))
        np.savez_compressed(os.path.join(OUT, f"hift_{tag}.npz"), seed=3, B=B, T=T, rng_seed=11,
                            f0=f0.numpy(), s=s.numpy().astype(np.float32), wav_inference=wav.numpy(),
                            wav_decode=wav_d.numpy(), s_stft=torch.cat([sr, si], 1).numpy(),
                            weight_checksum=np.array(csh))
    del hift

    # ---------------- integer length / alignment code ----------------
    ref_shims.install()
    from jyutvoice.utils.model import sequence_mask, generate_path
    from jyutvoice.utils.mask import make_pad_mask
    cases = {}
    g = torch.Generator().manual_seed(21)
    for ci, (B, Tx, ls) in enumerate([(1, 7, 1.0), (3, 11, 1.0), (2, 9, 3.0), (2, 5, 0.5), (1, 1, 1.0)]):
        logw = torch.randn(B, 1, Tx, generator=g) * 0.8 + 0.3
        x_lengths = torch.randint(1, Tx + 1, (B,), generator=g)
        x_lengths[0] = Tx
        x_mask = sequence_mask(x_lengths, Tx).unsqueeze(1).float()
        # jyutvoice_tts.py:184-196
        w = torch.exp(logw) * x_mask
        w_ceil = torch.ceil(w) * ls
        y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
        y_max = y_lengths.max()
        y_mask = sequence_mask(y_lengths, y_max).unsqueeze(1).to(x_mask.dtype)
        attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
        attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
        cases[f"c{ci}_logw"] = logw.numpy()
        cases[f"c{ci}_x_lengths"] = x_lengths.numpy()
        cases[f"c{ci}_length_scale"] = np.float32(ls)
        cases[f"c{ci}_y_lengths"] = y_lengths.numpy()
        cases[f"c{ci}_attn"] = attn.numpy()
        cases[f"c{ci}_pad_mask"] = make_pad_mask(y_lengths).numpy()
        cases[f"c{ci}_y_mask"] = y_mask.numpy()
    cases["n_cases"] = 5
    np.savez_compressed(os.path.join(OUT, "lengths.npz"), **cases)
    # ---------------- per-module fixtures ----------------
    module_golden()
    # ---------------- end to end: the reference's JyutVoiceTTS.synthesise (config 1 of BASELINE.json) ----------------
    synth_golden()
    # ---------------- speech-token encoder (prompt_h) ----------------
    flow_encoder_golden()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
