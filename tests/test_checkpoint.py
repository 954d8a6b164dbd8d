"""Checkpoint formats either side of the path (jyutvoice_b200/checkpoint.py): CPU only, synthetic files."""
import struct

import numpy as np
import pytest
import torch

from jyutvoice_b200 import (CausalConditionalCFM, CausalConditionalDecoder, DurationPredictor, HiFTGenerator, JyutVoiceTTS,
                            TextEncoder, checkpoint, synthetic)

ENC = dict(n_feats=80, n_channels=192, filter_channels=768, n_heads=2, n_layers=6, kernel_size=3, gin_channels=192, prenet=True,
           p_dropout=0.1)


def full_state_dict():
    sd = {}
    sd.update({"encoder." + k: v for k, v in synthetic.make_text_encoder_state_dict().items()})
    sd.update({"dp." + k: v for k, v in synthetic.make_duration_predictor_state_dict().items()})
    sd.update({"decoder." + k: v for k, v in synthetic.make_estimator_state_dict().items()})
    sd.update({"spk_embed_affine_layer." + k: v for k, v in synthetic.make_spk_affine_state_dict().items()})
    return sd


def make_tts(**kw):
    cfm = CausalConditionalCFM(estimator=CausalConditionalDecoder())
    return JyutVoiceTTS(encoder=TextEncoder("RoPE Encoder", ENC, 97, 4, 7), decoder=cfm, dp=DurationPredictor(576, 256, 3, 0.1, 192), **kw)


@pytest.mark.parametrize("wrapped", [False, True])
def test_load_pretrain_both_layouts(tmp_path, wrapped):
    """jyutvoice_tts.py:91-105: bare state_dict and Lightning {"state_dict": ...}; strict=False reports, not raises."""
    sd = full_state_dict()
    sd["training_only.step"] = torch.zeros(1)
    dropped = next(k for k in sd if k.startswith("dp."))
    kept = {k: v for k, v in sd.items() if k != dropped}
    path = tmp_path / "pretrain.pt"
    torch.save({"state_dict": kept, "epoch": 3} if wrapped else kept, path)
    tts = make_tts()
    res = tts.load_pretrain(str(path))
    assert res.missing_keys == [dropped] and res.unexpected_keys == ["training_only.step"]
    got = tts.state_dict()
    for k, v in kept.items():
        if k != "training_only.step":
            assert torch.equal(got[k], v), k
    tts2 = make_tts(pretrain_path=str(path))          # constructor route (jyutvoice_tts.py:55-57)
    assert torch.equal(tts2.state_dict()["decoder.estimator.final_proj.weight"], sd["decoder.estimator.final_proj.weight"])


def test_load_pretrain_missing_file(tmp_path):
    with pytest.raises(FileNotFoundError):
        make_tts().load_pretrain(str(tmp_path / "nope.pt"))
    with pytest.raises(FileNotFoundError):
        make_tts(pretrain_path=str(tmp_path / "nope.pt"))


def test_split_flow_checkpoint(tmp_path):
    """download_pretrain_weights.py:168-214: CosyVoice2 flow.pt -> flow_encoder.pt + flow_decoder.pt by prefix."""
    flow = {"decoder." + k: v for k, v in synthetic.make_estimator_state_dict().items()}
    flow.update({"spk_embed_affine_layer." + k: v for k, v in synthetic.make_spk_affine_state_dict().items()})
    flow.update({"encoder.embed.out.0.weight": torch.randn(4, 4), "input_embedding.weight": torch.randn(8, 4),
                 "encoder_proj.weight": torch.randn(4, 4), "length_regulator.model.0.weight": torch.randn(2, 2)})
    src = tmp_path / "flow.pt"
    torch.save(flow, src)
    enc_path, dec_path = checkpoint.split_flow_checkpoint(str(src), str(tmp_path / "out"))
    enc, dec = torch.load(enc_path, weights_only=True), torch.load(dec_path, weights_only=True)
    assert sorted(enc) == ["encoder.embed.out.0.weight", "encoder_proj.weight", "input_embedding.weight"]
    assert all(k.startswith(("decoder.", "spk_embed_affine_layer.")) for k in dec) and len(dec) == len(flow) - 4
    # the decoder file loads into the TTS mirror the way the reference's transfer learning does (strict=False)
    tts = make_tts()
    res = tts.load_pretrain(dec_path)
    assert not res.unexpected_keys and all(k.startswith(("encoder.", "dp.")) for k in res.missing_keys)
    # existing outputs are kept unless forced (:161-166)
    torch.save({"decoder.x": torch.zeros(1)}, src)
    assert checkpoint.split_flow_checkpoint(str(src), str(tmp_path / "out")) == (enc_path, dec_path)
    assert len(torch.load(dec_path, weights_only=True)) == len(dec)
    checkpoint.split_flow_checkpoint(str(src), str(tmp_path / "out"), force=True)
    assert list(torch.load(dec_path, weights_only=True)) == ["decoder.x"]
    torch.save({"encoder.x": torch.zeros(1)}, src)
    with pytest.raises(ValueError):
        checkpoint.split_flow_checkpoint(str(src), str(tmp_path / "out2"))


def test_load_hift_strict(tmp_path):
    sd = synthetic.make_hift_state_dict()
    torch.save(sd, tmp_path / "hift.pt")
    h = checkpoint.load_hift(HiFTGenerator(), str(tmp_path / "hift.pt"))
    assert not h.training and all(torch.equal(h.state_dict()[k], v) for k, v in sd.items())
    del sd[next(iter(sd))]
    torch.save(sd, tmp_path / "bad.pt")
    with pytest.raises(RuntimeError):
        checkpoint.load_hift(HiFTGenerator(), str(tmp_path / "bad.pt"))


def test_pickled_module_refused(tmp_path):
    torch.save(torch.nn.Linear(2, 2), tmp_path / "module.pt")
    with pytest.raises(Exception):
        checkpoint.load_checkpoint(str(tmp_path / "module.pt"))


@pytest.mark.parametrize("shape", [(4801,), (1, 4801), (2, 333)])
def test_write_wav_roundtrip(tmp_path, shape):
    import wave
    g = torch.Generator().manual_seed(3)
    wav = (torch.rand(shape, generator=g) * 2.4 - 1.2)          # some samples beyond the clip
    path = str(tmp_path / "out.wav")
    n = checkpoint.write_wav(path, wav, 24000)
    with wave.open(path, "rb") as f:
        assert (f.getframerate(), f.getsampwidth(), f.getnframes()) == (24000, 2, n)
        n_ch = f.getnchannels()
        pcm = np.frombuffer(f.readframes(n), dtype="<i2").reshape(n, n_ch).T
    ref = (wav.reshape(n_ch, -1).clamp(-1, 1) * 32767).round().short().numpy()
    assert n == shape[-1] and np.array_equal(pcm, ref)
    assert struct.unpack("<I", open(path, "rb").read(8)[4:])[0] == 36 + 2 * n * n_ch
