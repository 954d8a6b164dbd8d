"""Checkpoint formats either side of the path: what the reference reads before `synthesise` and writes after `hift.inference`.

  load_checkpoint        <- jyutvoice/models/jyutvoice_tts.py:73-105 (`load_pretrain`: a bare state_dict or a Lightning
                            {"state_dict": ...} file), infer.py:225-229 / :186-196 (flow_encoder.pt / hift.pt: bare state_dicts)
  split_flow_checkpoint  <- scripts/download_pretrain_weights.py:168-214 (`extract_flow_weights`: CosyVoice2 flow.pt ->
                            flow_encoder.pt + flow_decoder.pt by key prefix).  The download itself needs the network and is
                            out of scope; the split works on a file that is already on disk.
  load_hift              <- infer.py:186-196 (hift.pt into the vocoder)
  write_wav              <- infer.py:441 (`torchaudio.save(path, wav.cpu(), 24000)`): 16-bit PCM RIFF, no torchaudio needed

Host logic only: nothing here computes on the path.  Files are read with `weights_only=True` (tensors and plain
containers; a pickled module is refused) unless the caller opts out.
"""
import os
import struct

import torch

# download_pretrain_weights.py:183-196: which top-level prefixes go to which file
FLOW_ENCODER_PREFIXES = ("encoder.", "input_embedding.", "encoder_proj.")
FLOW_DECODER_PREFIXES = ("decoder.", "spk_embed_affine_layer.")


def load_checkpoint(path, weights_only=True):
    """state_dict of a .pt / .ckpt file; unwraps {"state_dict": ...} (jyutvoice_tts.py:91-98)."""
    if not os.path.exists(path):
        raise FileNotFoundError(f"Pretrain checkpoint not found: {path}")
    ckpt = torch.load(path, map_location="cpu", weights_only=weights_only)
    if isinstance(ckpt, dict) and "state_dict" in ckpt:
        ckpt = ckpt["state_dict"]
    if not isinstance(ckpt, dict):
        raise TypeError(f"{path} does not hold a state_dict")
    return ckpt


def split_flow_state_dict(flow_state_dict):
    """(flow_encoder, flow_decoder) state_dicts; keys keep their full names, anything else is dropped
    (download_pretrain_weights.py:180-196)."""
    enc = {k: v for k, v in flow_state_dict.items() if k.startswith(FLOW_ENCODER_PREFIXES)}
    dec = {k: v for k, v in flow_state_dict.items() if k.startswith(FLOW_DECODER_PREFIXES)}
    return enc, dec


def split_flow_checkpoint(flow_path, output_dir, force=False):
    """Writes <output_dir>/flow_encoder.pt and flow_decoder.pt from a CosyVoice2 flow.pt already on disk; existing
    outputs are kept unless `force` (download_pretrain_weights.py:161-166).  Returns the two paths."""
    enc_path = os.path.join(output_dir, "flow_encoder.pt")
    dec_path = os.path.join(output_dir, "flow_decoder.pt")
    if os.path.exists(enc_path) and os.path.exists(dec_path) and not force:
        return enc_path, dec_path
    enc, dec = split_flow_state_dict(load_checkpoint(flow_path))
    if not dec:
        raise ValueError(f"{flow_path} holds no decoder.* / spk_embed_affine_layer.* weights")
    os.makedirs(output_dir, exist_ok=True)
    torch.save(enc, enc_path)
    torch.save(dec, dec_path)
    return enc_path, dec_path


def load_pretrain(model, path):
    """jyutvoice_tts.py:73-105: strict=False load of whatever the file holds; returns the incompatible keys."""
    return model.load_state_dict(load_checkpoint(path), strict=False)


def load_hift(hift, path):
    """infer.py:186-196: hift.pt is a bare state_dict of HiFTGenerator (strict)."""
    hift.load_state_dict(load_checkpoint(path))
    return hift.eval()


def write_wav(path, wav, sample_rate=24000):
    """16-bit PCM RIFF file from a float waveform in [-1, 1]: [T], [1, T] or [channels, T] (torchaudio.save's layout,
    infer.py:441).  Values are clipped like the vocoder's own clamp (hifigan.py:576)."""
    w = torch.as_tensor(wav).detach().to("cpu", torch.float32)
    if w.dim() == 1:
        w = w.unsqueeze(0)
    if w.dim() != 2:
        raise ValueError("wav must be [T] or [channels, T]")
    n_ch, n = w.shape
    pcm = (w.clamp(-1.0, 1.0) * 32767.0).round().to(torch.int16).t().contiguous().numpy().tobytes()
    header = b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVEfmt " + struct.pack(
        "<IHHIIHH", 16, 1, n_ch, sample_rate, sample_rate * n_ch * 2, n_ch * 2, 16) + b"data" + struct.pack("<I", len(pcm))
    with open(path, "wb") as f:
        f.write(header)
        f.write(pcm)
    return n
