"""Drop-in mirror of the reference's top-level inference API.

  JyutVoiceTTS.synthesise  <- jyutvoice/models/jyutvoice_tts.py:108-253

Same constructor arguments and the same returned dict.  `encoder` / `dp` are jyutvoice_b200.TextEncoder /
DurationPredictor (text.py: the whole text front runs on the GPU, batched) or any modules with the reference's
signatures (e.g. the reference's own classes: they then run as the caller's PyTorch code in front of the path).
Length regulation: on CUDA the durations -> lengths -> alignment -> mu_y chain is three small kernels and a
gather (text.length_regulate: no [B,T,Tx] x [B,Tx,80] matmul); on CPU tensors the reference's torch ops are
kept one for one (host logic, used by the CPU tests).  Either way `mel_lengths` / `attn` are bit-exact.
The CFM solve runs on the sm_100a library.  Supersets: any batch size (the reference raises for
batch != 1: jyutvoice_tts.py:205-211) and `prompt_feat=None` / `prompt_h=None` defaults.
"""
import datetime as dt

import torch
import torch.nn as nn
from torch.nn import functional as F

from .flow_matching import CausalConditionalCFM
from . import checkpoint as _checkpoint
from . import text as _text


def sequence_mask(length, max_length=None):
    """jyutvoice/utils/model.py:7-11"""
    if max_length is None:
        max_length = length.max()
    x = torch.arange(max_length, dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def generate_path(duration, mask):
    """jyutvoice/utils/model.py:29-46: hard monotonic alignment from integer durations."""
    b, t_x, t_y = mask.shape
    cum_duration = torch.cumsum(duration, 1)
    path = sequence_mask(cum_duration.view(b * t_x), t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - F.pad(path, (0, 0, 1, 0, 0, 0))[:, :-1]
    return path * mask


def make_pad_mask(lengths, max_len=0):
    """jyutvoice/utils/mask.py:232-255"""
    max_len = max_len if max_len > 0 else int(lengths.max().item())
    seq = torch.arange(0, max_len, dtype=torch.int64, device=lengths.device)
    return seq.unsqueeze(0).expand(lengths.size(0), max_len) >= lengths.unsqueeze(-1)


class JyutVoiceTTS(nn.Module):
    def __init__(self, encoder, decoder, dp, output_size=80, spk_embed_dim=192, freeze_encoder=False,
                 freeze_decoder=False, optimizer=None, scheduler=None, pretrain_path=None, warmup_steps=100):
        super().__init__()
        if not isinstance(decoder, CausalConditionalCFM):
            raise TypeError("decoder must be a jyutvoice_b200.CausalConditionalCFM")
        self.encoder = encoder
        self.decoder = decoder
        self.dp = dp
        self.n_feats = getattr(encoder, "n_feats", output_size)
        self.spk_embed_affine_layer = torch.nn.Linear(spk_embed_dim, output_size)
        self.output_size = output_size
        self.freeze_decoder = freeze_decoder
        self.freeze_encoder = freeze_encoder
        if pretrain_path:
            self.load_pretrain(pretrain_path)

    def load_pretrain(self, pretrain_path):
        """jyutvoice_tts.py:73-105: bare or {"state_dict": ...} checkpoint, strict=False; returns the incompatible keys."""
        return _checkpoint.load_pretrain(self, pretrain_path)

    @torch.inference_mode()
    def synthesise(self, x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed, prompt_feat=None, prompt_h=None,
                   n_timesteps=10, temperature=1.0, length_scale=1.0):
        t0 = dt.datetime.now()
        # jyutvoice_tts.py:175-176
        c = self.spk_embed_affine_layer(F.normalize(spk_embed, dim=1))
        # :179-182
        x, mu_x, x_mask = self.encoder(x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed)
        logw = self.dp(x, x_mask, spk_embed)
        if logw.device.type == "cuda":
            # :184-203 as kernels: ceil / cumsum / lengths, then every mel frame gathers its token's 80-vector
            mu_y, y_lengths, frame_token, _ = _text.length_regulate(logw, x_mask, mu_x, length_scale)
            attn = _text.attn_from_frame_token(frame_token, x_mask.shape[-1], x_mask.dtype)
            y_max_length = y_lengths.max()
        else:
            # :184-187 (scale applied after the ceil; .long() truncates)
            w = torch.exp(logw) * x_mask
            w_ceil = torch.ceil(w) * length_scale
            y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
            y_max_length = y_lengths.max()
            # :190-196
            y_mask = sequence_mask(y_lengths, y_max_length).unsqueeze(1).to(x_mask.dtype)
            attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
            attn = generate_path(w_ceil.squeeze(1), attn_mask.squeeze(1)).unsqueeze(1)
            # :199-203
            mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
        encoder_outputs = mu_y[:, :, :y_max_length]
        B = mu_y.shape[0]
        lens = [int(v) for v in y_lengths.cpu()]
        # :213-229
        if prompt_feat is not None and prompt_h is not None:
            mel_len1 = prompt_feat.shape[1]
            mu_y = torch.cat([prompt_h.transpose(1, 2).to(mu_y.dtype), mu_y], dim=2)
            conds = torch.zeros([B, mu_y.shape[2], self.output_size], device=mu_y.device, dtype=mu_y.dtype)
            conds[:, :mel_len1] = prompt_feat
            conds = conds.transpose(1, 2).contiguous()
            lens = [mel_len1 + l for l in lens]
        else:
            mel_len1 = 0
            conds = None  # zeros (jyutvoice_tts.py:228)
        # :232-240
        decoder_outputs, _ = self.decoder(mu=mu_y.contiguous(), mask=None, spks=c, cond=conds, n_timesteps=n_timesteps,
                                          temperature=temperature, streaming=False, lengths=lens)
        decoder_outputs = decoder_outputs[:, :, mel_len1:]
        t = (dt.datetime.now() - t0).total_seconds()
        rtf = t * 24000 / (decoder_outputs.shape[-1] * 480)
        return {"encoder_outputs": encoder_outputs, "decoder_outputs": decoder_outputs, "attn": attn,
                "mel": decoder_outputs, "mel_lengths": y_lengths, "rtf": rtf}
