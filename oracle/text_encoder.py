"""ORACLE (test infrastructure, never shipped): fp32 CPU restatement of the text front of `synthesise`.

Restates, as plain functions over a state_dict, what these reference sites compute:
  * TextEncoder.forward                jyutvoice/models/text_encoder.py:401-451
      embeddings (:418-426), ConvReluNorm prenet (:31-83), [phoneme | speaker | language] concat (:439-447),
      Encoder: 6 x (MultiHeadAttention with partial RoPE, LayerNorm, FFN, LayerNorm) (:284-337, 172-260, 86-169), proj (:450-451)
  * channel LayerNorm (eps 1e-4)       text_encoder.py:12-28, duration_predictor.py:5-23
  * DurationPredictor.forward          jyutvoice/models/duration_predictor.py:48-60
  * durations -> alignment -> mu_y     jyutvoice/models/jyutvoice_tts.py:184-203 (with oracle/lengths.py for the integer part)

Pinned by tests/golden/synth_*.npz (enc_x, enc_mu, logw made by the UNMODIFIED reference with the synthetic weights of
jyutvoice_b200/synthetic.py) and live by tests/test_oracle_vs_reference.py when /root/reference exists.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import math

import torch
import torch.nn.functional as F

N_CH, GIN, HID, HEADS, LAYERS = 192, 192, 576, 2, 6


def channel_ln(x, gamma, beta, eps=1e-4):
    """text_encoder.py:21-28: LayerNorm over dim 1 of [B, C, T], biased variance."""
    mean = x.mean(1, keepdim=True)
    var = ((x - mean) ** 2).mean(1, keepdim=True)
    x = (x - mean) * torch.rsqrt(var + eps)
    return x * gamma.view(1, -1, 1) + beta.view(1, -1, 1)


def prenet(sd, x, x_mask, p="prenet."):
    """ConvReluNorm.forward, text_encoder.py:76-83 (dropout is the identity in eval mode)."""
    x_org = x
    for i in range(3):
        x = F.conv1d(x * x_mask, sd[f"{p}conv_layers.{i}.weight"], sd[f"{p}conv_layers.{i}.bias"], padding=2)
        x = channel_ln(x, sd[f"{p}norm_layers.{i}.gamma"], sd[f"{p}norm_layers.{i}.beta"])
        x = torch.relu(x)
    x = x_org + F.conv1d(x, sd[p + "proj.weight"], sd[p + "proj.bias"])
    return x * x_mask


def rope(x, d):
    """RotaryPositionalEmbeddings.forward, text_encoder.py:110-169: x [B, H, T, C]; the first d features are rotated."""
    T = x.shape[2]
    theta = 1.0 / (10000 ** (torch.arange(0, d, 2).float() / d))
    idx_theta = torch.einsum("n,d->nd", torch.arange(T).float(), theta)
    idx_theta2 = torch.cat([idx_theta, idx_theta], dim=1)
    cos, sin = idx_theta2.cos()[None, None], idx_theta2.sin()[None, None]
    x_rope, x_pass = x[..., :d], x[..., d:]
    neg_half = torch.cat([-x_rope[..., d // 2:], x_rope[..., : d // 2]], dim=-1)
    return torch.cat([x_rope * cos + neg_half * sin, x_pass], dim=-1)


def attention(sd, name, x, attn_mask):
    """MultiHeadAttention.forward (self-attention), text_encoder.py:218-252."""
    B, C, T = x.shape
    kc = C // HEADS
    q = F.conv1d(x, sd[name + ".conv_q.weight"], sd[name + ".conv_q.bias"])
    k = F.conv1d(x, sd[name + ".conv_k.weight"], sd[name + ".conv_k.bias"])
    v = F.conv1d(x, sd[name + ".conv_v.weight"], sd[name + ".conv_v.bias"])
    q, k, v = (z.view(B, HEADS, kc, T).transpose(2, 3) for z in (q, k, v))
    d = int(kc * 0.5)
    q, k = rope(q, d), rope(k, d)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(kc)
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    out = torch.matmul(F.softmax(scores, dim=-1), v)
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(out, sd[name + ".conv_o.weight"], sd[name + ".conv_o.bias"])


def ffn(sd, name, x, x_mask):
    """FFN.forward, text_encoder.py:275-281 (kernel 3)."""
    x = F.conv1d(x * x_mask, sd[name + ".conv_1.weight"], sd[name + ".conv_1.bias"], padding=1)
    x = torch.relu(x)
    x = F.conv1d(x * x_mask, sd[name + ".conv_2.weight"], sd[name + ".conv_2.bias"], padding=1)
    return x * x_mask


def encoder(sd, x, x_mask, p="encoder."):
    """Encoder.forward, text_encoder.py:327-337."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    for i in range(LAYERS):
        x = x * x_mask
        y = attention(sd, f"{p}attn_layers.{i}", x, attn_mask)
        x = channel_ln(x + y, sd[f"{p}norm_layers_1.{i}.gamma"], sd[f"{p}norm_layers_1.{i}.beta"])
        y = ffn(sd, f"{p}ffn_layers.{i}", x, x_mask)
        x = channel_ln(x + y, sd[f"{p}norm_layers_2.{i}.gamma"], sd[f"{p}norm_layers_2.{i}.beta"])
    return x * x_mask


def text_encoder_forward(sd, x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed):
    """TextEncoder.forward -> (x [B,576,T], mu [B,80,T], x_mask [B,1,T])."""
    h = (sd["emb.weight"][x] + sd["tone_emb.weight"][tone] + sd["word_pos_emb.weight"][word_pos]
         + sd["syllable_pos.weight"][syllable_pos]) * math.sqrt(N_CH)
    h = h.transpose(1, 2)
    T = h.shape[2]
    x_mask = (torch.arange(T)[None, :] < x_lengths[:, None]).unsqueeze(1).to(h.dtype)
    h = prenet(sd, h, x_mask)
    B = h.shape[0]
    spk = spk_embed.unsqueeze(-1).expand(B, GIN, T)
    lg = sd["lang_emb.weight"][lang].transpose(1, 2)
    h = torch.cat([h, spk, lg], dim=1)
    h = encoder(sd, h, x_mask)
    mu = F.conv1d(h, sd["proj.weight"], sd["proj.bias"]) * x_mask
    return h, mu, x_mask


def duration_predictor_forward(sd, x, x_mask, g):
    """DurationPredictor.forward, duration_predictor.py:48-60 -> logw [B,1,T]."""
    x = x + F.conv1d(g.unsqueeze(2), sd["cond.weight"], sd["cond.bias"])
    x = F.conv1d(x * x_mask, sd["conv_1.weight"], sd["conv_1.bias"], padding=1)
    x = channel_ln(torch.relu(x), sd["norm_1.gamma"], sd["norm_1.beta"])
    x = F.conv1d(x * x_mask, sd["conv_2.weight"], sd["conv_2.bias"], padding=1)
    x = channel_ln(torch.relu(x), sd["norm_2.gamma"], sd["norm_2.beta"])
    x = F.conv1d(x * x_mask, sd["proj.weight"], sd["proj.bias"])
    return x * x_mask


def regulate(logw, x_mask, mu_x, length_scale=1.0):
    """jyutvoice_tts.py:184-203 -> (mu_y [B,80,Ty], y_lengths [B] int64, attn [B,1,Tx,Ty])."""
    w = torch.exp(logw) * x_mask
    w_ceil = torch.ceil(w) * length_scale
    y_lengths = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long()
    Ty = int(y_lengths.max())
    y_mask = (torch.arange(Ty)[None, :] < y_lengths[:, None]).unsqueeze(1).to(x_mask.dtype)
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)
    dur = w_ceil.squeeze(1)
    b, t_x = dur.shape
    cum = torch.cumsum(dur, 1)
    path = (torch.arange(Ty, dtype=cum.dtype)[None, :] < cum.view(b * t_x, 1)).to(x_mask.dtype).view(b, t_x, Ty)
    path = path - F.pad(path, (0, 0, 1, 0, 0, 0))[:, :-1]
    attn = (path * attn_mask.squeeze(1)).unsqueeze(1)
    mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2)).transpose(1, 2)
    return mu_y, y_lengths, attn
