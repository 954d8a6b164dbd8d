"""ORACLE (test infrastructure, never shipped): fp32 CPU restatement of the speech-token encoder that produces `prompt_h`.

Restates, as plain functions over a state_dict (keys of flow_encoder.pt), what these reference sites compute:
  * FlowEncoder.forward                       infer.py:66-82 (embedding of clamped tokens * mask, encoder, encoder_proj)
  * UpsampleConformerEncoder.forward          jyutvoice/transformer/upsample_encoder.py:290-355 (streaming False / True)
  * LinearNoSubsampling + EspnetRelPositionalEncoding   transformer/subsampling.py:70-115, embedding.py:201-298
  * PreLookaheadLayer, Upsample1D             upsample_encoder.py:37-137
  * ConformerEncoderLayer (normalize_before, no macaron, no cnn module)   transformer/encoder_layer.py:236-330
  * RelPositionMultiHeadedAttention (+ rel_shift, masked softmax)          transformer/attention.py:70-110, 196-330
  * PositionwiseFeedForward with swish        transformer/positionwise_feed_forward.py:42-55
  * static chunk mask                         utils/mask.py:91-126, 161-200 (add_optional_chunk_mask, static_chunk_size > 0)

One utterance per call (the reference's only mode, infer.py:255-262): nothing is padded, so the convolutions see the
zero padding of an unbatched call.  Pinned by tests/golden/flow_encoder.npz (made by the UNMODIFIED reference with the
synthetic weights of jyutvoice_b200/synthetic.py) and live by tests/test_oracle_vs_reference.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.
"""
import math

import torch
import torch.nn.functional as F

C, HEADS, DK = 512, 8, 64


def rel_pos_table(T):
    """embedding.py:236-253 + :293-296: rows are positions T-1, T-2, ..., -(T-1)  -> [2T-1, 512]."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, C, 2, dtype=torch.float32) * -(math.log(10000.0) / C))
    pe = torch.zeros(2 * T - 1, C)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def embed(sd, p, x):
    """LinearNoSubsampling.forward: Linear -> LayerNorm(1e-5) -> * sqrt(512); pos_emb for the sequence length."""
    x = F.linear(x, sd[p + "out.0.weight"], sd[p + "out.0.bias"])
    x = F.layer_norm(x, (C,), sd[p + "out.1.weight"], sd[p + "out.1.bias"], 1e-5)
    return x * math.sqrt(C), rel_pos_table(x.shape[1]).unsqueeze(0)


def pre_lookahead(sd, p, x):
    """PreLookaheadLayer.forward (no context, no cache): conv k4 looking 3 frames ahead, leaky_relu, causal conv k3, + x."""
    y = x.transpose(1, 2)
    y = F.leaky_relu(F.conv1d(F.pad(y, (0, 3)), sd[p + "conv1.weight"], sd[p + "conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd[p + "conv2.weight"], sd[p + "conv2.bias"])
    return y.transpose(1, 2) + x


def chunk_mask(T, chunk):
    """add_optional_chunk_mask with static_chunk_size = chunk, all left chunks: query i sees keys < (i // chunk + 1) * chunk."""
    if chunk <= 0:
        return torch.ones(T, T, dtype=torch.bool)
    i = torch.arange(T)
    return i[None, :] < ((i // chunk + 1) * chunk)[:, None]


def rel_attention(sd, p, x, pos_emb, mask):
    """RelPositionMultiHeadedAttention.forward: scores[i, j] = ((q_i + u) . k_j + (q_i + v) . P(i - j)) / sqrt(d_k)."""
    B, T, _ = x.shape
    q = F.linear(x, sd[p + "linear_q.weight"], sd[p + "linear_q.bias"]).view(B, T, HEADS, DK)
    k = F.linear(x, sd[p + "linear_k.weight"], sd[p + "linear_k.bias"]).view(B, T, HEADS, DK).transpose(1, 2)
    v = F.linear(x, sd[p + "linear_v.weight"], sd[p + "linear_v.bias"]).view(B, T, HEADS, DK).transpose(1, 2)
    pp = F.linear(pos_emb, sd[p + "linear_pos.weight"]).view(1, -1, HEADS, DK).transpose(1, 2)   # [1, H, 2T-1, dk]
    qu = (q + sd[p + "pos_bias_u"]).transpose(1, 2)
    qv = (q + sd[p + "pos_bias_v"]).transpose(1, 2)
    ac = torch.matmul(qu, k.transpose(-2, -1))
    bd = torch.matmul(qv, pp.transpose(-2, -1))                                                   # [B, H, T, 2T-1]
    # rel_shift (attention.py:226-246): column c of row i holds position T-1-c; keep c = j - i + T - 1
    idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
    bd = torch.gather(bd, 3, idx.expand(B, HEADS, T, T))
    scores = (ac + bd) / math.sqrt(DK)
    dead = ~mask
    attn = torch.softmax(scores.masked_fill(dead, -float("inf")), dim=-1).masked_fill(dead, 0.0)
    out = torch.matmul(attn, v).transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(out, sd[p + "linear_out.weight"], sd[p + "linear_out.bias"])


def layer(sd, p, x, pos_emb, mask):
    """ConformerEncoderLayer.forward with normalize_before, ff_scale 1, LayerNorm eps 1e-12."""
    y = F.layer_norm(x, (C,), sd[p + "norm_mha.weight"], sd[p + "norm_mha.bias"], 1e-12)
    x = x + rel_attention(sd, p + "self_attn.", y, pos_emb, mask)
    y = F.layer_norm(x, (C,), sd[p + "norm_ff.weight"], sd[p + "norm_ff.bias"], 1e-12)
    y = F.linear(F.silu(F.linear(y, sd[p + "feed_forward.w_1.weight"], sd[p + "feed_forward.w_1.bias"])),
                 sd[p + "feed_forward.w_2.weight"], sd[p + "feed_forward.w_2.bias"])
    return x + y


def upsample(sd, p, x):
    """Upsample1D.forward (no cache): nearest x2, 4 zero frames on the left, conv k5."""
    y = F.interpolate(x.transpose(1, 2), scale_factor=2.0, mode="nearest")
    y = F.conv1d(F.pad(y, (4, 0)), sd[p + "conv.weight"], sd[p + "conv.bias"])
    return y.transpose(1, 2)


def conformer_forward(sd, xs, streaming=False, p="encoder.", static_chunk_size=25):
    """UpsampleConformerEncoder.forward for ONE utterance xs [1, T, 512] -> [1, 2T, 512]."""
    T = xs.shape[1]
    x, pos = embed(sd, p + "embed.", xs)
    mask = chunk_mask(T, static_chunk_size if streaming else 0)
    x = pre_lookahead(sd, p + "pre_lookahead_layer.", x)
    for i in range(6):
        x = layer(sd, f"{p}encoders.{i}.", x, pos, mask)
    x = upsample(sd, p + "up_layer.", x)
    x, pos = embed(sd, p + "up_embed.", x)
    mask = chunk_mask(2 * T, 2 * static_chunk_size if streaming else 0)
    for i in range(4):
        x = layer(sd, f"{p}up_encoders.{i}.", x, pos, mask)
    return F.layer_norm(x, (C,), sd[p + "after_norm.weight"], sd[p + "after_norm.bias"], 1e-5)


def flow_encoder_forward(sd, token, streaming=False):
    """infer.py:66-82 for ONE utterance: token [1, T] int64 -> (h [1, 2T, 80], hidden [1, 2T, 512])."""
    x = sd["input_embedding.weight"][torch.clamp(token, min=0)]
    hid = conformer_forward(sd, x, streaming)
    return F.linear(hid, sd["encoder_proj.weight"], sd["encoder_proj.bias"]), hid
