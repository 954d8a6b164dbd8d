"""ORACLE (test infrastructure, never shipped): fp32 CPU restatement of the CFM hot path.

Restates, as plain functions over a state_dict, what these reference sites compute:
  * CausalConditionalDecoder.forward            jyutvoice/flow/decoder.py:917-1018
  * SinusoidalPosEmb / TimestepEmbedding        decoder.py:15-30, 127-171
  * CausalResnetBlock1D / CausalBlock1D / CausalConv1d   decoder.py:110-115, 773-788, 737-770
  * BasicTransformerBlock.forward               jyutvoice/flow/transformer.py:355-443
    (+ diffusers==0.35.2 Attention/AttnProcessor2_0, GELU, LoRACompatibleLinear; absent dependency,
     published semantics restated in oracle/ref_shims.py header)
  * add_optional_chunk_mask / mask_to_bias      jyutvoice/utils/mask.py:129-207, utils/common.py:201-209
  * ConditionalCFM.solve_euler, CausalConditionalCFM.forward   jyutvoice/flow/flow_matching.py:215-265, 356-401

Pinned by: tests/golden/*.npz, produced by oracle/make_golden.py from the UNMODIFIED reference
imported from /root/reference (the reference's own tests hold no vector for this path:
SURVEY.md section 8c), and live in tests/test_oracle_vs_reference.py when /root/reference exists.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import math

import torch
import torch.nn.functional as F

CFG_RATE = 0.7  # configs/base.yaml:86 inference_cfg_rate


def time_embedding(sd, t, p="estimator."):
    """decoder.py:21-30 then :159-171.  t: [R] -> [R, 1024]."""
    half = 160
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = 1000 * t.unsqueeze(1) * f.unsqueeze(0)
    e = torch.cat((e.sin(), e.cos()), dim=-1).to(t.dtype)
    h = F.linear(e, sd[p + "time_mlp.linear_1.weight"], sd[p + "time_mlp.linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, sd[p + "time_mlp.linear_2.weight"], sd[p + "time_mlp.linear_2.bias"])


def causal_conv(sd, name, x):
    """decoder.py:767-770: left zero pad k-1, stride 1."""
    w = sd[name + ".weight"]
    return F.conv1d(F.pad(x, (w.shape[-1] - 1, 0)), w, sd[name + ".bias"])


def causal_block(sd, name, x, mask):
    """decoder.py:776-788: Mish(LN_channels(conv(x*mask))) * mask."""
    h = causal_conv(sd, name + ".block.0", x * mask)
    h = F.layer_norm(h.transpose(1, 2), (h.shape[1],), sd[name + ".block.2.weight"],
                     sd[name + ".block.2.bias"], 1e-5).transpose(1, 2)
    return F.mish(h) * mask


def resnet(sd, name, x, mask, temb):
    """decoder.py:110-115."""
    h = causal_block(sd, name + ".block1", x, mask)
    h = h + F.linear(F.mish(temb), sd[name + ".mlp.1.weight"], sd[name + ".mlp.1.bias"]).unsqueeze(-1)
    h = causal_block(sd, name + ".block2", h, mask)
    return h + F.conv1d(x * mask, sd[name + ".res_conv.weight"], sd[name + ".res_conv.bias"])


def attention(sd, name, x, bias, heads=8):
    """diffusers Attention + AttnProcessor2_0 as called at transformer.py:380-389."""
    b, t, _ = x.shape
    q = F.linear(x, sd[name + ".to_q.weight"])
    k = F.linear(x, sd[name + ".to_k.weight"])
    v = F.linear(x, sd[name + ".to_v.weight"])
    d = q.shape[-1] // heads
    q, k, v = (z.view(b, t, heads, d).transpose(1, 2) for z in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias.unsqueeze(1), dropout_p=0.0)
    o = o.transpose(1, 2).reshape(b, t, heads * d)
    return F.linear(o, sd[name + ".to_out.0.weight"], sd[name + ".to_out.0.bias"])


def tblock(sd, name, x, bias):
    """transformer.py:355-443 with norm_type layer_norm, no attn2, gelu FF."""
    n = F.layer_norm(x, (x.shape[-1],), sd[name + ".norm1.weight"], sd[name + ".norm1.bias"], 1e-5)
    x = attention(sd, name + ".attn1", n, bias) + x
    n = F.layer_norm(x, (x.shape[-1],), sd[name + ".norm3.weight"], sd[name + ".norm3.bias"], 1e-5)
    h = F.gelu(F.linear(n, sd[name + ".ff.net.0.proj.weight"], sd[name + ".ff.net.0.proj.bias"]))
    return F.linear(h, sd[name + ".ff.net.2.weight"], sd[name + ".ff.net.2.bias"]) + x


def attn_bias(mask, t_len, chunk=0):
    """decoder.py:950-959: key-padding mask repeated over queries (streaming=False), or AND-ed with the static chunk
    mask (streaming=True: query i sees keys < (i // chunk + 1) * chunk, mask.py:91-126, 180-187) -> (1-m)*-1e10."""
    m = mask.bool()  # [R,1,T]
    if chunk > 0:
        i = torch.arange(t_len)
        ending = torch.clamp((i // chunk + 1) * chunk, max=t_len)
        cm = torch.arange(t_len)[None, :] < ending[:, None]  # [T,T]
        m = m & cm[None]
    else:
        m = m.repeat(1, t_len, 1)
    if (m.sum(dim=-1) == 0).any():  # mask.py:202-206 repairs all-masked rows
        m = m.clone()
        m[m.sum(dim=-1) == 0] = True
    return (1.0 - m.to(mask.dtype)) * -1.0e10


def estimator_forward(sd, x, mask, mu, t, spks=None, cond=None, p="estimator.", chunk=0):
    """CausalConditionalDecoder.forward; chunk = static_chunk_size (50) for streaming=True, 0 for streaming=False.
    x,mu,cond [R,80,T]; mask [R,1,T]; t [R]."""
    temb = time_embedding(sd, t, p)
    h = torch.cat([x, mu], dim=1)
    if spks is not None:
        h = torch.cat([h, spks.unsqueeze(-1).expand(-1, -1, h.shape[-1])], dim=1)
    if cond is not None:
        h = torch.cat([h, cond], dim=1)
    T = h.shape[-1]
    bias = attn_bias(mask, T, chunk)

    def group(h, rname, tname):
        h = resnet(sd, rname, h, mask, temb)
        h = h.transpose(1, 2).contiguous()
        for j in range(4):
            h = tblock(sd, f"{tname}.{j}", h, bias)
        return h.transpose(1, 2).contiguous()

    h = group(h, p + "down_blocks.0.0", p + "down_blocks.0.1")
    skip = h
    h = causal_conv(sd, p + "down_blocks.0.2", h * mask)
    for i in range(12):
        h = group(h, p + f"mid_blocks.{i}.0", p + f"mid_blocks.{i}.1")
    h = torch.cat([h, skip], dim=1)
    h = group(h, p + "up_blocks.0.0", p + "up_blocks.0.1")
    h = causal_conv(sd, p + "up_blocks.0.2", h * mask)
    h = causal_block(sd, p + "final_block", h, mask)
    out = F.conv1d(h * mask, sd[p + "final_proj.weight"], sd[p + "final_proj.bias"])
    return out * mask


def t_span_cosine(n_timesteps, dtype=torch.float32):
    """flow_matching.py:387-389."""
    ts = torch.linspace(0, 1, n_timesteps + 1, dtype=dtype)
    return 1 - torch.cos(ts * 0.5 * torch.pi)


def cfm_forward(sd, noise_bank, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None,
                cfg_rate=CFG_RATE, p="estimator.", chunk=0):
    """CausalConditionalCFM.forward + solve_euler for ONE utterance (B=1, the reference's only mode)."""
    assert mu.shape[0] == 1
    T = mu.shape[2]
    x = noise_bank[:, :, :T].to(mu.dtype) * temperature
    t_span = t_span_cosine(n_timesteps, mu.dtype)
    t, dt = t_span[0], t_span[1] - t_span[0]
    t = t.unsqueeze(0)
    zeros = torch.zeros_like(mu)
    for step in range(1, n_timesteps + 1):
        x_in = torch.cat([x, x], dim=0)
        mask_in = torch.cat([mask, mask], dim=0)
        mu_in = torch.cat([mu, zeros], dim=0)
        t_in = torch.cat([t, t], dim=0)
        spks_in = torch.cat([spks, torch.zeros_like(spks)], dim=0)
        cond_in = torch.cat([cond, zeros], dim=0)
        v = estimator_forward(sd, x_in, mask_in, mu_in, t_in, spks_in, cond_in, p, chunk)
        v = (1.0 + cfg_rate) * v[0:1] - cfg_rate * v[1:2]
        x = x + dt * v
        t = t + dt
        if step < n_timesteps:
            dt = t_span[step + 1] - t
    return x.float()


def cfm_forward_batch(sd, noise_bank, mu, lengths, n_timesteps, temperature=1.0, spks=None, cond=None):
    """The batch oracle: a Python loop of B=1 unpadded calls."""
    B, _, Tmax = mu.shape
    out = torch.zeros_like(mu)
    for b in range(B):
        T = int(lengths[b])
        mask = torch.ones(1, 1, T, dtype=mu.dtype)
        c = cond[b:b + 1, :, :T] if cond is not None else torch.zeros(1, 80, T, dtype=mu.dtype)
        out[b:b + 1, :, :T] = cfm_forward(sd, noise_bank, mu[b:b + 1, :, :T], mask, n_timesteps,
                                          temperature, spks[b:b + 1], c)
    return out
