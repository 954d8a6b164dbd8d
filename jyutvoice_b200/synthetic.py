"""Deterministic random-init weights under the reference's own state_dict keys.

TEST INFRASTRUCTURE (also used by bench.py to build synthetic weights; it only creates tensors,
it runs no part of the path).  BASELINE.json asks for random-init weights of the reference
architecture; there is no network for checkpoints.  The reference's constructors reseed the global
RNG mid-construction (flow_matching.py:353), so instead of replaying constructor seeds we enumerate
the key/shape table ourselves and draw every tensor from one torch CPU generator in table order.
`oracle/make_golden.py` loads these dicts into the REAL reference with `strict=True`, which is the
proof that the table equals the reference's state_dict (910 tensors / 71,302,480 parameters for
`decoder.estimator.*`, README.md:171,233; 328 tensors for HiFT).

Key tables follow:
  estimator  jyutvoice/flow/decoder.py:798-915 (CausalConditionalDecoder.__init__),
             jyutvoice/flow/transformer.py:148-260 (BasicTransformerBlock), diffusers Attention/GELU
  hift       jyutvoice/hifigan/generator.py:239-355, f0_predictor.py:19-50
"""
from collections import OrderedDict
import math

import torch

from ._tables import (estimator_table, hift_table, text_encoder_table, duration_predictor_table,  # noqa: F401  (the single key/shape tables)
                      flow_encoder_table)


def _bias_shape(table, idx):
    # bias of a weight-normed conv: out channels. ConvTranspose (ups.*) has Cout = shape[1].
    key = table[idx][0]
    vshape = table[idx + 2][1]
    return (vshape[1],) if key.startswith("ups.") else (vshape[0],)


def _draw(table, seed, w_gain=math.sqrt(3.0), g_range=(0.8, 1.2), alpha_range=(0.5, 1.5)):
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    pending_g = None
    for idx, (key, shape, kind) in enumerate(table):
        if shape is None:
            shape = _bias_shape(table, idx)
        if kind == "w":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            a = 1.0 / math.sqrt(fan_in)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * a * w_gain
        elif kind == "b":
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "g":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "beta":
            sd[key] = 0.05 * torch.randn(shape, generator=g)
        elif kind == "emb":
            sd[key] = torch.randn(shape, generator=g) * shape[1] ** -0.5
        elif kind == "pb":  # xavier_uniform_ over [heads, d_k] (attention.py:222-223)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * math.sqrt(6.0 / (shape[0] + shape[1]))
        elif kind == "alpha":
            sd[key] = alpha_range[0] + (alpha_range[1] - alpha_range[0]) * torch.rand(shape, generator=g)
        elif kind == "wn_g":
            pending_g = (key, g_range[0] + (g_range[1] - g_range[0]) * torch.rand(shape, generator=g))  # ratio to ||v||
            sd[key] = None
        elif kind == "wn_v":
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            a = 1.0 / math.sqrt(fan_in)
            v = (torch.rand(shape, generator=g) * 2 - 1) * a * w_gain
            sd[key] = v
            gkey, ratio = pending_g
            norm = v.reshape(shape[0], -1).norm(dim=1).reshape(shape[0], 1, 1)
            sd[gkey] = ratio * norm
            pending_g = None
        else:
            raise ValueError(kind)
    return sd


def make_estimator_state_dict(seed=1234, prefix="estimator.", init="uniform", scale=1.0):
    """910 fp32 tensors keyed as CausalConditionalCFM.state_dict() (prefix 'estimator.').

    init="uniform" (default, what the goldens were made with): U(+-sqrt(3/fan_in)) weights, small random biases and
    LayerNorm affines.  init="reference": what the reference's own constructor leaves behind
    (decoder.py:414-430 `initialize_weights`: kaiming_normal_(relu) = N(0, 2/fan_in) on every Conv1d / Linear weight,
    zero biases; LayerNorm at PyTorch's default 1 / 0).  `scale` multiplies every Conv1d / Linear weight (stress sets).
    """
    table = estimator_table(prefix)
    if init == "uniform":
        sd = _draw(table, seed)
    elif init == "reference":
        g = torch.Generator(device="cpu").manual_seed(seed)
        sd = OrderedDict()
        for key, shape, kind in table:
            if kind == "w":
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                sd[key] = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
            elif kind == "g":
                sd[key] = torch.ones(shape)
            elif kind in ("b", "beta"):
                sd[key] = torch.zeros(shape)
            else:
                raise ValueError(kind)
    else:
        raise ValueError("init must be 'uniform' or 'reference'")
    if scale != 1.0:
        for (key, shape, kind) in table:
            if kind == "w":
                sd[key] = sd[key] * scale
    return sd


def make_hift_state_dict(seed=4321, f0_bias=None):
    """328 fp32 tensors keyed as HiFTGenerator.state_dict().

    f0_bias: optionally overwrite f0_predictor.classifier.bias (e.g. 200.0 -> every frame voiced;
    SURVEY.md section 7 item 5: random init gives f0 << voiced_threshold, all-unvoiced).
    """
    # PyTorch's default Conv init statistics (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in))), which is what the
    # reference's HiFTGenerator constructor leaves in place (its init_weights is a no-op under parametrised
    # weight-norm: SURVEY.md section 8c); g is perturbed around ||v|| and alpha around 1 so the folds are exercised.
    sd = _draw(hift_table(), seed, w_gain=1.0, g_range=(0.9, 1.1), alpha_range=(0.8, 1.2))
    if f0_bias is not None:
        sd["f0_predictor.classifier.bias"] = torch.full((1,), float(f0_bias))
    return sd


def make_text_encoder_state_dict(seed=2468, n_vocab=97, n_lang=4, n_tone=7):
    """TextEncoder weights under the reference's keys (text_encoder.py:340-420).  The reference zero-initialises
    prenet.proj (the prenet is the identity at construction); here it is drawn like every other conv so the branch counts."""
    return _draw(text_encoder_table(n_vocab, n_lang, n_tone), seed, w_gain=1.0)


def make_duration_predictor_state_dict(seed=1357):
    """DurationPredictor weights (duration_predictor.py:26-46).  The output layer is drawn small around a bias of 0.35,
    so exp(logw) lies in (1, 2) and every token lasts ceil(w) = 2 frames before `length_scale`: 50 tokens at
    length_scale 3.0 give the ~300-frame utterances of BASELINE.json's configs deterministically."""
    sd = _draw(duration_predictor_table(), seed, w_gain=1.0)
    sd["proj.weight"] = sd["proj.weight"] * 0.15
    sd["proj.bias"] = torch.full((1,), 0.35)
    return sd


def make_flow_encoder_state_dict(seed=9753, vocab_size=6561):
    """Speech-token encoder weights under the keys of flow_encoder.pt (infer.py:35-64, upsample_encoder.py:140-288).
    The token embedding is drawn N(0, 1) like nn.Embedding's default."""
    sd = _draw(flow_encoder_table(vocab_size), seed, w_gain=1.0)
    sd["input_embedding.weight"] = sd["input_embedding.weight"] * math.sqrt(512.0)
    return sd


def make_spk_affine_state_dict(seed=77):
    """spk_embed_affine_layer.{weight [80,192], bias} (jyutvoice_tts.py:45)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = 1.0 / math.sqrt(192)
    return OrderedDict(
        weight=(torch.rand((80, 192), generator=g) * 2 - 1) * a * math.sqrt(3.0),
        bias=(torch.rand((80,), generator=g) * 2 - 1) * 0.05,
    )


def noise_bank(n_frames=15000):
    """CausalConditionalCFM.rand_noise (flow_matching.py:353-354): seed 0, randn([1,80,50*300])."""
    torch_state = torch.random.get_rng_state()
    torch.manual_seed(0)
    z = torch.randn([1, 80, 50 * 300])
    torch.random.set_rng_state(torch_state)
    return z[:, :, :n_frames]


def checksum(sd):
    """fp64 sum of sums / sum of abs: a cheap fingerprint that two machines drew the same tensors."""
    s = 0.0
    a = 0.0
    for v in sd.values():
        s += float(v.double().sum())
        a += float(v.double().abs().sum())
    return s, a
