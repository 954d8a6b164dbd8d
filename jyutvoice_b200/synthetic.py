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

N_MID = 12
N_TBLOCKS = 4


def estimator_table(prefix="estimator."):
    """[(key, shape, kind)], kind in {w (fan-in scaled), b (bias), g (LN gamma), beta}."""
    t = []

    def lin(name, n_out, n_in, bias=True):
        t.append((name + ".weight", (n_out, n_in), "w"))
        if bias:
            t.append((name + ".bias", (n_out,), "b"))

    def conv(name, c_out, c_in, k):
        t.append((name + ".weight", (c_out, c_in, k), "w"))
        t.append((name + ".bias", (c_out,), "b"))

    def ln(name, c):
        t.append((name + ".weight", (c,), "g"))
        t.append((name + ".bias", (c,), "beta"))

    def resnet(name, c_in):
        lin(name + ".mlp.1", 256, 1024)
        conv(name + ".block1.block.0", 256, c_in, 3)
        ln(name + ".block1.block.2", 256)
        conv(name + ".block2.block.0", 256, 256, 3)
        ln(name + ".block2.block.2", 256)
        conv(name + ".res_conv", 256, c_in, 1)

    def tblock(name):
        ln(name + ".norm1", 256)
        lin(name + ".attn1.to_q", 512, 256, bias=False)
        lin(name + ".attn1.to_k", 512, 256, bias=False)
        lin(name + ".attn1.to_v", 512, 256, bias=False)
        lin(name + ".attn1.to_out.0", 256, 512)
        ln(name + ".norm3", 256)
        lin(name + ".ff.net.0.proj", 1024, 256)
        lin(name + ".ff.net.2", 256, 1024)

    lin("time_mlp.linear_1", 1024, 320)
    lin("time_mlp.linear_2", 1024, 1024)
    resnet("down_blocks.0.0", 320)
    for j in range(N_TBLOCKS):
        tblock(f"down_blocks.0.1.{j}")
    conv("down_blocks.0.2", 256, 256, 3)
    for i in range(N_MID):
        resnet(f"mid_blocks.{i}.0", 256)
        for j in range(N_TBLOCKS):
            tblock(f"mid_blocks.{i}.1.{j}")
    resnet("up_blocks.0.0", 512)
    for j in range(N_TBLOCKS):
        tblock(f"up_blocks.0.1.{j}")
    conv("up_blocks.0.2", 256, 256, 3)
    conv("final_block.block.0", 256, 256, 3)
    ln("final_block.block.2", 256)
    conv("final_proj", 80, 256, 1)
    return [(prefix + k, s, kind) for k, s, kind in t]


def hift_table():
    """[(key, shape, kind)], kind adds: wn_g / wn_v (weight-norm pair, g derived from v), alpha (Snake)."""
    t = []

    def wn_conv_new(name, shape):  # torch.nn.utils.parametrizations.weight_norm keys
        t.append((name + ".bias", None, "b"))  # shape filled below
        t.append((name + ".parametrizations.weight.original0", (shape[0], 1, 1), "wn_g"))
        t.append((name + ".parametrizations.weight.original1", shape, "wn_v"))

    def resblock(name, c, k):
        for grp in ("convs1", "convs2"):
            for i in range(3):
                wn_conv_new(f"{name}.{grp}.{i}", (c, c, k))
        for grp in ("activations1", "activations2"):
            for i in range(3):
                t.append((f"{name}.{grp}.{i}.alpha", (c,), "alpha"))

    t.append(("m_source.l_linear.weight", (1, 9), "w"))
    t.append(("m_source.l_linear.bias", (1,), "b"))
    wn_conv_new("conv_pre", (512, 80, 7))
    for i, (cin, cout, k) in enumerate([(512, 256, 16), (256, 128, 11), (128, 64, 7)]):
        wn_conv_new(f"ups.{i}", (cin, cout, k))  # ConvTranspose1d weight is [Cin, Cout, K]
    for i, (c, k) in enumerate([(256, 30), (128, 6), (64, 1)]):
        t.append((f"source_downs.{i}.weight", (c, 18, k), "w"))
        t.append((f"source_downs.{i}.bias", (c,), "b"))
    for i, (c, k) in enumerate([(256, 7), (128, 7), (64, 11)]):
        resblock(f"source_resblocks.{i}", c, k)
    for i, c in enumerate([256, 128, 64]):
        for j, k in enumerate([3, 7, 11]):
            resblock(f"resblocks.{3 * i + j}", c, k)
    wn_conv_new("conv_post", (18, 64, 7))
    for i, cin in zip((0, 2, 4, 6, 8), (80, 512, 512, 512, 512)):  # old-style weight_norm keys
        t.append((f"f0_predictor.condnet.{i}.bias", (512,), "b"))
        t.append((f"f0_predictor.condnet.{i}.weight_g", (512, 1, 1), "wn_g"))
        t.append((f"f0_predictor.condnet.{i}.weight_v", (512, cin, 3), "wn_v"))
    t.append(("f0_predictor.classifier.weight", (1, 512), "w"))
    t.append(("f0_predictor.classifier.bias", (1,), "b"))
    return t


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


def make_estimator_state_dict(seed=1234, prefix="estimator."):
    """910 fp32 tensors keyed as CausalConditionalCFM.state_dict() (prefix 'estimator.')."""
    return _draw(estimator_table(prefix), seed)


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
