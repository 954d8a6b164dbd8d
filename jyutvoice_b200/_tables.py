"""state_dict key/shape tables of the reference modules this package replaces.

Estimator: CausalConditionalDecoder.__init__ (jyutvoice/flow/decoder.py:798-915) with
configs/base.yaml:88-99 -> 910 tensors.  HiFT: HiFTGenerator.__init__ (jyutvoice/hifigan/generator.py:
239-355) + ConvRNNF0Predictor (f0_predictor.py:19-50) with configs/base.yaml:26-48 -> 328 tensors.
"""
import torch
import torch.nn as nn



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



def text_encoder_table(n_vocab=97, n_lang=4, n_tone=7, prefix=""):
    """TextEncoder.state_dict() (jyutvoice/models/text_encoder.py:340-420) with the base.yaml encoder_params:
    n_channels 192, gin_channels 192 -> hidden 576, filter 768, 2 heads, 6 layers, kernel 3, prenet (3 x conv k5).
    kind adds: emb (embedding table, N(0, C^-0.5))."""
    t = []
    C, H, Fc = 192, 576, 768
    for name, n in (("emb", n_vocab), ("lang_emb", n_lang), ("tone_emb", n_tone), ("word_pos_emb", 4), ("syllable_pos", 4)):
        t.append((f"{name}.weight", (n, C), "emb"))
    for i in range(3):
        t.append((f"prenet.conv_layers.{i}.weight", (C, C, 5), "w"))
        t.append((f"prenet.conv_layers.{i}.bias", (C,), "b"))
    for i in range(3):
        t.append((f"prenet.norm_layers.{i}.gamma", (C,), "g"))
        t.append((f"prenet.norm_layers.{i}.beta", (C,), "beta"))
    t.append(("prenet.proj.weight", (C, C, 1), "w"))
    t.append(("prenet.proj.bias", (C,), "b"))
    for i in range(6):
        for c in ("q", "k", "v", "o"):
            t.append((f"encoder.attn_layers.{i}.conv_{c}.weight", (H, H, 1), "w"))
            t.append((f"encoder.attn_layers.{i}.conv_{c}.bias", (H,), "b"))
    for i in range(6):
        t.append((f"encoder.norm_layers_1.{i}.gamma", (H,), "g"))
        t.append((f"encoder.norm_layers_1.{i}.beta", (H,), "beta"))
    for i in range(6):
        t.append((f"encoder.ffn_layers.{i}.conv_1.weight", (Fc, H, 3), "w"))
        t.append((f"encoder.ffn_layers.{i}.conv_1.bias", (Fc,), "b"))
        t.append((f"encoder.ffn_layers.{i}.conv_2.weight", (H, Fc, 3), "w"))
        t.append((f"encoder.ffn_layers.{i}.conv_2.bias", (H,), "b"))
    for i in range(6):
        t.append((f"encoder.norm_layers_2.{i}.gamma", (H,), "g"))
        t.append((f"encoder.norm_layers_2.{i}.beta", (H,), "beta"))
    t.append(("proj.weight", (80, H, 1), "w"))
    t.append(("proj.bias", (80,), "b"))
    return [(prefix + k, s_, kind) for k, s_, kind in t]


def duration_predictor_table(prefix=""):
    """DurationPredictor.state_dict() (jyutvoice/models/duration_predictor.py:26-46): in 576, filter 256, kernel 3, gin 192."""
    t = [("conv_1.weight", (256, 576, 3), "w"), ("conv_1.bias", (256,), "b"),
         ("norm_1.gamma", (256,), "g"), ("norm_1.beta", (256,), "beta"),
         ("conv_2.weight", (256, 256, 3), "w"), ("conv_2.bias", (256,), "b"),
         ("norm_2.gamma", (256,), "g"), ("norm_2.beta", (256,), "beta"),
         ("proj.weight", (1, 256, 1), "w"), ("proj.bias", (1,), "b"),
         ("cond.weight", (576, 192, 1), "w"), ("cond.bias", (576,), "b")]
    return [(prefix + k, s_, kind) for k, s_, kind in t]


def flow_encoder_table(vocab_size=6561, prefix=""):
    """state_dict of the speech-token encoder that produces `prompt_h` (flow_encoder.pt): infer.py:35-64 `FlowEncoder` =
    input_embedding + UpsampleConformerEncoder (jyutvoice/transformer/upsample_encoder.py:140-288 with 512 channels, 8 heads,
    2048 FFN units, 6 + 4 rel-pos layers, no macaron / cnn module) + encoder_proj.  kind adds: pb (pos_bias_u / v, xavier)."""
    C, Fc = 512, 2048
    t = [("input_embedding.weight", (vocab_size, C), "emb")]
    e = "encoder."
    for emb in ("embed", "up_embed"):
        t += [(f"{e}{emb}.out.0.weight", (C, C), "w"), (f"{e}{emb}.out.0.bias", (C,), "b"),
              (f"{e}{emb}.out.1.weight", (C,), "g"), (f"{e}{emb}.out.1.bias", (C,), "beta")]
    t += [(e + "after_norm.weight", (C,), "g"), (e + "after_norm.bias", (C,), "beta"),
          (e + "pre_lookahead_layer.conv1.weight", (C, C, 4), "w"), (e + "pre_lookahead_layer.conv1.bias", (C,), "b"),
          (e + "pre_lookahead_layer.conv2.weight", (C, C, 3), "w"), (e + "pre_lookahead_layer.conv2.bias", (C,), "b"),
          (e + "up_layer.conv.weight", (C, C, 5), "w"), (e + "up_layer.conv.bias", (C,), "b")]
    for stack, n in (("encoders", 6), ("up_encoders", 4)):
        for i in range(n):
            a = f"{e}{stack}.{i}."
            t += [(a + "self_attn.pos_bias_u", (8, 64), "pb"), (a + "self_attn.pos_bias_v", (8, 64), "pb")]
            for lin in ("q", "k", "v", "out"):
                t += [(a + f"self_attn.linear_{lin}.weight", (C, C), "w"), (a + f"self_attn.linear_{lin}.bias", (C,), "b")]
            t += [(a + "self_attn.linear_pos.weight", (C, C), "w"),
                  (a + "feed_forward.w_1.weight", (Fc, C), "w"), (a + "feed_forward.w_1.bias", (Fc,), "b"),
                  (a + "feed_forward.w_2.weight", (C, Fc), "w"), (a + "feed_forward.w_2.bias", (C,), "b"),
                  (a + "norm_ff.weight", (C,), "g"), (a + "norm_ff.bias", (C,), "beta"),
                  (a + "norm_mha.weight", (C,), "g"), (a + "norm_mha.bias", (C,), "beta")]
    t += [("encoder_proj.weight", (80, C), "w"), ("encoder_proj.bias", (80,), "b")]
    return [(prefix + k, s_, kind) for k, s_, kind in t]


def flow_encoder_keys(vocab_size=6561):
    return [(k, tuple(s_)) for k, s_, _ in flow_encoder_table(vocab_size)]


def text_encoder_keys(n_vocab=97, n_lang=4, n_tone=7):
    return [(k, tuple(s_)) for k, s_, _ in text_encoder_table(n_vocab, n_lang, n_tone)]


def duration_predictor_keys():
    return [(k, tuple(s_)) for k, s_, _ in duration_predictor_table()]


def _bias_shape(table, idx):
    # bias of a weight-normed conv: out channels.  ConvTranspose (ups.*) has Cout = shape[1].
    key = table[idx][0]
    vshape = table[idx + 2][1]
    return (vshape[1],) if key.startswith("ups.") else (vshape[0],)


def estimator_keys(num_mid_blocks=12, n_blocks=4):
    """[(key, shape)] without the 'estimator.' prefix (CausalConditionalDecoder.state_dict())."""
    assert (num_mid_blocks, n_blocks) == (N_MID, N_TBLOCKS)
    return [(k, tuple(s)) for k, s, _ in estimator_table(prefix="")]


def hift_keys():
    """[(key, shape)] of HiFTGenerator.state_dict()."""
    t = hift_table()
    return [(k, tuple(s) if s is not None else _bias_shape(t, i)) for i, (k, s, _) in enumerate(t)]


def build_param_tree(root, table):
    """Registers one frozen fp32 Parameter per table entry on nested bare Modules so that
    root.state_dict() has exactly the reference's keys (load_state_dict works unchanged)."""
    for key, shape in table:
        parts = key.split(".")
        mod = root
        for p in parts[:-1]:
            if p not in mod._modules:
                mod.add_module(p, nn.Module())
            mod = mod._modules[p]
        mod.register_parameter(parts[-1], nn.Parameter(torch.zeros(shape), requires_grad=False))
