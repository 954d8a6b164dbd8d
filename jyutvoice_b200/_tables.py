"""state_dict key/shape tables of the reference modules this package replaces.

Estimator: CausalConditionalDecoder.__init__ (jyutvoice/flow/decoder.py:798-915) with
configs/base.yaml:88-99 -> 910 tensors.  HiFT: HiFTGenerator.__init__ (jyutvoice/hifigan/generator.py:
239-355) + ConvRNNF0Predictor (f0_predictor.py:19-50) with configs/base.yaml:26-48 -> 328 tensors.
"""
import torch
import torch.nn as nn


def _resnet(name, cin):
    return [
        (f"{name}.mlp.1.weight", (256, 1024)), (f"{name}.mlp.1.bias", (256,)),
        (f"{name}.block1.block.0.weight", (256, cin, 3)), (f"{name}.block1.block.0.bias", (256,)),
        (f"{name}.block1.block.2.weight", (256,)), (f"{name}.block1.block.2.bias", (256,)),
        (f"{name}.block2.block.0.weight", (256, 256, 3)), (f"{name}.block2.block.0.bias", (256,)),
        (f"{name}.block2.block.2.weight", (256,)), (f"{name}.block2.block.2.bias", (256,)),
        (f"{name}.res_conv.weight", (256, cin, 1)), (f"{name}.res_conv.bias", (256,)),
    ]


def _tblock(name):
    return [
        (f"{name}.norm1.weight", (256,)), (f"{name}.norm1.bias", (256,)),
        (f"{name}.attn1.to_q.weight", (512, 256)), (f"{name}.attn1.to_k.weight", (512, 256)),
        (f"{name}.attn1.to_v.weight", (512, 256)),
        (f"{name}.attn1.to_out.0.weight", (256, 512)), (f"{name}.attn1.to_out.0.bias", (256,)),
        (f"{name}.norm3.weight", (256,)), (f"{name}.norm3.bias", (256,)),
        (f"{name}.ff.net.0.proj.weight", (1024, 256)), (f"{name}.ff.net.0.proj.bias", (1024,)),
        (f"{name}.ff.net.2.weight", (256, 1024)), (f"{name}.ff.net.2.bias", (256,)),
    ]


def estimator_keys(num_mid_blocks=12, n_blocks=4):
    t = [("time_mlp.linear_1.weight", (1024, 320)), ("time_mlp.linear_1.bias", (1024,)),
         ("time_mlp.linear_2.weight", (1024, 1024)), ("time_mlp.linear_2.bias", (1024,))]
    groups = [("down_blocks.0", 320)] + [(f"mid_blocks.{i}", 256) for i in range(num_mid_blocks)] + [("up_blocks.0", 512)]
    for name, cin in groups:
        t += _resnet(name + ".0", cin)
        for j in range(n_blocks):
            t += _tblock(f"{name}.1.{j}")
        if not name.startswith("mid"):
            t += [(f"{name}.2.weight", (256, 256, 3)), (f"{name}.2.bias", (256,))]
    t += [("final_block.block.0.weight", (256, 256, 3)), ("final_block.block.0.bias", (256,)),
          ("final_block.block.2.weight", (256,)), ("final_block.block.2.bias", (256,)),
          ("final_proj.weight", (80, 256, 1)), ("final_proj.bias", (80,))]
    return t


def _wn(name, vshape, cout):
    return [(f"{name}.bias", (cout,)),
            (f"{name}.parametrizations.weight.original0", (vshape[0], 1, 1)),
            (f"{name}.parametrizations.weight.original1", vshape)]


def _resblock(name, c, k):
    t = []
    for grp in ("convs1", "convs2"):
        for i in range(3):
            t += _wn(f"{name}.{grp}.{i}", (c, c, k), c)
    for grp in ("activations1", "activations2"):
        for i in range(3):
            t.append((f"{name}.{grp}.{i}.alpha", (c,)))
    return t


def hift_keys():
    t = [("m_source.l_linear.weight", (1, 9)), ("m_source.l_linear.bias", (1,))]
    t += _wn("conv_pre", (512, 80, 7), 512)
    for i, (cin, cout, k) in enumerate([(512, 256, 16), (256, 128, 11), (128, 64, 7)]):
        t += _wn(f"ups.{i}", (cin, cout, k), cout)
    for i, (c, k) in enumerate([(256, 30), (128, 6), (64, 1)]):
        t += [(f"source_downs.{i}.weight", (c, 18, k)), (f"source_downs.{i}.bias", (c,))]
    for i, (c, k) in enumerate([(256, 7), (128, 7), (64, 11)]):
        t += _resblock(f"source_resblocks.{i}", c, k)
    for i, c in enumerate([256, 128, 64]):
        for j, k in enumerate([3, 7, 11]):
            t += _resblock(f"resblocks.{3 * i + j}", c, k)
    t += _wn("conv_post", (18, 64, 7), 18)
    for i, cin in zip((0, 2, 4, 6, 8), (80, 512, 512, 512, 512)):
        t += [(f"f0_predictor.condnet.{i}.bias", (512,)), (f"f0_predictor.condnet.{i}.weight_g", (512, 1, 1)),
              (f"f0_predictor.condnet.{i}.weight_v", (512, cin, 3))]
    t += [("f0_predictor.classifier.weight", (1, 512)), ("f0_predictor.classifier.bias", (1,))]
    return t


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
