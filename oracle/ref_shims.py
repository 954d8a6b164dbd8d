"""Import shims that let the UNMODIFIED reference (/root/reference) run in this container.

TEST INFRASTRUCTURE ONLY.  Used by `oracle/make_golden.py` (to generate `tests/golden/`) and by
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent, e.g. on the GPU box).
Nothing in the product package imports this file.

No arithmetic of the reference's own files is restated here.  What IS restated is the arithmetic
of the third-party `diffusers==0.35.2` symbols the reference calls (requirements.txt:1; call sites
jyutvoice/flow/transformer.py:5-14,120-121,137,211-219 and jyutvoice/flow/decoder.py:8,146), because
that dependency is not installed and there is no network:

  * Attention(query_dim, heads, dim_head, bias=False): to_q/to_k/to_v = Linear(query_dim, heads*dim_head,
    bias=False); to_out = [Linear(heads*dim_head, query_dim), Dropout]; processor AttnProcessor2_0 =
    F.scaled_dot_product_attention over [B, heads, T, dim_head] with the additive mask broadcast over
    heads, scale 1/sqrt(dim_head); no residual, rescale_output_factor 1.
  * GELU(dim_in, dim_out): proj = Linear(dim_in, dim_out); F.gelu(proj(x), approximate="none").
  * LoRACompatibleLinear = nn.Linear (no LoRA layer attached).
  * get_activation("silu") = nn.SiLU().
"""
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

# Where the reference is imported from: $JYUTVOICE_REFERENCE, else the source tree of this container, else the
# byte-compiled copy oracle/build_ref.py made from it (oracle/_ref: what travels to the GPU box).
_BUILT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("JYUTVOICE_REFERENCE") or ("/root/reference" if os.path.isdir("/root/reference/jyutvoice") else _BUILT)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "jyutvoice"))


class _Attention(nn.Module):
    """diffusers 0.35.2 `Attention` restricted to the call the reference makes (self-attention)."""

    def __init__(self, query_dim, heads=8, dim_head=64, dropout=0.0, bias=False,
                 cross_attention_dim=None, upcast_attention=False, **_):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        b, t, _c = hidden_states.shape
        h = self.heads
        q = self.to_q(hidden_states)
        k = self.to_k(hidden_states)
        v = self.to_v(hidden_states)
        d = q.shape[-1] // h
        q = q.view(b, t, h, d).transpose(1, 2)
        k = k.view(b, t, h, d).transpose(1, 2)
        v = v.view(b, t, h, d).transpose(1, 2)
        if attention_mask is not None:
            # prepare_attention_mask: [B, Tq, Tk] -> repeat_interleave(heads) -> [B, heads, Tq, Tk]
            attention_mask = attention_mask.repeat_interleave(h, dim=0)
            attention_mask = attention_mask.view(b, h, -1, attention_mask.shape[-1])
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, t, h * d).to(q.dtype)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        return o


class _GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("symbol is import-only on the hot path")


def _get_activation(name):
    name = name.lower()
    table = {"silu": nn.SiLU, "swish": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}
    return table[name]()


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


_installed = False


class _BuiltRefFinder:
    """Imports `jyutvoice.*` from the byte-compiled tree oracle/build_ref.py wrote (<root>/jyutvoice/.../module.refbin,
    CPython .pyc format under another suffix)."""

    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if fullname != "jyutvoice" and not fullname.startswith("jyutvoice."):
            return None
        rel = os.path.join(self.root, *fullname.split("."))
        if os.path.isfile(os.path.join(rel, "__init__.refbin")):
            f = os.path.join(rel, "__init__.refbin")
            loader = importlib.machinery.SourcelessFileLoader(fullname, f)
            return importlib.util.spec_from_file_location(fullname, f, loader=loader, submodule_search_locations=[rel])
        if os.path.isfile(rel + ".refbin"):
            loader = importlib.machinery.SourcelessFileLoader(fullname, rel + ".refbin")
            return importlib.util.spec_from_file_location(fullname, rel + ".refbin", loader=loader)
        return None


def install():
    """Register stand-in modules so that `import jyutvoice.flow...` works from REF_ROOT."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    if os.path.isfile(os.path.join(REF_ROOT, "jyutvoice", "__init__.refbin")):
        sys.meta_path.insert(0, _BuiltRefFinder(REF_ROOT))  # the byte-compiled copy (oracle/_ref)
    elif REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # jyutvoice.utils: skip its __init__ (pulls hydra/lightning/...): bare package object.
    import importlib
    jv = importlib.import_module("jyutvoice")
    pkg = types.ModuleType("jyutvoice.utils")
    pkg.__path__ = [os.path.join(REF_ROOT, "jyutvoice", "utils")]
    import logging
    pkg.get_pylogger = lambda name=__name__: logging.getLogger(name)
    sys.modules["jyutvoice.utils"] = pkg
    jv.utils = pkg

    # third-party stand-ins
    _mod("conformer", ConformerBlock=type("ConformerBlock", (nn.Module,), {}))
    d = _mod("diffusers")
    dm = _mod("diffusers.models")
    _mod("diffusers.models.activations", get_activation=_get_activation)
    _mod("diffusers.models.attention", GEGLU=_Unused, GELU=_GELU, AdaLayerNorm=_Unused,
         AdaLayerNormZero=_Unused, ApproximateGELU=_Unused)
    _mod("diffusers.models.attention_processor", Attention=_Attention)
    _mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    _mod("diffusers.utils")
    _mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda cls: cls)
    d.models = dm

    # framework stubs for jyutvoice_tts
    class _LM(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass
    lt = _mod("lightning", LightningModule=_LM)
    _mod("lightning.pytorch")
    _mod("lightning.pytorch.utilities", grad_norm=lambda *a, **k: {})
    lt.pytorch = sys.modules["lightning.pytorch"]
    _mod("wandb")
    _mod("jyutvoice.utils.utils", plot_tensor=lambda *a, **k: None,
         intersperse=lambda lst, item: sum(([item, x] for x in lst), [])[0:] + [item])
    _mod("jyutvoice.utils.monotonic_align", maximum_path=None)
    _installed = True


def build_reference_cfm():
    """Reference CausalConditionalCFM with the configs/base.yaml:76-99 hyper-parameters (random init)."""
    install()
    from jyutvoice.flow.flow_matching import CausalConditionalCFM
    from jyutvoice.flow.decoder import CausalConditionalDecoder
    est = CausalConditionalDecoder(in_channels=320, out_channels=80, channels=[256], dropout=0.0,
                                   attention_head_dim=64, n_blocks=4, num_mid_blocks=12, num_heads=8,
                                   act_fn="gelu", static_chunk_size=50, num_decoding_left_chunks=-1)
    cfm_params = types.SimpleNamespace(sigma_min=1e-6, solver="euler", t_scheduler="cosine",
                                       training_cfg_rate=0.2, inference_cfg_rate=0.7, reg_loss_type="l1")
    cfm = CausalConditionalCFM(in_channels=240, n_spks=1, spk_emb_dim=80, cfm_params=cfm_params, estimator=est)
    return cfm.eval()


def build_reference_hift():
    """Reference HiFTGenerator with the configs/base.yaml:26-48 hyper-parameters (random init)."""
    install()
    from jyutvoice.hifigan.generator import HiFTGenerator
    from jyutvoice.hifigan.f0_predictor import ConvRNNF0Predictor
    f0 = ConvRNNF0Predictor(num_class=1, in_channels=80, cond_channels=512)
    hift = HiFTGenerator(in_channels=80, base_channels=512, nb_harmonics=8, sampling_rate=24000,
                         nsf_alpha=0.1, nsf_sigma=0.003, nsf_voiced_threshold=10,
                         upsample_rates=[8, 5, 3], upsample_kernel_sizes=[16, 11, 7],
                         istft_params={"n_fft": 16, "hop_len": 4},
                         resblock_kernel_sizes=[3, 7, 11],
                         resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
                         source_resblock_kernel_sizes=[7, 7, 11],
                         source_resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
                         lrelu_slope=0.1, audio_limit=0.99, f0_predictor=f0)
    return hift.eval()


def build_reference_tts(cfm=None):
    """Reference JyutVoiceTTS (jyutvoice/models/jyutvoice_tts.py) with its own TextEncoder / DurationPredictor and the
    configs/base.yaml encoder_params; random init (load the synthetic state dicts on top)."""
    install()
    from jyutvoice.models.jyutvoice_tts import JyutVoiceTTS
    from jyutvoice.models.text_encoder import TextEncoder
    from jyutvoice.models.duration_predictor import DurationPredictor
    enc_params = types.SimpleNamespace(n_feats=80, n_channels=192, filter_channels=768, filter_channels_dp=256, n_heads=2,
                                       n_layers=6, kernel_size=3, p_dropout=0.1, gin_channels=192, prenet=True)
    encoder = TextEncoder(encoder_type="RoPE Encoder", encoder_params=enc_params, n_vocab=97, n_lang=4, n_tone=7)
    dp = DurationPredictor(in_channels=576, filter_channels=256, kernel_size=3, p_dropout=0.1, gin_channels=192)
    tts = JyutVoiceTTS(encoder=encoder, decoder=cfm if cfm is not None else build_reference_cfm(), dp=dp, output_size=80,
                       spk_embed_dim=192)
    return tts.eval()


def build_reference_flow_encoder():
    """The reference's speech-token encoder: UpsampleConformerEncoder with the hyper-parameters of infer.py:44-60, plus the
    two layers infer.py's FlowEncoder wraps around it (input_embedding 6561 x 512, encoder_proj 512 -> 80), as a plain
    container with the same state_dict keys as flow_encoder.pt.  Random init."""
    install()
    from jyutvoice.transformer.upsample_encoder import UpsampleConformerEncoder

    class _FlowEncoder(nn.Module):
        def __init__(self):
            super().__init__()
            self.input_embedding = nn.Embedding(6561, 512)
            self.encoder = UpsampleConformerEncoder(
                output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
                positional_dropout_rate=0.1, attention_dropout_rate=0.1, normalize_before=True, input_layer="linear",
                pos_enc_layer_type="rel_pos_espnet", selfattention_layer_type="rel_selfattn", input_size=512,
                use_cnn_module=False, macaron_style=False, static_chunk_size=25)
            self.encoder_proj = nn.Linear(512, 80)

    return _FlowEncoder().eval()
