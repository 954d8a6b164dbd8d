"""Drop-in mirrors of the speech-token encoder that turns a prompt's speech tokens into `prompt_h` (voice cloning).

  UpsampleConformerEncoder  <- jyutvoice/transformer/upsample_encoder.py:140-355 (same constructor arguments, same keys)
  FlowEncoder               <- infer.py:35-82 (input_embedding + encoder + encoder_proj: the module flow_encoder.pt loads into)

Batched over ragged utterances; every utterance is computed as the reference's batch-1 call computes it (infer.py:255-262 is
the only call site and it is batch 1; the reference's own padded batch would let the look-ahead conv read padding).
fp32 on the GPU (3xTF32 tensor-core GEMMs).  Dropout is the identity (inference).  No CPU path: CUDA tensors only.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._tables import build_param_tree, flow_encoder_keys
from .flow_matching import _Workspace


def _p(z):
    return ctypes.c_void_p(z.data_ptr()) if z is not None else None


class _FlowEncModule(nn.Module):
    """One jv_flowenc handle per module; weights are handed over under the keys of flow_encoder.pt."""
    _prefix = ""

    def _init_native(self):
        self._handle = None
        self._handle_device = None
        self._ws = _Workspace()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_handle())

    def _drop_handle(self):
        if getattr(self, "_handle", None):
            _lib.lib().jv_flowenc_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._drop_handle()
        return super()._apply(fn, *a, **k)

    def handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return self._handle
        self._drop_handle()
        if device.type != "cuda":
            raise RuntimeError("jyutvoice_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        L = _lib.lib()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(L.jv_flowenc_create(device.index or 0, ctypes.byref(h)))
            try:
                _lib.set_weights(h, L.jv_flowenc_set_weight, ((self._prefix + k, v) for k, v in self.state_dict().items()))
                _lib.check(L.jv_flowenc_finalize(h))
            except Exception:
                L.jv_flowenc_destroy(h)
                raise
        self._handle, self._handle_device = h, device
        return h

    def _encode(self, token, xs, lengths, chunk, want_hidden, want_h):
        src = token if token is not None else xs
        dev = src.device
        B, T = src.shape[0], src.shape[1]
        lens = [int(v) for v in lengths.reshape(-1).cpu()]
        if len(lens) != B or min(lens) < 1 or max(lens) > T:
            raise ValueError("lengths must hold one value in [1, T] per utterance")
        h = self.handle(dev)
        L = _lib.lib()
        lens_c = _lib.i32_array(lens)
        tok = token.to(dev).contiguous().long() if token is not None else None
        feat = xs.contiguous().float() if xs is not None else None
        out_hidden = torch.empty((B, 2 * T, 512), dtype=torch.float32, device=dev) if want_hidden else None
        out_h = torch.empty((B, 2 * T, 80), dtype=torch.float32, device=dev) if want_h else None
        with torch.cuda.device(dev):
            ws = self._ws.get(L.jv_flowenc_workspace_bytes(h, B, T, lens_c), dev)
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(L.jv_flowenc_encode(h, B, T, lens_c, _p(tok), _p(feat), int(chunk), _p(out_hidden), _p(out_h), _p(ws), ws.numel(),
                                           stream))
        lens_t = torch.as_tensor(lens, device=dev) * 2
        masks = (torch.arange(2 * T, device=dev)[None, :] < lens_t[:, None]).unsqueeze(1)   # [B, 1, 2T] bool (upsample_encoder.py:335)
        return out_hidden, out_h, masks


class UpsampleConformerEncoder(_FlowEncModule):
    _prefix = "encoder."

    def __init__(self, input_size=512, output_size=512, attention_heads=8, linear_units=2048, num_blocks=6, dropout_rate=0.1,
                 positional_dropout_rate=0.1, attention_dropout_rate=0.1, input_layer="linear", pos_enc_layer_type="rel_pos_espnet",
                 normalize_before=True, static_chunk_size=25, use_dynamic_chunk=False, global_cmvn=None, use_dynamic_left_chunk=False,
                 positionwise_conv_kernel_size=1, macaron_style=False, selfattention_layer_type="rel_selfattn", activation_type="swish",
                 use_cnn_module=False, cnn_module_kernel=15, causal=False, cnn_module_norm="batch_norm", key_bias=True,
                 gradient_checkpointing=False):
        super().__init__()
        cfg = (input_size, output_size, attention_heads, linear_units, num_blocks, input_layer, pos_enc_layer_type, normalize_before,
               macaron_style, selfattention_layer_type, activation_type, use_cnn_module, key_bias, use_dynamic_chunk, global_cmvn)
        if cfg != (512, 512, 8, 2048, 6, "linear", "rel_pos_espnet", True, False, "rel_selfattn", "swish", False, True, False, None):
            raise ValueError("jyutvoice_b200 implements the CosyVoice2 flow-encoder configuration only (infer.py:44-60: 512 channels, "
                             "8 heads, 2048 units, 6 + 4 blocks, linear input layer, rel_pos_espnet, rel_selfattn, swish, no macaron, "
                             "no cnn module)")
        self._output_size = output_size
        self.static_chunk_size = static_chunk_size
        build_param_tree(self, [(k[len("encoder."):], s) for k, s in flow_encoder_keys() if k.startswith("encoder.")])
        self._init_native()

    def output_size(self):
        return self._output_size

    @torch.inference_mode()
    def forward(self, xs, xs_lens, decoding_chunk_size=0, num_decoding_left_chunks=-1, streaming=False):
        """Reference signature (upsample_encoder.py:290) -> (xs [B, 2T, 512], masks [B, 1, 2T] bool)."""
        hidden, _, masks = self._encode(None, xs, xs_lens, self.static_chunk_size if streaming else 0, True, False)
        return hidden, masks


class FlowEncoder(_FlowEncModule):
    def __init__(self, vocab_size=6561, input_size=512, output_size=80):
        super().__init__()
        if (input_size, output_size) != (512, 80):
            raise ValueError("jyutvoice_b200 implements FlowEncoder(input_size=512, output_size=80) only (infer.py:38)")
        self.vocab_size = vocab_size
        build_param_tree(self, flow_encoder_keys(vocab_size))
        self._init_native()

    @torch.inference_mode()
    def forward(self, token, token_len, streaming=False):
        """infer.py:66-82 -> (h [B, 2T, 80], masks [B, 1, 2T] bool).  `streaming` (not in the reference's wrapper, which always
        passes False) selects the encoder's static chunk mask (25 tokens, 50 after the upsampling)."""
        _, h, masks = self._encode(token, None, token_len, 25 if streaming else 0, False, True)
        return h, masks
