"""Drop-in mirrors of the text front of the reference's `synthesise`, backed by the sm_100a library.

  TextEncoder        <- jyutvoice/models/text_encoder.py:340-451   (same constructor arguments, same state_dict keys)
  DurationPredictor  <- jyutvoice/models/duration_predictor.py:26-60
  length_regulate    <- jyutvoice/models/jyutvoice_tts.py:184-203 + utils/model.py:29-46 (durations -> lengths -> alignment -> mu_y)

Batched over ragged utterances; fp32 on the GPU in both precision modes (the durations pass through ceil(), so this part
keeps the reference's arithmetic type).  Dropout is the identity (inference).  No CPU path: CUDA tensors only.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib
from ._tables import build_param_tree, duration_predictor_keys, text_encoder_keys
from .flow_matching import _Workspace


def _get(params, name, default=None):
    if isinstance(params, dict):
        return params.get(name, default)
    return getattr(params, name, default)


class _TextModule(nn.Module):
    """Shared handle management: one jv_text handle per module, weights passed under `self._prefix`."""
    _prefix = ""

    def _init_native(self):
        self._handle = None
        self._handle_device = None
        self._ws = _Workspace()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_handle())

    def _drop_handle(self):
        if getattr(self, "_handle", None):
            _lib.lib().jv_text_destroy(self._handle)
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
            _lib.check(L.jv_text_create(device.index or 0, ctypes.byref(h)))
            try:
                _lib.set_weights(h, L.jv_text_set_weight, ((self._prefix + k, v) for k, v in self.state_dict().items()))
                _lib.check(L.jv_text_finalize(h))
            except Exception:
                L.jv_text_destroy(h)
                raise
        self._handle, self._handle_device = h, device
        return h

    def _prep(self, B, Tx, x_lengths, dev):
        lens = [int(v) for v in x_lengths.reshape(-1).cpu()]
        if len(lens) != B or min(lens) < 1 or max(lens) > Tx:
            raise ValueError("x_lengths must hold one value in [1, Tx] per utterance")
        h = self.handle(dev)
        L = _lib.lib()
        lens_c = _lib.i32_array(lens)
        with torch.cuda.device(dev):
            ws = self._ws.get(L.jv_text_workspace_bytes(h, B, Tx, lens_c), dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        return h, L, lens, lens_c, ws, stream


def _p(z):
    return ctypes.c_void_p(z.data_ptr())


class TextEncoder(_TextModule):
    _prefix = "encoder."

    def __init__(self, encoder_type, encoder_params, n_vocab, n_lang, n_tone=7):
        super().__init__()
        cfg = tuple(_get(encoder_params, k) for k in ("n_feats", "n_channels", "filter_channels", "n_heads", "n_layers", "kernel_size",
                                                      "gin_channels"))
        if cfg != (80, 192, 768, 2, 6, 3, 192) or not _get(encoder_params, "prenet", True):
            raise ValueError("jyutvoice_b200 implements the configs/base.yaml text encoder only (n_feats 80, n_channels 192, "
                             "filter 768, 2 heads, 6 layers, kernel 3, gin 192, prenet)")
        self.encoder_type = encoder_type
        self.n_vocab, self.n_lang, self.n_tone = n_vocab, n_lang, n_tone
        self.n_feats, self.n_channels, self.gin_channels = 80, 192, 192
        self.hidden_channels = 576
        build_param_tree(self, text_encoder_keys(n_vocab, n_lang, n_tone))
        self._init_native()

    def output_size(self):
        return self.hidden_channels

    @torch.inference_mode()
    def forward(self, x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed):
        """Reference signature (text_encoder.py:401) -> (x [B,576,T], mu [B,80,T], x_mask [B,1,T])."""
        dev = x.device
        B, Tx = x.shape
        h, L, lens, lens_c, ws, stream = self._prep(B, Tx, x_lengths, dev)
        i64 = lambda z: z.to(dev).contiguous().long()
        x_, lang_, tone_, wp_, sp_ = i64(x), i64(lang), i64(tone), i64(word_pos), i64(syllable_pos)
        spk = spk_embed.to(dev).contiguous().float()
        out_x = torch.empty((B, 576, Tx), dtype=torch.float32, device=dev)
        out_mu = torch.empty((B, 80, Tx), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.jv_text_encode(h, B, Tx, lens_c, _p(x_), _p(lang_), _p(tone_), _p(wp_), _p(sp_), _p(spk), _p(out_x), _p(out_mu),
                                        _p(ws), ws.numel(), stream))
        lens_t = torch.as_tensor(lens, device=dev)
        x_mask = (torch.arange(Tx, device=dev)[None, :] < lens_t[:, None]).unsqueeze(1).to(torch.float32)
        return out_x, out_mu, x_mask


class DurationPredictor(_TextModule):
    _prefix = "dp."

    def __init__(self, in_channels, filter_channels, kernel_size, p_dropout, gin_channels):
        super().__init__()
        if (in_channels, filter_channels, kernel_size, gin_channels) != (576, 256, 3, 192):
            raise ValueError("jyutvoice_b200 implements the configs/base.yaml duration predictor only (576 -> 256, kernel 3, gin 192)")
        self.in_channels, self.filter_channels, self.p_dropout = in_channels, filter_channels, p_dropout
        build_param_tree(self, duration_predictor_keys())
        self._init_native()

    @torch.inference_mode()
    def forward(self, x, x_mask, g):
        """Reference signature (duration_predictor.py:48) -> logw [B,1,T]."""
        dev = x.device
        B, _, Tx = x.shape
        x_lengths = (x_mask[:, 0, :] != 0).sum(-1)
        h, L, lens, lens_c, ws, stream = self._prep(B, Tx, x_lengths, dev)
        x_ = x.contiguous().float()
        g_ = g.to(dev).contiguous().float()
        out = torch.empty((B, 1, Tx), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.jv_text_durations(h, B, Tx, lens_c, _p(x_), _p(g_), _p(out), _p(ws), ws.numel(), stream))
        return out


@torch.inference_mode()
def length_regulate(logw, x_mask, mu_x, length_scale=1.0):
    """jyutvoice_tts.py:184-203 on the GPU: (mu_y [B,80,Ty], y_lengths [B] int64, frame_token [B,Ty] int32, cum [B,Tx]).
    frame_token[b, t] is the token frame t copies (-1 beyond y_lengths[b]): `attn` is its one-hot form."""
    dev = logw.device
    if dev.type != "cuda":
        raise RuntimeError("jyutvoice_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    B, _, Tx = logw.shape
    L = _lib.lib()
    x_lens = (x_mask[:, 0, :] != 0).sum(-1).to(torch.int32).contiguous()
    logw_ = logw.contiguous().float()
    cum = torch.empty((B, Tx), dtype=torch.float32, device=dev)
    y_lengths = torch.empty((B,), dtype=torch.int64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    with torch.cuda.device(dev):
        _lib.check(L.jv_length_durations(B, Tx, _p(x_lens), _p(logw_), float(length_scale), _p(cum), _p(y_lengths), stream))
        Ty = int(y_lengths.max())  # the one host read of the path: the CFM workspace and launch shapes depend on it
        mu_y = torch.empty((B, 80, Ty), dtype=torch.float32, device=dev)
        frame_token = torch.empty((B, Ty), dtype=torch.int32, device=dev)
        mu_ = mu_x.contiguous().float()
        _lib.check(L.jv_length_align(B, Tx, Ty, _p(x_lens), _p(y_lengths), _p(cum), _p(mu_), _p(mu_y), _p(frame_token), stream))
    return mu_y, y_lengths, frame_token, cum


def attn_from_frame_token(frame_token, Tx, dtype=torch.float32):
    """The dense 0/1 alignment map the reference returns ([B,1,Tx,Ty], jyutvoice_tts.py:194-196) from the gather indices."""
    B, Ty = frame_token.shape
    attn = torch.zeros((B, Tx + 1, Ty), dtype=dtype, device=frame_token.device)
    idx = torch.where(frame_token >= 0, frame_token, torch.full_like(frame_token, Tx)).long()
    attn.scatter_(1, idx.unsqueeze(1), 1.0)
    return attn[:, :Tx].unsqueeze(1)
