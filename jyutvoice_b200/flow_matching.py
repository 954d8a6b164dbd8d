"""Drop-in mirrors of the reference's CFM classes, backed by the sm_100a library.

  CausalConditionalDecoder  <- jyutvoice/flow/decoder.py:798-1018   (same constructor kwargs, same state_dict keys)
  CausalConditionalCFM      <- jyutvoice/flow/flow_matching.py:343-401 (same constructor kwargs and forward signature)

Differences, all supersets: any batch size (the reference is batch-1: flow_matching.py:238-243), ragged
lengths given by the prefix `mask`, and a `precision` kwarg ("bf16" tcgen05 tensor cores / "fp32" FFMA).
`streaming=True` applies the reference's static chunk mask (chunk = static_chunk_size) inside the attention kernels.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from ._tables import estimator_keys, build_param_tree

DEFAULT_PRECISION = os.environ.get("JYUTVOICE_B200_PRECISION", "bf16")
_STREAM_FORMATS = {"fp16": 0, "bf16": 1, "fp32": 2}


def _lens_from_mask(mask):
    """[B,1,T] prefix mask -> python list of lengths (the only masks the reference builds:
    jyutvoice_tts.py:225,229 via make_pad_mask)."""
    m = mask[:, 0, :] != 0
    lens = m.sum(-1)
    prefix = m.to(torch.int32).cumprod(-1).sum(-1)
    if not torch.equal(lens, prefix):
        raise ValueError("mask must be a prefix (padding) mask")
    out = [int(v) for v in lens.cpu()]
    if min(out) < 1:
        raise ValueError("every utterance needs at least one valid frame")
    return out


class _Workspace:
    """Caller-owned arena for the C ABI, grown on demand (a torch uint8 tensor)."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes, device):
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        return self.buf


class CausalConditionalDecoder(nn.Module):
    def __init__(self, in_channels=320, out_channels=80, channels=(256,), dropout=0.0, attention_head_dim=64,
                 n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn="gelu", static_chunk_size=50,
                 num_decoding_left_chunks=-1, precision=None):
        super().__init__()
        cfg = (in_channels, out_channels, tuple(channels), attention_head_dim, n_blocks, num_mid_blocks, num_heads, act_fn)
        if cfg != (320, 80, (256,), 64, 4, 12, 8, "gelu") or dropout != 0.0:
            raise ValueError("jyutvoice_b200 implements the configs/base.yaml estimator only "
                             "(in 320, out 80, channels [256], 4 blocks, 12 mid, 8 heads x 64, gelu, dropout 0)")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.static_chunk_size = static_chunk_size
        self.num_decoding_left_chunks = num_decoding_left_chunks
        self.precision = precision or DEFAULT_PRECISION
        if self.precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {list(_lib.PREC)}")
        build_param_tree(self, estimator_keys(num_mid_blocks, n_blocks))
        self.stream_format = os.environ.get("JYUTVOICE_B200_STREAM", "fp16")
        self._sat_seen = 0
        self._handle = None
        self._handle_device = None
        self._ws = _Workspace()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_handle())

    # ---- native handle management
    def _drop_handle(self):
        if getattr(self, "_handle", None):
            _lib.lib().jv_estimator_destroy(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self._drop_handle()
        except Exception:
            pass

    def _apply(self, fn, *a, **k):
        self._drop_handle()  # parameters move / change dtype: re-upload lazily
        return super()._apply(fn, *a, **k)

    def handle(self, device):
        if self._handle is not None and self._handle_device == device:
            return self._handle
        self._drop_handle()
        if device.type != "cuda":
            raise RuntimeError("jyutvoice_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        L = _lib.lib()
        h = ctypes.c_void_p()
        with torch.cuda.device(device):  # the C ABI selects the handle's device and does not restore the caller's
            _lib.check(L.jv_estimator_create(device.index or 0, _lib.PREC[self.precision], ctypes.byref(h)))
            try:
                _lib.set_weights(h, L.jv_estimator_set_weight, self.state_dict().items())
                _lib.check(L.jv_estimator_finalize(h))
                _lib.check(L.jv_estimator_set_stream_format(h, _STREAM_FORMATS[self.stream_format]))
            except Exception:
                L.jv_estimator_destroy(h)
                raise
        self._handle, self._handle_device = h, device
        self._sat_seen = 0
        return h

    # ---- 16-bit residual stream of bf16 mode
    def saturation_count(self, synchronize=True):
        """Rows of the fp16 residual stream that may have reached +-65504 since the handle was built (0 = none can
        have).  synchronize=False reads the pinned copy of the last completed call without waiting for the device."""
        if self._handle is None:
            return 0
        n = ctypes.c_int64(0)
        with torch.cuda.device(self._handle_device):
            _lib.check(_lib.lib().jv_estimator_saturation_count(self._handle, 1 if synchronize else 0, ctypes.byref(n)))
        return int(n.value)

    def set_stream_format(self, fmt):
        """"fp16" (default: 11 significand bits, saturating), "bf16" (fp32 range, 8 bits) or "fp32"."""
        if fmt not in _STREAM_FORMATS:
            raise ValueError(f"stream format must be one of {list(_STREAM_FORMATS)}")
        self.stream_format = fmt
        if self._handle is not None:
            _lib.check(_lib.lib().jv_estimator_set_stream_format(self._handle, _STREAM_FORMATS[fmt]))

    def _guard_saturation(self):
        """Called before every launch: if an EARLIER call saturated the fp16 stream (weights with outlier activations),
        say so once and keep the stream in fp32 from now on (a stream that large has out-grown 16-bit resolution too).  Reads a pinned counter: no device synchronisation."""
        if self.precision != "bf16" or self.stream_format != "fp16" or self._handle is None:
            return
        n = self.saturation_count(synchronize=False)
        if n > self._sat_seen:
            import warnings
            warnings.warn(f"jyutvoice_b200: {n - self._sat_seen} residual-stream rows reached the fp16 range in the previous "
                          "call; its output may be clipped.  Switching this estimator's stream to fp32.", RuntimeWarning)
            self._sat_seen = n
            self.set_stream_format("fp32")

    @torch.inference_mode()
    def time_embedding(self, t):
        """Time conditioning of timesteps t [n] -> [n, 14, 256]: row i = resnet i's `mlp(time_mlp(time_embeddings(t)))`
        (decoder.py:15-30, 127-171, 101-103, 936); the solver computes this table once per solve."""
        dev = next(self.parameters()).device
        h = self.handle(dev)
        tl = [float(v) for v in t.reshape(-1).float().cpu()]
        out = torch.empty((len(tl), 14, 256), dtype=torch.float32, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().jv_estimator_time_embedding(h, _lib.f32_array(tl), len(tl), ctypes.c_void_p(out.data_ptr()), stream))
        return out

    @torch.inference_mode()
    def forward(self, x, mask, mu, t, spks=None, cond=None, streaming=False):
        """Reference signature (decoder.py:917).  x, mu, cond [R,80,T]; mask [R,1,T]; t [R]; spks [R,80]."""
        dev = x.device
        h = self.handle(dev)
        self._guard_saturation()
        R, _, T = x.shape
        lens = _lens_from_mask(mask)
        L = _lib.lib()
        lens_c = _lib.i32_array(lens)
        with torch.cuda.device(dev):
            nbytes = L.jv_cfm_workspace_bytes(h, R, lens_c)
            ws = self._ws.get(nbytes, dev)
        f = lambda z: None if z is None else z.contiguous().float()
        x_, mu_, spks_, cond_ = f(x), f(mu), f(spks), f(cond)
        t_host = _lib.f32_array(t.reshape(-1).float().cpu().tolist() if t.numel() == R else [float(t)] * R)
        out = torch.empty((R, 80, T), dtype=torch.float32, device=dev)
        p = lambda z: ctypes.c_void_p(0 if z is None else z.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.jv_estimator_set_chunk(h, int(self.static_chunk_size) if streaming else 0))
            _lib.check(L.jv_estimator_forward(h, R, T, lens_c, p(x_), p(mu_), t_host, p(spks_), p(cond_), p(out),
                                              p(ws), ws.numel(), stream))
        return out


class CausalConditionalCFM(nn.Module):
    def __init__(self, in_channels=240, cfm_params=None, n_spks=1, spk_emb_dim=80, estimator=None):
        super().__init__()
        self.n_feats = in_channels
        self.n_spks = n_spks
        self.spk_emb_dim = spk_emb_dim
        get = (lambda k, d: getattr(cfm_params, k, d)) if cfm_params is not None else (lambda k, d: d)
        self.solver = get("solver", "euler")
        self.sigma_min = get("sigma_min", 1e-6)
        self.t_scheduler = get("t_scheduler", "cosine")
        self.training_cfg_rate = get("training_cfg_rate", 0.2)
        self.inference_cfg_rate = get("inference_cfg_rate", 0.7)
        self.estimator = estimator if estimator is not None else CausalConditionalDecoder()
        # flow_matching.py:353-354: seed-0 noise bank, identical for every utterance.  (The reference
        # reseeds the GLOBAL RNG here; we draw from a private generator with the same seed instead.)
        g = torch.Generator(device="cpu").manual_seed(0)
        self.rand_noise = torch.randn([1, 80, 50 * 300], generator=g)
        self._noise_dev = None
        self._ws = _Workspace()

    def _noise(self, device):
        if self._noise_dev is None or self._noise_dev.device != device:
            self._noise_dev = self.rand_noise[0].to(device).contiguous()
        return self._noise_dev

    @torch.inference_mode()
    def forward(self, mu, mask, n_timesteps, temperature=1.0, spks=None, cond=None, streaming=False, lengths=None):
        """Reference signature (flow_matching.py:357-401) -> (mel fp32 [B,80,T], None).
        `lengths` (host ints) may replace `mask` to skip the device->host read of the mask."""
        if not isinstance(self.estimator, CausalConditionalDecoder):
            raise TypeError("estimator must be a jyutvoice_b200 CausalConditionalDecoder")
        dev = mu.device
        B, _, T = mu.shape
        if T > self.rand_noise.shape[2]:
            raise ValueError(f"{T} frames exceed the {self.rand_noise.shape[2]}-frame noise bank")
        if lengths is not None:
            lens = [int(v) for v in lengths]
            if len(lens) != B or min(lens) < 1 or max(lens) > T:
                raise ValueError("lengths must hold one value in [1, T] per utterance")
        else:
            lens = _lens_from_mask(mask)
        h = self.estimator.handle(dev)
        self.estimator._guard_saturation()
        L = _lib.lib()
        lens_c = _lib.i32_array(lens)
        t_span = torch.linspace(0, 1, n_timesteps + 1, dtype=torch.float32)
        if self.t_scheduler == "cosine":
            t_span = 1 - torch.cos(t_span * 0.5 * torch.pi)
        t_c = _lib.f32_array(t_span.tolist())
        with torch.cuda.device(dev):
            nbytes = L.jv_cfm_solve_workspace_bytes(h, B, lens_c)
            ws = self._ws.get(nbytes, dev)
        mu_ = mu.contiguous().float()
        spks_ = (spks if spks is not None else torch.zeros(B, 80, device=dev)).contiguous().float()
        cond_ = None if cond is None else cond.contiguous().float()
        noise = self._noise(dev)
        out = torch.empty((B, 80, T), dtype=torch.float32, device=dev)
        p = lambda z: ctypes.c_void_p(0 if z is None else z.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.cuda.device(dev):
            _lib.check(L.jv_estimator_set_chunk(h, int(self.estimator.static_chunk_size) if streaming else 0))
            _lib.check(L.jv_cfm_solve(h, B, T, lens_c, p(mu_), p(spks_), p(cond_), p(noise), noise.shape[1],
                                      float(temperature), int(n_timesteps), t_c, float(self.inference_cfg_rate),
                                      p(out), p(ws), ws.numel(), stream))
        return out, None
