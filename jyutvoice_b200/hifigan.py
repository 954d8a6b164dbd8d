"""Drop-in mirror of the reference's HiFT vocoder, backed by the sm_100a library.

  HiFTGenerator       <- jyutvoice/hifigan/generator.py:239-466  (same constructor kwargs, same state_dict keys,
                         inference(speech_feat, cache_source) -> (wav, s), decode(x, s) -> wav)
  ConvRNNF0Predictor  <- jyutvoice/hifigan/f0_predictor.py:19-55 (configuration carrier; its weights live under
                         `f0_predictor.*` of the generator's state_dict exactly as in the reference)

Supersets: `lengths=` for ragged batches (each utterance is decoded with its own zero boundary, i.e. equals
the reference's unpadded batch-1 call, which naive padding does not: SURVEY.md section 7 item 3) and `rng=` to
inject the source module's random draws (the parity seam).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
from torch.distributions.uniform import Uniform

from . import _lib
from ._tables import hift_keys, build_param_tree
from .flow_matching import DEFAULT_PRECISION, _Workspace


class ConvRNNF0Predictor(nn.Module):
    def __init__(self, num_class=1, in_channels=80, cond_channels=512):
        super().__init__()
        if (num_class, in_channels, cond_channels) != (1, 80, 512):
            raise ValueError("jyutvoice_b200 implements the configs/base.yaml F0 predictor only (1, 80, 512)")
        self.num_class = num_class


class HiFTGenerator(nn.Module):
    def __init__(self, in_channels=80, base_channels=512, nb_harmonics=8, sampling_rate=24000, nsf_alpha=0.1,
                 nsf_sigma=0.003, nsf_voiced_threshold=10, upsample_rates=(8, 5, 3), upsample_kernel_sizes=(16, 11, 7),
                 istft_params=None, resblock_kernel_sizes=(3, 7, 11),
                 resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), source_resblock_kernel_sizes=(7, 7, 11),
                 source_resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), lrelu_slope=0.1, audio_limit=0.99,
                 f0_predictor=None, precision=None):
        super().__init__()
        istft_params = istft_params or {"n_fft": 16, "hop_len": 4}
        cfg = (in_channels, base_channels, nb_harmonics, sampling_rate, float(nsf_alpha), float(nsf_sigma),
               float(nsf_voiced_threshold), tuple(upsample_rates), tuple(upsample_kernel_sizes),
               istft_params["n_fft"], istft_params["hop_len"], tuple(resblock_kernel_sizes),
               tuple(map(tuple, resblock_dilation_sizes)), tuple(source_resblock_kernel_sizes),
               tuple(map(tuple, source_resblock_dilation_sizes)), float(lrelu_slope), float(audio_limit))
        want = (80, 512, 8, 24000, 0.1, 0.003, 10.0, (8, 5, 3), (16, 11, 7), 16, 4, (3, 7, 11),
                ((1, 3, 5),) * 3, (7, 7, 11), ((1, 3, 5),) * 3, 0.1, 0.99)
        if cfg != want:
            raise ValueError("jyutvoice_b200 implements the configs/base.yaml:26-48 HiFT configuration only")
        self.sampling_rate = sampling_rate
        self.istft_params = istft_params
        self.lrelu_slope = lrelu_slope
        self.audio_limit = audio_limit
        self.nb_harmonics = nb_harmonics
        self.precision = precision or DEFAULT_PRECISION
        if self.precision not in _lib.PREC:
            raise ValueError(f"precision must be one of {list(_lib.PREC)}")
        build_param_tree(self, hift_keys())
        self._handle = None
        self._handle_device = None
        self._ws = _Workspace()
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._drop_handle())

    def _drop_handle(self):
        if getattr(self, "_handle", None):
            _lib.lib().jv_hift_destroy(self._handle)
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
        with torch.cuda.device(device):  # the C ABI selects the handle's device and does not restore the caller's
            _lib.check(L.jv_hift_create(device.index or 0, _lib.PREC[self.precision], ctypes.byref(h)))
            try:
                _lib.set_weights(h, L.jv_hift_set_weight, self.state_dict().items())
                _lib.check(L.jv_hift_finalize(h))
            except Exception:
                L.jv_hift_destroy(h)
                raise
        self._handle, self._handle_device = h, device
        return h

    # ---- pieces (each is one C-ABI call)
    def _prep(self, B, T, lengths, dev):
        lens = [T] * B if lengths is None else [int(v) for v in lengths]
        if len(lens) != B or min(lens) < 1 or max(lens) > T:
            raise ValueError("lengths must hold one value in [1, T] per utterance")
        h = self.handle(dev)
        L = _lib.lib()
        lens_c = _lib.i32_array(lens)
        with torch.cuda.device(dev):
            ws = self._ws.get(L.jv_hift_workspace_bytes(h, B, lens_c), dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        return h, L, lens_c, ws, stream

    @staticmethod
    def _p(z):
        return ctypes.c_void_p(z.data_ptr())

    def predict_f0(self, speech_feat, lengths=None):
        B, _, T = speech_feat.shape
        dev = speech_feat.device
        h, L, lens_c, ws, stream = self._prep(B, T, lengths, dev)
        mel = speech_feat.contiguous().float()
        f0 = torch.empty((B, T), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.jv_hift_f0(h, B, T, lens_c, self._p(mel), self._p(f0), self._p(ws), ws.numel(), stream))
        return f0

    def source(self, f0, rng, lengths=None):
        """f0 [B,T], rng {"phase": [B,9,1], "noise": [B,9,480T]} -> s [B,1,480T] (generator.py:459-461)."""
        B, T = f0.shape
        dev = f0.device
        h, L, lens_c, ws, stream = self._prep(B, T, lengths, dev)
        phase = rng["phase"].reshape(B, 9).to(dev).contiguous().float()
        noise = rng["noise"].to(dev).contiguous().float()
        if tuple(noise.shape) != (B, 9, 480 * T):
            raise ValueError("rng['noise'] must be [B, 9, 480*T]")
        s = torch.empty((B, 1, 480 * T), dtype=torch.float32, device=dev)
        f0 = f0.contiguous().float()
        with torch.cuda.device(dev):
            _lib.check(L.jv_hift_source(h, B, T, lens_c, self._p(f0), self._p(phase), self._p(noise), self._p(s),
                                        self._p(ws), ws.numel(), stream))
        return s

    def draw_source_rng(self, B, T, device):
        """The reference's three draws, same order / device / shapes (generator.py:155-158, :171, :235)."""
        phase = Uniform(low=-np.pi, high=np.pi).sample(sample_shape=(B, self.nb_harmonics + 1, 1))
        noise = torch.randn((B, self.nb_harmonics + 1, 480 * T), device=device)
        torch.randn((B, 480 * T, 1), device=device)  # drawn and discarded by the reference; keeps the RNG stream aligned
        return {"phase": phase, "noise": noise}

    @torch.inference_mode()
    def _stft(self, x, lengths=None):
        """Reference `_stft` (generator.py:371-381) on s.squeeze(1): x [B, 480T] -> (real, imag), each [B, 9, 120T+1].
        `lengths` (mel frames): every utterance is reflect-padded at its own end; frames beyond it are 0."""
        B, L_ = x.shape
        if L_ % 480 != 0:
            raise ValueError("the source must hold 480 samples per mel frame")
        T = L_ // 480
        dev = x.device
        h, L, lens_c, ws, stream = self._prep(B, T, lengths, dev)
        s_ = x.contiguous().float()
        out = torch.empty((B, 18, 120 * T + 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.jv_hift_stft(h, B, T, lens_c, self._p(s_), self._p(out), self._p(ws), ws.numel(), stream))
        return out[:, :9], out[:, 9:]

    @torch.inference_mode()
    def decode(self, x, s=None, lengths=None):
        """Reference signature decode(x, s) (generator.py:396-432): x [B,80,T], s [B,1,480T] -> wav [B,480T]."""
        B, _, T = x.shape
        dev = x.device
        if s is None or s.shape[-1] != 480 * T:
            raise ValueError("s must be [B, 1, 480*T]")
        h, L, lens_c, ws, stream = self._prep(B, T, lengths, dev)
        mel = x.contiguous().float()
        s_ = s.reshape(B, 480 * T).to(dev).contiguous().float()
        wav = torch.empty((B, 480 * T), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.jv_hift_decode(h, B, T, lens_c, self._p(mel), self._p(s_), self._p(wav), self._p(ws), ws.numel(), stream))
        return wav

    @torch.inference_mode()
    def inference(self, speech_feat, cache_source=None, lengths=None, rng=None):
        """Reference signature inference(speech_feat, cache_source) (generator.py:450-466) -> (wav, s)."""
        B, _, T = speech_feat.shape
        f0 = self.predict_f0(speech_feat, lengths)
        if rng is None:
            rng = self.draw_source_rng(B, T, speech_feat.device)
        s = self.source(f0, rng, lengths)
        if cache_source is not None and cache_source.shape[2] != 0:
            s[:, :, : cache_source.shape[2]] = cache_source.to(s.device)
        return self.decode(speech_feat, s, lengths), s
