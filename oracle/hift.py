"""ORACLE (test infrastructure, never shipped): fp32 CPU restatement of the HiFT vocoder.

Restates, as plain functions over a state_dict, what these reference sites compute:
  * ConvRNNF0Predictor.forward                 jyutvoice/hifigan/f0_predictor.py:52-55
  * SineGen.forward / SourceModuleHnNSF.forward jyutvoice/hifigan/generator.py:141-176, 220-236
  * HiFTGenerator._stft / _istft / decode / inference   generator.py:371-394, 396-432, 450-466
  * ResBlock.forward                            generator.py:90-97
  * Snake.forward                               jyutvoice/transformer/activation.py:73-84
  * get_padding                                 jyutvoice/utils/common.py:106-107
Weight-norm (torch parametrizations / old-style) is folded as W = g * v / ||v||, the norm taken over
all dims but 0.

STFT and iSTFT are restated as explicit 16-point DFT matrices + overlap-add instead of torch.stft /
torch.istft, so that the arithmetic the CUDA kernels implement is visible; the golden fixtures
(reference's torch.stft/istft path) pin them.

The RNG draws are explicit arguments (`rng`), produced by `draw_source_rng` in the reference's order
and shapes (generator.py:155-158 CPU Uniform.sample [B,9,1]; :171 randn_like [B,9,L]; :235
randn_like [B,L,1], drawn and discarded).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SR = 24000
UPSAMPLE = 480
N_FFT = 16
HOP = 4
N_HARM = 9
SINE_AMP = 0.1
NOISE_STD = 0.003
VOICED_THRESHOLD = 10.0
UPS = ((8, 16), (5, 11), (3, 7))  # (rate, kernel)
RES_K = (3, 7, 11)
RES_D = (1, 3, 5)
SRC_K = (7, 7, 11)
LRELU = 0.1
AUDIO_LIMIT = 0.99


def _wn(g, v):
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape([v.shape[0]] + [1] * (v.dim() - 1))
    return g * v / n


def conv_weight(sd, name):
    """Effective conv weight: folds either weight-norm flavour, or returns the plain weight."""
    if name + ".parametrizations.weight.original0" in sd:
        return _wn(sd[name + ".parametrizations.weight.original0"], sd[name + ".parametrizations.weight.original1"])
    if name + ".weight_g" in sd:
        return _wn(sd[name + ".weight_g"], sd[name + ".weight_v"])
    return sd[name + ".weight"]


def get_padding(k, d=1):
    return int((k * d - d) / 2)


def snake(x, alpha):
    a = alpha.view(1, -1, 1)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def f0_predict(sd, mel):
    """mel [B,80,T] -> f0 [B,T] (Hz)."""
    x = mel
    for i in (0, 2, 4, 6, 8):
        n = f"f0_predictor.condnet.{i}"
        x = F.elu(F.conv1d(x, conv_weight(sd, n), sd[n + ".bias"], padding=1))
    x = x.transpose(1, 2)
    return torch.abs(F.linear(x, sd["f0_predictor.classifier.weight"], sd["f0_predictor.classifier.bias"]).squeeze(-1))


def draw_source_rng(B, L, generator=None):
    """The three draws of one `inference` call, in the reference's order."""
    u = torch.rand((B, N_HARM, 1), generator=generator)
    phase = u * (2 * np.pi) + (-np.pi)  # Uniform(low,high).sample == low + rand*(high-low)
    noise = torch.randn((B, N_HARM, L), generator=generator)
    _unused = torch.randn((B, L, 1), generator=generator)
    return {"phase": phase, "noise": noise}


def source_module(sd, f0, rng):
    """f0 [B,T] -> s [B,1,L].  generator.py:459-461 (nearest upsample x480, SineGen, Linear+tanh)."""
    f0u = f0[:, None].repeat_interleave(UPSAMPLE, dim=2)  # nn.Upsample(scale_factor=480) nearest  [B,1,L]
    B, _, L = f0u.shape
    F_mat = torch.zeros((B, N_HARM, L))
    for i in range(N_HARM):
        F_mat[:, i:i + 1, :] = f0u * (i + 1) / SR
    theta = 2 * np.pi * (torch.cumsum(F_mat, dim=-1) % 1)
    phase = rng["phase"].clone()
    phase[:, 0, :] = 0
    sine = SINE_AMP * torch.sin(theta + phase)
    uv = (f0u > VOICED_THRESHOLD).type(torch.float32)
    noise_amp = uv * NOISE_STD + (1 - uv) * SINE_AMP / 3
    sine = sine * uv + noise_amp * rng["noise"]
    s = torch.tanh(F.linear(sine.transpose(1, 2), sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))
    return s.transpose(1, 2)


def hann16():
    n = torch.arange(N_FFT, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * n / N_FFT)).float()


def stft(s):
    """s [B,L] -> [B,18,L/4+1] = real(9) || imag(9).  Equals torch.stft(center, reflect) of generator.py:371-381."""
    w = hann16()
    xp = F.pad(s[:, None], (N_FFT // 2, N_FFT // 2), mode="reflect")[:, 0]
    frames = xp.unfold(1, N_FFT, HOP) * w  # [B, F, 16]
    n = torch.arange(N_FFT, dtype=torch.float64)
    k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * k[:, None] * n[None, :] / N_FFT
    re = frames @ torch.cos(ang).float().T
    im = frames @ (-torch.sin(ang)).float().T
    return torch.cat([re.transpose(1, 2), im.transpose(1, 2)], dim=1)


def istft(mag, phase):
    """mag, phase [B,9,F] -> [B,4(F-1)].  Equals torch.istft of generator.py:383-394."""
    w = hann16()
    mag = torch.clip(mag, max=1e2)
    re = mag * torch.cos(phase)
    im = mag * torch.sin(phase)
    n = torch.arange(N_FFT, dtype=torch.float64)
    k = torch.arange(N_FFT // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * k[:, None] * n[None, :] / N_FFT  # [9,16]
    c = torch.full((N_FFT // 2 + 1,), 2.0, dtype=torch.float64)
    c[0] = 1.0
    c[-1] = 1.0
    cr = (c[:, None] * torch.cos(ang) / N_FFT).float()
    ci = (-c[:, None] * torch.sin(ang) / N_FFT).float()
    ci[0] = 0
    ci[-1] = 0  # irfft ignores imag of DC and Nyquist
    fr = re.transpose(1, 2) @ cr + im.transpose(1, 2) @ ci  # [B,F,16]
    fr = fr * w
    B, Fr, _ = fr.shape
    Lp = HOP * (Fr - 1) + N_FFT
    y = torch.zeros(B, Lp)
    env = torch.zeros(Lp)
    w2 = w * w
    for j in range(N_FFT):  # overlap-add, 16 strided adds
        y[:, j:j + HOP * Fr:HOP] += fr[:, :, j]
        env[j:j + HOP * Fr:HOP] += w2[j]
    y = y[:, N_FFT // 2: Lp - N_FFT // 2] / env[N_FFT // 2: Lp - N_FFT // 2]
    return y


def resblock(sd, name, x, k):
    for i, d in enumerate(RES_D):
        xt = snake(x, sd[f"{name}.activations1.{i}.alpha"])
        xt = F.conv1d(xt, conv_weight(sd, f"{name}.convs1.{i}"), sd[f"{name}.convs1.{i}.bias"],
                      dilation=d, padding=get_padding(k, d))
        xt = snake(xt, sd[f"{name}.activations2.{i}.alpha"])
        xt = F.conv1d(xt, conv_weight(sd, f"{name}.convs2.{i}"), sd[f"{name}.convs2.{i}.bias"],
                      padding=get_padding(k, 1))
        x = xt + x
    return x


def decode(sd, mel, s):
    """mel [B,80,T], s [B,1,480T] -> wav [B,480T].  generator.py:396-432."""
    s_stft = stft(s.squeeze(1))
    x = F.conv1d(mel, conv_weight(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(UPS):
        x = F.leaky_relu(x, LRELU)
        x = F.conv_transpose1d(x, conv_weight(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i == len(UPS) - 1:
            x = F.pad(x, (1, 0), mode="reflect")
        dw = sd[f"source_downs.{i}.weight"]
        rate = (15, 3, 1)[i]
        si = F.conv1d(s_stft, dw, sd[f"source_downs.{i}.bias"], stride=rate, padding=0 if rate == 1 else rate // 2)
        si = resblock(sd, f"source_resblocks.{i}", si, SRC_K[i])
        x = x + si
        xs = None
        for j, kk in enumerate(RES_K):
            r = resblock(sd, f"resblocks.{3 * i + j}", x, kk)
            xs = r if xs is None else xs + r
        x = xs / len(RES_K)
    x = F.leaky_relu(x)  # slope 0.01 (generator.py:423)
    x = F.conv1d(x, conv_weight(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    mag = torch.exp(x[:, :N_FFT // 2 + 1])
    ph = torch.sin(x[:, N_FFT // 2 + 1:])
    y = istft(mag, ph)
    return torch.clamp(y, -AUDIO_LIMIT, AUDIO_LIMIT)


def inference(sd, mel, rng):
    """generator.py:450-466 with an empty cache_source.  Returns (wav [B,480T], s [B,1,480T])."""
    f0 = f0_predict(sd, mel)
    s = source_module(sd, f0, rng)
    return decode(sd, mel, s), s
