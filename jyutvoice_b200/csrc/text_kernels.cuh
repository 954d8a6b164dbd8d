// Kernels of the text front of `synthesise`: TextEncoder, DurationPredictor, length
// regulation.  Reference: jyutvoice/models/text_encoder.py, duration_predictor.py, jyutvoice_tts.py:184-203.
// Everything is fp32: the durations go through ceil(), so the encoder runs on the fp32-accurate engine (3xTF32 tcgen05,
// gemm_tf32.cuh; FFMA only for shapes it does not take) in both precision modes; it is ~0.2 % of the step's FLOPs.
// Layout: utterance b owns token rows [off_b, off_b + Tx_b) followed by TE_GAP zero rows (the k = 5 / k = 3 "same" convs
// of the prenet / FFN / duration predictor read them as their zero padding); frame_row[m] = b or -1.
#pragma once
#include "common.cuh"

namespace jv {

constexpr int TE_GAP = 2;
constexpr int TE_C = 192, TE_H = 576, TE_FC = 768, TE_HEADS = 2, TE_KC = 288, TE_ROPE = 144, TE_LAYERS = 6, TE_DP = 256;

// X0[m, 0:192] = (emb[x] + tone_emb[tone] + word_pos_emb[wp] + syllable_pos[sp]) * sqrt(192)   (text_encoder.py:418-426)
__global__ void te_embed_kernel(float* __restrict__ X0, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                                const long long* __restrict__ x, const long long* __restrict__ tone, const long long* __restrict__ wp,
                                const long long* __restrict__ sp, int Tx, const float* __restrict__ emb, const float* __restrict__ tone_emb,
                                const float* __restrict__ wp_emb, const float* __restrict__ sp_emb, int n_vocab, int n_tone) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * TE_C) return;
  const int m = (int)(idx / TE_C), c = (int)(idx % TE_C);
  const int b = frame_row[m];
  float v = 0.f;
  if (b >= 0) {
    const long i = (long)b * Tx + (m - row_off[b]);
    const int ix = min(max((int)x[i], 0), n_vocab - 1), it = min(max((int)tone[i], 0), n_tone - 1);
    const int iw = min(max((int)wp[i], 0), 3), is = min(max((int)sp[i], 0), 3);
    v = (emb[(long)ix * TE_C + c] + tone_emb[it * TE_C + c] + wp_emb[iw * TE_C + c] + sp_emb[is * TE_C + c]) * 13.856406460551018f;
  }
  X0[idx] = v;
}

// XH[m, 192:384] = spk_embed[b], XH[m, 384:576] = lang_emb[lang]   (text_encoder.py:439-447); zero on gap rows
__global__ void te_concat_kernel(float* __restrict__ XH, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                                 const float* __restrict__ spk, const long long* __restrict__ lang, int Tx,
                                 const float* __restrict__ lang_emb, int n_lang) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * 2 * TE_C) return;
  const int m = (int)(idx / (2 * TE_C)), c = (int)(idx % (2 * TE_C));
  const int b = frame_row[m];
  float v = 0.f;
  if (b >= 0) {
    if (c < TE_C) v = spk[b * TE_C + c];
    else {
      const int il = min(max((int)lang[(long)b * Tx + (m - row_off[b])], 0), n_lang - 1);
      v = lang_emb[il * TE_C + (c - TE_C)];
    }
  }
  XH[(long)m * TE_H + TE_C + c] = v;
}

// Channel LayerNorm of the reference's own LayerNorm class (eps 1e-4, biased variance; text_encoder.py:12-28):
//   y = LN(x [+ add]) * gamma + beta ; [relu] ; y = valid ? y : 0 -> out.  One warp per row, C <= 768.
__global__ void __launch_bounds__(256) te_ln_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ add, int ld_add,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta, int C, int relu,
                                                    const int* __restrict__ frame_row, float* __restrict__ out, int ldo, int M) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  float v[24];
  const int n = C / 32;  // 6, 8 or 18
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    if (j < n) {
      const int c = j * 32 + lane;
      float t = x[(long)m * ldx + c];
      if (add) t += add[(long)m * ld_add + c];
      v[j] = t;
      s += t;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    if (j < n) {
      const float d = v[j] - mean;
      q += d * d;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + 1e-4f);
  const bool valid = frame_row[m] >= 0;
#pragma unroll
  for (int j = 0; j < 24; ++j) {
    if (j < n) {
      const int c = j * 32 + lane;
      float y = (v[j] - mean) * rstd * gamma[c] + beta[c];
      if (relu) y = fmaxf(y, 0.f);
      out[(long)m * ldo + c] = valid ? y : 0.f;
    }
  }
}

// Partial rotary embedding of q and k in place (text_encoder.py:86-169, 200-202): per head the first 144 of 288 features,
// pairs (i, i + 72): x_i' = x_i cos - x_{i+72} sin, x_{i+72}' = x_{i+72} cos + x_i sin, angle = pos * 10000^(-2 i / 144).
// cs[pos, 0:72] = cos, cs[pos, 72:144] = sin (fp32 tables made on the host the way the reference builds its cache).
__global__ void te_rope_kernel(float* __restrict__ QKV, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                               const float* __restrict__ cs) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * 2 * TE_HEADS * 72) return;
  const int i = (int)(idx % 72);
  const int h = (int)((idx / 72) % TE_HEADS);
  const int qk = (int)((idx / (72 * TE_HEADS)) % 2);
  const int m = (int)(idx / (72 * TE_HEADS * 2));
  const int b = frame_row[m];
  if (b < 0) return;
  const int pos = m - row_off[b];
  const float c = cs[pos * 144 + i], s = cs[pos * 144 + 72 + i];
  float* p = QKV + (long)m * (3 * TE_H) + qk * TE_H + h * TE_KC;
  const float a = p[i], bb = p[i + 72];
  p[i] = a * c - bb * s;
  p[i + 72] = bb * c + a * s;
}

// Self-attention of the text encoder (text_encoder.py:231-252): 2 heads x 288, scores / sqrt(288), keys beyond the
// utterance masked (the reference fills -1e4: exp underflows to exactly 0 in fp32, i.e. exclusion), softmax, P V.
// One warp per (query row, head): two passes over the utterance's keys (scores kept in shared memory).  Tx is at most a
// few hundred tokens and the whole encoder is ~0.2 % of the step, so this stays a plain FFMA kernel.
__global__ void __launch_bounds__(128) te_attention_kernel(const float* __restrict__ QKV, float* __restrict__ ATT,
                                                           const int* __restrict__ frame_row, const int* __restrict__ row_off,
                                                           const int* __restrict__ row_len, int M, int Tmax) {
  extern __shared__ float te_sc[];  // [4 warps][Tmax]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  const int m = wid / TE_HEADS, h = wid % TE_HEADS;
  if (m >= M) return;
  const int b = frame_row[m];
  float* o = ATT + (long)m * TE_H + h * TE_KC;
  if (b < 0) {
    for (int c = lane; c < TE_KC; c += 32) o[c] = 0.f;
    return;
  }
  float* sc = te_sc + warp * Tmax;
  const int off = row_off[b], len = row_len[b];
  const float* q = QKV + (long)m * (3 * TE_H) + h * TE_KC;
  float qr[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) qr[j] = q[j * 32 + lane];
  float mx = -INFINITY;
  for (int t = 0; t < len; ++t) {
    const float* k = QKV + (long)(off + t) * (3 * TE_H) + TE_H + h * TE_KC;
    float d = 0.f;
#pragma unroll
    for (int j = 0; j < 9; ++j) d = fmaf(qr[j], k[j * 32 + lane], d);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) d += __shfl_xor_sync(0xffffffffu, d, s);
    d *= 0.058925565098878960f;  // 1 / sqrt(288)
    if (lane == 0) sc[t] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float acc[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float den = 0.f;
  for (int t = 0; t < len; ++t) {
    const float p = expf(sc[t] - mx);
    den += p;
    const float* v = QKV + (long)(off + t) * (3 * TE_H) + 2 * TE_H + h * TE_KC;
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = fmaf(p, v[j * 32 + lane], acc[j]);
  }
  const float inv = 1.0f / den;
#pragma unroll
  for (int j = 0; j < 9; ++j) o[j * 32 + lane] = acc[j] * inv;
}

// rows [M, C] (ld) -> out [B, C, Tx], zero beyond the utterance
__global__ void te_unpack_kernel(float* __restrict__ out, const float* __restrict__ X, int ld, int C, const int* __restrict__ row_off,
                                 const int* __restrict__ row_len, int B, int Tx) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * C * Tx) return;
  const int t = (int)(idx % Tx), c = (int)((idx / Tx) % C), b = (int)(idx / ((long)Tx * C));
  out[idx] = t < row_len[b] ? X[(long)(row_off[b] + t) * ld + c] : 0.f;
}

// XD[m, c] = x[b, c, t] + cond[b, c] on valid rows, 0 on gap rows   (duration_predictor.py:50 and the x * x_mask of :51)
__global__ void te_pack_cond_kernel(float* __restrict__ XD, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                                    const float* __restrict__ x, const float* __restrict__ cond, int C, int Tx) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx % C);
  const int b = frame_row[m];
  XD[idx] = b >= 0 ? x[((long)b * C + c) * Tx + (m - row_off[b])] + cond[b * C + c] : 0.f;
}

// jyutvoice_tts.py:184-187 and the cumulative sum of utils/model.py:37: one thread per utterance (Tx is small).
//   w = exp(logw) * mask ; w_ceil = ceil(w) * length_scale ; cum = cumsum(w_ceil) ; y_len = max(sum(w_ceil), 1) truncated.
// The running sum is kept in fp64 and every output rounded to fp32: that is what torch's CPU cumsum does, and it is exact
// for the integer-valued durations of length_scale = 1, 2, 3, 0.5, ...
__global__ void te_durations_kernel(const float* __restrict__ logw, const int* __restrict__ x_len, int B, int Tx, float length_scale,
                                    float* __restrict__ cum, long long* __restrict__ y_len) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double acc = 0.0;
  for (int i = 0; i < Tx; ++i) {
    const float w = i < x_len[b] ? expf(logw[(long)b * Tx + i]) : 0.f;
    const float wc = ceilf(w) * length_scale;
    acc += (double)wc;
    cum[(long)b * Tx + i] = (float)acc;
  }
  const float s = (float)acc;
  y_len[b] = (long long)fmaxf(s, 1.0f);
}

// generate_path + `attn^T @ mu_x` as a gather (utils/model.py:29-46, jyutvoice_tts.py:196-203): frame t of utterance b
// copies token i with cum[i-1] <= t < cum[i] (i < x_len, t < y_len); idx = -1 and mu_y = 0 where there is none.
__global__ void te_align_kernel(const float* __restrict__ cum, const int* __restrict__ x_len, const long long* __restrict__ y_len, int B,
                                int Tx, int Ty, const float* __restrict__ mu_x, float* __restrict__ mu_y, int* __restrict__ idx_out) {
  const long id = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long)B * Ty) return;
  const int b = (int)(id / Ty), t = (int)(id % Ty);
  int tok = -1;
  if (t < y_len[b]) {
    const float tf = (float)t;
    const float* c = cum + (long)b * Tx;
    int lo = 0, hi = x_len[b];  // first i in [0, x_len) with tf < c[i]
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (tf < c[mid]) hi = mid;
      else lo = mid + 1;
    }
    if (lo < x_len[b]) tok = lo;
  }
  idx_out[id] = tok;
  for (int ch = 0; ch < 80; ++ch) mu_y[((long)b * 80 + ch) * Ty + t] = tok >= 0 ? mu_x[((long)b * 80 + ch) * Tx + tok] : 0.f;
}

}  // namespace jv
