// Kernels of the speech-token encoder that produces `prompt_h`: the small row kernels
// between the GEMMs of flowenc.cu.  Reference: jyutvoice/transformer/upsample_encoder.py, attention.py, embedding.py.
// Everything is fp32 (the reference's arithmetic type; the output becomes part of the CFM's conditioning `mu`).
// Layout: utterance b owns rows [off_b, off_b + len_b) followed by FE_GAP zero rows, which are the zero padding of the
// look-ahead conv (3 frames to the right), the causal convs (2 / 4 frames to the left) of a batch-1 call; frame_row[m] = b or -1.
#pragma once
#include "common.cuh"

namespace jv {

constexpr int FE_C = 512, FE_HEADS = 8, FE_DK = 64, FE_FC = 2048, FE_GAP = 4, FE_LAYERS_A = 6, FE_LAYERS_B = 4;
constexpr int FE_PE_MAX = 5000;  // EspnetRelPositionalEncoding max_len (embedding.py:215): positions -(4999) .. 4999

// X[m, :] = emb[clamp(token[b, t], 0)] on valid rows, 0 on gap rows   (infer.py:77-78: embedding(clamp(token, min=0)) * mask)
__global__ void fe_embed_kernel(float* __restrict__ X, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                                const long long* __restrict__ token, int T, const float* __restrict__ emb, int vocab) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * (FE_C / 4)) return;
  const int m = (int)(idx / (FE_C / 4)), c4 = (int)(idx % (FE_C / 4));
  const int b = frame_row[m];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b >= 0) {
    long long id = token[(long)b * T + (m - row_off[b])];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);  // negative ids are clamped by the reference; ids >= vocab are a caller bug
    v = reinterpret_cast<const float4*>(emb + id * FE_C)[c4];
  }
  reinterpret_cast<float4*>(X + (long)m * FE_C)[c4] = v;
}

// X[m, :] = xs[b, t, :] on valid rows, 0 on gap rows   (the encoder called on features: upsample_encoder.py:290)
__global__ void fe_pack_kernel(float* __restrict__ X, const int* __restrict__ frame_row, const int* __restrict__ row_off, int M,
                               const float* __restrict__ xs, int T) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M * (FE_C / 4)) return;
  const int m = (int)(idx / (FE_C / 4)), c4 = (int)(idx % (FE_C / 4));
  const int b = frame_row[m];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b >= 0) v = reinterpret_cast<const float4*>(xs + ((long)b * T + (m - row_off[b])) * FE_C)[c4];
  reinterpret_cast<float4*>(X + (long)m * FE_C)[c4] = v;
}

// torch.nn.LayerNorm over the 512 channels (biased variance, two-pass), times `scale`, zero on gap rows.
// One warp per row: 16 values per lane in registers.
__global__ void __launch_bounds__(256) fe_ln_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float eps, float scale,
                                                    const int* __restrict__ frame_row, float* __restrict__ out, int M) {
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (m >= M) return;
  float4* o = reinterpret_cast<float4*>(out + (long)m * FE_C);
  if (frame_row[m] < 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float4* xr = reinterpret_cast<const float4*>(x + (long)m * FE_C);
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j] = xr[j * 32 + lane];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float mean = s * (1.0f / FE_C);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) q += __shfl_xor_sync(0xffffffffu, q, d);
  const float rstd = 1.0f / sqrtf(q * (1.0f / FE_C) + eps);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 g = reinterpret_cast<const float4*>(gamma)[j * 32 + lane];
    const float4 b = reinterpret_cast<const float4*>(beta)[j * 32 + lane];
    float4 r;
    r.x = ((v[j].x - mean) * rstd * g.x + b.x) * scale;
    r.y = ((v[j].y - mean) * rstd * g.y + b.y) * scale;
    r.z = ((v[j].z - mean) * rstd * g.z + b.z) * scale;
    r.w = ((v[j].w - mean) * rstd * g.w + b.w) * scale;
    o[j * 32 + lane] = r;
  }
}

// Relative-position self-attention (attention.py:283-330), 8 heads x 64:
//   score[i, j] = ((q_i + u_h) . k_j + (q_i + v_h) . P_h(i - j)) / 8 ;  keys j < limit(i) ;  softmax ;  out_i = sum_j p_ij v_j
// P (the layer's linear_pos applied to the sinusoid table) is stored by position: row r of PT holds position Tp - 1 - r, so
// the `rel_shift` of the reference is the index r = Tp - 1 - (i - j).  limit(i) = len_b, or with a static chunk mask
// (streaming, utils/mask.py:161-200) min(len_b, (i / chunk + 1) * chunk).  A query always sees at least key 0.
// One warp per (query row, head), two channels per lane; scores are kept in shared memory between the two passes.
// The encoder runs once per prompt over a few hundred frames, so this stays a plain FFMA kernel.
__global__ void __launch_bounds__(128) fe_rel_attention_kernel(const float* __restrict__ QKV, const float* __restrict__ PT, int Tp,
                                                               const float* __restrict__ bias_u, const float* __restrict__ bias_v,
                                                               float* __restrict__ ATT, const int* __restrict__ frame_row,
                                                               const int* __restrict__ row_off, const int* __restrict__ row_len, int M,
                                                               int Tlong, int chunk) {
  extern __shared__ float fe_sc[];  // [4 warps][Tlong]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + warp;
  const int m = wid / FE_HEADS, h = wid % FE_HEADS;
  if (m >= M) return;
  const int b = frame_row[m];
  float* o = ATT + (long)m * FE_C + h * FE_DK;
  if (b < 0) {
    o[lane] = 0.f;
    o[lane + 32] = 0.f;
    return;
  }
  float* sc = fe_sc + warp * Tlong;
  const int off = row_off[b], len = row_len[b];
  const int i = m - off;
  int limit = len;
  if (chunk > 0) limit = min(len, (i / chunk + 1) * chunk);
  const float* q = QKV + (long)m * (3 * FE_C) + h * FE_DK;
  const float q0 = q[lane], q1 = q[lane + 32];
  const float qu0 = q0 + bias_u[h * FE_DK + lane], qu1 = q1 + bias_u[h * FE_DK + lane + 32];
  const float qv0 = q0 + bias_v[h * FE_DK + lane], qv1 = q1 + bias_v[h * FE_DK + lane + 32];
  float mx = -INFINITY;
  for (int j = 0; j < limit; ++j) {
    const float* k = QKV + (long)(off + j) * (3 * FE_C) + FE_C + h * FE_DK;
    const float* p = PT + (long)(Tp - 1 - (i - j)) * FE_C + h * FE_DK;
    float ac = fmaf(qu0, k[lane], qu1 * k[lane + 32]);
    float bd = fmaf(qv0, p[lane], qv1 * p[lane + 32]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      ac += __shfl_xor_sync(0xffffffffu, ac, s);
      bd += __shfl_xor_sync(0xffffffffu, bd, s);
    }
    const float d = (ac + bd) * 0.125f;  // / sqrt(64)
    if (lane == 0) sc[j] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float a0 = 0.f, a1 = 0.f, den = 0.f;
  for (int j = 0; j < limit; ++j) {
    const float pr = expf(sc[j] - mx);
    den += pr;
    const float* v = QKV + (long)(off + j) * (3 * FE_C) + 2 * FE_C + h * FE_DK;
    a0 = fmaf(pr, v[lane], a0);
    a1 = fmaf(pr, v[lane + 32], a1);
  }
  const float inv = 1.0f / den;
  o[lane] = a0 * inv;
  o[lane + 32] = a1 * inv;
}

// Upsample1D's F.interpolate(scale 2, nearest) (upsample_encoder.py:68): XU[row of (b, u)] = X[row of (b, u / 2)], 0 on gap rows
__global__ void fe_repeat_kernel(float* __restrict__ XU, const int* __restrict__ frame_row_up, const int* __restrict__ row_off_up, int M_up,
                                 const float* __restrict__ X, const int* __restrict__ row_off) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M_up * (FE_C / 4)) return;
  const int m = (int)(idx / (FE_C / 4)), c4 = (int)(idx % (FE_C / 4));
  const int b = frame_row_up[m];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (b >= 0) v = reinterpret_cast<const float4*>(X + (long)(row_off[b] + (m - row_off_up[b]) / 2) * FE_C)[c4];
  reinterpret_cast<float4*>(XU + (long)m * FE_C)[c4] = v;
}

// rows [M, C] -> out [B, T, C] (channel-last, as the reference returns it), zero beyond each utterance
__global__ void fe_unpack_kernel(float* __restrict__ out, const float* __restrict__ X, int C, const int* __restrict__ row_off,
                                 const int* __restrict__ row_len, int B, int T) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * T * C) return;
  const int c = (int)(idx % C), t = (int)((idx / C) % T), b = (int)(idx / ((long)C * T));
  out[idx] = t < row_len[b] ? X[(long)(row_off[b] + t) * C + c] : 0.f;
}

}  // namespace jv
