// Non-GEMM kernels of the CFM estimator path: input packing, LayerNorm(+Mish) rows, masked
// multi-head attention (fp32 FFMA flash-style), time-embedding GEMVs, CFG + Euler update.
#pragma once
#include "common.cuh"

namespace jv {

// ------------------------------------------------------------------------------------------
// Packed activation layout: estimator row r (one utterance, or one half of a CFG pair) owns flat
// frames [row_off[r], row_off[r] + row_len[r]); GAP zero frames follow every row so that a causal
// k=3 tap shifted by -1/-2 reads the conv's zero padding.  frame_row[m] = r, or -1 on gap frames.
// ------------------------------------------------------------------------------------------
constexpr int EST_GAP = 2;

// A0[m, 0:320] = [x | mu | spks | cond] (decoder.py:937-943).  cfg != 0: row r reads utterance r>>1 and
// odd rows get mu = spks = cond = 0 (flow_matching.py:246-251).
template <typename TA>
__global__ void pack_input_kernel(TA* __restrict__ A0, const int* __restrict__ frame_row, const int* __restrict__ row_off,
                                  int M_alloc, const float* __restrict__ x, const float* __restrict__ mu,
                                  const float* __restrict__ spks, const float* __restrict__ cond, int Tmax, int cfg) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)M_alloc * 320) return;
  int m = (int)(idx / 320), c = (int)(idx % 320);
  int r = frame_row[m];
  float v = 0.f;
  if (r >= 0) {
    int t = m - row_off[r];
    int b = cfg ? (r >> 1) : r;
    bool uncond = cfg && (r & 1);
    int sec = c / 80, cc = c % 80;
    if (sec == 0) v = x[((long)b * 80 + cc) * Tmax + t];
    else if (!uncond) {
      if (sec == 1) v = mu[((long)b * 80 + cc) * Tmax + t];
      else if (sec == 2) v = spks ? spks[b * 80 + cc] : 0.f;
      else v = cond ? cond[((long)b * 80 + cc) * Tmax + t] : 0.f;
    }
  }
  A0[idx] = DT<TA>::from_f(v);
}

// x0[b, c, t] = noise[c, t] * temperature for t < len_b, else 0 (flow_matching.py:385)
__global__ void init_noise_kernel(float* __restrict__ x, const float* __restrict__ noise, long noise_stride,
                                  const int* __restrict__ lens, int B, int Tmax, float temperature) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * 80 * Tmax) return;
  int t = (int)(idx % Tmax);
  int c = (int)((idx / Tmax) % 80);
  int b = (int)(idx / ((long)Tmax * 80));
  x[idx] = t < lens[b] ? noise[c * noise_stride + t] * temperature : 0.f;
}

// x[b,c,t] += dt * ((1+cfg) * v[row 2b] - cfg * v[row 2b+1])   (flow_matching.py:255-259)
// dt = dts[step[0]]: the Euler step index lives on the device (row_tidx of the solver), so one captured launch sequence
// serves every step of the solve.
__global__ void cfg_euler_kernel(float* __restrict__ x, const float* __restrict__ v, int ldv, const int* __restrict__ row_off,
                                 const int* __restrict__ lens, int B, int Tmax, const float* __restrict__ dts,
                                 const int* __restrict__ step, float cfg) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * 80 * Tmax) return;
  const float dt = dts[step[0]];
  int t = (int)(idx % Tmax);
  int c = (int)((idx / Tmax) % 80);
  int b = (int)(idx / ((long)Tmax * 80));
  if (t >= lens[b]) return;
  float vc = v[(long)(row_off[2 * b] + t) * ldv + c];
  float vu = v[(long)(row_off[2 * b + 1] + t) * ldv + c];
  float d = (1.0f + cfg) * vc - cfg * vu;
  x[idx] = x[idx] + dt * d;
}

// next Euler step: every estimator row moves on to the time-embedding row of step k + 1
__global__ void step_advance_kernel(int* __restrict__ row_tidx, int R) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < R) row_tidx[i] += 1;
}

// out[r, c, t] = v[row_off[r] + t, c] for t < len_r else 0
__global__ void unpack_output_kernel(float* __restrict__ out, const float* __restrict__ v, int ldv, const int* __restrict__ row_off,
                                     const int* __restrict__ lens, int R, int Tmax) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)R * 80 * Tmax) return;
  int t = (int)(idx % Tmax);
  int c = (int)((idx / Tmax) % 80);
  int r = (int)(idx / ((long)Tmax * 80));
  out[idx] = t < lens[r] ? v[(long)(row_off[r] + t) * ldv + c] : 0.f;
}

// ------------------------------------------------------------------------------------------
// LayerNorm over 256 channels of each frame (one warp per frame), optional Mish, optional per-row
// vector add (time embedding), optional matrix add (res_conv branch), validity mask.
//   y = LN(x) * g + b ; y = act(y) ; y += add_row[row_tidx[r] * add_row_stride + c] ; y = valid ? y : 0 ;
//   y += add_mat[m, c] ; out_f32 / out_act
// ------------------------------------------------------------------------------------------
struct LnArgs {
  const float* x; int ldx;
  const float* gamma; const float* beta;
  int act;   // activation after the affine (ACT_MISH for CausalBlock1D, ACT_NONE for norm1 / norm3)
  const float* add_row; const int* row_tidx; int add_row_stride;
  const float* add_mat; int ld_add;
  const int* frame_row;
  float* out_f32; int ldo;
  void* out_act; int ldo2;
  int M;
};

template <typename TA>
__global__ void __launch_bounds__(256) ln256_kernel(const LnArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= a.M) return;
  const int m = warp;
  const int r = a.frame_row ? a.frame_row[m] : 0;
  const int c0 = lane * 8;
  float v[8];
  {
    const float4 p0 = *reinterpret_cast<const float4*>(a.x + (long)m * a.ldx + c0);
    const float4 p1 = *reinterpret_cast<const float4*>(a.x + (long)m * a.ldx + c0 + 4);
    v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w; v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / 256.0f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { float d = v[j] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.0f / sqrtf(q * (1.0f / 256.0f) + 1e-5f);
  const bool valid = r >= 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = (v[j] - mean) * rstd * __ldg(a.gamma + c0 + j) + __ldg(a.beta + c0 + j);
    y = apply_act(y, a.act, 0.f, 0.f);
    if (a.add_row && valid) y += a.add_row[(long)a.row_tidx[r] * a.add_row_stride + c0 + j];
    y = valid ? y : 0.f;
    if (a.add_mat) y += a.add_mat[(long)m * a.ld_add + c0 + j];
    v[j] = y;
  }
  if (a.out_f32) {
    float4* p = reinterpret_cast<float4*>(a.out_f32 + (long)m * a.ldo + c0);
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (a.out_act) {
    TA* p = (TA*)a.out_act + (long)m * a.ldo2 + c0;
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = DT<TA>::from_f(v[j]);
  }
}

// ------------------------------------------------------------------------------------------
// Masked multi-head attention, 8 heads x 64, non-causal, key padding by row length
// (decoder.py:955-959: bias -1e10 on padded keys == excluding them).  fp32 FFMA, online softmax.
// grid (q_tiles, heads, rows), 256 threads, 64 queries x 64 keys per step.
// ------------------------------------------------------------------------------------------
constexpr int ATT_PAD = 68;
constexpr int ATT_SMEM_BYTES = 4 * 64 * ATT_PAD * (int)sizeof(float);

template <typename TA>
__global__ void __launch_bounds__(256) attention_simt_kernel(const TA* __restrict__ qkv, int ld, TA* __restrict__ out, int ldo,
                                                             const int* __restrict__ row_off, const int* __restrict__ row_len,
                                                             float scale, int chunk) {
  extern __shared__ float att_smem[];
  float (*Qt)[ATT_PAD] = reinterpret_cast<float (*)[ATT_PAD]>(att_smem);                      // [d][q]
  float (*Kt)[ATT_PAD] = reinterpret_cast<float (*)[ATT_PAD]>(att_smem + 64 * ATT_PAD);       // [d][k]
  float (*Vs)[ATT_PAD] = reinterpret_cast<float (*)[ATT_PAD]>(att_smem + 2 * 64 * ATT_PAD);   // [k][d]
  float (*Ps)[ATT_PAD] = reinterpret_cast<float (*)[ATT_PAD]>(att_smem + 3 * 64 * ATT_PAD);   // [q][k]
  const int r = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
  const int len = row_len[r];
  if (q0 >= len) return;
  const long off = row_off[r];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int warp = tid >> 5, lane = tid & 31;

  // load Q tile (transposed into smem); rows beyond len read as 0
  for (int i = warp; i < 64; i += 8) {
    const int t = q0 + i;
#pragma unroll
    for (int dd = 0; dd < 2; ++dd) {
      const int d = lane + 32 * dd;
      Qt[d][i] = t < len ? DT<TA>::to_f(qkv[(off + t) * ld + h * 64 + d]) : 0.f;
    }
  }
  float o[4][4];
  float mrow[4], lrow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    mrow[i] = -INFINITY;
    lrow[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }

  // streaming=True (decoder.py:950-953): query t sees keys < min(len, (t / chunk + 1) * chunk); chunk = 0: all keys
  int klim[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = q0 + ty * 4 + i;
    klim[i] = chunk > 0 ? min(len, (t / chunk + 1) * chunk) : len;
  }
  const int kend = chunk > 0 ? min(len, ((q0 + 63) / chunk + 1) * chunk) : len;
  for (int k0 = 0; k0 < kend; k0 += 64) {
    __syncthreads();  // previous iteration done with Kt/Vs/Ps (and Q stores visible on first pass)
    for (int i = warp; i < 64; i += 8) {
      const int t = k0 + i;
#pragma unroll
      for (int dd = 0; dd < 2; ++dd) {
        const int d = lane + 32 * dd;
        float kv = 0.f, vv = 0.f;
        if (t < len) {
          kv = DT<TA>::to_f(qkv[(off + t) * ld + 512 + h * 64 + d]);
          vv = DT<TA>::to_f(qkv[(off + t) * ld + 1024 + h * 64 + d]);
        }
        Kt[d][i] = kv;
        Vs[i][d] = vv;
      }
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int d = 0; d < 64; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&Qt[d][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Kt[d][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }
    // scale, mask, online softmax (row statistics shared by the 16 threads with equal ty)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kidx = k0 + tx * 4 + j;
        s[i][j] = kidx < klim[i] ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, w));
      const float mnew = fmaxf(mrow[i], mx);  // finite: the first tile holds >= 1 visible key for every row
      const float corr = expf(mrow[i] - mnew);
      float ps = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = expf(s[i][j] - mnew);
        s[i][j] = p;
        ps += p;
      }
#pragma unroll
      for (int w = 8; w > 0; w >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, w);
      lrow[i] = lrow[i] * corr + ps;
      mrow[i] = mnew;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
      *reinterpret_cast<float4*>(&Ps[ty * 4 + i][tx * 4]) = make_float4(s[i][0], s[i][1], s[i][2], s[i][3]);
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 64; ++k) {
      const float4 b = *reinterpret_cast<const float4*>(&Vs[k][tx * 4]);
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = Ps[ty * 4 + i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = fmaf(p, bv[j], o[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = q0 + ty * 4 + i;
    if (t >= len) continue;
    const float inv = 1.0f / lrow[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) out[(off + t) * ldo + h * 64 + tx * 4 + j] = DT<TA>::from_f(o[i][j] * inv);
  }
}

// ------------------------------------------------------------------------------------------
// Small dense rows: out[i, j] = act_out( dot(W[j, :K], act_in(in[i, :K])) + b[j] ), one warp per (i, j).
// Used for the time-embedding MLPs (decoder.py:934-935, :101-103), which depend on t only.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gemv_rows_kernel(const float* __restrict__ W, const float* __restrict__ b,
                                                        const float* __restrict__ in, float* __restrict__ out, int n_in_rows,
                                                        int J, int K, int act_in, int act_out, int ld_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_in_rows * J) return;
  const int i = warp / J, j = warp % J;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) {
    float x = in[(long)i * K + k];
    x = apply_act(x, act_in, 0.f, 0.f);
    s = fmaf(W[(long)j * K + k], x, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    s += b ? b[j] : 0.f;
    out[(long)i * ld_out + j] = apply_act(s, act_out, 0.f, 0.f);
  }
}

}  // namespace jv
