// Non-GEMM kernels of the HiFT vocoder: mel packing, f0 head, NSF source, STFT, Snake / leaky-ReLU
// rows, reflection fix-up, conv_post head + inverse STFT.
#pragma once
#include "common.cuh"

namespace jv {

// Packed layout: utterance b owns mel frames [off_b, off_b + T_b) followed by HIFT_GAP zero frames;
// a stage with rate r (1, 8, 40, 120) keeps frame (r*off_b + p).  The zero gap is >= the largest
// conv reach at every stage (k=11, dilation 5 -> 25 rows <= 8*4), so every utterance sees the zero
// padding of the reference's unpadded batch-1 call.
constexpr int HIFT_GAP = 4;

struct HiftSeq {
  const int* off;  // [B+1] mel-frame offsets
  const int* len;  // [B] mel frames
  int B;
};

// frame_row for a stage: rows [rate*off_b, rate*off_b + rate*len_b + extra) -> b, else -1
__global__ void hift_frame_row_kernel(int* __restrict__ fr, int rows, HiftSeq sq, int rate, int extra) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= rows) return;
  int v = -1;
  for (int b = 0; b < sq.B; ++b) {
    int s = rate * sq.off[b];
    if (m >= s && m < s + rate * sq.len[b] + extra) { v = b; break; }
  }
  fr[m] = v;
}

// MEL[m, 0:128] = mel[b, c, t] (c < 80), zero elsewhere
template <typename TA>
__global__ void hift_pack_mel_kernel(TA* __restrict__ MEL, const int* __restrict__ fr, HiftSeq sq, int rows,
                                     const float* __restrict__ mel, int Tmax) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)rows * 128) return;
  int m = (int)(idx >> 7), c = (int)(idx & 127);
  int b = fr[m];
  float v = 0.f;
  if (b >= 0 && c < 80) v = mel[((long)b * 80 + c) * Tmax + (m - sq.off[b])];
  MEL[idx] = DT<TA>::from_f(v);
}

// f0[b, t] = | w . h[m, :512] + bias |   (f0_predictor.py:52-55), one warp per frame
template <typename TA>
__global__ void __launch_bounds__(256) hift_f0_head_kernel(const TA* __restrict__ Hh, const int* __restrict__ fr, HiftSeq sq, int rows,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ f0, int Tmax) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int b = fr[warp];
  if (b < 0) return;
  float s = 0.f;
  for (int k = lane; k < 512; k += 32) s = fmaf(DT<TA>::to_f(Hh[(long)warp * 512 + k]), __ldg(w + k), s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) f0[(long)b * Tmax + (warp - sq.off[b])] = fabsf(s + bias[0]);
}

__global__ void zero_f32_kernel(float* __restrict__ p, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.f;
}

// ---------------------------------------------------------------- NSF source (generator.py:141-176, 220-236)
// Phase prefix per (b, harmonic): D[b,h,t] = sum_{t'<t} 480 * F[b,h,t'] accumulated in fp64 — torch's CPU
// cumsum accumulates fp32 inputs in double and rounds each output to float (ATen cumsum_cpu_kernel, acc_type).
// D has row stride Dld = the longest utterance (<= Tmax, the row stride of the caller's tensors): the workspace query
// sees only the lengths.
__global__ void hift_phase_prefix_kernel(const float* __restrict__ f0, int Tmax, const int* __restrict__ len, int B,
                                         double* __restrict__ D, int Dld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 9) return;
  int b = i / 9, h = i % 9;
  double acc = 0.0;
  const int T = len[b];
  for (int t = 0; t < T; ++t) {
    D[((long)b * 9 + h) * Dld + t] = acc;
    float F = f0[(long)b * Tmax + t] * (float)(h + 1) / 24000.0f;
    acc += 480.0 * (double)F;
  }
}

// Four consecutive samples per thread (480 = 4 * 120: they share a mel frame): the nine noise planes are read and s is
// written as 16-byte vectors, which is what makes the kernel bandwidth-bound (36 B in + 4 B out per sample).  Unvoiced
// frames (uv = 0) skip the nine sines: `sine * 0 + noise` is the noise.  Voiced frames evaluate sin with the MUFU after an
// exact range reduction to [-pi, pi] (|error| < 1e-6 on a 0.1-amplitude term; the phase itself is the fp64 prefix rounded
// to fp32 exactly as torch's cumsum does).
__global__ void __launch_bounds__(256) hift_source_kernel(const float* __restrict__ f0, const double* __restrict__ D,
                                                          const float* __restrict__ phase, const float* __restrict__ noise,
                                                          const float* __restrict__ lw, const float* __restrict__ lb,
                                                          const int* __restrict__ len, int B, int Tmax, float* __restrict__ s, int Dld) {
  const long L = 480L * Tmax;
  const long idx4 = (long)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 samples
  if (idx4 * 4 >= (long)B * L) return;
  const long idx = idx4 * 4;
  const int b = (int)(idx / L);
  const long j = idx - (long)b * L;
  const int t = (int)(j / 480), jj = (int)(j - 480L * t);
  float4* out = reinterpret_cast<float4*>(s + idx);
  if (t >= len[b]) {
    *out = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float f = f0[(long)b * Tmax + t];
  const bool voiced = f > 10.0f;
  const float noise_amp = voiced ? 0.003f : (0.1f / 3.0f);
  const float bias = lb[0];
  float acc[4] = {bias, bias, bias, bias};
#pragma unroll
  for (int h = 0; h < 9; ++h) {
    const float4 nz = __ldcs(reinterpret_cast<const float4*>(noise + ((long)b * 9 + h) * L + j));  // streamed: read once
    const float w = lw[h];
    float v[4] = {noise_amp * nz.x, noise_amp * nz.y, noise_amp * nz.z, noise_amp * nz.w};
    if (voiced) {  // warp-uniform except at frame boundaries (120 threads per frame)
      const float F = f * (float)(h + 1) / 24000.0f;
      const double d0 = D[((long)b * 9 + h) * Dld + t];
      const double Fd = (double)F;
      const float ph = h == 0 ? 0.f : phase[b * 9 + h];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float cf = (float)(d0 + (double)(jj + i + 1) * Fd);
        const float frac = cf - floorf(cf);  // fp32 `% 1` of a non-negative value
        float x = 6.283185307179586f * frac + ph;  // in (-pi, 3 pi)
        if (x > 3.14159265358979f) x -= 6.283185307179586f;
        v[i] += 0.1f * sin_ftz(x);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fmaf(w, v[i], acc[i]);
  }
  *out = make_float4(tanhf(acc[0]), tanhf(acc[1]), tanhf(acc[2]), tanhf(acc[3]));
}

// cos(2 pi m / 16): with fully unrolled loops every use below is a compile-time constant, i.e. an FFMA immediate.
__device__ constexpr float COS16[16] = {1.0f, 0.9238795325112867f, 0.7071067811865476f, 0.3826834323650898f, 0.0f,
                                        -0.3826834323650898f, -0.7071067811865476f, -0.9238795325112867f, -1.0f,
                                        -0.9238795325112867f, -0.7071067811865476f, -0.3826834323650898f, 0.0f,
                                        0.3826834323650898f, 0.7071067811865476f, 0.9238795325112867f};
#define JV_COS16(m) (COS16[(m) & 15])
#define JV_SIN16(m) (COS16[((m) + 12) & 15])          /* sin x = cos(x - pi / 2) */
#define JV_HANN16(n) (0.5f - 0.5f * COS16[(n) & 15])  /* periodic hann window of 16 */

// ---------------------------------------------------------------- STFT of the source (generator.py:371-381)
// 16-point DFT, hop 4, periodic hann, center + reflect padding (per utterance at its own length).
struct StftTables {
  float wc[9][16];  // w[n] * cos(2 pi k n / 16)
  float ws[9][16];  // -w[n] * sin(2 pi k n / 16)
};

// One thread per frame computes its 18 values; the block's 128 x 18 outputs are contiguous in SST, so they leave through
// shared memory as 16-byte vectors (one thread per frame storing 18 scalars at a 36-byte stride ran at 8 % of HBM peak).
// The 16 input samples of neighbouring frames overlap by 12: they are read through L1 (__ldg), 4 B per sample from DRAM.
template <typename TA>
__global__ void __launch_bounds__(128) hift_stft_kernel(TA* __restrict__ SST, const int* __restrict__ fr, HiftSeq sq, int rows,
                                                        const float* __restrict__ s, int Tmax, const StftTables tb) {
  __shared__ __align__(16) TA tile[128 * 18];
  // The analysis tables stay in the kernel-parameter constant bank: after full unrolling every tb.wc[k][n] is a compile-time
  // address, i.e. an FFMA operand.  (A shared-memory copy costs one LDS per FFMA: 288 per frame bound the kernel at 83 us.)
  const int m0 = blockIdx.x * 128;
  const int m = m0 + threadIdx.x;
  const int b = m < rows ? fr[m] : -1;
  TA* o = tile + threadIdx.x * 18;
  if (b < 0) {
#pragma unroll
    for (int k = 0; k < 18; ++k) o[k] = DT<TA>::from_f(0.f);
  } else {
    const int f = m - 120 * sq.off[b];
    const int L = 480 * sq.len[b];
    const float* sb = s + (long)b * 480 * Tmax;
    float x[16];
    if (f >= 2 && 4 * f + 8 <= L) {  // interior frame: four aligned 16-byte loads
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(sb + 4 * f - 8) + q);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int n = 0; n < 16; ++n) {
        int i = 4 * f - 8 + n;
        if (i < 0) i = -i;
        if (i >= L) i = 2 * (L - 1) - i;
        x[n] = sb[i];
      }
    }
    // 16-point real DFT of the windowed frame with the n <-> 16 - n and k <-> 8 - k symmetries folded in (~120 FMAs
    // with immediate twiddles instead of 288 table FMAs: the kernel was bound by those, not by HBM)
    float xw[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) xw[n] = x[n] * JV_HANN16(n);
    float a[9], bb[8];  // a_n = xw_n + xw_{16-n}, b_n = xw_n - xw_{16-n}
    a[0] = xw[0];
    a[8] = xw[8];
#pragma unroll
    for (int n = 1; n < 8; ++n) {
      a[n] = xw[n] + xw[16 - n];
      bb[n] = xw[n] - xw[16 - n];
    }
    float re[9], im[9];
#pragma unroll
    for (int k = 0; k <= 4; ++k) {  // Re_k = E + O, Re_{8-k} = E - O (E: even n, O: odd n)
      float E = 0.f, O = 0.f;
#pragma unroll
      for (int n = 0; n <= 8; n += 2) E = fmaf(a[n], JV_COS16(k * n), E);
#pragma unroll
      for (int n = 1; n < 8; n += 2) O = fmaf(a[n], JV_COS16(k * n), O);
      re[k] = E + O;
      re[8 - k] = E - O;
    }
    im[0] = 0.f;
    im[8] = 0.f;
#pragma unroll
    for (int k = 1; k <= 4; ++k) {  // Im_k = -(SE + SO), Im_{8-k} = -(SO - SE)
      float SE = 0.f, SO = 0.f;
#pragma unroll
      for (int n = 2; n < 8; n += 2) SE = fmaf(bb[n], JV_SIN16(k * n), SE);
#pragma unroll
      for (int n = 1; n < 8; n += 2) SO = fmaf(bb[n], JV_SIN16(k * n), SO);
      im[k] = -(SE + SO);
      im[8 - k] = SE - SO;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      o[k] = DT<TA>::from_f(re[k]);
      o[9 + k] = DT<TA>::from_f(im[k]);
    }
  }
  __syncthreads();
  const int valid = min(128, rows - m0);
  constexpr int VEC = 16 / (int)sizeof(TA);  // elements per 16-byte vector
  const int n_el = valid * 18;
  TA* dst = SST + (long)m0 * 18;             // 128 * 18 * sizeof(TA) is a multiple of 16: every block starts aligned
  for (int i = threadIdx.x * VEC; i + VEC <= n_el; i += 128 * VEC)
    *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(tile + i);
  for (int i = (n_el / VEC) * VEC + threadIdx.x; i < n_el; i += 128) dst[i] = tile[i];
}

// out[b, k, f] = SST[120 * off_b + f, k] for f < 120 * len_b + 1, else 0  (inspection hook: jv_hift_stft)
template <typename TA>
__global__ void hift_unpack_stft_kernel(float* __restrict__ out, const TA* __restrict__ SST, HiftSeq sq, int Fmax) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)sq.B * 18 * Fmax) return;
  const int f = (int)(idx % Fmax);
  const int k = (int)((idx / Fmax) % 18);
  const int b = (int)(idx / ((long)Fmax * 18));
  out[idx] = f < 120 * sq.len[b] + 1 ? DT<TA>::to_f(SST[(120L * sq.off[b] + f) * 18 + k]) : 0.f;
}

// ---------------------------------------------------------------- elementwise rows: out = act(in) (fp32 in, TA out)
template <typename TA>
__global__ void act_rows_kernel(const float* __restrict__ in, TA* __restrict__ out, long n, int Cn, int act, float p,
                                const float* __restrict__ vec) {
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(in + i);
  const int c = (int)(i % Cn);
  float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float a = vec ? __ldg(vec + c + j) : 0.f;
    r[j] = sizeof(TA) == 2 ? apply_act_fast(r[j], act, p, a) : apply_act(r[j], act, p, a);  // bf16 mode: single-MUFU forms
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) out[i + j] = DT<TA>::from_f(r[j]);
}

// ReflectionPad1d((1,0)) at stage 2 (generator.py:407-408): x[0] := u[1] (= position 2) before `x + si`.
// X already holds u + si (si was the GEMM's residual), so position 0 = (X[2] - si[2]) + si[0].
__global__ void hift_reflect_fix_kernel(float* __restrict__ X, const float* __restrict__ SI, int Cn, HiftSeq sq, int rate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sq.B * Cn) return;
  int b = i / Cn, c = i % Cn;
  long r0 = (long)rate * sq.off[b];
  X[r0 * Cn + c] = (X[(r0 + 2) * Cn + c] - SI[(r0 + 2) * Cn + c]) + SI[r0 * Cn + c];
}

// im2col of the strided source_downs convolutions (bf16 mode): IM[m, j] = SST_flat[(m*u - u/2)*18 + j], j < klen,
// zero beyond; the window of a stride-u, kernel-2u conv over the 18 STFT channels is contiguous in the channel-last
// layout, so this is a shifted copy.  It turns K_tap = 18 (not TMA-addressable) into one dense K = Kp GEMM.
__global__ void hift_im2col_kernel(bf16* __restrict__ IM, const bf16* __restrict__ SST, long rows_out, int Kp, int klen, int u,
                                   long sst_elems) {
  const int groups = Kp >> 3;
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows_out * groups) return;
  const long m = idx / groups;
  const int j0 = (int)(idx % groups) * 8;
  const long base = (m * u - u / 2) * 18 + j0;
  bf16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long e = base + j;
    v[j] = (j0 + j < klen && e >= 0 && e < sst_elems) ? SST[e] : __float2bfloat16_rn(0.f);
  }
  *reinterpret_cast<uint4*>(IM + m * Kp + j0) = *reinterpret_cast<const uint4*>(v);
}

// x = (r0 + r1 + r2) / 3 (generator.py:415-421), then the leaky_relu of the next consumer -> conv input
template <typename TA>
__global__ void mean3_act_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                                 TA* __restrict__ out, long n, float slope) {
  long i = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float4 x = *reinterpret_cast<const float4*>(a + i);
  const float4 y = *reinterpret_cast<const float4*>(b + i);
  const float4 z = *reinterpret_cast<const float4*>(c + i);
  float r[4] = {(x.x + y.x + z.x) / 3.0f, (x.y + y.y + z.y) / 3.0f, (x.z + y.z + z.z) / 3.0f, (x.w + y.w + z.w) / 3.0f};
#pragma unroll
  for (int j = 0; j < 4; ++j) out[i + j] = DT<TA>::from_f(r[j] > 0.f ? r[j] : r[j] * slope);
}

// ---------------------------------------------------------------- output head (generator.py:425-431, 383-394)
// SPEC[m, 0:9] = log-magnitude, SPEC[m, 9:18] = pre-sin phase for frame m.  y[n] = OLA / envelope, clamp.
constexpr int SPEC_LD = 24;  // conv_post output: 18 channels padded to 24 (the tcgen05 GEMM wants N % 8 == 0; 96-byte rows)

struct IstftTables {
  float cr[9][16];  // c_k cos(2 pi k j / 16) / 16 * w[j]
  float ci[9][16];  // -c_k sin(2 pi k j / 16) / 16 * w[j]  (0 for k = 0, 8)
  float w2[16];
};

constexpr int ISTFT_THREADS = 288;
// One block = 1024 output samples of one utterance = 256 hops (+ 3 frames of overlap) = 259 frames: 288 threads, so that
// phase 1 is ONE round of loads (with 256 threads the 3 overlap frames cost a second, nearly empty, latency round).
// Phase 1, one thread per frame:
// read the frame's 18 values (rows are `ld` = SPEC_LD floats apart: 16-byte loads), magnitude / phase with three MUFU ops
// per bin (ex2, sin, cos .approx after an exact range reduction; libm's expf / sinf / sincosf bound the old kernel), then the
// windowed 16-point inverse DFT as FFMAs whose table operands come straight from the constant bank, and the 16 samples
// go to shared memory.  Phase 2, four samples per thread: overlap-add of the four frames that cover a sample, envelope,
// clamp, one 16-byte store.
__global__ void __launch_bounds__(ISTFT_THREADS) hift_istft_kernel(const float* __restrict__ SPEC, int ld, HiftSeq sq, int Tmax,
                                                         float* __restrict__ wav, const IstftTables tb, float limit, long spec_rows) {
  constexpr int NF = 259;             // frames a block touches: f_lo .. f_lo + 258
  __shared__ float Y[16][NF + 1];     // Y[j][frame]: sample j of the frame's windowed inverse DFT (conflict-free in both phases)
  const int b = blockIdx.y;
  const int T = sq.len[b];
  const int n0 = blockIdx.x * 1024;   // first output sample of this block
  const long Lmax = 480L * Tmax;
  if (n0 >= 480 * Tmax) return;
  float4* out = reinterpret_cast<float4*>(wav + (long)b * Lmax + n0) + threadIdx.x;
  const int n = n0 + 4 * threadIdx.x;
  const bool writer = threadIdx.x < 256;
  if (n0 >= 480 * T) {  // block entirely beyond the utterance
    if (writer && n < 480 * Tmax) *out = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int F = 120 * T + 1;
  const int f_lo = n0 / 4 - 1;        // sample n is covered by frames (n + 8) / 4 - 3 .. (n + 8) / 4
  const long row0 = 120L * sq.off[b];
  if (const int fi = threadIdx.x; fi < NF) {
    const int f = f_lo + fi;
    float y[16];
    if (f >= 0 && f < F && row0 + f < spec_rows) {
      const float4* sp = reinterpret_cast<const float4*>(SPEC + (row0 + f) * ld);
      float v[20];
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const float4 t = __ldcs(sp + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
      float re[9], im[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float mag = fminf(exp_fast(v[k]), 100.0f);
        const float px = v[9 + k];
        const float ph = sin_ftz(px - 6.283185307179586f * rintf(px * 0.15915494309189535f));  // sin(x), x reduced to [-pi, pi]
        float cs;
        asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(ph));  // |ph| <= 1
        re[k] = mag * cs;
        im[k] = mag * sin_ftz(ph);
      }
      // windowed inverse real DFT, y[j] = w[j] / 16 * sum_k c_k (Re_k cos(2 pi k j / 16) - Im_k sin(2 pi k j / 16)), c = 1 for
      // k in {0, 8} else 2, with the j <-> 16 - j and j <-> 8 - j symmetries folded in (immediate twiddles, ~115 FMAs)
      float cre[9], cim[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float c = (k == 0 || k == 8) ? 1.0f : 2.0f;
        cre[k] = c * re[k];
        cim[k] = c * im[k];
      }
      float A[9], Bs[9];  // A symmetric, Bs antisymmetric in j <-> 16 - j
#pragma unroll
      for (int j = 0; j <= 4; ++j) {  // A[j] = AE + AO, A[8 - j] = AE - AO (even / odd k)
        float AE = 0.f, AO = 0.f;
#pragma unroll
        for (int k = 0; k <= 8; k += 2) AE = fmaf(cre[k], JV_COS16(k * j), AE);
#pragma unroll
        for (int k = 1; k < 8; k += 2) AO = fmaf(cre[k], JV_COS16(k * j), AO);
        A[j] = AE + AO;
        A[8 - j] = AE - AO;
      }
      Bs[0] = 0.f;
      Bs[8] = 0.f;
#pragma unroll
      for (int j = 1; j <= 4; ++j) {  // Bs[j] = BE + BO, Bs[8 - j] = BO - BE
        float BE = 0.f, BO = 0.f;
#pragma unroll
        for (int k = 2; k < 8; k += 2) BE = fmaf(cim[k], JV_SIN16(k * j), BE);
#pragma unroll
        for (int k = 1; k < 8; k += 2) BO = fmaf(cim[k], JV_SIN16(k * j), BO);
        Bs[j] = BE + BO;
        Bs[8 - j] = BO - BE;
      }
      y[0] = (JV_HANN16(0) * 0.0625f) * A[0];
      y[8] = (JV_HANN16(8) * 0.0625f) * A[8];
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        y[j] = (JV_HANN16(j) * 0.0625f) * (A[j] - Bs[j]);
        y[16 - j] = (JV_HANN16(16 - j) * 0.0625f) * (A[j] + Bs[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) Y[j][fi] = y[j];
  }
  __syncthreads();
  if (!writer || n >= 480 * Tmax) return;
  float r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ni = n + i;
    float yv = 0.f;
    if (ni < 480 * T) {
      const int f_hi = (ni + 8) / 4;
      float acc = 0.f, env = 0.f;
#pragma unroll
      for (int d = 3; d >= 0; --d) {
        const int f = f_hi - d;
        if (f < 0 || f >= F) continue;
        const int j = ni + 8 - 4 * f;  // = 4 d + i: compile-time after unrolling
        acc += Y[j][f - f_lo];
        env += JV_HANN16(j) * JV_HANN16(j);
      }
      yv = acc / env;
      yv = fminf(fmaxf(yv, -limit), limit);
    }
    r[i] = yv;
  }
  *out = make_float4(r[0], r[1], r[2], r[3]);
}

}  // namespace jv
