// fp32-FFMA GEMM-with-taps: the exact-arithmetic engine of JV_PREC_FP32, and the engine for the
// few shapes TMA cannot address (strided source_downs convs, K_tap not a multiple of 64).
// Operands of type TA (float or bf16) are widened to fp32; accumulation is fp32 FFMA.
#pragma once
#include "common.cuh"

namespace jv {

// Shared epilogue: one output element.
template <typename TA>
__device__ __forceinline__ void gemm_epilogue_store(const GemmDesc& g, int m, int n, float acc) {
  long orow = (long)m * g.o_stride + g.o_off;
  if (orow >= g.o_rows) return;
  float v = acc;
  if (g.bias) v += __ldg(g.bias + n);
  v = apply_act(v, g.act, g.act_param, g.act_vec ? __ldg(g.act_vec + n) : 0.f);
  const int fr = g.frame_row ? g.frame_row[orow] : 0;
  if (g.add_row && fr >= 0) v += g.add_row[(long)g.row_tidx[fr] * g.add_row_stride + n];
  if (fr < 0) v = 0.f;
  if (g.resid) v += g.resid[orow * g.ldr + n];
  if (g.out_f32) g.out_f32[orow * g.ldo + n] = v;
  if (g.out_act) {
    float w = apply_act(v, g.act2, g.act2_param, g.act2_vec ? __ldg(g.act2_vec + n) : 0.f);
    ((TA*)g.out_act)[orow * g.ldo2 + n] = DT<TA>::from_f(w);
  }
}

template <typename TA>
__global__ void __launch_bounds__(256) gemm_taps_simt_kernel(const GemmDesc g) {
  constexpr int BM = 128, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int a_r = tid >> 1, a_k = (tid & 1) * 8;   // A tile: row, k segment of 8
  const int w_n = tid >> 2, w_k = (tid & 3) * 4;   // W tile: n, k segment of 4
  const int Ktot = g.n_taps * g.K_tap;
  const TA* Wp = (const TA*)g.W;

  for (int s = 0; s < g.n_taps; ++s) {
    const int src = g.tap_src[s];
    const TA* Ap = (const TA*)g.A[src];
    const int lda = g.lda[src];
    const long arow = (long)(m0 + a_r) * g.a_stride + g.tap_shift[s];
    const bool arow_ok = (m0 + a_r) < g.M && arow >= 0 && arow < g.a_rows[src];
    for (int k0 = 0; k0 < g.K_tap; k0 += BK) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int k = k0 + a_k + i;
        float v = 0.f;
        if (arow_ok && k < g.K_tap) v = DT<TA>::to_f(Ap[arow * lda + k]);
        As[a_k + i][a_r] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + w_k + i;
        float v = 0.f;
        if ((n0 + w_n) < g.N && k < g.K_tap) v = DT<TA>::to_f(Wp[(long)(n0 + w_n) * Ktot + s * g.K_tap + k]);
        Ws[w_k + i][w_n] = v;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[8], b[4];
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
        a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
        b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int m = m0 + ty * 8 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < g.N) gemm_epilogue_store<TA>(g, m, n, acc[i][j]);
    }
  }
}

template <typename TA>
static void launch_gemm_simt(const GemmDesc& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  dim3 grid(cdiv(g.M, 128), cdiv(g.N, 64));
  gemm_taps_simt_kernel<TA><<<grid, 256, 0, st>>>(g);
  JV_LAUNCHED();
}

}  // namespace jv
