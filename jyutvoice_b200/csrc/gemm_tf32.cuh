// fp32-accurate GEMM-with-taps on the tensor cores: tcgen05.mma kind::tf32 with 3xTF32 error compensation
// (SURVEY.md section 7 hard part 1: single-pass TF32 gives 5e-3 mel error and fails the 1e-3 gate, the hi / lo split
// gives ~5e-6).  This is the contraction engine of precision mode fp32 and of the text front, in place of the FFMA kernel:
//     A = A_hi + A_lo,  W = W_hi + W_lo   (X_hi = X with the low 13 mantissa bits cleared: exactly a TF32 number)
//     A W^T  ~=  A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T          (the dropped A_lo W_lo^T term is ~2^-22 relative)
// accumulated in fp32 in TMEM.  Weights are split once at finalize (W_hi / W_lo in global memory); activations arrive as
// plain fp32 by TMA and are split IN SHARED MEMORY by two converter warps (hi written back in place, lo into a second tile,
// fence.proxy.async), so no producer has to know about the split and no extra global traffic exists.
//   warp 0 lane 0 : TMA producer   (A tile 128 x 32 fp32, W_hi / W_lo tiles block_n x 32 fp32: 128-byte rows, SWIZZLE_128B)
//   warp 1 lane 0 : MMA issuer     (12 tcgen05.mma.kind::tf32 per K block: 4 k-steps of 8 x 3 products)
//   warps 2, 3    : converters     (warp 2 also owns the TMEM allocation)
//   warps 4 .. 7  : epilogue       (one thread per output row; same epilogue steps as gemm_epilogue_store, vector stores)
// LayerNorm epilogues stay separate row kernels in fp32 mode (Engine::gemm_lowered), exactly as with the FFMA engine.
#pragma once
#include "gemm_tc.cuh"

namespace jv {
namespace tf32 {

using namespace tc;

constexpr int BK = 32;                 // fp32 elements per K block: 128 bytes = one SWIZZLE_128B row
constexpr int A_BYTES = BLOCK_M * 128;  // 16 KB
constexpr int THREADS = 256;
constexpr int MAX_ST = 4;

struct Maps {
  CUtensorMap a0, a1, w_hi, w_lo;
};
struct Params {
  int block_n, n_tiles_n, num_tiles, stages, w_bytes, k_blocks_per_tap;
  int vec_ok;  // 16-byte epilogue loads / stores are legal (pitches and pointers aligned)
};

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One chunk of 32 consecutive columns of one output row: the epilogue of gemm_epilogue_store<float>, 16-byte accesses.
__device__ __forceinline__ void epilogue_chunk(const GemmDesc& g, const Params& p, long orow, int fr, int n0, const uint32_t (&acc)[32]) {
  const int nv = min(32, g.N - n0);
  if (nv <= 0) return;
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  if (g.bias) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nv) v[j] += __ldg(g.bias + n0 + j);
  }
  if (g.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nv) v[j] = apply_act(v[j], g.act, g.act_param, g.act_vec ? __ldg(g.act_vec + n0 + j) : 0.f);
  }
  if (g.add_row && fr >= 0) {
    const float* ar = g.add_row + (long)g.row_tidx[fr] * g.add_row_stride + n0;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nv) v[j] += ar[j];
  }
  if (fr < 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
  }
  const bool vec = p.vec_ok && nv == 32;
  if (g.resid) {
    const float* r = g.resid + orow * g.ldr + n0;
    if (vec) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 t = *reinterpret_cast<const float4*>(r + 4 * j4);
        v[4 * j4] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) v[j] += r[j];
    }
  }
  if (g.out_f32) {
    float* o = g.out_f32 + orow * g.ldo + n0;
    if (vec) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) o[j] = v[j];
    }
  }
  if (g.out_act) {
    if (g.act2 != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) v[j] = apply_act(v[j], g.act2, g.act2_param, g.act2_vec ? __ldg(g.act2_vec + n0 + j) : 0.f);
    }
    float* o = (float*)g.out_act + orow * g.ldo2 + n0;
    if (vec) {
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) *reinterpret_cast<float4*>(o + 4 * j4) = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nv) o[j] = v[j];
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1) gemm_taps_tf32_kernel(const __grid_constant__ Maps tm, const GemmDesc g, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int stage_bytes = 2 * A_BYTES + 2 * p.w_bytes;  // A (raw -> hi) | A_lo | W_hi | W_lo
  const uint32_t bars = base + p.stages * stage_bytes;
  const uint32_t full_bar = bars, conv_bar = bars + 8 * MAX_ST, empty_bar = bars + 16 * MAX_ST;
  const uint32_t tfull_bar = bars + 24 * MAX_ST, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_iters = g.n_taps * p.k_blocks_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.w_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.w_lo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(conv_bar + 8 * i, 64);
      mbar_init(empty_bar + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot_ptr;
  auto sA = [&](int st) { return base + st * stage_bytes; };
  auto sAl = [&](int st) { return base + st * stage_bytes + A_BYTES; };
  auto sWh = [&](int st) { return base + st * stage_bytes + 2 * A_BYTES; };
  auto sWl = [&](int st) { return base + st * stage_bytes + 2 * A_BYTES + p.w_bytes; };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      const uint32_t tx = A_BYTES + 2u * (uint32_t)p.block_n * 128u;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles_n) * BLOCK_M, n0 = (tile % p.n_tiles_n) * p.block_n;
        for (int s = 0; s < g.n_taps; ++s) {
          const CUtensorMap* tmA = g.tap_src[s] ? &tm.a1 : &tm.a0;
          for (int kb = 0; kb < p.k_blocks_per_tap; ++kb) {
            mbar_wait(empty_bar + 8 * st, ph ^ 1, 61);
            mbar_expect_tx(full_bar + 8 * st, tx);
            tma_load_2d(tmA, full_bar + 8 * st, sA(st), kb * BK, m0 + g.tap_shift[s]);
            tma_load_2d(&tm.w_hi, full_bar + 8 * st, sWh(st), s * g.K_tap + kb * BK, n0);
            tma_load_2d(&tm.w_lo, full_bar + 8 * st, sWl(st), s * g.K_tap + kb * BK, n0);
            if (++st == p.stages) { st = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D fp32 (bit 4), A and B TF32 (code 2 at [7,10) and [10,13)), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      int st = 0, acc = 0;
      uint32_t ph = 0, aph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + 8 * acc, aph ^ 1, 62);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(conv_bar + 8 * st, ph, 63);  // A split into hi / lo (which also implies the TMA data has landed)
          tc_fence_after();
          const uint64_t ah = make_smem_desc(sA(st)), al = make_smem_desc(sAl(st)), wh = make_smem_desc(sWh(st)), wl = make_smem_desc(sWl(st));
#pragma unroll
          for (int k = 0; k < BK / 8; ++k) {  // 8 fp32 = 32 bytes per k-step: +2 in the descriptor's address field
            umma_tf32(d, al + 2 * k, wh + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);  // small terms first
            umma_tf32(d, ah + 2 * k, wl + 2 * k, idesc, 1u);
            umma_tf32(d, ah + 2 * k, wh + 2 * k, idesc, 1u);
          }
          umma_commit(empty_bar + 8 * st);
          if (++st == p.stages) { st = 0; ph ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc);
        if (++acc == 2) { acc = 0; aph ^= 1; }
      }
    }
  } else if (warp < 4) {
    // ===================== converters: A -> (A_hi in place, A_lo) =====================
    const int t = threadIdx.x - 64;  // 0 .. 63
    int st = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int it = 0; it < k_iters; ++it) {
        mbar_wait(full_bar + 8 * st, ph, 64);
        const uint32_t a = sA(st), al = sAl(st);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const uint32_t off = (uint32_t)(i * 64 + t) * 16u;
          const float4 v = lds128(a + off);
          const float hx = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), hy = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
          const float hz = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u), hw = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
          sts128(a + off, hx, hy, hz, hw);
          sts128(al + off, v.x - hx, v.y - hy, v.z - hz, v.w - hw);
        }
        fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
        mbar_arrive(conv_bar + 8 * st);
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: one thread per output row =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t aph = 0;
    const int n_chunks = (p.block_n + 31) >> 5;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.n_tiles_n) * BLOCK_M, n0 = (tile % p.n_tiles_n) * p.block_n;
      const int m = m0 + q * 32 + lane;
      const long orow = (long)m * g.o_stride + g.o_off;
      const bool in_range = m < g.M && orow < g.o_rows;
      const int fr = in_range ? (g.frame_row ? g.frame_row[orow] : 0) : -1;
      mbar_wait(tfull_bar + 8 * acc, aph, 65);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256;
      for (int c = 0; c < n_chunks; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        if (in_range) epilogue_chunk(g, p, orow, fr, n0 + c * 32, r);
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == 2) { acc = 0; aph ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tf32

// JYUTVOICE_B200_TF32=0: precision mode fp32 and the text front run every contraction on the FFMA engine (gemm_simt.cuh)
static inline bool use_tf32() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_TF32");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static inline bool gemm_tf32_supported(const GemmDesc& g) {
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (!g.W_hi || !g.W_lo || g.a_stride != 1) return false;
  if (g.K_tap % tf32::BK != 0 || g.n_taps < 1 || g.n_taps > MAX_TAPS) return false;
  for (int s = 0; s < 2; ++s) {
    if (!g.A[s]) continue;
    if (!al16(g.A[s]) || g.lda[s] % 4 != 0 || g.lda[s] < g.K_tap) return false;
  }
  if (g.ln1_gamma || g.ln2_gamma || g.x_bf16) return false;  // LayerNorms are separate row kernels on this path
  return g.M > 0 && g.N > 0;
}

static inline void launch_gemm_tf32(const GemmDesc& g, TmapCache& cache, int num_sms, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(tf32::gemm_taps_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
  tf32::Params p;
  p.block_n = round_up(std::min(g.N, 256), 16);
  p.n_tiles_n = cdiv(g.N, p.block_n);
  const int m_tiles = cdiv(g.M, tc::BLOCK_M);
  p.num_tiles = m_tiles * p.n_tiles_n;
  p.w_bytes = round_up(p.block_n * 128, 1024);
  p.k_blocks_per_tap = g.K_tap / tf32::BK;
  const int stage_bytes = 2 * tf32::A_BYTES + 2 * p.w_bytes;
  p.stages = std::min(tf32::MAX_ST, (tc::SMEM_LIMIT - 2048) / stage_bytes);
  JV_REQUIRE(p.stages >= 2, JV_ERR_STATE, "not enough shared memory for the TF32 GEMM pipeline");
  auto al16 = [](const void* q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
  p.vec_ok = (al16(g.resid) && al16(g.out_f32) && al16(g.out_act) && (!g.resid || g.ldr % 4 == 0) && (!g.out_f32 || g.ldo % 4 == 0) &&
              (!g.out_act || g.ldo2 % 4 == 0) && g.o_stride * (long)1 >= 1)
                 ? 1
                 : 0;
  if (cache.maps.size() > 4096) cache.maps.clear();
  const long Ktot = (long)g.n_taps * g.K_tap;
  tf32::Maps tm;
  // kind 1 = fp32, SWIZZLE_128B (box inner 32 elements = 128 bytes)
  tm.a0 = cache.get(g.A[0], g.K_tap, g.a_rows[0], (long)g.lda[0] * 4, tf32::BK, tc::BLOCK_M, 1);
  tm.a1 = g.A[1] ? cache.get(g.A[1], g.K_tap, g.a_rows[1], (long)g.lda[1] * 4, tf32::BK, tc::BLOCK_M, 1) : tm.a0;
  tm.w_hi = cache.get(g.W_hi, Ktot, g.N, Ktot * 4, tf32::BK, p.block_n, 1);
  tm.w_lo = cache.get(g.W_lo, Ktot, g.N, Ktot * 4, tf32::BK, p.block_n, 1);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(std::min(p.num_tiles, num_sms));
  cfg.blockDim = dim3(tf32::THREADS);
  cfg.dynamicSmemBytes = 1024 + p.stages * stage_bytes + 256;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  JV_CUDA(cudaLaunchKernelEx(&cfg, tf32::gemm_taps_tf32_kernel, tm, g, p));
  JV_LAUNCHED();
}

}  // namespace jv
