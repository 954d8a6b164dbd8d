// tcgen05 / TMA / TMEM GEMM-with-taps for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// One persistent, warp-specialised kernel:
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n, K=16)
//   warp 2        : TMEM allocator (512 columns = 2 accumulator stages of up to 256 fp32 columns)
//   warps 4..7    : epilogue      (tcgen05.ld 32x32b.x32 -> registers -> bias/residual/act/mask -> global)
// "Taps" make Conv1d a GEMM without im2col: tap s reads the A tile shifted by tap_shift[s] rows; TMA
// zero-fills rows outside the tensor, and the packed activation layout keeps >= |shift| zero rows
// between utterances, which is exactly the conv's zero padding.
#pragma once
#include "common.cuh"

namespace jv {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements: 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int B_STAGE_BYTES = 256 * BLOCK_K * 2;      // 32 KB (block_n <= 256)
constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (with a message) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("jyutvoice_b200: mbarrier wait timed out (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B)
//   [46,48) version=1 | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Epilogue for 8 consecutive columns of one output row (vectorised global access).
__device__ __forceinline__ void epilogue8(const GemmDesc& g, long orow, int n, bool row_valid, const uint32_t* acc) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[j]);
  if (g.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(g.bias + n + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (g.resid) {
    const float4 r0 = *reinterpret_cast<const float4*>(g.resid + orow * g.ldr + n);
    const float4 r1 = *reinterpret_cast<const float4*>(g.resid + orow * g.ldr + n + 4);
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (g.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], g.act, g.act_param, g.act_vec ? __ldg(g.act_vec + n + j) : 0.f);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = row_valid ? v[j] * g.out_scale : 0.f;
  if (g.out_f32) {
    float4* p = reinterpret_cast<float4*>(g.out_f32 + orow * g.ldo + n);
    if (g.accumulate) {
      const float4 o0 = p[0], o1 = p[1];
      v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
      v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
    }
    p[0] = make_float4(v[0], v[1], v[2], v[3]);
    p[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (g.out_act) {
    if (g.act2 != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], g.act2, g.act2_param, g.act2_vec ? __ldg(g.act2_vec + n + j) : 0.f);
    }
    __nv_bfloat162 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>((bf16*)g.out_act + orow * g.ldo2 + n) = *reinterpret_cast<uint4*>(h);
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_taps_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                    const __grid_constant__ CUtensorMap tmW, const GemmDesc g, const int block_n,
                    const int n_tiles_n, const int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t smem_a = base;
  const uint32_t smem_b = base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_b + STAGES * B_STAGE_BYTES;
  const uint32_t full_bar = bars;                      // STAGES x 8 B
  const uint32_t empty_bar = bars + 8 * STAGES;        // STAGES x 8 B
  const uint32_t tfull_bar = bars + 16 * STAGES;       // 2 x 8 B
  const uint32_t tempty_bar = bars + 16 * STAGES + 16; // 2 x 8 B
  const uint32_t tmem_slot = bars + 16 * STAGES + 32;  // 4 B
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blocks_per_tap = g.K_tap / BLOCK_K;
  const int k_iters = g.n_taps * k_blocks_per_tap;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmW) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + 8 * i, 1);
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
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + block_n * BLOCK_K * 2;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * BLOCK_M;
        const int n0 = (tile % n_tiles_n) * block_n;
        for (int s = 0; s < g.n_taps; ++s) {
          const CUtensorMap* tmA = g.tap_src[s] ? &tmA1 : &tmA0;
          const int arow = m0 + g.tap_shift[s];
          for (int kb = 0; kb < k_blocks_per_tap; ++kb) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1, 1);
            mbar_expect_tx(full_bar + 8 * stage, tx_bytes);
            tma_load_2d(tmA, full_bar + 8 * stage, smem_a + stage * A_STAGE_BYTES, kb * BLOCK_K, arow);
            tma_load_2d(&tmW, full_bar + 8 * stage, smem_b + stage * B_STAGE_BYTES, s * g.K_tap + kb * BLOCK_K, n0);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
      // a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar + 8 * acc_stage, acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * 256;
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(full_bar + 8 * stage, phase, 3);
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + stage * A_STAGE_BYTES);
          const uint64_t bdesc = make_smem_desc(smem_b + stage * B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 B (= 16 bf16) inside the 128 B swizzle row: +2 in the >>4 address field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + 8 * stage);  // frees the smem slot when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc_stage);  // accumulator complete -> epilogue
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles_n) * BLOCK_M;
      const int n0 = (tile % n_tiles_n) * block_n;
      mbar_wait(tfull_bar + 8 * acc_stage, acc_phase, 4);
      tc_fence_after();
      const int m = m0 + q * 32 + lane;
      const long orow = (long)m * g.o_stride + g.o_off;
      const bool in_range = m < g.M && orow < g.o_rows;
      const bool row_valid = in_range && (g.frame_row == nullptr || g.frame_row[orow] >= 0);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc_stage * 256;
      for (int c0 = 0; c0 < block_n; c0 += 32) {
        uint32_t acc[32];
        tmem_ld32(taddr + c0, acc);  // warp-collective: every lane participates
        if (in_range) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const int n = n0 + c0 + j;
            if (n + 8 <= g.N) epilogue8(g, orow, n, row_valid, acc + j);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar + 8 * acc_stage);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled();

struct TmapKey {
  const void* ptr;
  long inner, rows, pitch;
  int box_rows;
  bool operator<(const TmapKey& o) const {
    if (ptr != o.ptr) return ptr < o.ptr;
    if (inner != o.inner) return inner < o.inner;
    if (rows != o.rows) return rows < o.rows;
    if (pitch != o.pitch) return pitch < o.pitch;
    return box_rows < o.box_rows;
  }
};

// Cache of encoded tensor maps (encoding costs ~1 us; the same buffers recur every layer / step).
struct TmapCache {
  std::map<TmapKey, CUtensorMap> maps;
  const CUtensorMap& get(const void* ptr, long inner_elems, long rows, long pitch_elems, int box_rows);
};

// True when the tcgen05 kernel can run this problem (else the caller uses the FFMA engine).
struct ProfileState {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // pairs
  double flops = 0.0;
};
ProfileState& profile_state();

bool gemm_tc_supported(const GemmDesc& g);
void launch_gemm_tc(const GemmDesc& g, TmapCache& cache, int num_sms, cudaStream_t st);

}  // namespace jv
