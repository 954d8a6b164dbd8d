// tcgen05 / TMA / TMEM GEMM-with-taps for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
// One persistent, warp-specialised kernel, 384 threads, one CTA per SM:
//   warp 0 lane 0 : TMA producer  (cp.async.bulk.tensor.2d -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1 lane 0 : MMA issuer    (tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n, K=16)
//   warp 2        : TMEM allocator (512 columns = 2 accumulator stages of up to 256 fp32 columns)
//   warps 4..7    : epilogue group 0 (accumulator stage 0: even tiles of this CTA)
//   warps 8..11   : epilogue group 1 (accumulator stage 1: odd tiles)
// Epilogue: one thread owns one output row (TMEM lane).  Per 32-column chunk: tcgen05.ld -> registers ->
// bias / LayerNorm / activation / time-embedding / mask / residual -> swizzled smem staging -> TMA store.
// The residual tile arrives by TMA load into the same staging buffer, so all global traffic of the
// epilogue is coalesced by the TMA engine.  LayerNorm over the 256 channels of a frame is row-local:
// statistics are taken in extra sweeps over TMEM (which doubles as the scratch for the post-LN value).
//
// "Taps" make Conv1d a GEMM without im2col: tap s reads the A tile shifted by tap_shift[s] rows; TMA
// zero-fills rows outside the tensor, and the packed activation layout keeps >= |shift| zero rows
// between utterances, which is exactly the conv's zero padding.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace jv {
namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // bf16 elements: 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 8;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
constexpr int EPI_WARPS = 8;       // epilogue warps of the general kernels (two groups of four)
constexpr int EPI_WARPS_MAX = 16;  // bf16-only epilogues run two warps per (group, lane quarter): even / odd column chunks
constexpr int EPI_F32_BYTES = 32 * 32 * 4;  // 4 KB: 32 rows x 128 B, SWIZZLE_128B
constexpr int EPI_B16_BYTES = 32 * 32 * 2;  // 2 KB: 32 rows x 64 B, SWIZZLE_64B
constexpr int EPI_BYTES_PER_WARP = 2 * EPI_F32_BYTES + EPI_B16_BYTES;  // residual-in 4 KB | fp32-out 4 KB | bf16-out 2 KB
constexpr int BAR_BYTES = 640;   // mbarriers
constexpr int MAX_ACC = 8;       // accumulator stages in TMEM (512 columns / block_n, at most 8)
constexpr int TAB_BYTES = 1024;  // slab mode: a_off[32] + b_desc[<=96] (uint64)
constexpr int XCH_BYTES = 4096;  // row-statistics exchange of LayerNorm epilogues: 4 quarters x (2 or 4) warps x 32 lanes x float2
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 384;
constexpr int NUM_THREADS_WIDE = 128 + 32 * EPI_WARPS_MAX;  // 640

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
  uint32_t polls = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 4095u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > 4000000000LL) {
        printf("jyutvoice_b200: mbarrier wait timed out (tag %d, block %d, thread %d)\n", tag, blockIdx.x, threadIdx.x);
        __trap();
      }
    }
  }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* tm, uint32_t bar, uint32_t dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster issue ONE M = 256 MMA; each feeds its own 128 activation rows and
// HALF of the weight tile, so an SM ingests half the weight bytes per flop (the N = 256 kernels are bound by L2 -> SM
// operand traffic, ~43 B / clk / SM: DESIGN.md section 6).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {  // same smem offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// TMA load whose mbarrier may live in the peer CTA (the leader collects both CTAs' bytes on one barrier)
__device__ __forceinline__ void tma_load_2d_cg2(const CUtensorMap* tm, uint32_t bar_cluster, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tm), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {  // arrives on `bar`'s offset in every CTA of `mask`
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// Remote arrive with release semantics at cluster scope: publishes this thread's prior writes to the peer CTA.  It is a
// cluster-scope fence: measured (ncu, fused feed-forward) it stalls the warp for several hundred cycles -- use it once per
// CTA and event, and the relaxed form below where nothing has to be published (e.g. "TMEM drained").
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)tm), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read4() { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Programmatic dependent launch: the prologue (barrier init, TMEM allocation, smem vector cache) runs while the
// previous kernel in the stream drains; nothing produced by that kernel is touched before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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
// issue only: several loads in flight, one tmem_ld_wait() for all of them
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- row-chunk helpers: 32 consecutive columns of one row per thread.
// Every optional epilogue step is its own unswitched loop over the 32 registers, so a feature that is off costs one
// uniform branch per chunk instead of one per element.  Per-column vectors are read as float4 (same address in every
// lane -> one broadcast transaction), guarded in groups of 8 columns for N that is not a multiple of 32.
__device__ __forceinline__ void add_vec32(float (&v)[32], const float* __restrict__ pv, int n_valid) {
#pragma unroll
  for (int j8 = 0; j8 < 4; ++j8) {
    if (j8 * 8 + 8 <= n_valid) {
      const float4 a = *reinterpret_cast<const float4*>(pv + j8 * 8);
      const float4 b = *reinterpret_cast<const float4*>(pv + j8 * 8 + 4);
      v[j8 * 8 + 0] += a.x; v[j8 * 8 + 1] += a.y; v[j8 * 8 + 2] += a.z; v[j8 * 8 + 3] += a.w;
      v[j8 * 8 + 4] += b.x; v[j8 * 8 + 5] += b.y; v[j8 * 8 + 6] += b.z; v[j8 * 8 + 7] += b.w;
    }
  }
}
__device__ __forceinline__ void ln_affine32(float (&v)[32], float mean, float rstd, const float* __restrict__ gamma,
                                            const float* __restrict__ beta) {
  // two FMAs per value: v * rstd - mean * rstd, then the affine (the epilogues are instruction-bound; the rounding of
  // mean * rstd costs 6e-8 * |mean / std| absolute, far below the bf16 rounding of the result)
  const float nmr = -mean * rstd;
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 gm = *reinterpret_cast<const float4*>(gamma + j4 * 4);
    const float4 bt = *reinterpret_cast<const float4*>(beta + j4 * 4);
    v[j4 * 4 + 0] = fmaf(fmaf(v[j4 * 4 + 0], rstd, nmr), gm.x, bt.x);
    v[j4 * 4 + 1] = fmaf(fmaf(v[j4 * 4 + 1], rstd, nmr), gm.y, bt.y);
    v[j4 * 4 + 2] = fmaf(fmaf(v[j4 * 4 + 2], rstd, nmr), gm.z, bt.z);
    v[j4 * 4 + 3] = fmaf(fmaf(v[j4 * 4 + 3], rstd, nmr), gm.w, bt.w);
  }
}
__device__ __forceinline__ void act32(float (&v)[32], int act, float prm, const float* __restrict__ vec, int n_valid) {
  switch (act) {
    case ACT_GELU:
      // bf16 mode: GELU through the hardware tanh (tanh.approx.f32).  |GELU_tanh - GELU_erf| <= 4.7e-4, measured
      // end effect on the estimator output 6e-4 max-abs vs 3e-2 from the bf16 operands.  fp32 mode uses erff.
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = v[j];
        const float u = x * fmaf(x * x, 0.0356774081f, 0.7978845608f);
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(u));
        const float hx = 0.5f * x;
        v[j] = fmaf(hx, th, hx);
      }
      break;
    case ACT_ELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + (exp_fast(fminf(v[j], 0.f)) - 1.0f);
      break;
    case ACT_LRELU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f) + prm * fminf(v[j], 0.f);
      break;
    case ACT_MISH:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fast_mish(v[j]);
      break;
    case ACT_SILU:
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = v[j] * rcp_ftz(1.f + exp_fast(-v[j]));
      break;
    case ACT_SNAKE:
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        if (j8 * 8 + 8 <= n_valid) {
          const float4 a0 = *reinterpret_cast<const float4*>(vec + j8 * 8);
          const float4 a1 = *reinterpret_cast<const float4*>(vec + j8 * 8 + 4);
          const float al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float sn = sin_ftz(v[j8 * 8 + j] * al[j]);
            v[j8 * 8 + j] = fmaf(rcp_ftz(al[j] + 1e-9f) * sn, sn, v[j8 * 8 + j]);
          }
        }
      }
      break;
    default:
      break;
  }
}
__device__ __forceinline__ void acc_to_f32(const uint32_t (&acc)[32], float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
}

struct TcMaps {
  CUtensorMap a0, a1, w;        // operands (bf16, box 64 x rows, SWIZZLE_128B)
  CUtensorMap resid;            // fp32 [view rows, N], box 32 x 32, SWIZZLE_128B
  CUtensorMap out_f32;          // fp32 [view rows, N], box 32 x 32, SWIZZLE_128B
  CUtensorMap out_act;          // bf16 [view rows, N], box 32 x 32, SWIZZLE_64B
  CUtensorMap out_ln;           // bf16 [view rows, N], box 32 x 32, SWIZZLE_64B
};

struct TcParams {
  int block_n, n_tiles_n, num_tiles, stages, b_stage_bytes;
  int cluster;    // CTAs per cluster (1 or 2): the CTAs of a cluster work on consecutive m-tiles of the same n-tile and
                  // each loads 1/cluster of the weight tile, multicast to all of them (halves the L2 -> SM weight traffic)
  int num_units;  // ceil(m_tiles / cluster) * n_tiles_n   (weight-resident mode: m_tiles)
  int pair;       // cluster == 2 and the two CTAs form one cta_group::2 MMA (M = 256): each loads half of the weight tile,
                  // nothing is multicast; only rank 0 issues MMAs, and it collects both CTAs' TMA bytes and epilogue arrivals
  // Weight-resident mode (K_total * block_n * 2 <= 128 KB, bf16-only output: QKV, FF1): a CTA is bound to one n-tile,
  // loads its whole weight tile into smem ONCE and streams only activation tiles through the ring.  The kernel is
  // bound by SM <-> L2 traffic, and the weight tile is 2/3 of a 128 x 256 tile's operand bytes.
  int wres;
  // Slab mode (implies wres; stride-1 convs, one activation source): per 64-channel K block ONE activation slab of
  // 128 + (max_shift - min_shift) rows is loaded, and every tap multiplies a row-shifted view of it (descriptor start address) with its resident weight tile.  A conv tap then costs no activation traffic at all.
  int slab, slab_rows, min_shift, a_stage_bytes;
  // Accumulator stages in TMEM: 512 / block_n (2 for N = 256 ... 8 for N <= 64).  Tile-parallel epilogue: with few
  // 32-column chunks per tile (narrow N) each share of the epilogue warps takes whole tiles (tile t -> share t % N_SUB)
  // instead of splitting one tile's chunks, so several tiles are in flight and the per-tile TMA round trips overlap.
  int n_acc, acc_cols, tile_par;
  int tab_off;  // slab mode: smem offset of the per-tap descriptor tables (the MMA thread must issue one MMA per ~32 cycles)
  int b_region_bytes;      // ring: stages * b_stage_bytes; resident: k_iters * b_stage_bytes
  int epi_bytes_per_warp;  // staging per epilogue warp; pieces at off_R (residual in, 4 KB), off_OF (fp32 out, 4 KB or the
  int off_R, off_OF, off_OB;  // first bf16 buffer), off_OB (second bf16 buffer, 2 KB); equal offsets = shared / single buffer
  // Per-column vectors cached in smem once per CTA (byte offsets from the aligned smem base; 0 = read from global).
  // Only when the CTA's n-tile is fixed (one n-tile, or weight-resident mode).  With 227 KB of smem there is no L1,
  // so every uncached read of bias / gamma / beta / alpha is an L2 round trip inside the epilogue's dependency chain.
  int vec_bias, vec_ln1, vec_ln2, vec_act, vec_act2;
  int debug;               // experiments only (JYUTVOICE_B200_DEBUG): 1 = no output stores, 2 = no epilogue at all
  int epi16;               // XB + residual + LN2 kernels with N == block_n == 256: the sixteen-warp epilogue (WLN instantiations)
  long long* trace;        // debug (tools/gemm_trace.py): per-CTA clock64 sums, 8 slots per CTA: MMA thread total | wait tempty |
                           // wait full | units ; epilogue warp 0: total | wait tfull | wait residual | -
};

// byte offset of 16-byte chunk j of row r inside a staging buffer
__device__ __forceinline__ uint32_t swz128(int r, int j) { return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t swz64(int r, int j) { return (uint32_t)(r * 64 + ((j ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// two floats -> packed fp16, round to nearest, saturating at +-65504 (no inf can enter the residual stream)
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Stage a 32x32 bf16 block held as one row per lane and TMA-store it.
__device__ __forceinline__ void stage_store_b16(const CUtensorMap* tm, uint32_t hbuf, int lane, const float (&v)[32], int col, int row0) {
#pragma unroll
  for (int j = 0; j < 4; ++j)
    sts128u(hbuf + swz64(lane, j), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
            pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
  fence_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tm, hbuf, col, row0);
    bulk_commit();
  }
}

// EPI: compile-time epilogue feature mask (specialised instantiations keep the per-chunk instruction stream short);
// EPI < 0 is the generic kernel that tests the descriptor at run time.
// Bisecting aids (JYUTVOICE_B200_DEBUG bit mask, the wait-time counters of tools/gemm_trace.py / mlp_trace.py /
// attention_trace.py) exist only in builds with -DJV_TRACE: as run-time flags inside the epilogue and MMA loops they cost
// 3.5-4.5 % end to end (same-box A/B: 156.2 vs 149.3-150.8 ms per step).
#ifdef JV_TRACE
#define JV_DBG(p) ((p).debug)
#else
#define JV_DBG(p) 0
#endif
constexpr int EPI_LN1 = 1, EPI_RESID = 2, EPI_F32 = 4, EPI_OACT = 8, EPI_LN2 = 16;
// EPI_XB: the residual stream is bf16 (GemmDesc::x_bf16): `resid` in and the main output are 2 KB bf16 chunks, so the
// same 8 KB of staging per warp holds two residual chunks in flight and two output buffers (no store is ever waited for
// right after it was issued).  The fp32-stream kernels move twice the bytes with half the loads in flight.
constexpr int EPI_XB = 32;

// PAIR: the cta_group::2 variant (TcParams::pair).  A separate instantiation, not a run-time switch: a kernel that contains
// cta_group::2 instructions can only be launched with an even cluster size ("cluster misconfiguration" otherwise).
template <int EPI, bool PAIR = false, bool WLN = false>
__global__ void __launch_bounds__((EPI == EPI_OACT || WLN) ? NUM_THREADS_WIDE : NUM_THREADS, 1)
gemm_taps_tc_kernel(const __grid_constant__ TcMaps tm, const GemmDesc g, const TcParams p) {
  const bool F_LN1 = EPI >= 0 ? (EPI & EPI_LN1) != 0 : g.ln1_gamma != nullptr;
  const bool F_RESID = EPI >= 0 ? (EPI & EPI_RESID) != 0 : g.resid != nullptr;
  const bool F_F32 = EPI >= 0 ? (EPI & EPI_F32) != 0 : g.out_f32 != nullptr;
  const bool F_OACT = EPI >= 0 ? (EPI & EPI_OACT) != 0 : g.out_act != nullptr;
  const bool F_LN2 = EPI >= 0 ? (EPI & EPI_LN2) != 0 : g.ln2_gamma != nullptr;
  constexpr bool XB = EPI >= 0 && (EPI & EPI_XB) != 0;
  // WLN: sixteen epilogue warps for the residual + LayerNorm kernels of the 16-bit stream (register-resident epilogue below)
  constexpr int N_EPI_WARPS = (EPI == EPI_OACT || WLN) ? EPI_WARPS_MAX : EPI_WARPS;
  constexpr int N_SUB = N_EPI_WARPS / 4;  // warps sharing one TMEM lane quarter: they split the 32-column chunks of a tile
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024 B alignment
  const uint32_t smem_a = base;
  const uint32_t smem_b = smem_a + p.stages * p.a_stage_bytes;
  const uint32_t smem_epi = smem_b + p.b_region_bytes;  // EPI_WARPS staging regions, 1024-aligned pieces
  const uint32_t bars = smem_epi + N_EPI_WARPS * p.epi_bytes_per_warp;
  const uint32_t full_bar = bars;                        // MAX_STAGES x 8 B
  const uint32_t empty_bar = bars + 8 * MAX_STAGES;      // MAX_STAGES x 8 B
  const uint32_t tfull_bar = bars + 16 * MAX_STAGES;     // MAX_ACC x 8 B
  const uint32_t tempty_bar = tfull_bar + 8 * MAX_ACC;   // MAX_ACC x 8 B
  const uint32_t epi_bar = tempty_bar + 8 * MAX_ACC;     // EPI_WARPS_MAX x 16 B (residual-load barrier per warp)
  const uint32_t tmem_slot = epi_bar + 16 * EPI_WARPS_MAX;   // 4 B
  const uint32_t wres_bar = tmem_slot + 8;               // 8 B: resident weight tile landed
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_blocks_per_tap = g.K_tap / BLOCK_K;
  const int k_iters = g.n_taps * k_blocks_per_tap;
  const int csz = p.cluster;
  const int cta_rank = csz > 1 ? (int)cluster_cta_rank() : 0;
  const uint16_t cmask = (uint16_t)((1u << csz) - 1u);
  const bool wres = p.wres != 0;
  const int unit0 = wres ? blockIdx.x / p.n_tiles_n : blockIdx.x / csz;
  const int unit_step = wres ? gridDim.x / p.n_tiles_n : gridDim.x / csz;
  auto tile_m0 = [&](int unit) { return wres ? unit * BLOCK_M : ((unit / p.n_tiles_n) * csz + cta_rank) * BLOCK_M; };
  auto tile_n0 = [&](int unit) { return wres ? (int)(blockIdx.x % p.n_tiles_n) * p.block_n : (unit % p.n_tiles_n) * p.block_n; };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.a1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tm.w) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(full_bar + 8 * i, 1);
      mbar_init(empty_bar + 8 * i, PAIR ? 1 : csz);
    }
    for (int i = 0; i < p.n_acc; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      // PAIR: one elected, relaxed remote arrive per epilogue warp of either CTA (nothing but "TMEM drained" is signalled)
      mbar_init(tempty_bar + 8 * i, PAIR ? 2 * (p.tile_par ? 4 : N_EPI_WARPS) : (p.tile_par ? 128 : 32 * N_EPI_WARPS));
    }
    for (int i = 0; i < 2 * EPI_WARPS_MAX; ++i) mbar_init(epi_bar + 8 * i, 1);
    mbar_init(wres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  if (p.slab && warp == 3) {  // descriptor tables for the MMA thread: everything that does not depend on the ring stage
    uint64_t* tab = reinterpret_cast<uint64_t*>(smem_raw + (base - raw) + p.tab_off);
    for (int s2 = lane; s2 < g.n_taps; s2 += 32) {
      const uint32_t d = (uint32_t)(g.tap_shift[s2] - p.min_shift);
      // start address += d rows (128 B = 8 x 16 B).  No descriptor base offset: measured on B200, the tensor core applies the
      // 128B swizzle to the absolute smem address, exactly as TMA wrote the slab, so a row-shifted view needs nothing else.
      tab[s2] = (uint64_t)(d * 8u);
    }
    for (int i = lane; i < g.n_taps * k_blocks_per_tap; i += 32) tab[32 + i] = make_smem_desc(smem_b + i * p.b_stage_bytes);
  }
  {  // per-column vectors -> smem (see TcParams)
    const int n_fix = wres ? (int)(blockIdx.x % p.n_tiles_n) * p.block_n : 0;
    float* sm = reinterpret_cast<float*>(smem_raw + (base - raw));
    for (int i = threadIdx.x; i < p.block_n; i += blockDim.x) {
      const bool ok = n_fix + i < g.N;
      if (p.vec_bias) sm[p.vec_bias / 4 + i] = (ok && g.bias) ? g.bias[n_fix + i] : 0.f;
      if (p.vec_ln1) {
        sm[p.vec_ln1 / 4 + i] = ok ? g.ln1_gamma[n_fix + i] : 0.f;
        sm[p.vec_ln1 / 4 + p.block_n + i] = ok ? g.ln1_beta[n_fix + i] : 0.f;
      }
      if (p.vec_ln2) {
        sm[p.vec_ln2 / 4 + i] = ok ? g.ln2_gamma[n_fix + i] : 0.f;
        sm[p.vec_ln2 / 4 + p.block_n + i] = ok ? g.ln2_beta[n_fix + i] : 0.f;
      }
      if (p.vec_act) sm[p.vec_act / 4 + i] = ok ? g.act_vec[n_fix + i] : 1.f;
      if (p.vec_act2) sm[p.vec_act2 / 4 + i] = ok ? g.act2_vec[n_fix + i] : 1.f;
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (csz > 1) cluster_sync_all();  // peers' barriers are initialised before any multicast / remote arrive
  tc_fence_after();
  if (warp == 0 && lane == 0 && wres && unit0 < p.num_units) {
    // The resident weight tile of this CTA's n-tile does not depend on the previous kernel: its 64-128 KB load runs
    // while that kernel drains (a CTA that starts on an SM the previous grid has already left sits in pdl_wait anyway).
    mbar_expect_tx(wres_bar, (uint32_t)k_iters * p.block_n * BLOCK_K * 2);
    for (int it = 0; it < k_iters; ++it)
      tma_load_2d(&tm.w, wres_bar, smem_b + it * p.b_stage_bytes, it * BLOCK_K, tile_n0(unit0));
  }
  pdl_wait();  // from here on the previous kernel's results (activations, residual stream) may be read / overwritten
  const uint32_t tmem_base = *tmem_slot_ptr;

  // Register re-partitioning (general kernels, 384 threads x 168 registers): the producer / MMA / allocator warpgroup
  // drops to 80 registers, the two epilogue warpgroups grow to 208 (128 x 80 + 256 x 208 <= 384 x 168; a larger request
  // would block in setmaxnreg.inc for ever).  setmaxnreg sits inside the role branches so that ptxas allocates per role.
  if (warp < 4) {
  if (N_EPI_WARPS == EPI_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
  if (WLN) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");  // 128 x (96 - 64) freed = 512 x (104 - 96) requested below
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = A_STAGE_BYTES + p.block_n * BLOCK_K * 2;
      const bool slab = p.slab != 0;
      const int b_rows = p.block_n / csz;  // weight rows this CTA fetches (and multicasts)
      for (int unit = unit0; unit < p.num_units; unit += unit_step) {
        const int m0 = tile_m0(unit);
        const int n0 = tile_n0(unit);
        if (slab) {  // one slab per 64-channel block, shared by all taps
          for (int kb = 0; kb < k_blocks_per_tap; ++kb) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1, 1);
            mbar_expect_tx(full_bar + 8 * stage, (uint32_t)p.slab_rows * BLOCK_K * 2);
            tma_load_2d(&tm.a0, full_bar + 8 * stage, smem_a + stage * p.a_stage_bytes, kb * BLOCK_K, m0 + p.min_shift);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        if constexpr (PAIR) {  // both CTAs load (own activation rows, own half of the weight tile); all bytes land on the leader's barrier
          const uint32_t full_leader = mapa_u32(full_bar, 0);
          for (int s = 0; s < g.n_taps; ++s) {
            const CUtensorMap* tmA = g.tap_src[s] ? &tm.a1 : &tm.a0;
            const int arow = m0 + g.tap_shift[s];
            for (int kb = 0; kb < k_blocks_per_tap; ++kb) {
              mbar_wait(empty_bar + 8 * stage, phase ^ 1, 1);
              if (cta_rank == 0) mbar_expect_tx(full_bar + 8 * stage, 2u * (A_STAGE_BYTES + (uint32_t)b_rows * BLOCK_K * 2));
              tma_load_2d_cg2(tmA, full_leader + 8 * stage, smem_a + stage * p.a_stage_bytes, kb * BLOCK_K, arow);
              tma_load_2d_cg2(&tm.w, full_leader + 8 * stage, smem_b + stage * p.b_stage_bytes, s * g.K_tap + kb * BLOCK_K,
                              n0 + cta_rank * b_rows);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        for (int s = 0; s < g.n_taps; ++s) {
          const CUtensorMap* tmA = g.tap_src[s] ? &tm.a1 : &tm.a0;
          const int arow = m0 + g.tap_shift[s];
          for (int kb = 0; kb < k_blocks_per_tap; ++kb) {
            mbar_wait(empty_bar + 8 * stage, phase ^ 1, 1);
            mbar_expect_tx(full_bar + 8 * stage, wres ? (uint32_t)A_STAGE_BYTES : tx_bytes);
            tma_load_2d(tmA, full_bar + 8 * stage, smem_a + stage * p.a_stage_bytes, kb * BLOCK_K, arow);
            if (wres) {
            } else if (csz == 1)
              tma_load_2d(&tm.w, full_bar + 8 * stage, smem_b + stage * p.b_stage_bytes, s * g.K_tap + kb * BLOCK_K, n0);
            else
              tma_load_2d_mcast(&tm.w, full_bar + 8 * stage, smem_b + stage * p.b_stage_bytes + cta_rank * b_rows * BLOCK_K * 2,
                                s * g.K_tap + kb * BLOCK_K, n0 + cta_rank * b_rows, cmask);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && !(PAIR && cta_rank != 0)) {  // CTA pair: the leader issues for both
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
      // a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)  (M = 256 across the CTA pair)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)((PAIR ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc_stage = 0;
      uint32_t acc_phase = 0;
      // wait-time counters: only in builds with -DJV_TRACE (JYUTVOICE_B200_NVCC_FLAGS=-DJV_TRACE python -m jyutvoice_b200.build --force);
      // a run-time flag around every wait of this single-thread loop is not free
#ifdef JV_TRACE
      long long* tr = p.trace ? p.trace + 8L * blockIdx.x : nullptr;
#else
      constexpr long long* tr = nullptr;
#endif
      long long w_tempty = 0, w_full = 0, n_units = 0;
      const long long t_begin = tr ? clock64() : 0;
#define TC_TWAIT(counter, call)               \
  do {                                        \
    const long long c0_ = tr ? clock64() : 0; \
    call;                                     \
    if (tr) counter += clock64() - c0_;       \
  } while (0)
      if (wres && unit0 < p.num_units) mbar_wait(wres_bar, 0, 6);
      for (int unit = unit0; unit < p.num_units; unit += unit_step) {
        ++n_units;
        TC_TWAIT(w_tempty, mbar_wait(tempty_bar + 8 * acc_stage, acc_phase ^ 1, 2));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc_stage * p.acc_cols;
        if (p.slab) {
          const uint64_t* tab = reinterpret_cast<const uint64_t*>(smem_raw + (base - raw) + p.tab_off);
          for (int kb = 0; kb < k_blocks_per_tap; ++kb) {
            TC_TWAIT(w_full, mbar_wait(full_bar + 8 * stage, phase, 3));
            tc_fence_after();
            const uint64_t a_stage_desc = make_smem_desc(smem_a + stage * p.a_stage_bytes);  // 1024-aligned: base offset 0
            for (int s = 0; s < g.n_taps; ++s) {
              const uint64_t adesc = a_stage_desc + tab[s];
              const uint64_t bdesc = tab[32 + s * k_blocks_per_tap + kb];
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || s > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(empty_bar + 8 * stage);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit(tfull_bar + 8 * acc_stage);
          if (++acc_stage == p.n_acc) { acc_stage = 0; acc_phase ^= 1; }
          continue;
        }
        if constexpr (PAIR) {
          for (int it = 0; it < k_iters; ++it) {
            TC_TWAIT(w_full, mbar_wait(full_bar + 8 * stage, phase, 3));  // both CTAs' tiles of this stage have landed
            tc_fence_after();
            const uint64_t adesc = make_smem_desc(smem_a + stage * p.a_stage_bytes);
            const uint64_t bdesc = make_smem_desc(smem_b + stage * p.b_stage_bytes);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16_cg2(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma_commit_cg2(empty_bar + 8 * stage, 3);  // frees the slot in both CTAs
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          umma_commit_cg2(tfull_bar + 8 * acc_stage, 3);  // both CTAs' epilogues own a half of the accumulator
          if (++acc_stage == p.n_acc) { acc_stage = 0; acc_phase ^= 1; }
          continue;
        }
        for (int it = 0; it < k_iters; ++it) {
          TC_TWAIT(w_full, mbar_wait(full_bar + 8 * stage, phase, 3));
          tc_fence_after();
          const uint64_t adesc = make_smem_desc(smem_a + stage * p.a_stage_bytes);
          const uint64_t bdesc = make_smem_desc(smem_b + (wres ? it : stage) * p.b_stage_bytes);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 B (= 16 bf16) inside the 128 B swizzle row: +2 in the >>4 address field
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in every CTA of the cluster: the peers multicast into it) when these MMAs retire
          if (csz == 1) umma_commit(empty_bar + 8 * stage);
          else umma_commit_mcast(empty_bar + 8 * stage, cmask);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc_stage);  // accumulator complete -> epilogue
        if (++acc_stage == p.n_acc) { acc_stage = 0; acc_phase ^= 1; }
      }
      if (tr) { tr[0] = clock64() - t_begin; tr[1] = w_tempty; tr[2] = w_full; tr[3] = n_units; }
#undef TC_TWAIT
    }
  }
  } else {
    if (N_EPI_WARPS == EPI_WARPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    if (WLN) asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ===================== epilogue: every epilogue warp works on the current tile =====================
    // Warp e owns TMEM lane quarter q (rows q*32 .. q*32+31 of the tile) and the chunks c == sub_id (mod N_SUB), so a
    // row is shared by N_SUB threads; LayerNorm statistics are combined through smem + a named barrier per quarter.
    // A single warp runs this long dependent instruction stream at ~0.2 IPC, so the epilogue latency of a tile scales
    // with 1 / N_SUB; the accumulator stage alternates per tile, overlapping the next tile's MMAs.
    const int e = warp - 4;
    const int q = e & 3;        // TMEM lane quarter (warp index % 4)
    const int sub_id = e >> 2;  // which share of the column chunks
    float* xch = reinterpret_cast<float*>(smem_raw + (bars + 512 - raw)) + q * (4 * 32 * 2);  // [N_SUB <= 4][32][2] for this quarter
    auto row_sum2 = [&](float& a, float& b) {  // (a, b) summed over the N_SUB threads that share a row
      if (N_SUB == 1) return;
      xch[(sub_id * 32 + lane) * 2] = a;
      xch[(sub_id * 32 + lane) * 2 + 1] = b;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * N_SUB) : "memory");
      float ta = 0.f, tb = 0.f;
#pragma unroll
      for (int i = 0; i < N_SUB; ++i) {
        ta += xch[(i * 32 + lane) * 2];
        tb += xch[(i * 32 + lane) * 2 + 1];
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * N_SUB) : "memory");  // all have read: the slot may be reused
      a = ta;
      b = tb;
    };
    // Staging per warp: R = residual in (TMA load), OF = fp32 out, OB = bf16 out (TMA stores).  R is refilled
    // with chunk c + 1 as soon as every lane has read chunk c, independent of the stores; the output buffers
    // are waited for (cp.async.bulk.wait_group.read) only right before they are rewritten, i.e. after the math.
    // Kernels without fp32 output alternate OF / OB as two bf16 buffers and never wait for the latest store.
    const uint32_t sR = smem_epi + e * p.epi_bytes_per_warp + p.off_R;
    const uint32_t sOF = smem_epi + e * p.epi_bytes_per_warp + p.off_OF;
    const uint32_t sOB = smem_epi + e * p.epi_bytes_per_warp + p.off_OB;  // kernels without a bf16 copy: == sR (LN2 sweep alternates R / OF)
    const bool one_buf = p.off_OB == p.off_OF;                            // a single bf16 staging buffer
    const uint32_t ebar = epi_bar + 16 * e;
    uint32_t ephase = 0;
    uint32_t acc_phase = 0;
    int grp = 0;  // accumulator stage of the current tile
    const bool tile_par = p.tile_par != 0;
    const int c_first = tile_par ? 0 : sub_id, c_step = tile_par ? 1 : N_SUB;
    int t_local = 0;   // tiles seen by this CTA
    uint32_t n_out = 0;  // output chunks staged by this warp (alternates the bf16 staging buffers)
    const int n_chunks = p.block_n >> 5;
    const float* smf = reinterpret_cast<const float*>(smem_raw + (base - raw));
#ifdef JV_TRACE
    long long* etr = (p.trace && e == 0 && lane == 0) ? p.trace + 8L * blockIdx.x + 4 : nullptr;  // epilogue warp 0's view
#else
    constexpr long long* etr = nullptr;
#endif
    long long ew_tfull = 0, ew_resid = 0;
    const long long et_begin = etr ? clock64() : 0;
#define TC_EWAIT(counter, call)                \
  do {                                         \
    const long long c0_ = etr ? clock64() : 0; \
    call;                                      \
    if (etr) counter += clock64() - c0_;       \
  } while (0)
    for (int unit = unit0; unit < p.num_units; unit += unit_step, ++t_local) {
      if (tile_par && (t_local % N_SUB) != sub_id) {  // another share's tile: just keep the stage / phase counters in step
        if (++grp == p.n_acc) { grp = 0; acc_phase ^= 1; }
        continue;
      }
      const int m0 = tile_m0(unit);
      const int n0 = tile_n0(unit);
      const int row0 = m0 + q * 32;  // first row (view coordinates) of this warp's 32-row slab
      // per-column vectors, indexed by the absolute column n: smem copy (n-tile fixed per CTA) or global
      const float* v_bias = g.bias ? (p.vec_bias ? smf + p.vec_bias / 4 - n0 : g.bias) : nullptr;
      const float* v_g1 = p.vec_ln1 ? smf + p.vec_ln1 / 4 - n0 : g.ln1_gamma;
      const float* v_b1 = p.vec_ln1 ? smf + p.vec_ln1 / 4 + p.block_n - n0 : g.ln1_beta;
      const float* v_g2 = p.vec_ln2 ? smf + p.vec_ln2 / 4 - n0 : g.ln2_gamma;
      const float* v_b2 = p.vec_ln2 ? smf + p.vec_ln2 / 4 + p.block_n - n0 : g.ln2_beta;
      const float* v_act = g.act_vec ? (p.vec_act ? smf + p.vec_act / 4 - n0 : g.act_vec) : nullptr;
      const float* v_act2 = g.act2_vec ? (p.vec_act2 ? smf + p.vec_act2 / 4 - n0 : g.act2_vec) : nullptr;
      constexpr bool EPI16_OK = WLN && XB && EPI >= 0 && (EPI & EPI_LN2) != 0 && (EPI & EPI_RESID) != 0 && N_SUB == 4;
      constexpr int NCHK = 8 / N_SUB;  // 32-column chunks per thread: columns (sub_id + N_SUB * i) * 32
      if constexpr (WLN && EPI == (EPI_LN1 | EPI_OACT)) {
        // ================= sixteen-warp epilogue of conv + LayerNorm + activation -> bf16 (CausalBlock1D) =================
        // K = 3 x 256 gives 9 k clk of MMAs per tile and the eight-warp epilogue took 15.7 k (tools/gemm_trace.py): same remedy
        // as below.  Statistics sweep, then the main pass; the two bf16 staging buffers alternate.
        const int m = row0 + lane;
        const long orow = (long)m * g.o_stride + g.o_off;
        const bool in_range = m < g.M && orow < g.o_rows;
        const int fr = (in_range && g.frame_row) ? g.frame_row[orow] : (in_range ? 0 : -1);
        const bool row_valid = fr >= 0;
        const float* add_row = (g.add_row && row_valid) ? g.add_row + (long)g.row_tidx[fr] * g.add_row_stride : nullptr;
        TC_EWAIT(ew_tfull, mbar_wait(tfull_bar + 8 * grp, acc_phase, 4));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * p.acc_cols;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int i = 0; i < NCHK; ++i) {
          const int c = sub_id + N_SUB * i;
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          if (v_bias) add_vec32(v, v_bias + n0 + c * 32, 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s1 += v[j];
            s2 = fmaf(v[j], v[j], s2);
          }
        }
        row_sum2(s1, s2);
        const float mean1 = s1 * (1.0f / 256.0f);
        const float rstd1 = rsqrtf(fmaxf(s2 * (1.0f / 256.0f) - mean1 * mean1, 0.f) + 1e-5f);
#pragma unroll 1
        for (int i = 0; i < NCHK; ++i) {
          const int c = sub_id + N_SUB * i;
          const int n = n0 + c * 32;
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          if (v_bias) add_vec32(v, v_bias + n, 32);
          ln_affine32(v, mean1, rstd1, v_g1 + n, v_b1 + n);
          if (g.act != ACT_NONE) act32(v, g.act, g.act_param, v_act ? v_act + n : nullptr, 32);
          if (add_row) add_vec32(v, add_row + n, 32);
          if (!row_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (g.act2 != ACT_NONE) act32(v, g.act2, g.act2_param, v_act2 ? v_act2 + n : nullptr, 32);
          if (lane == 0) bulk_wait_read1();  // the store that last used this buffer (two groups back) has read it
          __syncwarp();
          stage_store_b16(&tm.out_act, (n_out & 1) ? sOB : sOF, lane, v, n, row0);
          ++n_out;
        }
        tc_fence_before();
        if constexpr (PAIR) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(tempty_bar + 8 * grp, 0));
        } else {
          mbar_arrive(tempty_bar + 8 * grp);
        }
        if (++grp == p.n_acc) { grp = 0; acc_phase ^= 1; }
        continue;
      }
      if constexpr (EPI16_OK) {  // (WLN instantiations are only launched for these kernels: TcParams::epi16)
        // ================= sixteen-warp epilogue (out-proj / FF2 / conv2 + residual + LayerNorm, 16-bit stream) =================
        // tools/gemm_trace.py: with eight epilogue warps the epilogue of a tile takes 16 k clk (27 k with LN1 + Mish) against
        // ~6-9 k for the tile's loads + MMAs, and ncu shows it is the warps' own instruction streams (~800-2000 instructions per
        // thread and tile at 0.2 IPC per scheduler), not hand-offs.  Sixteen warps halve the stream per warp and double the
        // warps per scheduler.  Same algorithm as the chunked path below (values stashed in TMEM between the two LayerNorm
        // passes: keeping a thread's 64 values in registers spills at the 96-register budget of a 640-thread CTA, and with
        // 227 KB of shared memory carved out a spill is an L2 round trip), but both residual chunks are in flight since before
        // the accumulator wait and the residual buffers double as the staging of both outputs: 4 KB of staging per warp.
        if (lane == 0) {
          bulk_wait_read0();  // the previous tile's stores were staged in R
          mbar_expect_tx(ebar, NCHK * EPI_B16_BYTES);
#pragma unroll
          for (int i = 0; i < NCHK; ++i) tma_load_2d(&tm.resid, ebar, sR + i * EPI_B16_BYTES, n0 + (sub_id + N_SUB * i) * 32, row0);
        }
        const int m = row0 + lane;
        const long orow = (long)m * g.o_stride + g.o_off;
        const bool in_range = m < g.M && orow < g.o_rows;
        const int fr = (in_range && g.frame_row) ? g.frame_row[orow] : (in_range ? 0 : -1);
        const bool row_valid = fr >= 0;
        const float* add_row = (F_LN1 && g.add_row && row_valid) ? g.add_row + (long)g.row_tidx[fr] * g.add_row_stride : nullptr;
        TC_EWAIT(ew_tfull, mbar_wait(tfull_bar + 8 * grp, acc_phase, 4));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * p.acc_cols;
        float mean1 = 0.f, rstd1 = 1.f;
        if (F_LN1) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
          for (int i = 0; i < NCHK; ++i) {
            const int c = sub_id + N_SUB * i;
            uint32_t acc[32];
            tmem_ld32(taddr + c * 32, acc);
            float v[32];
            acc_to_f32(acc, v);
            if (v_bias) add_vec32(v, v_bias + n0 + c * 32, 32);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              s1 += v[j];
              s2 = fmaf(v[j], v[j], s2);
            }
          }
          row_sum2(s1, s2);
          mean1 = s1 * (1.0f / 256.0f);
          rstd1 = rsqrtf(fmaxf(s2 * (1.0f / 256.0f) - mean1 * mean1, 0.f) + 1e-5f);
        }
        float sum2 = 0.f, sq2 = 0.f, amax = 0.f;
#pragma unroll 1
        for (int i = 0; i < NCHK; ++i) {
          const int c = sub_id + N_SUB * i;
          const int n = n0 + c * 32;
          const uint32_t rbuf = sR + i * EPI_B16_BYTES;
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          if (v_bias) add_vec32(v, v_bias + n, 32);
          if (F_LN1) ln_affine32(v, mean1, rstd1, v_g1 + n, v_b1 + n);
          if (g.act != ACT_NONE) act32(v, g.act, g.act_param, v_act ? v_act + n : nullptr, 32);
          if (add_row) add_vec32(v, add_row + n, 32);
          if (!row_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (i == 0) {
            TC_EWAIT(ew_resid, mbar_wait(ebar, ephase & 1, 5));
            ephase ^= 1;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 r = lds128(rbuf + swz64(lane, j));
            const uint32_t w[4] = {__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w)};
            if (g.x_in_half) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                v[8 * j + 2 * k] += f.x;
                v[8 * j + 2 * k + 1] += f.y;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
                v[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] = __float_as_uint(v[j]);
            sum2 += v[j];
            sq2 = fmaf(v[j], v[j], sq2);
            amax = fmaxf(amax, fabsf(v[j]));
          }
          tmem_st32(taddr + c * 32, acc);  // kept for the LayerNorm pass
          // the stream value, rounded to fp16 (saturating) or bf16, over the residual chunk this lane has just read
          if (g.x_out_half) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128u(rbuf + swz64(lane, j), pack_f16_sat(v[8 * j], v[8 * j + 1]), pack_f16_sat(v[8 * j + 2], v[8 * j + 3]),
                      pack_f16_sat(v[8 * j + 4], v[8 * j + 5]), pack_f16_sat(v[8 * j + 6], v[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128u(rbuf + swz64(lane, j), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                      pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm.out_f32, rbuf, n, row0);
            bulk_commit();
          }
        }
        row_sum2(sum2, sq2);
        if (g.x_out_half && g.sat_flag && amax >= 65504.f) atomicAdd(g.sat_flag, 1);
        const float mean2 = sum2 * (1.0f / 256.0f);
        const float rstd2 = rsqrtf(fmaxf(sq2 * (1.0f / 256.0f) - mean2 * mean2, 0.f) + 1e-5f);
#pragma unroll 1
        for (int i = 0; i < NCHK; ++i) {
          const int c = sub_id + N_SUB * i;
          const int n = n0 + c * 32;
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          ln_affine32(v, mean2, rstd2, v_g2 + n, v_b2 + n);
          if (!row_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (lane == 0) bulk_wait_read1();  // the stream store of this chunk (two groups back at least) has read R[i]
          __syncwarp();
          stage_store_b16(&tm.out_ln, sR + i * EPI_B16_BYTES, lane, v, n, row0);
        }
        tc_fence_before();
        if constexpr (PAIR) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(tempty_bar + 8 * grp, 0));
        } else {
          mbar_arrive(tempty_bar + 8 * grp);
        }
        if (++grp == p.n_acc) { grp = 0; acc_phase ^= 1; }
        continue;
      }
      if (XB) {  // residual of the first two chunks: in flight while the accumulator is still being computed
        if (lane == 0) {
          mbar_expect_tx(ebar, EPI_B16_BYTES);
          tma_load_2d(&tm.resid, ebar, sR, n0 + c_first * 32, row0);
          if ((c_first + c_step) * 32 < g.N - n0) {
            mbar_expect_tx(ebar + 8, EPI_B16_BYTES);
            tma_load_2d(&tm.resid, ebar + 8, sR + EPI_B16_BYTES, n0 + (c_first + c_step) * 32, row0);
          }
        }
      } else if (F_RESID && lane == 0 && !(JV_DBG(p) & 512)) {  // residual of the first chunk: in flight while the accumulator is still being computed
        if (F_LN2) bulk_wait_read0();  // the previous tile's post-LayerNorm stores may still be reading R
        mbar_expect_tx(ebar, EPI_F32_BYTES);
        tma_load_2d(&tm.resid, ebar, sR, n0 + c_first * 32, row0);
      }
      // row bookkeeping (two dependent global loads, ~1 us each from L2) is issued BEFORE waiting for the accumulator:
      // measured 15 % of the QKV kernel's warp samples sat on frame_row[] when it was loaded after the wait
      const int m = row0 + lane;
      const long orow = (long)m * g.o_stride + g.o_off;
      const bool in_range = m < g.M && orow < g.o_rows;
      const int fr = (in_range && g.frame_row) ? g.frame_row[orow] : (in_range ? 0 : -1);
      const bool row_valid = fr >= 0;
      const float* add_row = (F_LN1 && g.add_row && row_valid) ? g.add_row + (long)g.row_tidx[fr] * g.add_row_stride : nullptr;
      TC_EWAIT(ew_tfull, mbar_wait(tfull_bar + 8 * grp, acc_phase, 4));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * p.acc_cols;

      // ---- optional pre-LayerNorm statistics of (acc + bias) over the whole row (block_n == N == 256)
      float mean1 = 0.f, rstd1 = 1.f;
      if (F_LN1 && !(JV_DBG(p) & 32)) {  // one sweep: sum and sum of squares in fp32 (256 O(1) values; bf16-mode tolerance)
        float s1 = 0.f, s2 = 0.f;
        for (int c = c_first; c < n_chunks; c += c_step) {
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          if (v_bias) add_vec32(v, v_bias + n0 + c * 32, 32);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            s1 += v[j];
            s2 = fmaf(v[j], v[j], s2);
          }
        }
        if (!tile_par) row_sum2(s1, s2);
        mean1 = s1 * (1.0f / (float)g.N);
        rstd1 = rsqrtf(fmaxf(s2 * (1.0f / (float)g.N) - mean1 * mean1, 0.f) + 1e-5f);
      }

      // ---- main pass
      float sum2 = 0.f, sq2 = 0.f;
      float amax = 0.f;  // largest |value| this thread stores into the 16-bit stream (saturation counter)
      const int n_chunks_valid = (g.N - n0 + 31) / 32 < n_chunks ? (g.N - n0 + 31) / 32 : n_chunks;
      int i_chunk = 0;
      for (int c = c_first; c < ((JV_DBG(p) & 2) ? 0 : n_chunks_valid); c += c_step, ++i_chunk) {
        const int n = n0 + c * 32;
        const int n_valid = g.N - n < 32 ? g.N - n : 32;
        uint32_t acc[32];
        tmem_ld32(taddr + c * 32, acc);
        float v[32];
        acc_to_f32(acc, v);
        if (JV_DBG(p) & 4) {  // experiment: TMEM reads only
          if (v[0] == 123.456f) printf("x");
          continue;
        }
        if (v_bias && !(JV_DBG(p) & 8)) add_vec32(v, v_bias + n, n_valid);
        if (F_LN1 && !(JV_DBG(p) & 128)) ln_affine32(v, mean1, rstd1, v_g1 + n, v_b1 + n);
        if (g.act != ACT_NONE && !(JV_DBG(p) & 64)) act32(v, g.act, g.act_param, v_act ? v_act + n : nullptr, n_valid);
        if (add_row) add_vec32(v, add_row + n, n_valid);
        if (!row_valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        if (XB) {
          const int ib = i_chunk & 1;
          TC_EWAIT(ew_resid, mbar_wait(ebar + 8 * ib, (ephase >> ib) & 1, 5));
          ephase ^= 1u << ib;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 r = lds128(sR + ib * EPI_B16_BYTES + swz64(lane, j));
            const uint32_t w[4] = {__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z), __float_as_uint(r.w)};
            if (g.x_in_half) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
                v[8 * j + 2 * k] += f.x;
                v[8 * j + 2 * k + 1] += f.y;
              }
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // bf16 -> fp32 is a 16-bit shift
                v[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
                v[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
              }
            }
          }
          __syncwarp();  // every lane has read this R buffer: refill it with the residual two chunks ahead
          if (lane == 0 && c + 2 * c_step < n_chunks_valid) {
            mbar_expect_tx(ebar + 8 * ib, EPI_B16_BYTES);
            tma_load_2d(&tm.resid, ebar + 8 * ib, sR + ib * EPI_B16_BYTES, n + 64 * c_step, row0);
          }
        } else if (F_RESID && !(JV_DBG(p) & 512)) {
          TC_EWAIT(ew_resid, mbar_wait(ebar, ephase, 5));
          ephase ^= 1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = lds128(sR + swz128(lane, j));
            v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
          }
          __syncwarp();  // every lane has read R: refill it with the next chunk's residual
          if (lane == 0 && c + c_step < n_chunks_valid) {
            mbar_expect_tx(ebar, EPI_F32_BYTES);
            tma_load_2d(&tm.resid, ebar, sR, n + 32 * c_step, row0);
          }
        }
        if (F_LN2) {  // keep the final value in TMEM for the post-LayerNorm sweep
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            acc[j] = __float_as_uint(v[j]);
            sum2 += v[j];
            sq2 = fmaf(v[j], v[j], sq2);
            if (XB) amax = fmaxf(amax, fabsf(v[j]));
          }
          tmem_st32(taddr + c * 32, acc);
        }
        // output staging: wait (late) for the stores that last used the buffers
        const uint32_t hbuf = (F_F32 && !XB) ? sOB : ((n_out & 1) ? sOB : sOF);
        ++n_out;
        if (lane == 0) {
          if ((F_F32 && !XB) || one_buf) bulk_wait_read0();
          else bulk_wait_read1();
        }
        __syncwarp();
        if (XB) {  // the stream value itself, rounded to fp16 (stream) or bf16 (the copy that feeds the next conv)
          if (g.x_out_half) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128u(hbuf + swz64(lane, j), pack_f16_sat(v[8 * j], v[8 * j + 1]), pack_f16_sat(v[8 * j + 2], v[8 * j + 3]),
                      pack_f16_sat(v[8 * j + 4], v[8 * j + 5]), pack_f16_sat(v[8 * j + 6], v[8 * j + 7]));
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128u(hbuf + swz64(lane, j), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                      pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
          }
        } else if (F_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) sts128(sOF + swz128(lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (F_OACT) {
          if (g.act2 != ACT_NONE) act32(v, g.act2, g.act2_param, v_act2 ? v_act2 + n : nullptr, n_valid);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128u(hbuf + swz64(lane, j), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                    pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
        }
        if (!(JV_DBG(p) & 16)) fence_async_smem();
        __syncwarp();
        if (lane == 0 && !(JV_DBG(p) & 1)) {
          if (XB) tma_store_2d(&tm.out_f32, hbuf, n, row0);
          else if (F_F32) tma_store_2d(&tm.out_f32, sOF, n, row0);
          if (F_OACT) tma_store_2d(&tm.out_act, hbuf, n, row0);
          bulk_commit();
        }
      }

      // ---- optional post-LayerNorm of the final value (norm1 / norm3 of the next transformer block)
      if (F_LN2) {
        if (!tile_par) row_sum2(sum2, sq2);
        // fp16 stream: count the (row, column-share) pairs in which a stored value reached the format's largest finite number
        if (XB && g.x_out_half && g.sat_flag && amax >= 65504.f) atomicAdd(g.sat_flag, 1);
        const float mean2 = sum2 * (1.0f / (float)g.N);
        const float rstd2 = rsqrtf(fmaxf(sq2 * (1.0f / (float)g.N) - mean2 * mean2, 0.f) + 1e-5f);
        int n_ln = 0;
        for (int c = c_first; c < n_chunks; c += c_step, ++n_ln) {
          const int n = n0 + c * 32;
          uint32_t acc[32];
          tmem_ld32(taddr + c * 32, acc);
          float v[32];
          acc_to_f32(acc, v);
          ln_affine32(v, mean2, rstd2, v_g2 + n, v_b2 + n);
          if (!row_valid) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          if (lane == 0) {  // alternate OF / OB as bf16 buffers; the first two chunks wait for the main pass's stores
            if (n_ln < 2 && !XB) bulk_wait_read0();
            else bulk_wait_read1();
          }
          __syncwarp();
          if (XB) {  // the main pass's alternation simply continues
            stage_store_b16(&tm.out_ln, (n_out & 1) ? sOB : sOF, lane, v, n, row0);
            ++n_out;
          } else {
            stage_store_b16(&tm.out_ln, (n_ln & 1) ? sOB : sOF, lane, v, n, row0);
          }
        }
      }
      tc_fence_before();
      if constexpr (PAIR) {  // the leader's MMA thread waits for both halves
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_u32(tempty_bar + 8 * grp, 0));
      } else {
        mbar_arrive(tempty_bar + 8 * grp);
      }
      if (++grp == p.n_acc) { grp = 0; acc_phase ^= 1; }
    }
    if (etr) { etr[0] = clock64() - et_begin; etr[1] = ew_tfull; etr[2] = ew_resid; etr[3] = t_local; }
#undef TC_EWAIT
    if (lane == 0) bulk_wait0();  // smem must outlive the last TMA store's reads
  }

  tc_fence_before();
  __syncthreads();
  if (csz > 1) cluster_sync_all();  // no CTA leaves while a peer can still multicast into it or arrive on its barriers
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc

// ---------------------------------------------------------------- host side
// development aid (tools/gemm_trace.py, jv_debug_gemm_trace): launches whose epilogue mask equals `epi` fill `buf`
struct GemmTrace {
  long long* buf = nullptr;
  int epi = -1;
};
inline GemmTrace& gemm_trace() {
  static GemmTrace t;
  return t;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled();

struct TmapKey {
  const void* ptr;
  long inner, rows, pitch_bytes;
  int box_inner, box_rows, kind;  // kind: 0 = bf16 SW128 (operand), 1 = f32 SW128, 2 = bf16 SW64
  bool operator<(const TmapKey& o) const {
    if (ptr != o.ptr) return ptr < o.ptr;
    if (inner != o.inner) return inner < o.inner;
    if (rows != o.rows) return rows < o.rows;
    if (pitch_bytes != o.pitch_bytes) return pitch_bytes < o.pitch_bytes;
    if (box_inner != o.box_inner) return box_inner < o.box_inner;
    if (box_rows != o.box_rows) return box_rows < o.box_rows;
    return kind < o.kind;
  }
};

// Cache of encoded tensor maps (encoding costs ~1 us; the same buffers recur every layer / step).
struct TmapCache {
  std::map<TmapKey, CUtensorMap> maps;
  const CUtensorMap& get(const void* ptr, long inner_elems, long rows, long pitch_bytes, int box_inner, int box_rows, int kind);
};

struct ProfileState {
  bool on = false;
  std::vector<cudaEvent_t> ev;  // pairs
  double flops = 0.0;
};
ProfileState& profile_state();

bool use_pdl();  // JYUTVOICE_B200_PDL=0 disables programmatic dependent launch for every tcgen05 kernel

// True when the tcgen05 kernel can run this problem (else the caller uses the FFMA engine).
bool gemm_tc_supported(const GemmDesc& g);
void launch_gemm_tc(const GemmDesc& g, TmapCache& cache, int num_sms, cudaStream_t st);

}  // namespace jv
