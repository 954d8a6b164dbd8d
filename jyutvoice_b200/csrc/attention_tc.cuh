// tcgen05 multi-head attention for the estimator's transformer blocks (bf16 mode).
// 8 heads x 64, non-causal, keys masked by the row length (decoder.py:955-959: -1e10 bias on padded keys).
// One CTA = 128 queries of one (row, head); keys are consumed in tiles of 64 with an online softmax:
//   S = Q K^T      tcgen05.mma M=128 N=64  K=64      (Q, K tiles by TMA, 128B swizzle, K-major)
//   P = softmax    one thread per query row: the 64 scores of a key tile in registers (one tcgen05.ld round trip), exp2,
//                  running max / sum with lazy rescaling, P -> smem (bf16, K-major)
//   O += P V       tcgen05.mma M=128 N=64  K=64      (V tile by TMA is the MN-major B operand as it lies in memory)
// O stays in TMEM across key tiles and is rescaled in place (tcgen05.ld / st) when the running max moves.
// Four CTAs fit per SM (48 KB smem, 128 TMEM columns each): while one CTA's softmax threads work, the others' MMAs,
// TMA loads and barrier round trips are in flight.
#pragma once
#include "gemm_tc.cuh"

namespace jv {
namespace attn {

constexpr int TQ = 128, TK = 64, HD = 64;
constexpr int Q_BYTES = TQ * HD * 2;   // 16 KB: 128 x 64 bf16 (also the P tile: 128 queries x 64 keys)
constexpr int KV_BYTES = TK * HD * 2;  // 8 KB: 64 x 64 bf16
constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + 2 * KV_BYTES + 128;  // Q, P, K, V: 48 KB -> four CTAs per SM
constexpr int THREADS = 160;
constexpr int TMEM_COLS = 128;  // S: columns [0,64), O: columns [64,128)

// MN-major SWIZZLE_128B descriptor for the V tile [keys][64 d]: rows of 128 B (64 d), 8-key groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 16;  // LBO: next 64-wide MN block (unused: N = 64 is one block)
  d |= (uint64_t)(1024 >> 4) << 32;  // SBO: next group of 8 K rows (keys)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / ALU pipes for x <= 0 (softmax arguments): round-to-nearest split x = n + f, |f| <= 0.5, 2^f by a
// degree-3 polynomial (max relative error 7.5e-5, far below bf16's 3.9e-3 rounding of P), exponent added with one LEA.
// Every POLY_EVERY-th exponential of a key tile goes here instead of the MUFU, whose 16 results per clock per SM bound
// the exponential phase of a tile (measured: 15 cycles per ex2 per warp with four softmax warps per scheduler).
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;                       // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const int n = __float_as_int(t) - 0x4B400000;          // round(x)
  const float f = x - (t - 12582912.0f);                 // [-0.5, 0.5]
  float p = fmaf(f, 0.05517149344f, 0.2426111102f);      // minimax fit of 2^f on [-0.5, 0.5]: relative error <= 7.5e-5
  p = fmaf(p, f, 0.6932610273f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (n << 23));
}
#ifndef JV_ATTN_POLY_EVERY
#define JV_ATTN_POLY_EVERY 0
#endif

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 4)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV, bf16* __restrict__ out, int ldo, const int* __restrict__ row_off,
                    const int* __restrict__ row_len, float scale_log2e, int chunk) {
  using namespace tc;
  const int r = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int len = row_len[r];
  if (q0 >= len) return;
  const int off = row_off[r];
  // streaming=True (decoder.py:950-953): query t sees keys < min(len, (t / chunk + 1) * chunk); chunk = 0: all keys
  const int kend = chunk > 0 ? min(len, ((q0 + TQ - 1) / chunk + 1) * chunk) : len;
  const int nt = (kend + TK - 1) / TK;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sQ = base, sP = base + Q_BYTES, sK = base + 2 * Q_BYTES, sV = sK + KV_BYTES;
  const uint32_t bars = sV + KV_BYTES;
  const uint32_t bar_q = bars, bar_k = bars + 8, bar_v = bars + 16, bar_s = bars + 24, bar_p = bars + 32, bar_o = bars + 40;
  const uint32_t tmem_slot = bars + 48;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 128);
      mbar_init(bar_o, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQKV) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // QKV (previous kernel's output) is read only after this point
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tS = tmem_base, tO = tmem_base + 64;

  if (warp == 0) {
    if (lane == 0) {
      // instruction descriptors: bf16 x bf16 -> fp32, M = 128; S: N = 128 both K-major; PV: N = 64, B MN-major (bit 16)
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      mbar_expect_tx(bar_q, Q_BYTES);
      tma_load_2d(&tmQKV, bar_q, sQ, h * HD, off + q0);
      mbar_expect_tx(bar_k, KV_BYTES);
      tma_load_2d(&tmKV, bar_k, sK, 512 + h * HD, off);
      mbar_expect_tx(bar_v, KV_BYTES);
      tma_load_2d(&tmKV, bar_v, sV, 1024 + h * HD, off);
      mbar_wait(bar_q, 0, 10);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T; afterwards the K buffer is reloaded with tile j + 1
        mbar_wait(bar_k, j & 1, 11);
        tc_fence_after();
        const uint64_t adesc = make_smem_desc(sQ), bdesc = make_smem_desc(sK);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tS, adesc + 2 * k, bdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s);
      };
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        const uint32_t ph = j & 1;
        if (j + 1 < nt) {  // K buffer is free once S_j has been computed
          mbar_wait(bar_s, ph, 12);
          mbar_expect_tx(bar_k, KV_BYTES);
          tma_load_2d(&tmKV, bar_k, sK, 512 + h * HD, off + (j + 1) * TK);
        }
        mbar_wait(bar_p, ph, 13);  // softmax j done: P_j in smem, O rescaled, S buffer free
        tc_fence_after();
        mbar_wait(bar_v, ph, 14);
#pragma unroll
        for (int s = 0; s < TK / 16; ++s) {
          // A = P: one K-major 128 x 64 tile; step s covers keys [16 s, 16 s + 16)
          const uint64_t adesc = make_smem_desc(sP) + 2 * s;
          const uint64_t bdesc = make_smem_desc_mn(sV + s * 2048);
          umma_bf16(tO, adesc, bdesc, idesc_o, (j > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(bar_o);
        if (j + 1 < nt) {
          issue_s(j + 1);  // queued right behind PV_j: the next softmax starts while O is still accumulating
          mbar_wait(bar_o, ph, 15);  // V buffer (and P) free once O has been updated
          mbar_expect_tx(bar_v, KV_BYTES);
          tma_load_2d(&tmKV, bar_v, sV, 1024 + h * HD, off + (j + 1) * TK);
        }
      }
    }
  } else {
    // ===================== softmax: one thread per query row =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    // A warp whose 32 query rows all lie beyond the utterance (last query tile) does no math: its P rows only feed O rows
    // that are never stored.  It still follows the barrier protocol tile by tile.
    const bool live = q0 + q * 32 < len;
    const int klim = chunk > 0 ? min(len, ((q0 + row) / chunk + 1) * chunk) : len;  // per query row in streaming mode
    for (int j = 0; j < nt; ++j) {
      const uint32_t ph = j & 1;
      const int k0 = j * TK;
      const int kvalid = klim - k0;  // visible keys of this tile for this row (<= 0: none, >= TK: all)
      mbar_wait(bar_s, ph, 16);
      tc_fence_after();
      if (!live) {
        if (j > 0) mbar_wait(bar_o, ph ^ 1, 17);  // stay in step: bar_p's previous phase has completed once PV_{j-1} has
        mbar_arrive(bar_p);
        continue;
      }
      // the whole 64-key row of S in registers: one TMEM round trip per key tile
      uint32_t s0[32], s1[32];
      tmem_ld32_issue(tS + lane_addr, s0);
      tmem_ld32_issue(tS + lane_addr + 32, s1);
      tmem_ld_wait();
      if (kvalid < TK) {  // full context: warp-uniform, only the last key tile masks
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kvalid) s0[i] = 0xff800000u;  // -inf
          if (i + 32 >= kvalid) s1[i] = 0xff800000u;
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s0[i]));
        mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s1[i]));
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // lazy rescaling: keep the reference max unless the row max grew by more than 2^8 (p <= 256 then: harmless)
      const float m_tile = mx * scale_log2e;
      float m_new = m_run, alpha = 1.f;
      bool rescale = false;
      if (m_tile > m_run + 8.f) {
        m_new = m_tile;
        alpha = fast_exp2(m_run - m_new);  // 0 at j == 0
        rescale = j > 0;
      }
      const bool any_rescale = __any_sync(0xffffffffu, rescale);  // tcgen05.ld / st are warp-collective
      if (j > 0) {
        mbar_wait(bar_o, ph ^ 1, 17);  // O_{j-1} accumulated: O is stable and the P buffer is free
        tc_fence_after();
      }
      if (j > 0 && any_rescale) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[16];
          tmem_ld16(tO + lane_addr + c * 16, o);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st16(tO + lane_addr + c * 16, o);
        }
      }
      float ls4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = fmaf(__uint_as_float(s0[i]), scale_log2e, -m_new);  // ex2(-inf) = 0 for masked keys
        const float x1 = fmaf(__uint_as_float(s1[i]), scale_log2e, -m_new);
        const bool poly = JV_ATTN_POLY_EVERY > 0 && (i % (JV_ATTN_POLY_EVERY > 0 ? JV_ATTN_POLY_EVERY : 1)) == 0;
        const float p0 = poly ? poly_exp2(x0) : fast_exp2(x0);
        const float p1 = poly ? poly_exp2(x1) : fast_exp2(x1);
        ls4[i & 3] += p0;
        ls4[i & 3] += p1;
        s0[i] = __float_as_uint(p0);
        s1[i] = __float_as_uint(p1);
      }
      // P[row][key]: K-major SWIZZLE_128B, 16-byte chunk = key / 8
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        sts128u(sP + swz128(row, g8), pack_bf16(__uint_as_float(s0[8 * g8]), __uint_as_float(s0[8 * g8 + 1])),
                pack_bf16(__uint_as_float(s0[8 * g8 + 2]), __uint_as_float(s0[8 * g8 + 3])),
                pack_bf16(__uint_as_float(s0[8 * g8 + 4]), __uint_as_float(s0[8 * g8 + 5])),
                pack_bf16(__uint_as_float(s0[8 * g8 + 6]), __uint_as_float(s0[8 * g8 + 7])));
        sts128u(sP + swz128(row, 4 + g8), pack_bf16(__uint_as_float(s1[8 * g8]), __uint_as_float(s1[8 * g8 + 1])),
                pack_bf16(__uint_as_float(s1[8 * g8 + 2]), __uint_as_float(s1[8 * g8 + 3])),
                pack_bf16(__uint_as_float(s1[8 * g8 + 4]), __uint_as_float(s1[8 * g8 + 5])),
                pack_bf16(__uint_as_float(s1[8 * g8 + 6]), __uint_as_float(s1[8 * g8 + 7])));
      }
      l_run = l_run * alpha + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
      m_run = m_new;
      fence_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      mbar_arrive(bar_p);
    }
    mbar_wait(bar_o, (nt - 1) & 1, 18);
    tc_fence_after();
    const int t = q0 + row;
    const float inv = 1.0f / l_run;
    bf16* dst = out + (long)(off + t) * ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(tO + lane_addr + c * 32, o);
      if (t < len) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[8 * g8]) * inv, __uint_as_float(o[8 * g8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[8 * g8 + 2]) * inv, __uint_as_float(o[8 * g8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[8 * g8 + 4]) * inv, __uint_as_float(o[8 * g8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[8 * g8 + 6]) * inv, __uint_as_float(o[8 * g8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + g8 * 8) = u;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace attn

static inline void launch_attention_tc(TmapCache& cache, const void* qkv, void* out, const int* row_off, const int* row_len,
                                       long M_alloc, int R, int Tmax_len, int chunk, cudaStream_t st) {
  static unsigned long long attr = 0;
  if (first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(attn::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn::SMEM_BYTES));
  const CUtensorMap tm = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, attn::TQ, 0);
  const CUtensorMap tmkv = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, attn::TK, 0);
  dim3 grid(cdiv(Tmax_len, attn::TQ), 8, R);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(attn::THREADS);
  cfg.dynamicSmemBytes = attn::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  JV_CUDA(cudaLaunchKernelEx(&cfg, attn::attention_tc_kernel, tm, tmkv, (bf16*)out, 512, row_off, row_len,
                             0.125f * 1.4426950408889634f, chunk));
  JV_LAUNCHED();
}

}  // namespace jv
