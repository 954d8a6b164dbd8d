// tcgen05 multi-head attention, second generation (bf16 mode default).  Same contract as attention_tc.cuh:
// 8 heads x 64, non-causal, keys masked by the row length (decoder.py:955-959), optional static chunk mask (streaming).
//
// What changed against attention_tc_kernel (81 us per call at batch 64 x 300: tensor 21 %, MUFU 35 % busy, every warp
// waiting on the softmax -> P -> smem -> PV -> S chain of its one query tile):
//   * S is double-buffered in TMEM (two 96-key tiles): S_{j+1} = Q K_{j+1}^T is computed while the softmax threads work
//     on S_j, so they never wait for the tensor core in steady state;
//   * P never goes through shared memory: the softmax threads write it back into the S columns (bf16 pairs, in place) and
//     the PV MMA takes its A operand from TMEM -- no st.shared, no fence.proxy.async, no P buffer;
//   * K / V of up to 384 keys are resident (a 4-stage ring of 96-key tiles, loaded up front), so no TMA latency sits
//     inside the per-tile chain at the benchmark's lengths; longer rows recycle the ring;
//   * key tiles are 96 wide: five hand-offs per query tile become four, and fully masked 32-key chunks of the last
//     tile are skipped (their exponentials are not evaluated, their PV k-steps are not issued).
// One CTA = 128 queries of one (row, head); two CTAs per SM (256 TMEM columns, 112 KB smem each).
//   warps 0..3 : softmax, one thread per query row (TMEM lane quarter = warp index)
//   warp 4     : MMA issuer (lane 0)
//   warp 5     : TMEM allocation, TMA producer (lane 0)
#pragma once
#include "attention_tc.cuh"

namespace jv {
namespace attn2 {

using namespace tc;
using attn::fast_exp2;
using attn::make_smem_desc_mn;
using attn::tmem_ld16;
using attn::tmem_ld32_issue;
using attn::tmem_ld_wait;
using attn::tmem_st16;

// Two shapes (NCH = 32-key chunks per key tile):
//   NCH = 3: 96-key tiles, 256 TMEM columns, 112 KB smem -> two CTAs per SM (few hand-offs per query tile)
//   NCH = 1: 32-key tiles, 128 TMEM columns,  48 KB smem -> four CTAs per SM (the other three cover a CTA's fixed costs)
constexpr int TQ = 128, HD = 64, NST = 4;            // NST K / V stages of TK keys
constexpr int Q_BYTES = TQ * HD * 2;                 // 16 KB
constexpr int THREADS = 192;
template <int NCH>
struct Shape {
  static constexpr int TK = 32 * NCH;
  static constexpr int KV_BYTES = TK * HD * 2;       // 12 KB / 4 KB
  static constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + NST * KV_BYTES, OFF_BAR = OFF_V + NST * KV_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256;
  static constexpr int TMEM_COLS = NCH == 1 ? 128 : 256;  // S0 [0,TK) | S1 [TK,2TK) | O [2TK, 2TK+64); P_b aliases the first TK/2 columns of S_b
  static constexpr int CTAS_PER_SM = NCH == 1 ? 4 : 2;
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int NCH>
__global__ void __launch_bounds__(THREADS, Shape<NCH>::CTAS_PER_SM)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, bf16* __restrict__ out, int ldo,
                     const int* __restrict__ row_off, const int* __restrict__ row_len, float scale_log2e, int chunk,
                     long long* __restrict__ trace) {
  constexpr int TK = Shape<NCH>::TK, KV_BYTES = Shape<NCH>::KV_BYTES, OFF_K = Shape<NCH>::OFF_K, OFF_V = Shape<NCH>::OFF_V,
                OFF_BAR = Shape<NCH>::OFF_BAR, TMEM_COLS = Shape<NCH>::TMEM_COLS;
  // rows in reverse order: the QKV GEMM has just written 119 MB (about the size of L2) in increasing row order, so the LAST rows
  // are the ones still in L2 when this kernel starts
  const int r = gridDim.z - 1 - blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * TQ;
  const int len = row_len[r];
  // optional per-CTA timeline (jv_debug_attention_trace): 8 clock64 values written by thread 0 (softmax warp 0)
  // (only in -DJV_TRACE builds: the run-time flag around the waits costs a few per cent)
#ifdef JV_TRACE
  long long* tr = trace ? trace + 8L * ((long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) : nullptr;
  const bool tracer = tr != nullptr && threadIdx.x == 0;
#else
  constexpr long long* tr = nullptr;
  constexpr bool tracer = false;
#endif
  long long w_s = 0, w_o = 0;
  if (tracer) tr[0] = clock64();
  if (q0 >= len) return;
  const int off = row_off[r];
  // streaming=True (decoder.py:950-953): query t sees keys < min(len, (t / chunk + 1) * chunk); chunk = 0: all keys
  const int kend = chunk > 0 ? min(len, ((q0 + TQ - 1) / chunk + 1) * chunk) : len;
  const int nt = (kend + TK - 1) / TK;

  extern __shared__ __align__(1024) uint8_t smem_attn2[];
  const uint32_t base = smem_u32(smem_attn2);
  const uint32_t sQ = base, sK = base + OFF_K, sV = base + OFF_V;
  const uint32_t bars = base + OFF_BAR;
  const uint32_t bar_q = bars, bar_o = bars + 8;
  const uint32_t bar_s = bars + 16;          // [2] S buffer b holds a fresh S tile
  const uint32_t bar_p = bars + 32;          // [2] P written into buffer b (128 arrivals)
  const uint32_t bar_k = bars + 48;          // [NST] K stage loaded
  const uint32_t bar_v = bar_k + 8 * NST;    // [NST] V stage loaded
  const uint32_t bar_kf = bar_v + 8 * NST;   // [NST] K stage consumed (S MMAs retired)
  const uint32_t bar_vf = bar_kf + 8 * NST;  // [NST] V stage consumed (PV MMAs retired)
  const uint32_t tmem_slot = bar_vf + 8 * NST;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_attn2 + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 5) {
    if (lane == 0) {
      if (base & 1023u) {  // SWIZZLE_128B tiles need 1 KB alignment; the smem budget has no room for an alignment slack
        printf("jyutvoice_b200: attention smem base not 1 KB aligned\n");
        __trap();
      }
      mbar_init(bar_q, 1);
      mbar_init(bar_o, 1);
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_s + 8 * i, 1);
        mbar_init(bar_p + 8 * i, 128);
      }
      for (int i = 0; i < NST; ++i) {
        mbar_init(bar_k + 8 * i, 1);
        mbar_init(bar_v + 8 * i, 1);
        mbar_init(bar_kf + 8 * i, 1);
        mbar_init(bar_vf + 8 * i, 1);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmQ) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmKV) : "memory");
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
  const uint32_t tO = tmem_base + 2 * TK;
  if (tracer) tr[1] = clock64();

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_expect_tx(bar_q, Q_BYTES);
      tma_load_2d(&tmQ, bar_q, sQ, h * HD, off + q0);
      for (int t = 0; t < nt; ++t) {
        const int s = t % NST;
        if (t >= NST) mbar_wait(bar_kf + 8 * s, ((t / NST) - 1) & 1, 21);  // S_{t - NST} has retired: the K stage is free
        mbar_expect_tx(bar_k + 8 * s, KV_BYTES);
        tma_load_2d(&tmKV, bar_k + 8 * s, sK + s * KV_BYTES, 512 + h * HD, off + t * TK);
        if (t >= NST) mbar_wait(bar_vf + 8 * s, ((t / NST) - 1) & 1, 22);
        mbar_expect_tx(bar_v + 8 * s, KV_BYTES);
        tma_load_2d(&tmKV, bar_v + 8 * s, sV + s * KV_BYTES, 1024 + h * HD, off + t * TK);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptors: bf16 x bf16 -> fp32, M = 128.  S: N = 96, both operands K-major.  PV: N = 64, A (= P) from
      // TMEM, B (= V) MN-major (bit 16).
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      const uint64_t qdesc = make_smem_desc(sQ);
      auto issue_s = [&](int t) {  // S_t = Q K_t^T into buffer t & 1
        const int s = t % NST;
        mbar_wait(bar_k + 8 * s, (t / NST) & 1, 23);
        tc_fence_after();
        const uint64_t kdesc = make_smem_desc(sK + s * KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + (t & 1) * TK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s + 8 * (t & 1));
        umma_commit(bar_kf + 8 * s);
      };
      mbar_wait(bar_q, 0, 24);
      issue_s(0);
      if (nt > 1) issue_s(1);
      for (int j = 0; j < nt; ++j) {
        const int b = j & 1, s = j % NST;
        mbar_wait(bar_p + 8 * b, (j >> 1) & 1, 25);  // softmax j done: P_j in TMEM (over S_j), O rescaled if it had to be
        tc_fence_after();
        mbar_wait(bar_v + 8 * s, (j / NST) & 1, 26);
        tc_fence_after();
        const int kv = min(kend - j * TK, TK);
        const int ksteps = ((kv + 31) / 32) * 2;  // 16 keys per step; whole 32-key chunks (the softmax threads zero-fill a chunk's masked keys)
        const uint32_t tP = tmem_base + b * TK;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16_ts(tO, tP + ks * 8, make_smem_desc_mn(sV + s * KV_BYTES + ks * 2048), idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_o);
        umma_commit(bar_vf + 8 * s);
        if (j + 2 < nt) issue_s(j + 2);  // executes after PV_j (in-order pipe): buffer b is free again
      }
    }
  } else {
    // ===================== softmax: one thread per query row =====================
    const int q = warp;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    // A warp whose 32 query rows all lie beyond the utterance (last query tile) does no math: its P rows only feed O rows
    // that are never stored.  It still follows the barrier protocol tile by tile.
    const bool live = q0 + q * 32 < len;
    const int klim = chunk > 0 ? min(len, ((q0 + row) / chunk + 1) * chunk) : len;  // per query row in streaming mode
    for (int j = 0; j < nt; ++j) {
      const int b = j & 1;
      const int k0 = j * TK;
      const int kvalid = klim - k0;                            // visible keys of this tile for this row (<= 0: none)
      const int nch = (min(kend - k0, TK) + 31) >> 5;          // 32-key chunks the CTA processes in this tile (1..3)
      const uint32_t tS = tmem_base + lane_addr + b * TK;
      {
        const long long c0 = tracer ? clock64() : 0;
        mbar_wait(bar_s + 8 * b, (j >> 1) & 1, 27);
        if (tracer) {
          const long long c1 = clock64();
          if (j == 0) tr[2] = c1;
          else w_s += c1 - c0;
        }
      }
      tc_fence_after();
      if (!live) {
        if (j > 0) mbar_wait(bar_o, (j - 1) & 1, 28);  // stay in step with bar_o's phases
        mbar_arrive(bar_p + 8 * b);
        continue;
      }
      uint32_t s0[32], s1[NCH > 1 ? 32 : 1], s2[NCH > 2 ? 32 : 1];
      tmem_ld32_issue(tS, s0);
      if constexpr (NCH > 1) {
        if (nch > 1) tmem_ld32_issue(tS + 32, s1);
      }
      if constexpr (NCH > 2) {
        if (nch > 2) tmem_ld32_issue(tS + 64, s2);
      }
      tmem_ld_wait();
      if constexpr (NCH > 1) {
        if (nch < 2) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s1[i] = 0xff800000u;
        }
      }
      if constexpr (NCH > 2) {
        if (nch < 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) s2[i] = 0xff800000u;
        }
      }
      if (kvalid < TK) {  // full context: warp-uniform, only the last key tile masks
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kvalid) s0[i] = 0xff800000u;  // -inf
          if constexpr (NCH > 1) { if (i + 32 >= kvalid) s1[i] = 0xff800000u; }
          if constexpr (NCH > 2) { if (i + 64 >= kvalid) s2[i] = 0xff800000u; }
        }
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float v = __uint_as_float(s0[i]);
        if constexpr (NCH > 1) v = fmaxf(v, __uint_as_float(s1[i]));
        if constexpr (NCH > 2) v = fmaxf(v, __uint_as_float(s2[i]));
        mx4[i & 3] = fmaxf(mx4[i & 3], v);
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
      bool waited_o = false;
      if (j > 0 && any_rescale) {
        mbar_wait(bar_o, (j - 1) & 1, 29);  // O_{j-1} accumulated: O is stable
        tc_fence_after();
        waited_o = true;
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
      uint32_t pk[16];
      // chunk 0 is always processed; chunks 1, 2 only when the tile reaches them (ex2(-inf) = 0 covers per-row masks)
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(s0[i]), scale_log2e, -m_new));
        const float p1 = fast_exp2(fmaf(__uint_as_float(s0[i + 1]), scale_log2e, -m_new));
        ls4[(i >> 1) & 3] += p0 + p1;
        pk[i >> 1] = pack_bf16(p0, p1);
      }
      tmem_st16_nowait(tS, pk);
      if constexpr (NCH > 1) {
        if (nch > 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(s1[i]), scale_log2e, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(s1[i + 1]), scale_log2e, -m_new));
            ls4[(i >> 1) & 3] += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16_nowait(tS + 16, pk);
        }
      }
      if constexpr (NCH > 2) {
        if (nch > 2) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(s2[i]), scale_log2e, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(s2[i + 1]), scale_log2e, -m_new));
            ls4[(i >> 1) & 3] += p0 + p1;
            pk[i >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16_nowait(tS + 32, pk);
        }
      }
      l_run = l_run * alpha + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
      m_run = m_new;
      if (j > 0 && !waited_o) {  // long done by now: keeps this thread in step with bar_o's phases
        const long long c0 = tracer ? clock64() : 0;
        mbar_wait(bar_o, (j - 1) & 1, 30);
        if (tracer) w_o += clock64() - c0;
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar_p + 8 * b);
    }
    if (tracer) tr[3] = clock64();
    mbar_wait(bar_o, (nt - 1) & 1, 31);
    tc_fence_after();
    if (tracer) tr[4] = clock64();
    const int t = q0 + row;
    const float inv = 1.0f / l_run;
    bf16* dst = out + (long)(off + t) * ldo + h * HD;
    if (live) {
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
  }
  if (tracer) {
    tr[5] = clock64();
    tr[6] = w_s;
    tr[7] = w_o;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace attn2

// debug: device buffer of 8 int64 per attention CTA (grid order) that the next launches fill with their timeline
inline long long*& attention_trace_buffer() {
  static long long* p = nullptr;
  return p;
}

// JYUTVOICE_B200_ATTN_NCH = 1 | 3: key-tile shape of attention_tc2_kernel (see Shape)
static inline int attention_nch() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_ATTN_NCH");
    v = e ? atoi(e) : 1;
    if (v != 1 && v != 3) v = 1;
  }
  return v;
}

template <int NCH>
static inline void launch_attention_tc2_shape(TmapCache& cache, const void* qkv, void* out, const int* row_off, const int* row_len,
                                              long M_alloc, int R, int Tmax_len, int chunk, cudaStream_t st) {
  using S = attn2::Shape<NCH>;
  static unsigned long long attr = 0;
  if (first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(attn2::attention_tc2_kernel<NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM_BYTES));
  const CUtensorMap tm = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, attn2::TQ, 0);
  const CUtensorMap tmkv = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, S::TK, 0);
  dim3 grid(cdiv(Tmax_len, attn2::TQ), 8, R);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(attn2::THREADS);
  cfg.dynamicSmemBytes = S::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  JV_CUDA(cudaLaunchKernelEx(&cfg, attn2::attention_tc2_kernel<NCH>, tm, tmkv, (bf16*)out, 512, row_off, row_len,
                             0.125f * 1.4426950408889634f, chunk, attention_trace_buffer()));
  JV_LAUNCHED();
}

static inline void launch_attention_tc2(TmapCache& cache, const void* qkv, void* out, const int* row_off, const int* row_len,
                                        long M_alloc, int R, int Tmax_len, int chunk, cudaStream_t st) {
  if (attention_nch() == 1) launch_attention_tc2_shape<1>(cache, qkv, out, row_off, row_len, M_alloc, R, Tmax_len, chunk, st);
  else launch_attention_tc2_shape<3>(cache, qkv, out, row_off, row_len, M_alloc, R, Tmax_len, chunk, st);
}

}  // namespace jv
