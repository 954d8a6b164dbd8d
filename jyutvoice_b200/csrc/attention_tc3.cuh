// tcgen05 multi-head attention, persistent variant (bf16 mode default when every row has <= NST * TK = 384 keys).
// Same contract as attention_tc.cuh / attention_tc2.cuh.
//
// attention_tc2_kernel's per-CTA timeline on B200 (tools/attention_trace.py, batch 64 x 300, clock cycles of one CTA =
// 128 queries of one (row, head)): set-up 1400, first S tile ready 2640, softmax loop 8135, last PV 676, O read-out and
// store 1325 -- 43 % of a CTA's life is fixed cost that the one other CTA on the SM cannot cover, and the MUFU (the
// bounding unit: one ex2 per score) is 30 % busy.  Here a CTA is persistent: it owns a contiguous range of (row, head,
// query tile) items and pipelines ACROSS items:
//   * barriers, TMEM and descriptors are set up once;
//   * the K / V tiles of a (row, head) are loaded once and serve all its query tiles; the next (row, head)'s tiles are
//     prefetched stage by stage as the last query tile of the current one retires its MMAs;
//   * the next Q tile is loaded as soon as the current item's last S MMA has retired, and the MMA thread runs two S tiles
//     ahead across the item boundary, so the softmax threads find the next item's first S tile waiting for them;
//   * only the O read-out of an item is still in the softmax threads' way.
// Tile pipeline inside an item as in attention_tc2_kernel: S double-buffered in TMEM, P written over S and consumed by the
// PV MMA from TMEM, 96-key tiles, masked 32-key chunks skipped.
//   warps 0..3 : softmax, one thread per query row (TMEM lane quarter = warp index)
//   warp 4     : MMA issuer (lane 0)
//   warp 5     : TMEM allocation, TMA producer (lane 0)
#pragma once
#include "attention_tc2.cuh"

namespace jv {
namespace attn3 {

using namespace tc;
using attn::fast_exp2;
using attn::make_smem_desc_mn;
using attn::tmem_ld16;
using attn::tmem_ld32_issue;
using attn::tmem_ld_wait;
using attn::tmem_st16;
using attn2::tmem_st16_nowait;
using attn2::tmem_st_wait;
using attn2::umma_bf16_ts;

constexpr int TQ = 128, TK = 96, HD = 64, NST = 4;
constexpr int Q_BYTES = TQ * HD * 2;   // 16 KB
constexpr int KV_BYTES = TK * HD * 2;  // 12 KB
constexpr int OFF_K = Q_BYTES, OFF_V = OFF_K + NST * KV_BYTES, OFF_BAR = OFF_V + NST * KV_BYTES;
constexpr int MAX_ITEMS = 64;                                // items per CTA (decoded once into smem; the host sizes the grid accordingly)
constexpr int OFF_ITEMS = OFF_BAR + 256;
constexpr int SMEM_BYTES = OFF_ITEMS + MAX_ITEMS * 12;       // 112 KB + 1 KB = 115712 B: exactly two CTAs per SM
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;  // S0 [0,96) | S1 [96,192) | O [192,256); P_b aliases the first 48 columns of S_b

// One work item = 128 queries of one (row, head).  items[] holds (row << 8 | head << 4 | query tile), ordered, non-empty.
struct Item {
  int code, h, q0, len, off, kend, nt;
  bool last_in_group;  // last item of its (row, head) inside this CTA's range: its MMAs release the K / V stages
  bool first_in_group;
};

// Decoded once per CTA into shared memory (three words per item): every role walks the same item list, and four dependent
// global loads per item on the MMA thread's and the softmax threads' paths cost more than the whole S hand-off (measured).
__device__ __forceinline__ void decode_item(int* __restrict__ tab, const int* __restrict__ items, int i, int beg, int end,
                                            const int* __restrict__ row_off, const int* __restrict__ row_len, int chunk) {
  const int code = items[i];
  const int r = code >> 8;
  const int q0 = (code & 15) * TQ;
  const int len = row_len[r];
  // streaming=True (decoder.py:950-953): query t sees keys < min(len, (t / chunk + 1) * chunk); chunk = 0: all keys
  const int kend = chunk > 0 ? min(len, ((q0 + TQ - 1) / chunk + 1) * chunk) : len;
  const int nt = (kend + TK - 1) / TK;
  const int last = ((i + 1 == end) || ((items[i + 1] >> 4) != (code >> 4))) ? 1 : 0;
  const int first = (i == beg || (items[i - 1] >> 4) != (code >> 4)) ? 1 : 0;
  int* e = tab + 3 * (i - beg);
  e[0] = row_off[r];
  e[1] = len | (kend << 16);
  e[2] = (code & 0xff) | (nt << 8) | (last << 16) | (first << 17);
}
struct ItemTab {
  const int* tab;
  int beg;
};
__device__ __forceinline__ void load_item(Item& it, const ItemTab& t, int i) {
  const int* e = t.tab + 3 * (i - t.beg);
  const int w1 = e[1], w2 = e[2];
  it.off = e[0];
  it.len = w1 & 0xffff;
  it.kend = w1 >> 16;
  it.code = w2 & 0xff;  // head << 4 | query tile
  it.h = (w2 >> 4) & 15;
  it.q0 = (w2 & 15) * TQ;
  it.nt = (w2 >> 8) & 0xff;
  it.last_in_group = (w2 >> 16) & 1;
  it.first_in_group = (w2 >> 17) & 1;
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, bf16* __restrict__ out, int ldo,
                     const int* __restrict__ row_off, const int* __restrict__ row_len, const int* __restrict__ items, int n_items,
                     float scale_log2e, int chunk, long long* __restrict__ trace) {
  // optional per-CTA timeline (jv_debug_attention_trace): sums over the CTA's items, written by thread 0 (softmax warp 0)
#ifdef JV_TRACE
  long long* tr = trace ? trace + 16L * blockIdx.x : nullptr;
#else
  constexpr long long* tr = nullptr;
#endif
  const bool tracer = tr != nullptr && threadIdx.x == 0;
  long long t_decode = 0, t_first_s = 0, t_loop = 0, t_wait_s = 0, t_wait_o = 0, t_final = 0, t_store = 0;
  const long long t_entry = tracer ? clock64() : 0;
  // contiguous, balanced item range of this CTA
  const int beg = (int)(((long)blockIdx.x * n_items) / gridDim.x);
  const int end = (int)(((long)(blockIdx.x + 1) * n_items) / gridDim.x);

  extern __shared__ __align__(1024) uint8_t smem_attn3[];
  const uint32_t base = smem_u32(smem_attn3);
  const uint32_t sQ = base, sK = base + OFF_K, sV = base + OFF_V;
  const uint32_t bars = base + OFF_BAR;
  const uint32_t bar_q = bars;               // Q tile loaded
  const uint32_t bar_qf = bars + 8;          // Q tile consumed (last S MMA of the item retired)
  const uint32_t bar_o = bars + 16;          // PV of a tile retired (one completion per tile, in order)
  const uint32_t bar_s = bars + 24;          // [2] S buffer b holds a fresh S tile
  const uint32_t bar_p = bars + 40;          // [2] P written into buffer b (128 arrivals)
  const uint32_t bar_k = bars + 56;          // [NST] K stage loaded
  const uint32_t bar_v = bar_k + 8 * NST;    // [NST] V stage loaded
  const uint32_t bar_kf = bar_v + 8 * NST;   // [NST] K stage released (last S MMA that reads it retired)
  const uint32_t bar_vf = bar_kf + 8 * NST;  // [NST] V stage released
  const uint32_t tmem_slot = bar_vf + 8 * NST;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_attn3 + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* item_tab = reinterpret_cast<int*>(smem_attn3 + OFF_ITEMS);
  // item list, lengths and offsets are written at layout time, not by the previous kernel: decoded before pdl_wait()
  for (int i = beg + (int)threadIdx.x; i < end; i += THREADS) decode_item(item_tab, items, i, beg, end, row_off, row_len, chunk);
  const ItemTab itab{item_tab, beg};
  if (warp == 5) {
    if (lane == 0) {
      if (base & 1023u) {  // SWIZZLE_128B tiles need 1 KB alignment; the smem budget has no room for an alignment slack
        printf("jyutvoice_b200: attention smem base not 1 KB aligned\n");
        __trap();
      }
      mbar_init(bar_q, 1);
      mbar_init(bar_qf, 1);
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

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int n_k[NST] = {0, 0, 0, 0}, n_v[NST] = {0, 0, 0, 0};  // loads issued so far into each stage
      int n_q = 0;
      Item it;
      auto load_q = [&]() {
        if (n_q > 0) mbar_wait(bar_qf, (n_q - 1) & 1, 40);  // the previous item's last S MMA has retired
        mbar_expect_tx(bar_q, Q_BYTES);
        tma_load_2d(&tmQ, bar_q, sQ, it.h * HD, it.off + it.q0);
        ++n_q;
      };
      auto load_k = [&](int s) {
        if (n_k[s] > 0) mbar_wait(bar_kf + 8 * s, (n_k[s] - 1) & 1, 41);
        mbar_expect_tx(bar_k + 8 * s, KV_BYTES);
        tma_load_2d(&tmKV, bar_k + 8 * s, sK + s * KV_BYTES, 512 + it.h * HD, it.off + s * TK);
        ++n_k[s];
      };
      auto load_v = [&](int s) {
        if (n_v[s] > 0) mbar_wait(bar_vf + 8 * s, (n_v[s] - 1) & 1, 42);
        mbar_expect_tx(bar_v + 8 * s, KV_BYTES);
        tma_load_2d(&tmKV, bar_v + 8 * s, sV + s * KV_BYTES, 1024 + it.h * HD, it.off + s * TK);
        ++n_v[s];
      };
      for (int i = beg; i < end; ++i) {
        load_item(it, itab, i);
        if (it.first_in_group) {
          // tiles this (row, head) needs inside the range = those of its last item here (kend grows with the query tile)
          Item li = it;
          for (int last = i; !li.last_in_group;) load_item(li, itab, ++last);
          const int nkv = li.nt;
          if (i == beg) load_q();  // nothing to wait for: Q first
          // in the order the previous group's last item releases things (its S tiles run two ahead of its PV tiles):
          // K0 K1 K2 V0 K3 V1 [Q] V2 V3
          for (int s = 0; s < nkv; ++s) {
            load_k(s);
            if (s >= 2) load_v(s - 2);
          }
          if (i != beg) load_q();
          if (nkv >= 2) load_v(nkv - 2);
          load_v(nkv - 1);
        } else {
          load_q();
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0 && beg < end) {
      // instruction descriptors: bf16 x bf16 -> fp32, M = 128.  S: N = 96, both operands K-major.  PV: N = 64, A (= P) from
      // TMEM, B (= V) MN-major (bit 16).
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TK >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TQ >> 4) << 24);
      const uint64_t qdesc = make_smem_desc(sQ);
      // two cursors over the CTA's tile stream: S runs two tiles ahead of PV
      struct Cursor {
        int i, j, g;  // item index, tile of the item, global tile count
        Item it;
        int grp;      // loads of the current group's stages consumed: parity source for bar_k / bar_v
        int n_item;   // items entered (parity source for bar_q)
      };
      Cursor cs, cp;
      auto enter = [&](Cursor& c, int i) {
        c.i = i;
        c.j = 0;
        if (i < end) {
          load_item(c.it, itab, i);
          if (c.it.first_in_group) ++c.grp;
          ++c.n_item;
        }
      };
      auto advance = [&](Cursor& c) {
        ++c.g;
        if (++c.j == c.it.nt) enter(c, c.i + 1);
      };
      cs.g = cp.g = 0;
      cs.grp = cp.grp = 0;
      cs.n_item = cp.n_item = 0;
      enter(cs, beg);
      enter(cp, beg);
      // Every stage is loaded exactly once per group that uses it, and a group uses stages 0 .. nkv-1; a stage's load count
      // therefore is NOT the group count when groups differ in length.  Track per-stage counts in both cursors.
      int ks_cnt[NST] = {0, 0, 0, 0}, vp_cnt[NST] = {0, 0, 0, 0};
      int ks_grp[NST] = {0, 0, 0, 0}, vp_grp[NST] = {0, 0, 0, 0};  // group in which the stage was last counted
      auto test = [&](uint32_t bar, uint32_t parity) {  // non-blocking: has that phase completed?
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        return ok != 0;
      };
      // The loop below never blocks on one of the two streams while the other could make progress: a late Q tile (it can only
      // be fetched once the previous item's last S MMA has retired, one tile before it is wanted) must not hold up the PV of
      // the tiles in flight.
      auto s_ready = [&]() {  // inputs of the next S tile have landed (also registers the stage's use in this group)
        const int s = cs.j;
        if (cs.j == 0 && !test(bar_q, (cs.n_item - 1) & 1)) return false;
        if (ks_grp[s] != cs.grp) { ks_grp[s] = cs.grp; ++ks_cnt[s]; }  // first use of this stage in this group
        return test(bar_k + 8 * s, (ks_cnt[s] - 1) & 1);
      };
      auto issue_s = [&]() {
        const int s = cs.j, b = cs.g & 1;
        tc_fence_after();
        const uint64_t kdesc = make_smem_desc(sK + s * KV_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(tmem_base + b * TK, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        umma_commit(bar_s + 8 * b);
        if (cs.j == cs.it.nt - 1) umma_commit(bar_qf);         // the Q tile may be replaced
        if (cs.it.last_in_group) umma_commit(bar_kf + 8 * s);  // ... and so may this K stage
        advance(cs);
      };
      auto p_ready = [&]() {
        const int s = cp.j, b = cp.g & 1;
        if (!test(bar_p + 8 * b, (cp.g >> 1) & 1)) return false;  // softmax of this tile done: P in TMEM (over S), O rescaled if needed
        if (vp_grp[s] != cp.grp) { vp_grp[s] = cp.grp; ++vp_cnt[s]; }
        return test(bar_v + 8 * s, (vp_cnt[s] - 1) & 1);
      };
      auto issue_pv = [&]() {
        const int s = cp.j, b = cp.g & 1;
        tc_fence_after();
        const int kv = min(cp.it.kend - cp.j * TK, TK);
        const int ksteps = ((kv + 31) / 32) * 2;  // 16 keys per step; whole 32-key chunks (masked keys of a chunk hold P = 0)
        const uint32_t tP = tmem_base + b * TK;
        for (int ks = 0; ks < ksteps; ++ks)
          umma_bf16_ts(tO, tP + ks * 8, make_smem_desc_mn(sV + s * KV_BYTES + ks * 2048), idesc_o, (cp.j > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_o);
        if (cp.it.last_in_group) umma_commit(bar_vf + 8 * s);
        advance(cp);
      };
      uint32_t idle = 0;
      long long t0 = 0;
      while (cp.i < end) {
        bool progressed = false;
        // S tile g may be written once PV of tile g - 2 has been ISSUED (in-order tensor pipe): at most two S tiles ahead
        if (cs.i < end && cs.g < cp.g + 2 && s_ready()) { issue_s(); progressed = true; }
        if (p_ready()) { issue_pv(); progressed = true; }
        if (progressed) { idle = 0; t0 = 0; continue; }
        // Nothing to issue: SLEEP on the next P tile (try_wait suspends the thread until the phase completes or the
        // hardware's time limit passes).  A spinning MMA thread out-prioritises the softmax warp of its scheduler
        // (highest warp id first) -- measured: the softmax threads then wait ~2000 clk per item for their own PV.
        mbar_try_wait(bar_p + 8 * (cp.g & 1), (cp.g >> 1) & 1);
        if ((++idle & 0xfffu) == 0) {  // bounded: a protocol bug traps instead of hanging the GPU
          const long long now = clock64();
          if (t0 == 0) t0 = now;
          if (now - t0 > 4000000000LL) {
            printf("jyutvoice_b200: attention MMA loop stalled (block %d, S tile %d, PV tile %d)\n", blockIdx.x, cs.g, cp.g);
            __trap();
          }
        }
      }
    }
  } else {
    // ===================== softmax: one thread per query row =====================
    const int q = warp;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int g = 0;          // global tile count of this CTA
    int o_seen = -1;    // last tile whose PV completion this thread has waited for (bar_o completes once per tile, in order)
    Item it;
    const long long t_setup = tracer ? clock64() : 0;
    for (int i = beg; i < end; ++i) {
      const long long c_item = tracer ? clock64() : 0;
      load_item(it, itab, i);
      long long c_loop0 = 0;
      if (tracer) {
        c_loop0 = clock64() + (long long)(it.len & 0);  // (depends on the loaded item: ordered after the loads)
        t_decode += c_loop0 - c_item;
      }
      float m_run = -INFINITY, l_run = 0.f;
      // A warp whose 32 query rows all lie beyond the utterance (last query tile) does no math: its P rows only feed O rows
      // that are never stored.  It still follows the barrier protocol tile by tile.
      const bool live = it.q0 + q * 32 < it.len;
      const int klim = chunk > 0 ? min(it.len, ((it.q0 + row) / chunk + 1) * chunk) : it.len;  // per query row in streaming mode
      for (int j = 0; j < it.nt; ++j, ++g) {
        const int b = g & 1;
        const int k0 = j * TK;
        const int kvalid = klim - k0;                        // visible keys of this tile for this row (<= 0: none)
        const int nch = (min(it.kend - k0, TK) + 31) >> 5;   // 32-key chunks the CTA processes in this tile (1..3)
        const uint32_t tS = tmem_base + lane_addr + b * TK;
        {
          const long long c0 = tracer ? clock64() : 0;
          mbar_wait(bar_s + 8 * b, (g >> 1) & 1, 47);
          if (tracer) {
            const long long c1 = clock64();
            if (j == 0) { t_first_s += c1 - c0; c_loop0 = c1; }
            else t_wait_s += c1 - c0;
          }
        }
        tc_fence_after();
        if (!live) {
          if (g > 0 && o_seen < g - 1) { mbar_wait(bar_o, (g - 1) & 1, 48); o_seen = g - 1; }  // stay in step with bar_o's phases
          mbar_arrive(bar_p + 8 * b);
          continue;
        }
        uint32_t s0[32], s1[32], s2[32];
        tmem_ld32_issue(tS, s0);
        if (nch > 1) tmem_ld32_issue(tS + 32, s1);
        if (nch > 2) tmem_ld32_issue(tS + 64, s2);
        tmem_ld_wait();
        if (nch < 2) {
#pragma unroll
          for (int e = 0; e < 32; ++e) s1[e] = 0xff800000u;
        }
        if (nch < 3) {
#pragma unroll
          for (int e = 0; e < 32; ++e) s2[e] = 0xff800000u;
        }
        if (kvalid < TK) {  // full context: warp-uniform, only the last key tile masks
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            if (e >= kvalid) s0[e] = 0xff800000u;  // -inf
            if (e + 32 >= kvalid) s1[e] = 0xff800000u;
            if (e + 64 >= kvalid) s2[e] = 0xff800000u;
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int e = 0; e < 32; ++e)
          mx4[e & 3] = fmaxf(mx4[e & 3], fmaxf(__uint_as_float(s0[e]), fmaxf(__uint_as_float(s1[e]), __uint_as_float(s2[e]))));
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
        if (j > 0 && any_rescale) {
          if (o_seen < g - 1) { mbar_wait(bar_o, (g - 1) & 1, 49); o_seen = g - 1; }  // O accumulated up to the previous tile: stable
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[16];
            tmem_ld16(tO + lane_addr + c * 16, o);
#pragma unroll
            for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
            tmem_st16(tO + lane_addr + c * 16, o);
          }
        }
        float ls4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(s0[e]), scale_log2e, -m_new));
          const float p1 = fast_exp2(fmaf(__uint_as_float(s0[e + 1]), scale_log2e, -m_new));
          ls4[(e >> 1) & 3] += p0 + p1;
          pk[e >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16_nowait(tS, pk);
        if (nch > 1) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(s1[e]), scale_log2e, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(s1[e + 1]), scale_log2e, -m_new));
            ls4[(e >> 1) & 3] += p0 + p1;
            pk[e >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16_nowait(tS + 16, pk);
        }
        if (nch > 2) {
#pragma unroll
          for (int e = 0; e < 32; e += 2) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(s2[e]), scale_log2e, -m_new));
            const float p1 = fast_exp2(fmaf(__uint_as_float(s2[e + 1]), scale_log2e, -m_new));
            ls4[(e >> 1) & 3] += p0 + p1;
            pk[e >> 1] = pack_bf16(p0, p1);
          }
          tmem_st16_nowait(tS + 32, pk);
        }
        l_run = l_run * alpha + ((ls4[0] + ls4[1]) + (ls4[2] + ls4[3]));
        m_run = m_new;
        if (g > 0 && o_seen < g - 1) {  // long done: keeps bar_o's phases in step
          const long long c0 = tracer ? clock64() : 0;
          mbar_wait(bar_o, (g - 1) & 1, 50);
          o_seen = g - 1;
          if (tracer) t_wait_o += clock64() - c0;
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar_p + 8 * b);
      }
      // ---- item done: O / l -> bf16 -> global (the MMA thread is already computing the next item's first S tiles)
      const long long c_end = tracer ? clock64() : 0;
      mbar_wait(bar_o, (g - 1) & 1, 51);
      o_seen = g - 1;
      tc_fence_after();
      const long long c_fin = tracer ? clock64() : 0;
      if (tracer) {
        t_loop += c_end - c_loop0;
        t_final += c_fin - c_end;
      }
      if (live) {
        const int t = it.q0 + row;
        const float inv = 1.0f / l_run;
        bf16* dst = out + (long)(it.off + t) * ldo + it.h * HD;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld32(tO + lane_addr + c * 32, o);
          if (t < it.len) {
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
      tc_fence_before();  // the O reads above are ordered before this thread's next P arrival, which gates the PV that overwrites O
      if (tracer) t_store += clock64() - c_fin;
    }
    if (tracer) {
      tr[0] = t_setup - t_entry; tr[1] = t_decode; tr[2] = t_first_s; tr[3] = t_loop; tr[4] = t_wait_s; tr[5] = t_wait_o;
      tr[6] = t_final; tr[7] = t_store; tr[8] = clock64() - t_entry; tr[9] = end - beg; tr[10] = g;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

}  // namespace attn3

// JYUTVOICE_B200_ATTN: 2 (default) = attention_tc2_kernel, 3 = persistent kernel where it applies (measured slower:
// DESIGN.md section 4), 1 = attention_tc_kernel
static inline int attention_version() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_ATTN");
    v = e ? atoi(e) : 2;
    if (v < 1 || v > 3) v = 2;
  }
  return v;
}

// items: device array of n_items codes (row << 8 | head << 4 | query tile), ordered by (row, head, tile), only tiles with
// q0 < len (built once per layout: estimator.cu upload_layout).
static inline void launch_attention(TmapCache& cache, const void* qkv, void* out, const int* row_off, const int* row_len,
                                    const int* items, int n_items, long M_alloc, int R, int Tmax_len, int chunk, int num_sms,
                                    cudaStream_t st) {
  const int ver = attention_version();
  if (ver == 1) {
    launch_attention_tc(cache, qkv, out, row_off, row_len, M_alloc, R, Tmax_len, chunk, st);
    return;
  }
  if (ver == 2 || Tmax_len > attn3::NST * attn3::TK || items == nullptr || n_items <= 0) {  // long rows recycle the K / V ring per item
    launch_attention_tc2(cache, qkv, out, row_off, row_len, M_alloc, R, Tmax_len, chunk, st);
    return;
  }
  static unsigned long long attr = 0;
  if (first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(attn3::attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, attn3::SMEM_BYTES));
  const CUtensorMap tm = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, attn3::TQ, 0);
  const CUtensorMap tmkv = cache.get(qkv, 1536, M_alloc, 1536 * 2, 64, attn3::TK, 0);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  int grid = n_items < 2 * num_sms ? n_items : 2 * num_sms;
  grid = std::max(grid, cdiv(n_items, attn3::MAX_ITEMS));  // a CTA's item table holds MAX_ITEMS entries
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(attn3::THREADS);
  cfg.dynamicSmemBytes = attn3::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = lattr;
  cfg.numAttrs = use_pdl() ? 1 : 0;
  JV_CUDA(cudaLaunchKernelEx(&cfg, attn3::attention_tc3_kernel, tm, tmkv, (bf16*)out, 512, row_off, row_len, items, n_items,
                             0.125f * 1.4426950408889634f, chunk, attention_trace_buffer()));
  JV_LAUNCHED();
}

}  // namespace jv
