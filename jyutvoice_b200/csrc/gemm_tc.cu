// Host side of the tcgen05 GEMM: tensor-map encoding (driver entry point fetched at run time, so the
// library does not link libcuda), tile-shape / pipeline-depth choice and launch.
#include <stdlib.h>
#include <algorithm>

#include "gemm_tc.cuh"

namespace jv {

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    JV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    JV_REQUIRE(p != nullptr && qres == cudaDriverEntryPointSuccess, JV_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

const CUtensorMap& TmapCache::get(const void* ptr, long inner_elems, long rows, long pitch_bytes, int box_inner, int box_rows,
                                  int kind) {
  TmapKey key{ptr, inner_elems, rows, pitch_bytes, box_inner, box_rows, kind};
  auto it = maps.find(key);
  if (it != maps.end()) return it->second;
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)inner_elems, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = kind == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = get_encode_tiled()(&m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JV_REQUIRE(r == CUDA_SUCCESS, JV_ERR_CUDA,
             "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%ld rows=%ld pitch=%ld box=%dx%d kind=%d", (int)r, ptr, inner_elems, rows,
             pitch_bytes, box_inner, box_rows, kind);
  return maps.emplace(key, m).first->second;
}

ProfileState& profile_state() {
  static ProfileState s;
  return s;
}

// JYUTVOICE_B200_CLUSTER=n: CTAs per cluster for the weight multicast (0 / 1 disables it; debugging aid)
static int cluster_size() {  // CTAs per cluster for the weight multicast: 0 / 1 = off (default)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_CLUSTER");
    v = e ? atoi(e) : 1;  // measured on B200: 1 >= 2 >> 4 for every estimator shape (profiles/r1j_cluster_sweep.txt)
    if (v < 1) v = 1;
    if (v > 8) v = 8;
  }
  return v;
}

// n-tile width of the N > 256 kernels (QKV, FF1): 256 unless JYUTVOICE_B200_BLOCKN overrides it (128 / 192 / 256 and N % it == 0)
static int wide_block_n(int N) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_BLOCKN");
    v = e ? atoi(e) : 256;
    if (v != 128 && v != 192 && v != 256) v = 256;
  }
  return (N % v == 0) ? v : 256;
}

typedef void (*KernelFn)(const tc::TcMaps, const GemmDesc, const tc::TcParams);
struct KernelEntry {
  int epi;
  KernelFn fn;
  KernelFn fn_pair;  // cta_group::2 variant of the same epilogue
  KernelFn fn_wln = nullptr, fn_wln_pair = nullptr;  // sixteen-warp sixteen-warp LayerNorm epilogue (16-bit stream kernels)
};
constexpr int N_KERNELS = 12;
// the epilogue shapes the estimator / HiFT graphs actually use, plus the run-time generic kernel (last)
static const KernelEntry* kernel_table() {
  using namespace tc;
  static const KernelEntry t[N_KERNELS] = {
      {EPI_OACT, gemm_taps_tc_kernel<EPI_OACT>, gemm_taps_tc_kernel<EPI_OACT, true>},                                              // QKV, FF1, plain convs
      {EPI_RESID | EPI_F32 | EPI_LN2, gemm_taps_tc_kernel<EPI_RESID | EPI_F32 | EPI_LN2>, gemm_taps_tc_kernel<EPI_RESID | EPI_F32 | EPI_LN2, true>},    // out-proj, FF2 (+ next norm)
      {EPI_RESID | EPI_F32 | EPI_OACT, gemm_taps_tc_kernel<EPI_RESID | EPI_F32 | EPI_OACT>, gemm_taps_tc_kernel<EPI_RESID | EPI_F32 | EPI_OACT, true>},  // last FF2 of a group, HiFT conv2
      {EPI_LN1 | EPI_OACT, gemm_taps_tc_kernel<EPI_LN1 | EPI_OACT>, gemm_taps_tc_kernel<EPI_LN1 | EPI_OACT, true>,                           // CausalBlock1D (block1, final_block)
       gemm_taps_tc_kernel<EPI_LN1 | EPI_OACT, false, true>, gemm_taps_tc_kernel<EPI_LN1 | EPI_OACT, true, true>},
      {EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, gemm_taps_tc_kernel<EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2>, gemm_taps_tc_kernel<EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, true>},  // block2 + res + norm1
      {EPI_F32, gemm_taps_tc_kernel<EPI_F32>, gemm_taps_tc_kernel<EPI_F32, true>},                                                // res_conv, final_proj, conv_post
      {EPI_RESID | EPI_F32, gemm_taps_tc_kernel<EPI_RESID | EPI_F32>, gemm_taps_tc_kernel<EPI_RESID | EPI_F32, true>},                        // HiFT ups + source, last conv2
      {EPI_F32 | EPI_OACT, gemm_taps_tc_kernel<EPI_F32 | EPI_OACT>, gemm_taps_tc_kernel<EPI_F32 | EPI_OACT, true>},                          // HiFT source_downs (im2col)
      // bf16 residual stream (estimator, bf16 mode)
      {EPI_XB | EPI_RESID | EPI_F32 | EPI_LN2, gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32 | EPI_LN2>, gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32 | EPI_LN2, true>,
       gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32 | EPI_LN2, false, true>, gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32 | EPI_LN2, true, true>},
      {EPI_XB | EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, gemm_taps_tc_kernel<EPI_XB | EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2>, gemm_taps_tc_kernel<EPI_XB | EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, true>,
       gemm_taps_tc_kernel<EPI_XB | EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, false, true>, gemm_taps_tc_kernel<EPI_XB | EPI_LN1 | EPI_RESID | EPI_F32 | EPI_LN2, true, true>},
      {EPI_XB | EPI_RESID | EPI_F32, gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32>, gemm_taps_tc_kernel<EPI_XB | EPI_RESID | EPI_F32, true>},
      {-1, gemm_taps_tc_kernel<-1>, gemm_taps_tc_kernel<-1, true>},
  };
  return t;
}

bool use_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// JYUTVOICE_B200_EPI16=0: eight instead of sixteen epilogue warps for the residual + LayerNorm kernels of the 16-bit stream
static bool use_epi16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_EPI16");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// JYUTVOICE_B200_FORCE_MODES=1: use the weight-resident / slab variants even for small problems (parity tests)
static bool force_modes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_FORCE_MODES");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

static bool use_tile_par() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_TILEPAR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool use_slab() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_SLAB");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool use_wres() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_WRES");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// JYUTVOICE_B200_PAIR=0: no cta_group::2 CTA pairs (every MMA is cta_group::1, M = 128)
static bool use_pair() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

bool gemm_tc_supported(const GemmDesc& g) {
  if (g.a_stride != 1) return false;
  if (g.K_tap % tc::BLOCK_K != 0) return false;
  if (g.N % 8 != 0) return false;
  if (g.n_taps < 1 || g.n_taps > MAX_TAPS) return false;
  for (int s = 0; s < 2; ++s) {
    if (!g.A[s]) continue;
    if (!aligned16(g.A[s]) || g.lda[s] % 8 != 0 || g.lda[s] < g.K_tap) return false;
  }
  if (!aligned16(g.W)) return false;
  if (g.resid && (!aligned16(g.resid) || g.ldr % (g.x_bf16 ? 8 : 4) != 0)) return false;
  if (g.out_f32 && (!aligned16(g.out_f32) || g.ldo % (g.x_bf16 ? 8 : 4) != 0)) return false;
  if (g.x_bf16 && (!g.resid || !g.out_f32 || g.out_act)) return false;  // bf16 stream: residual in, stream out (+ LayerNorm out)
  if (g.out_act && (!aligned16(g.out_act) || g.ldo2 % 8 != 0)) return false;
  if (g.out_ln && (!aligned16(g.out_ln) || g.ldo3 % 8 != 0)) return false;
  if ((g.ln1_gamma || g.ln2_gamma || g.add_row) && g.N != 256) return false;
  if (g.ln2_gamma && !g.out_ln) return false;
  return true;
}

// view of an output-like tensor through (o_stride, o_off): row i of the view = tensor row i*o_stride + o_off
static const CUtensorMap& out_view(TmapCache& cache, const GemmDesc& g, const void* ptr, int ld, int esize, int kind) {
  const long view_rows = (g.o_rows - g.o_off + g.o_stride - 1) / g.o_stride;
  const char* base = (const char*)ptr + (long)g.o_off * ld * esize;
  return cache.get(base, g.N, view_rows, (long)g.o_stride * ld * esize, 32, 32, kind);
}

void launch_gemm_tc(const GemmDesc& g, TmapCache& cache, int num_sms, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(attr_set)) {
    for (int i = 0; i < N_KERNELS; ++i) {
      JV_CUDA(cudaFuncSetAttribute((const void*)kernel_table()[i].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
      JV_CUDA(cudaFuncSetAttribute((const void*)kernel_table()[i].fn_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
      if (kernel_table()[i].fn_wln) {
        JV_CUDA(cudaFuncSetAttribute((const void*)kernel_table()[i].fn_wln, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
        JV_CUDA(cudaFuncSetAttribute((const void*)kernel_table()[i].fn_wln_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_LIMIT));
      }
    }
  }
  tc::TcParams p;
  p.block_n = g.N <= 256 ? round_up(g.N, 32) : wide_block_n(g.N);
  p.n_tiles_n = cdiv(g.N, p.block_n);
  const int m_tiles = cdiv(g.M, tc::BLOCK_M);
  p.num_tiles = m_tiles * p.n_tiles_n;
  p.b_stage_bytes = round_up(p.block_n * tc::BLOCK_K * 2, 1024);
  const long Ktot = (long)g.n_taps * g.K_tap;
  const int k_iters = (int)(Ktot / tc::BLOCK_K);
  const int epi = (g.ln1_gamma ? tc::EPI_LN1 : 0) | (g.resid ? tc::EPI_RESID : 0) | (g.out_f32 ? tc::EPI_F32 : 0) |
                  (g.out_act ? tc::EPI_OACT : 0) | (g.ln2_gamma ? tc::EPI_LN2 : 0) | (g.x_bf16 ? tc::EPI_XB : 0);
  // Slab mode (stride-1 conv, one source, resident weights): see TcParams.
  int min_shift = 0, max_shift = 0;
  bool one_src = true;
  for (int t = 0; t < g.n_taps; ++t) {
    min_shift = t == 0 ? g.tap_shift[t] : std::min(min_shift, g.tap_shift[t]);
    max_shift = t == 0 ? g.tap_shift[t] : std::max(max_shift, g.tap_shift[t]);
    if (g.tap_src[t] != 0) one_src = false;
  }
  const int slab_rows = round_up(tc::BLOCK_M + (max_shift - min_shift), 8);
  const bool w_fits = k_iters * p.b_stage_bytes <= 131072;
  // ... and only if at least two slab stages fit next to the resident weights and the epilogue staging of this kernel
  const int staging_guess = (epi == tc::EPI_OACT) ? tc::EPI_WARPS_MAX * tc::EPI_B16_BYTES : tc::EPI_WARPS * tc::EPI_BYTES_PER_WARP;
  const bool slab_fits = tc::SMEM_LIMIT - (1024 + staging_guess + tc::BAR_BYTES + tc::XCH_BYTES + tc::TAB_BYTES) - k_iters * p.b_stage_bytes >=
                         2 * round_up(slab_rows * tc::BLOCK_K * 2, 1024);
  p.slab = (g.n_taps >= 2 && one_src && w_fits && slab_fits && p.n_tiles_n == 1 && slab_rows <= 256 &&
            (m_tiles >= 4 * num_sms || force_modes()) && use_slab())
               ? 1
               : 0;
  p.slab_rows = slab_rows;
  p.min_shift = min_shift;
  p.a_stage_bytes = p.slab ? round_up(slab_rows * tc::BLOCK_K * 2, 1024) : tc::A_STAGE_BYTES;
  // weight-resident mode: whole weight tile <= 128 KB, bf16-only epilogue (or slab mode), enough m-tiles per CTA
  const int ctas_per_ntile = num_sms / p.n_tiles_n;
  p.wres = (p.slab || (epi == tc::EPI_OACT && w_fits && ctas_per_ntile >= 1 && (m_tiles >= 4 * ctas_per_ntile || force_modes()) && use_wres())) ? 1 : 0;
  // CTA pair (cta_group::2, M = 256 over two SMs): for the kernels that stream their weights (not weight-resident, not
  // slab): each CTA of the pair loads half of the weight tile.  N per instruction must be a multiple of 16.
  p.pair = (!p.wres && !p.slab && (p.block_n == 256 || p.block_n == 128 || p.block_n == 64) && m_tiles >= 2 && num_sms % 2 == 0 &&
            cluster_size() == 1 && use_pair())
               ? 1
               : 0;
  if (p.pair) p.b_stage_bytes = round_up((p.block_n / 2) * tc::BLOCK_K * 2, 1024);
  const bool epi16 = ((g.x_bf16 && g.resid && g.ln2_gamma) || (epi == (tc::EPI_LN1 | tc::EPI_OACT) && !p.wres && !p.slab)) && g.N == 256 &&
                     p.block_n == 256 && use_epi16();
  const bool wide = epi == tc::EPI_OACT || epi16;  // 16 epilogue warps, small staging: bf16-only epilogues and the sixteen-warp LayerNorm epilogue
  const int n_epi_warps = wide ? tc::EPI_WARPS_MAX : tc::EPI_WARPS;
  {  // staging layout per epilogue warp, exactly what this epilogue kind needs
    const bool f_resid = epi & tc::EPI_RESID, f_f32 = epi & tc::EPI_F32, f_oact = epi & tc::EPI_OACT;
    int off = 0;
    p.off_R = 0;
    p.epi16 = epi16 ? 1 : 0;
    if (p.epi16) {  // sixteen warps: two 2 KB buffers (residual in / stream out / LayerNorm out, or two alternating bf16 outputs)
      p.off_OF = 0;
      p.off_OB = tc::EPI_B16_BYTES;
      off = 2 * tc::EPI_B16_BYTES;
    } else if (g.x_bf16) {  // R0 R1 | O0 | O1, 2 KB each
      p.off_OF = 2 * tc::EPI_B16_BYTES;
      p.off_OB = 3 * tc::EPI_B16_BYTES;
      off = 4 * tc::EPI_B16_BYTES;
    } else {
    if (f_resid) off += tc::EPI_F32_BYTES;
    p.off_OF = off;
    if (f_f32) {
      off += tc::EPI_F32_BYTES;
      p.off_OB = f_oact ? off : p.off_R;       // bf16 copy next to the fp32 buffer, or (LN2 only) alternate with R
      if (f_oact) off += tc::EPI_B16_BYTES;
      if (!f_oact && !f_resid) { p.off_OB = off; off += tc::EPI_B16_BYTES; }  // fp32-only kernel with an LN2 output (unused today)
    } else {                                   // bf16 outputs only: two buffers (one in weight-resident wide kernels: smem is tight)
      off += tc::EPI_B16_BYTES;
      if (wide && p.wres) p.off_OB = p.off_OF;
      else { p.off_OB = off; off += tc::EPI_B16_BYTES; }
    }
    }
    p.epi_bytes_per_warp = round_up(off, 1024);
  }
  // per-column vector cache (only when the CTA's n-tile is fixed), in priority order while smem remains
  const bool n_fixed = p.n_tiles_n == 1 || p.wres;
  const int vbytes = p.block_n * 4;
  int want_bias = (n_fixed && g.bias) ? vbytes : 0, want_ln1 = (n_fixed && g.ln1_gamma) ? 2 * vbytes : 0,
      want_ln2 = (n_fixed && g.ln2_gamma) ? 2 * vbytes : 0, want_act = (n_fixed && g.act_vec) ? vbytes : 0,
      want_act2 = (n_fixed && g.act2_vec) ? vbytes : 0;
  {
    const int n_sub = n_epi_warps / 4;
    const int n_chunks = p.block_n / 32;
    p.acc_cols = p.block_n;
    p.n_acc = std::min(tc::MAX_ACC, tc::TMEM_COLS / p.block_n);
    // tile-parallel epilogue for narrow tiles without LayerNorm (row statistics need the chunk split)
    p.tile_par = (n_chunks < 2 * n_sub && !g.ln1_gamma && !g.ln2_gamma && p.n_acc >= n_sub && use_tile_par()) ? 1 : 0;
    if (!p.tile_par) p.n_acc = 2;
  }
  const int bar_bytes = tc::BAR_BYTES + ((g.ln1_gamma || g.ln2_gamma) ? tc::XCH_BYTES : 0) + (p.slab ? tc::TAB_BYTES : 0);
  JV_REQUIRE(!p.slab || g.n_taps * (g.K_tap / tc::BLOCK_K) <= 96, JV_ERR_INVALID, "too many (tap, K block) pairs for slab mode");
  const int fixed_novec = 1024 + n_epi_warps * p.epi_bytes_per_warp + bar_bytes;
  const int ring = p.wres ? p.a_stage_bytes : p.a_stage_bytes + p.b_stage_bytes;
  {
    // smem left after the pipeline the kernel would get WITHOUT any vector cache (never trade a pipeline stage for it)
    int avail = tc::SMEM_LIMIT - fixed_novec - (p.wres ? k_iters * p.b_stage_bytes : 0);
    int st = avail / ring;
    if (st > tc::MAX_STAGES) st = tc::MAX_STAGES;
    int spare = avail - st * ring;
    auto take = [&](int& w) { if (w <= spare) spare -= w; else w = 0; };
    take(want_bias); take(want_ln1); take(want_ln2); take(want_act); take(want_act2);
  }
  const int vec_total = want_bias + want_ln1 + want_ln2 + want_act + want_act2;
  const int fixed = fixed_novec + vec_total;
  p.b_region_bytes = p.wres ? k_iters * p.b_stage_bytes : 0;
  p.stages = (tc::SMEM_LIMIT - fixed - p.b_region_bytes) / ring;
  if (p.stages > tc::MAX_STAGES) p.stages = tc::MAX_STAGES;
  JV_REQUIRE(p.stages >= 2, JV_ERR_STATE, "not enough shared memory for the GEMM pipeline (N=%d K_tap=%d taps=%d epi=%d slab=%d wres=%d fixed=%d bregion=%d ring=%d)", g.N, g.K_tap, g.n_taps, epi, p.slab, p.wres, fixed, p.b_region_bytes, ring);
  if (!p.wres) p.b_region_bytes = p.stages * p.b_stage_bytes;
  const int smem = fixed + p.stages * p.a_stage_bytes + p.b_region_bytes;
  {  // vector cache sits after the barrier block: offsets from the aligned base
    p.tab_off = p.stages * p.a_stage_bytes + p.b_region_bytes + n_epi_warps * p.epi_bytes_per_warp + bar_bytes - tc::TAB_BYTES;
    int off = p.stages * p.a_stage_bytes + p.b_region_bytes + n_epi_warps * p.epi_bytes_per_warp + bar_bytes;
    auto place = [&](int want) { int o = want ? off : 0; off += want; return o; };
    p.vec_bias = place(want_bias);
    p.vec_ln1 = place(want_ln1);
    p.vec_ln2 = place(want_ln2);
    p.vec_act = place(want_act);
    p.vec_act2 = place(want_act2);
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("JYUTVOICE_B200_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  p.trace = (gemm_trace().buf && gemm_trace().epi == epi) ? gemm_trace().buf : nullptr;
  if (cache.maps.size() > 4096) cache.maps.clear();  // before the gets: references must stay valid below
  tc::TcMaps tm;
  tm.a0 = cache.get(g.A[0], g.K_tap, g.a_rows[0], (long)g.lda[0] * 2, tc::BLOCK_K, p.slab ? p.slab_rows : tc::BLOCK_M, 0);
  tm.a1 = g.A[1] ? cache.get(g.A[1], g.K_tap, g.a_rows[1], (long)g.lda[1] * 2, tc::BLOCK_K, tc::BLOCK_M, 0) : tm.a0;
  p.cluster = p.pair ? 2 : 1;
  for (int cs = cluster_size(); cs >= 2 && !p.pair; cs >>= 1)
    if (!p.wres && m_tiles >= cs && p.block_n % (8 * cs) == 0 && num_sms % cs == 0) { p.cluster = cs; break; }
  p.num_units = p.wres ? m_tiles : cdiv(m_tiles, p.cluster) * p.n_tiles_n;
  tm.w = cache.get(g.W, Ktot, g.N, Ktot * 2, tc::BLOCK_K, p.block_n / p.cluster, 0);
  tm.resid = g.resid ? out_view(cache, g, g.resid, g.ldr, g.x_bf16 ? 2 : 4, g.x_bf16 ? 2 : 1) : tm.a0;
  tm.out_f32 = g.out_f32 ? out_view(cache, g, g.out_f32, g.ldo, g.x_bf16 ? 2 : 4, g.x_bf16 ? 2 : 1) : tm.a0;
  tm.out_act = g.out_act ? out_view(cache, g, g.out_act, g.ldo2, 2, 2) : tm.a0;
  tm.out_ln = g.out_ln ? out_view(cache, g, g.out_ln, g.ldo3, 2, 2) : tm.a0;
  int grid = p.num_units * p.cluster < num_sms ? p.num_units * p.cluster : num_sms;
  grid -= grid % p.cluster;
  if (p.wres) grid = ctas_per_ntile * p.n_tiles_n;
  ProfileState& ps = profile_state();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ps.on) {
    JV_CUDA(cudaEventCreate(&e0));
    JV_CUDA(cudaEventCreate(&e1));
    JV_CUDA(cudaEventRecord(e0, st));
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(wide ? tc::NUM_THREADS_WIDE : tc::NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see pdl_wait() in the kernel
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  const KernelEntry* ke = &kernel_table()[N_KERNELS - 1];  // generic
  bool found = false;
  for (int i = 0; i < N_KERNELS - 1; ++i)
    if (kernel_table()[i].epi == epi) { ke = &kernel_table()[i]; found = true; }
  KernelFn fn = p.epi16 ? (p.pair ? ke->fn_wln_pair : ke->fn_wln) : (p.pair ? ke->fn_pair : ke->fn);
  JV_REQUIRE(fn != nullptr, JV_ERR_STATE, "no kernel for epilogue kind %d", epi);
  JV_REQUIRE(found || !g.x_bf16, JV_ERR_INVALID, "no bf16-stream kernel for epilogue kind %d", epi);
  JV_CUDA(cudaLaunchKernelEx(&cfg, fn, tm, g, p));
  JV_LAUNCHED();
  if (ps.on) {
    JV_CUDA(cudaEventRecord(e1, st));
    ps.ev.push_back(e0);
    ps.ev.push_back(e1);
    ps.flops += g.algo_flops;
  }
}

}  // namespace jv
