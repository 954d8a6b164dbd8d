// Host side of the tcgen05 GEMM: tensor-map encoding (driver entry point fetched at run time, so the
// library does not link libcuda), tile-shape choice and launch.
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

const CUtensorMap& TmapCache::get(const void* ptr, long inner_elems, long rows, long pitch_elems, int box_rows) {
  TmapKey key{ptr, inner_elems, rows, pitch_elems, box_rows};
  auto it = maps.find(key);
  if (it != maps.end()) return it->second;
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)inner_elems, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)tc::BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_tiled()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  JV_REQUIRE(r == CUDA_SUCCESS, JV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%ld rows=%ld pitch=%ld box_rows=%d",
             (int)r, ptr, inner_elems, rows, pitch_elems, box_rows);
  return maps.emplace(key, m).first->second;
}

ProfileState& profile_state() {
  static ProfileState s;
  return s;
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
  if (g.bias && !aligned16(g.bias)) return false;
  if (g.resid && (!aligned16(g.resid) || g.ldr % 4 != 0)) return false;
  if (g.out_f32 && (!aligned16(g.out_f32) || g.ldo % 4 != 0)) return false;
  if (g.out_act && (!aligned16(g.out_act) || g.ldo2 % 8 != 0)) return false;
  return true;
}

void launch_gemm_tc(const GemmDesc& g, TmapCache& cache, int num_sms, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return;
  static bool attr_set = false;
  if (!attr_set) {
    JV_CUDA(cudaFuncSetAttribute(tc::gemm_taps_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  int block_n = g.N <= 256 ? round_up(g.N, 32) : 256;
  const int n_tiles_n = cdiv(g.N, block_n);
  const int m_tiles = cdiv(g.M, tc::BLOCK_M);
  const int num_tiles = m_tiles * n_tiles_n;
  const long Ktot = (long)g.n_taps * g.K_tap;
  if (cache.maps.size() > 4096) cache.maps.clear();  // before the gets: references must stay valid below
  const CUtensorMap& tA0 = cache.get(g.A[0], g.K_tap, g.a_rows[0], g.lda[0], tc::BLOCK_M);
  const CUtensorMap& tA1 = g.A[1] ? cache.get(g.A[1], g.K_tap, g.a_rows[1], g.lda[1], tc::BLOCK_M) : tA0;
  const CUtensorMap& tW = cache.get(g.W, Ktot, g.N, Ktot, block_n);
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  ProfileState& ps = profile_state();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ps.on) {
    JV_CUDA(cudaEventCreate(&e0));
    JV_CUDA(cudaEventCreate(&e1));
    JV_CUDA(cudaEventRecord(e0, st));
  }
  tc::gemm_taps_tc_kernel<<<grid, tc::NUM_THREADS, tc::SMEM_BYTES, st>>>(tA0, tA1, tW, g, block_n, n_tiles_n, num_tiles);
  JV_LAUNCHED();
  if (ps.on) {
    JV_CUDA(cudaEventRecord(e1, st));
    ps.ev.push_back(e0);
    ps.ev.push_back(e1);
    ps.flops += g.algo_flops;
  }
}

}  // namespace jv
