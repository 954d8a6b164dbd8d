// Library-wide C-ABI pieces: error text, launch counter, version, GEMM test hook.
#include "engine.cuh"
#include "weights.cuh"

namespace jv {
static thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
std::atomic<uint64_t> g_graph_launches{0};
std::atomic<uint64_t> g_simt_fallbacks{0};
void set_last_error(const std::string& msg) { g_last_error = msg; }

__global__ void cvt_f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2bfloat16_rn(in[i]);
}
}  // namespace jv

using namespace jv;

extern "C" {

int jv_version(void) { return 1; }
const char* jv_last_error(void) { return g_last_error.c_str(); }
uint64_t jv_launch_count(void) { return g_launch_count.load(); }
uint64_t jv_graph_launch_count(void) { return g_graph_launches.load(); }
uint64_t jv_simt_fallback_count(void) { return g_simt_fallbacks.load(); }

int jv_profile_begin(void) {
  JV_API_BEGIN
  ProfileState& ps = profile_state();
  for (cudaEvent_t e : ps.ev) cudaEventDestroy(e);
  ps.ev.clear();
  ps.flops = 0.0;
  ps.on = true;
  JV_API_END
}

int jv_profile_end(double* kernel_ms, double* algo_flops, int64_t* launches) {
  JV_API_BEGIN
  ProfileState& ps = profile_state();
  ps.on = false;
  JV_CUDA(cudaDeviceSynchronize());
  double ms = 0.0;
  for (size_t i = 0; i + 1 < ps.ev.size(); i += 2) {
    float t = 0.f;
    JV_CUDA(cudaEventElapsedTime(&t, ps.ev[i], ps.ev[i + 1]));
    ms += t;
  }
  if (kernel_ms) *kernel_ms = ms;
  if (algo_flops) *algo_flops = ps.flops;
  if (launches) *launches = (int64_t)(ps.ev.size() / 2);
  for (cudaEvent_t e : ps.ev) cudaEventDestroy(e);
  ps.ev.clear();
  JV_API_END
}

int jv_bench_gemm(int M, int N, int K_tap, int taps, int mode, int iters, double* ms_out) {
  JV_API_BEGIN
  JV_REQUIRE(M > 0 && N > 0 && K_tap > 0 && taps > 0 && taps <= MAX_TAPS && iters > 0 && ms_out, JV_ERR_INVALID, "bad arguments");
  int dev = 0;
  JV_CUDA(cudaGetDevice(&dev));
  Engine eng;
  eng.init(dev, JV_PREC_BF16);
  const int Mp = round_up(M, 128);
  DeviceAlloc mem;  // frees on every exit path, error throws included
  void* A = mem.raw((size_t)Mp * K_tap * 2);
  void* W = mem.raw((size_t)N * K_tap * taps * 2);
  void* OA = mem.raw((size_t)Mp * N * 2);
  void* OL = mem.raw((size_t)Mp * N * 2);
  float* R = (float*)mem.raw((size_t)Mp * N * 4);
  float* OF = (float*)mem.raw((size_t)Mp * N * 4);
  float* vec = (float*)mem.raw((size_t)N * 4);
  JV_CUDA(cudaMemset(A, 0, (size_t)Mp * K_tap * 2));
  JV_CUDA(cudaMemset(W, 0, (size_t)N * K_tap * taps * 2));
  JV_CUDA(cudaMemset(R, 0, (size_t)Mp * N * 4));
  JV_CUDA(cudaMemset(vec, 0, (size_t)N * 4));
  GemmDesc g = gemm_desc_default();
  g.A[0] = A; g.lda[0] = K_tap; g.a_rows[0] = Mp;
  g.n_taps = taps; g.K_tap = K_tap;
  for (int t = 0; t < taps; ++t) g.tap_shift[t] = t - (taps - 1);
  g.W = W; g.M = M; g.N = N; g.bias = vec; g.o_rows = Mp;
  if (mode & 1) { g.resid = R; g.ldr = N; g.out_f32 = OF; g.ldo = N; }
  if (mode & 2) g.act = ACT_GELU;
  if (mode & 4) { g.ln1_gamma = vec; g.ln1_beta = vec; g.act = ACT_MISH; }
  if (mode & 8) { g.ln2_gamma = vec; g.ln2_beta = vec; g.out_ln = OL; g.ldo3 = N; }
  if (mode & 16) { g.out_act = OA; g.ldo2 = N; }
  if (mode & 32) g.x_bf16 = 1;  // R / OF buffers are simply over-allocated
  JV_REQUIRE(gemm_tc_supported(g), JV_ERR_INVALID, "shape not supported by the tcgen05 engine");
  struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } ev;
  JV_CUDA(cudaEventCreate(&ev.a));
  JV_CUDA(cudaEventCreate(&ev.b));
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  for (int i = 0; i < 3; ++i) eng.gemm(g, 0);
  JV_CUDA(cudaEventRecord(e0, 0));
  for (int i = 0; i < iters; ++i) eng.gemm(g, 0);
  JV_CUDA(cudaEventRecord(e1, 0));
  JV_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  JV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_out = ms / iters;
  JV_API_END
}

int jv_test_gemm(int precision, int M, int N, int K, const float* A, const float* W, const float* bias, float* C, void* stream) {
  try {
    JV_REQUIRE(M > 0 && N > 0 && K > 0 && A && W && C, JV_ERR_INVALID, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    JV_CUDA(cudaGetDevice(&dev));
    Engine eng;
    eng.init(dev, precision);
    GemmDesc g = gemm_desc_default();
    g.lda[0] = K;
    g.a_rows[0] = M;
    g.n_taps = 1;
    g.K_tap = K;
    g.M = M;
    g.N = N;
    g.bias = bias;
    g.out_f32 = C;
    g.ldo = N;
    g.o_rows = M;
    DeviceAlloc mem;
    bf16 *Ab = nullptr, *Wb = nullptr;
    if (precision == JV_PREC_BF16) {
      Ab = (bf16*)mem.raw((size_t)M * K * 2);
      Wb = (bf16*)mem.raw((size_t)N * K * 2);
      cvt_f32_to_bf16_kernel<<<(unsigned)(((long)M * K + 255) / 256), 256, 0, st>>>(A, Ab, (long)M * K);
      JV_LAUNCHED();
      cvt_f32_to_bf16_kernel<<<(unsigned)(((long)N * K + 255) / 256), 256, 0, st>>>(W, Wb, (long)N * K);
      JV_LAUNCHED();
      g.A[0] = Ab;
      g.W = Wb;
      JV_REQUIRE(gemm_tc_supported(g), JV_ERR_INVALID, "shape not supported by the tcgen05 engine (need K %% 64 == 0, N %% 8 == 0)");
    } else {
      g.A[0] = A;
      g.W = W;
    }
    eng.gemm(g, st);
    JV_CUDA(cudaStreamSynchronize(st));
  } catch (const jv::Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return JV_ERR_CUDA;
  }
  return JV_OK;
}

}  // extern "C"
