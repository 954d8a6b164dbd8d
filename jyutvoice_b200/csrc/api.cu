// Library-wide C-ABI pieces: error text, launch counter, version, GEMM test hook.
#include "engine.cuh"

namespace jv {
static thread_local std::string g_last_error;
std::atomic<uint64_t> g_launch_count{0};
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
    bf16 *Ab = nullptr, *Wb = nullptr;
    if (precision == JV_PREC_BF16) {
      JV_CUDA(cudaMalloc(&Ab, (size_t)M * K * 2));
      JV_CUDA(cudaMalloc(&Wb, (size_t)N * K * 2));
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
    if (Ab) cudaFree(Ab);
    if (Wb) cudaFree(Wb);
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
