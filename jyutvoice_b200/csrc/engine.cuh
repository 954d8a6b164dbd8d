// Precision-dispatching GEMM engine shared by the estimator and HiFT handles.
#pragma once
#include "common.cuh"
#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace jv {

struct Engine {
  int device = 0;
  int precision = JV_PREC_FP32;
  int num_sms = 148;
  TmapCache tmaps;
  bool is_bf16() const { return precision == JV_PREC_BF16; }
  size_t act_size() const { return is_bf16() ? 2 : 4; }

  void init(int dev, int prec) {
    JV_REQUIRE(prec == JV_PREC_FP32 || prec == JV_PREC_BF16, JV_ERR_INVALID, "unknown precision %d", prec);
    int count = 0;
    JV_CUDA(cudaGetDeviceCount(&count));
    JV_REQUIRE(dev >= 0 && dev < count, JV_ERR_INVALID, "device %d out of range (have %d)", dev, count);
    device = dev;
    precision = prec;
    JV_CUDA(cudaSetDevice(dev));
    cudaDeviceProp p;
    JV_CUDA(cudaGetDeviceProperties(&p, dev));
    JV_REQUIRE(p.major == 10, JV_ERR_CUDA, "jyutvoice_b200 needs an sm_100a device (B200); found sm_%d%d", p.major, p.minor);
    num_sms = p.multiProcessorCount;
  }

  void gemm(const GemmDesc& g, cudaStream_t st) {
    if (is_bf16()) {
      if (gemm_tc_supported(g)) launch_gemm_tc(g, tmaps, num_sms, st);
      else launch_gemm_simt<bf16>(g, st);
    } else {
      launch_gemm_simt<float>(g, st);
    }
  }
};

}  // namespace jv
