// Unity translation unit: the whole library is one nvcc compilation (kernels live in headers).
#include "api.cu"
#include "gemm_tc.cu"
#include "estimator.cu"
#include "hift.cu"
#include "text.cu"
#include "flowenc.cu"
