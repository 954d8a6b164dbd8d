// Shared host/device helpers for the jyutvoice_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <stdexcept>

#include "../../include/jyutvoice_b200.h"

namespace jv {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- errors
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;
extern std::atomic<uint64_t> g_graph_launches;  // CUDA-graph replays of an Euler step (jv_cfm_solve)
extern std::atomic<uint64_t> g_simt_fallbacks;  // bf16-mode GEMMs that were lowered to the FFMA engine (unsupported tcgen05 shape)

#define JV_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      char _buf[512];                                                                              \
      snprintf(_buf, sizeof(_buf), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw ::jv::Error(JV_ERR_CUDA, _buf);                                                        \
    }                                                                                              \
  } while (0)

#define JV_REQUIRE(cond, code, ...)                                                                \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      char _buf[512];                                                                              \
      snprintf(_buf, sizeof(_buf), __VA_ARGS__);                                                   \
      throw ::jv::Error(code, _buf);                                                               \
    }                                                                                              \
  } while (0)

// Count + check a kernel launch (the check is cudaPeekAtLastError: no sync).
#define JV_LAUNCHED()                                                                              \
  do {                                                                                             \
    ::jv::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                  \
    JV_CUDA(cudaPeekAtLastError());                                                                \
  } while (0)

// ---------------------------------------------------------------- workspace arena (caller-owned memory)
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  bool dry;  // dry run: only measure
  Arena(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0), dry(p == nullptr) {}
  template <typename T>
  T* alloc(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    size_t o = off;
    off += bytes;
    if (dry) return nullptr;
    JV_REQUIRE(off <= cap, JV_ERR_STATE, "workspace too small: need >= %zu bytes, have %zu", off, cap);
    return (T*)(base + o);
  }
};

// cudaFuncSetAttribute is per device: true the first time `site` is asked about the CURRENT device (handles on several
// devices may live in one process; a process-wide "done" flag would leave the second device without its smem opt-in)
static inline bool first_use_on_device(unsigned long long& site) {
  static std::mutex mu;  // handles on several host threads may reach a site at the same time
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  if (site & (1ull << dev)) return false;
  site |= 1ull << dev;
  return true;
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return cdiv(a, b) * b; }

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct DT;
template <> struct DT<float> {
  static constexpr int id = 0;
  __device__ __forceinline__ static float to_f(float v) { return v; }
  __device__ __forceinline__ static float from_f(float v) { return v; }
};
template <> struct DT<bf16> {
  static constexpr int id = 1;
  __device__ __forceinline__ static float to_f(bf16 v) { return __bfloat162float(v); }
  __device__ __forceinline__ static bf16 from_f(float v) { return __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------- activations (epilogue codes)
enum Act : int {
  ACT_NONE = 0,
  ACT_GELU = 1,    // exact erf GELU (diffusers GELU, approximate="none")
  ACT_ELU = 2,     // alpha = 1
  ACT_LRELU = 3,   // slope = act_param
  ACT_SNAKE = 4,   // x + sin^2(a x)/(a + 1e-9), a = act_vec[n]
  ACT_MISH = 5,
  ACT_SILU = 6,
};

__device__ __forceinline__ float act_mish(float x) {
  // x * tanh(softplus(x)); softplus with torch's threshold 20
  float sp = x > 20.f ? x : log1pf(expf(x));
  return x * tanhf(sp);
}

__device__ __forceinline__ float apply_act(float v, int act, float p, float a) {
  switch (act) {
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_ELU: return v > 0.f ? v : expm1f(v);
    case ACT_LRELU: return v > 0.f ? v : v * p;
    case ACT_SNAKE: {
      float s = sinf(v * a);
      return v + (1.0f / (a + 1e-9f)) * s * s;
    }
    case ACT_MISH: return act_mish(v);
    case ACT_SILU: return v / (1.f + expf(-v));
    default: return v;
  }
}

// ---------------------------------------------------------------- fast epilogue math (bf16 mode only:
// results are rounded to bf16 or feed bf16 operands, so ~1e-6 absolute error is invisible)
// Single-MUFU primitives with flush-to-zero: the non-ftz forms (__expf, __fdividef, __sinf) add a denormal range check
// per element whose predicates serialise the 32 independent element chains of a chunk (measured: Mish 25 us of a 49 us conv).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sin_ftz(float x) {
  float y;
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float exp_fast(float x) { return ex2_ftz(x * 1.4426950408889634f); }
__device__ __forceinline__ float fast_erf(float x) {
  // Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7
  const float ax = fabsf(x);
  const float t = rcp_ftz(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * exp_fast(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float fast_mish(float x) {
  // x * tanh(softplus(x)) = x * (e^2 + 2e) / (e^2 + 2e + 2), e = exp(x)
  // branch-free: for x >= 20 the ratio rounds to 1 (e^2 ~ 2.4e17, no overflow), so clamping the exponent is exact
  const float e = exp_fast(fminf(x, 20.f));
  const float n = e * (e + 2.0f);
  return x * (n * rcp_ftz(n + 2.0f));
}
__device__ __forceinline__ float apply_act_fast(float v, int act, float p, float a) {
  switch (act) {
    case ACT_GELU: return 0.5f * v * (1.f + fast_erf(v * 0.70710678118654752440f));
    case ACT_ELU: return fmaxf(v, 0.f) + (exp_fast(fminf(v, 0.f)) - 1.0f);
    case ACT_LRELU: return fmaxf(v, 0.f) + p * fminf(v, 0.f);
    case ACT_SNAKE: {
      const float s = sin_ftz(v * a);
      return fmaf(rcp_ftz(a + 1e-9f) * s, s, v);
    }
    case ACT_MISH: return fast_mish(v);
    case ACT_SILU: return v * rcp_ftz(1.f + exp_fast(-v));
    default: return v;
  }
}

// ---------------------------------------------------------------- GEMM-with-taps problem description
// C[out_row(m), n] = epilogue( sum_s sum_k A_src(s)[m*a_stride + shift_s, k] * W[n, s*K_tap + k] )
// This one form covers Linear, (causal / dilated / strided) Conv1d and the polyphase ConvTranspose1d.
constexpr int MAX_TAPS = 32;

struct GemmDesc {
  // A operands: up to two sources (channel concat, e.g. the U-Net skip), element type = activation type
  const void* A[2];
  int lda[2];          // row pitch in elements
  long a_rows[2];      // rows outside [0, a_rows) read as zero
  int a_stride;        // a_row = m * a_stride + shift
  int n_taps;
  int tap_src[MAX_TAPS];
  int tap_shift[MAX_TAPS];
  int K_tap;           // K per tap (same for all taps)
  const void* W;       // [N, n_taps*K_tap] K-major, activation type
  const float* W_hi;   // fp32 engines only: W split for 3xTF32 (W_hi = W with the low 13 mantissa bits cleared, W_lo = W - W_hi);
  const float* W_lo;   //   null -> the FFMA engine runs the contraction
  int M, N;
  // epilogue, in this order (every step optional):
  //   v = acc + bias[n]
  //   v = LN1(v) over the N channels of the row (gamma/beta; needs N == 256)
  //   v = act(v)
  //   v += add_row[row_tidx[frame_row[row]] * add_row_stride + n]      (time embedding)
  //   v = frame_row[row] >= 0 ? v : 0                                   (padding mask)
  //   v += resid[row, n]                                                (fp32 residual / skip branch)
  //   out_f32[row, n] = v ; out_act[row, n] = act2(v) ; out_ln[row, n] = mask(LN2(v))
  const float* bias;   // [N] or null
  const float* ln1_gamma;
  const float* ln1_beta;
  int act;
  float act_param;
  const float* act_vec;
  const float* add_row;
  const int* row_tidx;
  int add_row_stride;
  const int* frame_row;  // [rows] >= 0 valid (value = utterance row), < 0 -> 0; null = all valid (indexed by out_row)
  const float* resid;    // fp32 [*, ldr] indexed by out_row, or null
  int ldr;
  float* out_f32;        // optional fp32 output [*, ldo]
  void* out_act;         // optional activation-typed output [*, ldo2]
  int ldo, ldo2;
  int act2;              // activation applied to the out_act copy only
  float act2_param;
  const float* act2_vec;
  const float* ln2_gamma;
  const float* ln2_beta;
  void* out_ln;          // activation-typed LN2 output [*, ldo3]
  int ldo3;
  int o_stride, o_off;   // out_row = m * o_stride + o_off
  long o_rows;           // out_row must be < o_rows
  int x_bf16;            // tcgen05 engine only: `resid` and `out_f32` point to 16-bit tensors (16-bit residual stream) ...
  int x_in_half, x_out_half;  // ... holding fp16 (11-bit significand, stores saturate) instead of bf16
  int* sat_flag;         // fp16 stream only: incremented when a value stored into the stream reached +-65504 (saturating convert)
  double algo_flops;     // algorithmic FLOPs of this launch (valid frames, true N and K); profiling only
};

static inline GemmDesc gemm_desc_default() {
  GemmDesc g;
  memset(&g, 0, sizeof(g));
  g.a_stride = 1;
  g.o_stride = 1;
  return g;
}

}  // namespace jv

// C-ABI wrappers: no exception crosses the boundary
#define JV_API_BEGIN try {
#define JV_API_END                                   \
  }                                                  \
  catch (const jv::Error& e) {                       \
    jv::set_last_error(e.what());                    \
    return e.code;                                   \
  }                                                  \
  catch (const std::exception& e) {                  \
    jv::set_last_error(e.what());                    \
    return JV_ERR_CUDA;                              \
  }                                                  \
  return JV_OK;
