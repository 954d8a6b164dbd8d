// Host-side weight store and packers shared by the estimator and HiFT handles.
// set_weight() keeps an fp32 host copy per reference state_dict key; finalize() packs them into the
// K-major [N, taps*K_tap] slabs the GEMM engines read (fp32 or bf16) and uploads them once.
#pragma once
#include <set>
#include "common.cuh"

namespace jv {

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  size_t numel() const { return data.size(); }
};

struct WeightStore {
  std::map<std::string, HostTensor> t;
  void set(const char* key, const float* data, const int64_t* shape, int ndim) {
    JV_REQUIRE(key && data && shape && ndim >= 1 && ndim <= 4, JV_ERR_INVALID, "set_weight: bad arguments");
    HostTensor h;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
      JV_REQUIRE(shape[i] > 0, JV_ERR_INVALID, "set_weight(%s): non-positive dim", key);
      h.shape.push_back(shape[i]);
      n *= (size_t)shape[i];
    }
    h.data.resize(n);
    JV_CUDA(cudaMemcpy(h.data.data(), data, n * sizeof(float), cudaMemcpyDefault));  // host or device source
    t[key] = std::move(h);
  }
  mutable std::set<std::string> used;  // keys the architecture asked for (finalize: every stored key must be one of them)
  // load_state_dict(strict=True) semantics: a key the architecture never reads is an error, not silently ignored
  void require_all_used() const {
    for (const auto& kv : t) JV_REQUIRE(used.count(kv.first) != 0, JV_ERR_INVALID, "unexpected weight '%s'", kv.first.c_str());
  }
  const HostTensor& get(const std::string& key, std::initializer_list<int64_t> shape) const {
    auto it = t.find(key);
    JV_REQUIRE(it != t.end(), JV_ERR_STATE, "missing weight '%s'", key.c_str());
    used.insert(key);
    const HostTensor& h = it->second;
    bool ok = h.shape.size() == shape.size();
    size_t i = 0;
    for (int64_t s : shape) {
      if (ok && h.shape[i] != s) ok = false;
      ++i;
    }
    JV_REQUIRE(ok, JV_ERR_INVALID, "weight '%s' has the wrong shape", key.c_str());
    return h;
  }
  bool has(const std::string& key) const { return t.count(key) != 0; }
};

static inline uint16_t f32_to_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                            // round to nearest even
  return (uint16_t)(u >> 16);
}

// Device allocation owned by a handle.
struct DeviceAlloc {
  std::vector<void*> ptrs;
  size_t bytes = 0;
  ~DeviceAlloc() {
    for (void* p : ptrs) cudaFree(p);
  }
  void* raw(size_t nbytes) {
    void* p = nullptr;
    JV_CUDA(cudaMalloc(&p, nbytes ? nbytes : 16));
    ptrs.push_back(p);
    bytes += nbytes;
    return p;
  }
  float* upload_f32(const std::vector<float>& v) {
    float* p = (float*)raw(v.size() * sizeof(float));
    JV_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return p;
  }
  // 3xTF32 split of an fp32 weight: hi = w with the low 13 mantissa bits cleared (exactly a TF32 number), lo = w - hi
  void upload_tf32_split(const std::vector<float>& v, float** hi, float** lo) {
    std::vector<float> h(v.size()), l(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
      uint32_t u;
      memcpy(&u, &v[i], 4);
      u &= 0xffffe000u;
      memcpy(&h[i], &u, 4);
      l[i] = v[i] - h[i];
    }
    *hi = upload_f32(h);
    *lo = upload_f32(l);
  }
  // upload as activation type: fp32 or bf16
  void* upload_act(const std::vector<float>& v, bool as_bf16) {
    if (!as_bf16) return upload_f32(v);
    std::vector<uint16_t> h(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = f32_to_bf16_bits(v[i]);
    void* p = raw(h.size() * 2);
    JV_CUDA(cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    return p;
  }
};

// One packed GEMM weight: W[N_pad, n_taps*K_tap] (+ bias[N_pad]); N_pad rows beyond N are zero.
struct PackedW {
  void* W = nullptr;
  float* W_hi = nullptr;  // fp32 packs: the 3xTF32 split of W (gemm_tf32.cuh)
  float* W_lo = nullptr;
  float* bias = nullptr;
  int N = 0, N_pad = 0, K_tap = 0, n_taps = 0;
};

struct TapSrc {
  int k;      // kernel index into the conv weight's last dim
  int c_off;  // first input channel of this tap's slab
  int c_len;  // channels taken (<= K_tap; the rest of the slab is zero)
};

// Conv1d weight w[Cout, Cin, Kw] -> W[n, s*K_tap + c] = w[n, taps[s].c_off + c, taps[s].k]
static inline std::vector<float> pack_conv_taps(const float* w, int Cout, int Cin, int Kw, const std::vector<TapSrc>& taps,
                                                int K_tap, int N_pad) {
  std::vector<float> out((size_t)N_pad * taps.size() * K_tap, 0.f);
  const size_t Kt = taps.size() * (size_t)K_tap;
  for (int n = 0; n < Cout; ++n)
    for (size_t s = 0; s < taps.size(); ++s)
      for (int c = 0; c < taps[s].c_len; ++c)
        out[n * Kt + s * K_tap + c] = w[((size_t)n * Cin + taps[s].c_off + c) * Kw + taps[s].k];
  return out;
}

// ConvTranspose1d weight w[Cin, Cout, Kw] -> W[n=cout, s*K_tap + c=cin] = w[c, n, taps[s].k]
static inline std::vector<float> pack_convT_taps(const float* w, int Cin, int Cout, int Kw, const std::vector<TapSrc>& taps,
                                                 int K_tap, int N_pad) {
  std::vector<float> out((size_t)N_pad * taps.size() * K_tap, 0.f);
  const size_t Kt = taps.size() * (size_t)K_tap;
  for (int n = 0; n < Cout; ++n)
    for (size_t s = 0; s < taps.size(); ++s)
      for (int c = 0; c < taps[s].c_len; ++c)
        out[n * Kt + s * K_tap + c] = w[((size_t)(taps[s].c_off + c) * Cout + n) * Kw + taps[s].k];
  return out;
}

static inline std::vector<float> pad_vec(const float* b, int N, int N_pad) {
  std::vector<float> v((size_t)N_pad, 0.f);
  if (b)
    for (int i = 0; i < N; ++i) v[i] = b[i];
  return v;
}

}  // namespace jv
