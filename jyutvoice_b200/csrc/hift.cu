// HiFT vocoder on B200: F0 predictor, NSF source, STFT, conv_pre, 3 x (polyphase ConvTranspose + source
// fusion + 3 Snake ResBlocks), conv_post, exp/sin head and inverse STFT.
// Reference: jyutvoice/hifigan/generator.py:239-466, jyutvoice/hifigan/f0_predictor.py:19-55,
// jyutvoice/transformer/activation.py:73-84 (Snake).
#include <cmath>
#include <memory>
#include <algorithm>

#include "engine.cuh"
#include "hift_kernels.cuh"
#include "weights.cuh"

namespace jv {

struct ResBlockW {
  PackedW c1[3], c2[3];
  float* a1[3];
  float* a2[3];
  int C = 0, k = 0;
};
struct UpPhase {
  PackedW w;
  std::vector<int> shifts;
};
struct UpW {
  int u = 0, k = 0, pad = 0, cin = 0, cout = 0;
  std::vector<UpPhase> phases;
};

static const int UPS_U[3] = {8, 5, 3};
static const int UPS_K[3] = {16, 11, 7};
static const int STAGE_C[3] = {256, 128, 64};
static const int STAGE_RATE[3] = {8, 40, 120};
static const int SRC_U[3] = {15, 3, 1};
static const int RES_K[3] = {3, 7, 11};
static const int SRC_K[3] = {7, 7, 11};
static const int RES_D[3] = {1, 3, 5};

}  // namespace jv

using namespace jv;

struct jv_hift {
  Engine eng;
  WeightStore store;
  DeviceAlloc mem;
  bool finalized = false;
  PackedW f0conv[5];
  float *cls_w = nullptr, *cls_b = nullptr, *src_w = nullptr, *src_b = nullptr;
  PackedW conv_pre, conv_post;
  UpW ups[3];
  PackedW src_down[3];     // [N, taps*18] for the FFMA path (fp32 mode)
  PackedW src_down_im[3];  // [N, Kp] zero-padded rows for the im2col + tcgen05 path (bf16 mode)
  ResBlockW src_rb[3], rb[9];
  StftTables stft_tb;
  IstftTables istft_tb;
};

namespace jv {

// effective conv weight: folds torch weight-norm (g * v / ||v||, norm over all dims but 0) when present
static std::vector<float> effective_weight(const WeightStore& st, const std::string& name, std::initializer_list<int64_t> shape) {
  const bool wn_new = st.has(name + ".parametrizations.weight.original0");
  const bool wn_old = st.has(name + ".weight_g");
  if (!wn_new && !wn_old) return st.get(name + ".weight", shape).data;
  const int64_t d0 = *shape.begin();
  const HostTensor& g = st.get(name + (wn_new ? ".parametrizations.weight.original0" : ".weight_g"), {d0, 1, 1});
  const HostTensor& v = st.get(name + (wn_new ? ".parametrizations.weight.original1" : ".weight_v"), shape);
  std::vector<float> w(v.data.size());
  const size_t inner = v.data.size() / (size_t)d0;
  for (int64_t i = 0; i < d0; ++i) {
    double ss = 0.0;
    for (size_t j = 0; j < inner; ++j) ss += (double)v.data[i * inner + j] * v.data[i * inner + j];
    const float nrm = (float)std::sqrt(ss);
    for (size_t j = 0; j < inner; ++j) w[i * inner + j] = g.data[i] * v.data[i * inner + j] / nrm;
  }
  return w;
}

static PackedW hpack(jv_hift* h, std::vector<float>&& w, const std::vector<float>& bias, int N, int K_tap, int n_taps) {
  PackedW p;
  p.N = N;
  p.N_pad = N;
  p.K_tap = K_tap;
  p.n_taps = n_taps;
  p.W = h->mem.upload_act(w, h->eng.is_bf16());
  if (!h->eng.is_bf16()) h->mem.upload_tf32_split(w, &p.W_hi, &p.W_lo);
  p.bias = h->mem.upload_f32(bias);
  return p;
}

// Conv1d [Cout, Cin, Kw] -> taps k = 0..Kw-1, K_tap >= Cin (zero padded), N_pad >= Cout
static PackedW hpack_conv(jv_hift* h, const std::string& name, int Cout, int Cin, int Kw, int K_tap, int N_pad) {
  std::vector<float> w = effective_weight(h->store, name, {Cout, Cin, Kw});
  const HostTensor& b = h->store.get(name + ".bias", {Cout});
  std::vector<TapSrc> taps;
  for (int k = 0; k < Kw; ++k) taps.push_back({k, 0, Cin});
  return hpack(h, pack_conv_taps(w.data(), Cout, Cin, Kw, taps, K_tap, N_pad), pad_vec(b.data.data(), Cout, N_pad), N_pad, K_tap, Kw);
}

static ResBlockW hpack_resblock(jv_hift* h, const std::string& name, int Cn, int k) {
  ResBlockW r;
  r.C = Cn;
  r.k = k;
  for (int i = 0; i < 3; ++i) {
    r.c1[i] = hpack_conv(h, name + ".convs1." + std::to_string(i), Cn, Cn, k, Cn, Cn);
    r.c2[i] = hpack_conv(h, name + ".convs2." + std::to_string(i), Cn, Cn, k, Cn, Cn);
    r.a1[i] = h->mem.upload_f32(h->store.get(name + ".activations1." + std::to_string(i) + ".alpha", {Cn}).data);
    r.a2[i] = h->mem.upload_f32(h->store.get(name + ".activations2." + std::to_string(i) + ".alpha", {Cn}).data);
  }
  return r;
}

static void hift_finalize_impl(jv_hift* h) {
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "hift already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  const int f0_cin[5] = {80, 512, 512, 512, 512};
  for (int i = 0; i < 5; ++i)
    h->f0conv[i] = hpack_conv(h, "f0_predictor.condnet." + std::to_string(2 * i), 512, f0_cin[i], 3, i == 0 ? 128 : 512, 512);
  h->cls_w = h->mem.upload_f32(h->store.get("f0_predictor.classifier.weight", {1, 512}).data);
  h->cls_b = h->mem.upload_f32(h->store.get("f0_predictor.classifier.bias", {1}).data);
  h->src_w = h->mem.upload_f32(h->store.get("m_source.l_linear.weight", {1, 9}).data);
  h->src_b = h->mem.upload_f32(h->store.get("m_source.l_linear.bias", {1}).data);
  h->conv_pre = hpack_conv(h, "conv_pre", 512, 80, 7, 128, 512);
  h->conv_post = hpack_conv(h, "conv_post", 18, 64, 7, 64, SPEC_LD);
  int cin = 512;
  for (int i = 0; i < 3; ++i) {
    UpW& U = h->ups[i];
    U.u = UPS_U[i];
    U.k = UPS_K[i];
    U.pad = (U.k - U.u) / 2;
    U.cin = cin;
    U.cout = STAGE_C[i];
    const std::string name = "ups." + std::to_string(i);
    std::vector<float> w = effective_weight(h->store, name, {U.cin, U.cout, U.k});
    const HostTensor& b = h->store.get(name + ".bias", {U.cout});
    for (int ph = 0; ph < U.u; ++ph) {
      // out[m*u + ph] = sum_delta W[:, :, ph + pad - delta*u]^T x[m + delta]   (ConvTranspose1d, stride u)
      UpPhase P;
      std::vector<TapSrc> taps;
      for (int delta = (ph + U.pad) / U.u; ph + U.pad - delta * U.u < U.k; --delta) {
        taps.push_back({ph + U.pad - delta * U.u, 0, U.cin});
        P.shifts.push_back(delta);
      }
      P.w = hpack(h, pack_convT_taps(w.data(), U.cin, U.cout, U.k, taps, U.cin, U.cout), pad_vec(b.data.data(), U.cout, U.cout),
                  U.cout, U.cin, (int)taps.size());
      U.phases.push_back(std::move(P));
    }
    cin = U.cout;
    // source_downs: plain Conv1d(18 -> C, kernel 2u (1 when u == 1), stride u)
    const int su = SRC_U[i], sk = su == 1 ? 1 : 2 * su;
    h->src_down[i] = hpack_conv(h, "source_downs." + std::to_string(i), STAGE_C[i], 18, sk, 18, STAGE_C[i]);
    {  // same weights as one dense K = Kp row per output channel (tap-major, channel-minor == the im2col window order)
      const int klen = sk * 18, Kp = round_up(klen, 64);
      const HostTensor& w = h->store.get("source_downs." + std::to_string(i) + ".weight", {STAGE_C[i], 18, sk});
      const HostTensor& b = h->store.get("source_downs." + std::to_string(i) + ".bias", {STAGE_C[i]});
      std::vector<float> wp((size_t)STAGE_C[i] * Kp, 0.f);
      for (int n = 0; n < STAGE_C[i]; ++n)
        for (int t = 0; t < sk; ++t)
          for (int c = 0; c < 18; ++c) wp[(size_t)n * Kp + t * 18 + c] = w.data[((size_t)n * 18 + c) * sk + t];
      h->src_down_im[i] = hpack(h, std::move(wp), pad_vec(b.data.data(), STAGE_C[i], STAGE_C[i]), STAGE_C[i], Kp, 1);
    }
    h->src_rb[i] = hpack_resblock(h, "source_resblocks." + std::to_string(i), STAGE_C[i], SRC_K[i]);
    for (int j = 0; j < 3; ++j) h->rb[3 * i + j] = hpack_resblock(h, "resblocks." + std::to_string(3 * i + j), STAGE_C[i], RES_K[j]);
  }
  // DFT tables (periodic hann, n_fft 16)
  double w[16];
  for (int n = 0; n < 16; ++n) w[n] = 0.5 - 0.5 * std::cos(2.0 * M_PI * n / 16.0);
  for (int k = 0; k < 9; ++k)
    for (int n = 0; n < 16; ++n) {
      const double ang = 2.0 * M_PI * k * n / 16.0;
      const float wf = (float)w[n];
      h->stft_tb.wc[k][n] = wf * (float)std::cos(ang);
      h->stft_tb.ws[k][n] = wf * (float)(-std::sin(ang));
      const double ck = (k == 0 || k == 8) ? 1.0 : 2.0;
      h->istft_tb.cr[k][n] = (float)(ck * std::cos(ang) / 16.0) * wf;
      h->istft_tb.ci[k][n] = (k == 0 || k == 8) ? 0.f : (float)(-ck * std::sin(ang) / 16.0) * wf;
    }
  for (int n = 0; n < 16; ++n) h->istft_tb.w2[n] = (float)w[n] * (float)w[n];
  h->store.require_all_used();  // names the first unexpected key (load_state_dict(strict=True) semantics)
  JV_REQUIRE(h->store.t.size() == 328, JV_ERR_STATE, "expected 328 HiFT tensors, got %zu (unexpected keys present)", h->store.t.size());
  h->store.t.clear();
  JV_CUDA(cudaDeviceSynchronize());
  h->finalized = true;
}

// ------------------------------------------------------------------------------------------ layout / workspace
struct HiftLayout {
  int B = 0, Tm = 0;
  std::vector<int> off, len;
  int rows[4], rows_alloc[4];  // mel, stage0, stage1, stage2
};

static HiftLayout hift_layout(int B, const int32_t* lens) {
  HiftLayout L;
  L.B = B;
  L.off.resize(B + 1);
  L.len.assign(lens, lens + B);
  int off = 0;
  for (int b = 0; b < B; ++b) {
    JV_REQUIRE(lens[b] >= 1, JV_ERR_INVALID, "lens[%d] = %d must be >= 1", b, lens[b]);
    L.off[b] = off;
    off += lens[b] + HIFT_GAP;
  }
  L.off[B] = off;
  L.Tm = off;
  const int rates[4] = {1, 8, 40, 120};
  for (int i = 0; i < 4; ++i) {
    L.rows[i] = rates[i] * off;
    L.rows_alloc[i] = round_up(L.rows[i], 128);
  }
  return L;
}

struct HiftBuffers {
  int *off, *len;
  int* fr[4];
  void *MEL, *H1, *H2;           // f0 predictor
  double* D;                     // source phase prefix [B, 9, Tmax]
  void* SST;                     // [rows2, 18]
  void* IM;                      // im2col of the source_downs windows (bf16 mode): max over stages of rows * Kp
  void* P[4];                    // conv inputs per level: [Tm,512] [8Tm,256] [40Tm,128] [120Tm,64]
  float* X[3];                   // per stage: x after ups + source fusion (fp32)
  float* S[3][3];                // per stage: the three ResBlock streams (S[i][0] first carries the source branch)
  void *XT[3], *XT2[3];          // per stage activation-typed conv inputs
  float* SPEC;                   // [rows2, SPEC_LD]
};

// `Tlong` = longest utterance of the batch (NOT the caller's tensor stride: jv_hift_workspace_bytes sees only the lengths)
static HiftBuffers hift_carve(Arena& ar, const Engine& eng, const HiftLayout& L, int Tlong) {
  HiftBuffers b;
  const size_t es = eng.act_size();
  b.off = ar.alloc<int>(L.B + 1);
  b.len = ar.alloc<int>(L.B);
  for (int i = 0; i < 4; ++i) b.fr[i] = ar.alloc<int>(L.rows_alloc[i]);
  b.MEL = ar.alloc<char>((size_t)L.rows_alloc[0] * 128 * es);
  b.H1 = ar.alloc<char>((size_t)L.rows_alloc[0] * 512 * es);
  b.H2 = ar.alloc<char>((size_t)L.rows_alloc[0] * 512 * es);
  b.D = ar.alloc<double>((size_t)L.B * 9 * Tlong);
  b.SST = ar.alloc<char>((size_t)L.rows_alloc[3] * 18 * es);
  {
    size_t im = 0;
    const int kp[3] = {576, 128, 64};
    for (int i = 0; i < 3; ++i) im = std::max(im, (size_t)L.rows_alloc[i + 1] * kp[i] * 2);
    b.IM = ar.alloc<char>(eng.is_bf16() ? im : 16);
  }
  const int pc[4] = {512, 256, 128, 64};
  for (int i = 0; i < 4; ++i) b.P[i] = ar.alloc<char>((size_t)L.rows_alloc[i] * pc[i] * es);
  for (int i = 0; i < 3; ++i) {
    const size_t n = (size_t)L.rows_alloc[i + 1] * STAGE_C[i];
    b.X[i] = ar.alloc<float>(n);
    for (int j = 0; j < 3; ++j) b.S[i][j] = ar.alloc<float>(n);
    b.XT[i] = ar.alloc<char>(n * es);
    b.XT2[i] = ar.alloc<char>(n * es);
  }
  b.SPEC = ar.alloc<float>((size_t)L.rows_alloc[3] * SPEC_LD);
  return b;
}

struct HCtx {
  jv_hift* h;
  HiftLayout L;
  HiftBuffers b;
  HiftSeq sq;
  int Tmax;   // row stride of the caller's [B, *, Tmax] tensors
  int Tlong;  // longest utterance: row stride of the phase-prefix scratch D
  long valid_rows[4];  // valid frames per level (algorithmic FLOP accounting)
  cudaStream_t st;
};

static void hift_setup(HCtx& c, jv_hift* h, int B, int Tmax, const int32_t* lens, void* ws, size_t ws_bytes, void* stream) {
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "hift not finalised");
  JV_REQUIRE(B >= 1 && Tmax >= 1 && lens && ws, JV_ERR_INVALID, "bad arguments");
  for (int b = 0; b < B; ++b)
    JV_REQUIRE(lens[b] >= 1 && lens[b] <= Tmax, JV_ERR_INVALID, "lens[%d] = %d outside [1, Tmax = %d]", b, lens[b], Tmax);
  JV_CUDA(cudaSetDevice(h->eng.device));
  c.h = h;
  c.L = hift_layout(B, lens);
  Arena ar(ws, ws_bytes);
  c.Tlong = 0;
  for (int b = 0; b < B; ++b) c.Tlong = std::max(c.Tlong, lens[b]);
  c.b = hift_carve(ar, h->eng, c.L, c.Tlong);
  c.Tmax = Tmax;
  c.st = (cudaStream_t)stream;
  JV_CUDA(cudaMemcpyAsync(c.b.off, c.L.off.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.len, c.L.len.data(), B * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaStreamSynchronize(c.st));  // c.L's vectors are copied again when HCtx moves; keep it simple
  {
    long tsum = 0;
    for (int b = 0; b < B; ++b) tsum += lens[b];
    c.valid_rows[0] = tsum;
    c.valid_rows[1] = 8 * tsum;
    c.valid_rows[2] = 40 * tsum;
    c.valid_rows[3] = 120 * tsum + B;
  }
  c.sq.off = c.b.off;
  c.sq.len = c.b.len;
  c.sq.B = B;
}

static void hift_frame_rows(const HCtx& c, int level) {
  const int rates[4] = {1, 8, 40, 120};
  const int rows = c.L.rows_alloc[level];
  hift_frame_row_kernel<<<cdiv(rows, 256), 256, 0, c.st>>>(c.b.fr[level], rows, c.sq, rates[level], level == 3 ? 1 : 0);
  JV_LAUNCHED();
}

// generic "same-length" Conv1d as GEMM taps at one level
static GemmDesc hconv_desc(const HCtx& c, const PackedW& w, const void* A, int level, int dilation, int pad) {
  GemmDesc g = gemm_desc_default();
  g.A[0] = A;
  g.lda[0] = w.K_tap;
  g.a_rows[0] = c.L.rows_alloc[level];
  g.n_taps = w.n_taps;
  g.K_tap = w.K_tap;
  for (int k = 0; k < w.n_taps; ++k) {
    g.tap_src[k] = 0;
    g.tap_shift[k] = k * dilation - pad;
  }
  g.W = w.W;
  g.W_hi = w.W_hi;
  g.W_lo = w.W_lo;
  g.M = c.L.rows[level];
  g.N = w.N;
  g.bias = w.bias;
  g.frame_row = c.b.fr[level];
  g.o_rows = c.L.rows_alloc[level];
  g.algo_flops = 2.0 * (double)c.valid_rows[level] * w.N * w.n_taps * w.K_tap;  // callers fix up padded N / K
  return g;
}

template <typename TA>
static void launch_act_rows(const HCtx& c, const float* in, void* out, long n, int Cn, int act, float p, const float* vec) {
  act_rows_kernel<TA><<<(unsigned)((n / 4 + 255) / 256), 256, 0, c.st>>>(in, (TA*)out, n, Cn, act, p, vec);
  JV_LAUNCHED();
}
static void run_act_rows(const HCtx& c, const float* in, void* out, long n, int Cn, int act, float p, const float* vec) {
  if (c.h->eng.is_bf16()) launch_act_rows<bf16>(c, in, out, n, Cn, act, p, vec);
  else launch_act_rows<float>(c, in, out, n, Cn, act, p, vec);
}

// ResBlock (generator.py:90-97).  On entry XT holds Snake_{a1[0]}(x0).  The block's own stream lives in S.
// Final value (x after the third pair) * final_scale goes to final_out (accumulating if asked) and,
// optionally, act2(final_out) to final_act.
static void run_resblock(const HCtx& c, const ResBlockW& w, int stage, const float* x0, float* S, void* XT, void* XT2) {
  Engine& e = c.h->eng;
  const int level = stage + 1;
  const int Cn = w.C;
  for (int i = 0; i < 3; ++i) {
    const int d = RES_D[i];
    GemmDesc g = hconv_desc(c, w.c1[i], XT, level, d, (w.k * d - d) / 2);
    g.act = ACT_SNAKE;
    g.act_vec = w.a2[i];
    g.out_act = XT2;
    g.ldo2 = Cn;
    e.gemm(g, c.st);
    g = hconv_desc(c, w.c2[i], XT2, level, 1, (w.k - 1) / 2);
    g.resid = i == 0 ? x0 : S;
    g.ldr = Cn;
    g.out_f32 = S;
    g.ldo = Cn;
    if (i < 2) {
      g.out_act = XT;
      g.ldo2 = Cn;
      g.act2 = ACT_SNAKE;
      g.act2_vec = w.a1[i + 1];
    }
    e.gemm(g, c.st);
  }
}

template <typename TA>
static void launch_pack_mel(const HCtx& c, const float* mel) {
  const long n = (long)c.L.rows_alloc[0] * 128;
  hift_pack_mel_kernel<TA><<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>((TA*)c.b.MEL, c.b.fr[0], c.sq, c.L.rows_alloc[0], mel, c.Tmax);
  JV_LAUNCHED();
}
static void run_pack_mel(const HCtx& c, const float* mel) {
  if (c.h->eng.is_bf16()) launch_pack_mel<bf16>(c, mel);
  else launch_pack_mel<float>(c, mel);
}

static void run_f0(const HCtx& c, const float* mel, float* f0) {
  jv_hift* h = c.h;
  hift_frame_rows(c, 0);
  run_pack_mel(c, mel);
  const void* in = c.b.MEL;
  void* bufs[2] = {c.b.H1, c.b.H2};
  for (int i = 0; i < 5; ++i) {
    GemmDesc g = hconv_desc(c, h->f0conv[i], in, 0, 1, 1);
    if (i == 0) g.algo_flops *= 80.0 / 128.0;
    g.act = ACT_ELU;
    g.out_act = bufs[i & 1];
    g.ldo2 = 512;
    h->eng.gemm(g, c.st);
    in = bufs[i & 1];
  }
  const long n = (long)c.L.B * c.Tmax;
  zero_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(f0, n);
  JV_LAUNCHED();
  const int rows = c.L.rows_alloc[0];
  if (h->eng.is_bf16())
    hift_f0_head_kernel<bf16><<<cdiv(rows * 32, 256), 256, 0, c.st>>>((const bf16*)in, c.b.fr[0], c.sq, rows, h->cls_w, h->cls_b, f0, c.Tmax);
  else
    hift_f0_head_kernel<float><<<cdiv(rows * 32, 256), 256, 0, c.st>>>((const float*)in, c.b.fr[0], c.sq, rows, h->cls_w, h->cls_b, f0, c.Tmax);
  JV_LAUNCHED();
}

// s_stft (generator.py:371-381, 399-400) -> SST [rows(level 3), 18], activation type; needs fr[3]
static void run_stft(const HCtx& c, const float* s) {
  jv_hift* h = c.h;
  const int rows = c.L.rows_alloc[3];
  if (h->eng.is_bf16()) hift_stft_kernel<bf16><<<cdiv(rows, 128), 128, 0, c.st>>>((bf16*)c.b.SST, c.b.fr[3], c.sq, rows, s, c.Tmax, h->stft_tb);
  else hift_stft_kernel<float><<<cdiv(rows, 128), 128, 0, c.st>>>((float*)c.b.SST, c.b.fr[3], c.sq, rows, s, c.Tmax, h->stft_tb);
  JV_LAUNCHED();
}

static void run_decode(const HCtx& c, const float* mel, const float* s, float* wav) {
  jv_hift* h = c.h;
  Engine& e = h->eng;
  for (int lv = 0; lv < 4; ++lv) hift_frame_rows(c, lv);
  run_pack_mel(c, mel);
  run_stft(c, s);
  // conv_pre, then leaky_relu(0.1) for ups[0] (generator.py:402-404)
  {
    GemmDesc g = hconv_desc(c, h->conv_pre, c.b.MEL, 0, 1, 3);
    g.algo_flops *= 80.0 / 128.0;
    g.out_act = c.b.P[0];
    g.ldo2 = 512;
    g.act2 = ACT_LRELU;
    g.act2_param = 0.1f;
    e.gemm(g, c.st);
  }
  for (int i = 0; i < 3; ++i) {
    const UpW& U = h->ups[i];
    const int lv_in = i, lv_out = i + 1;
    const int Cn = STAGE_C[i];
    // si = source_resblocks[i](source_downs[i](s_stft)) -> S[i][0]   (generator.py:411-412; independent of x)
    {
      const int su = SRC_U[i];
      const PackedW& w = h->src_down[i];
      GemmDesc g = gemm_desc_default();
      if (e.is_bf16()) {  // im2col (a shifted copy) + one dense tcgen05 GEMM
        const PackedW& wi = h->src_down_im[i];
        const long rows_out = c.L.rows_alloc[lv_out];
        const long n = rows_out * (wi.K_tap / 8);
        hift_im2col_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>((bf16*)c.b.IM, (const bf16*)c.b.SST, rows_out, wi.K_tap,
                                                                        w.n_taps * 18, su, (long)c.L.rows_alloc[3] * 18);
        JV_LAUNCHED();
        g.A[0] = c.b.IM;
        g.lda[0] = wi.K_tap;
        g.a_rows[0] = rows_out;
        g.n_taps = 1;
        g.K_tap = wi.K_tap;
        g.W = wi.W;
        g.W_hi = wi.W_hi;
        g.W_lo = wi.W_lo;
      } else {
        g.A[0] = c.b.SST;
        g.lda[0] = 18;
        g.a_rows[0] = c.L.rows_alloc[3];
        g.a_stride = su;
        g.n_taps = w.n_taps;
        g.K_tap = 18;
        for (int t = 0; t < w.n_taps; ++t) {
          g.tap_src[t] = 0;
          g.tap_shift[t] = t - su / 2;
        }
        g.W = w.W;
        g.W_hi = w.W_hi;
        g.W_lo = w.W_lo;
      }
      g.M = c.L.rows[lv_out];
      g.N = Cn;
      g.bias = w.bias;
      g.frame_row = c.b.fr[lv_out];
      g.out_f32 = c.b.S[i][0];
      g.ldo = Cn;
      g.out_act = c.b.XT[i];
      g.ldo2 = Cn;
      g.act2 = ACT_SNAKE;
      g.act2_vec = h->src_rb[i].a1[0];
      g.o_rows = c.L.rows_alloc[lv_out];
      g.algo_flops = 2.0 * (double)c.valid_rows[lv_out] * Cn * w.n_taps * 18;
      e.gemm(g, c.st);
      run_resblock(c, h->src_rb[i], i, c.b.S[i][0], c.b.S[i][0], c.b.XT[i], c.b.XT2[i]);
    }
    // x = ups[i](leaky_relu(x)) + si: one GEMM per output phase, si added as the epilogue residual
    for (int ph = 0; ph < U.u; ++ph) {
      const UpPhase& P = U.phases[ph];
      GemmDesc g = gemm_desc_default();
      g.A[0] = c.b.P[i];
      g.lda[0] = U.cin;
      g.a_rows[0] = c.L.rows_alloc[lv_in];
      g.n_taps = P.w.n_taps;
      g.K_tap = U.cin;
      for (int t = 0; t < P.w.n_taps; ++t) {
        g.tap_src[t] = 0;
        g.tap_shift[t] = P.shifts[t];
      }
      g.W = P.w.W;
      g.W_hi = P.w.W_hi;
      g.W_lo = P.w.W_lo;
      g.M = c.L.rows[lv_in];
      g.N = Cn;
      g.bias = P.w.bias;
      g.frame_row = c.b.fr[lv_out];
      g.resid = c.b.S[i][0];
      g.ldr = Cn;
      g.out_f32 = c.b.X[i];
      g.ldo = Cn;
      g.o_stride = U.u;
      g.o_off = ph + (i == 2 ? 1 : 0);  // stage 2: ReflectionPad1d((1,0)) shifts the signal by one
      g.o_rows = c.L.rows_alloc[lv_out];
      g.algo_flops = 2.0 * (double)c.valid_rows[lv_in] * Cn * P.w.n_taps * U.cin;
      e.gemm(g, c.st);
    }
    if (i == 2) {
      hift_reflect_fix_kernel<<<cdiv(c.L.B * Cn, 128), 128, 0, c.st>>>(c.b.X[i], c.b.S[i][0], Cn, c.sq, STAGE_RATE[i]);
      JV_LAUNCHED();
    }
    // x = mean_j resblocks[3i+j](x); then leaky_relu for the next consumer (0.1 before ups, 0.01 before conv_post)
    const long n = (long)c.L.rows_alloc[lv_out] * Cn;
    for (int j = 0; j < 3; ++j) {
      const ResBlockW& w = h->rb[3 * i + j];
      run_act_rows(c, c.b.X[i], c.b.XT[i], n, Cn, ACT_SNAKE, 0.f, w.a1[0]);
      run_resblock(c, w, i, c.b.X[i], c.b.S[i][j], c.b.XT[i], c.b.XT2[i]);
    }
    const float slope = i == 2 ? 0.01f : 0.1f;
    if (e.is_bf16())
      mean3_act_kernel<bf16><<<(unsigned)((n / 4 + 255) / 256), 256, 0, c.st>>>(c.b.S[i][0], c.b.S[i][1], c.b.S[i][2], (bf16*)c.b.P[i + 1], n, slope);
    else
      mean3_act_kernel<float><<<(unsigned)((n / 4 + 255) / 256), 256, 0, c.st>>>(c.b.S[i][0], c.b.S[i][1], c.b.S[i][2], (float*)c.b.P[i + 1], n, slope);
    JV_LAUNCHED();
  }
  {
    GemmDesc g = hconv_desc(c, h->conv_post, c.b.P[3], 3, 1, 3);
    g.algo_flops *= 18.0 / SPEC_LD;
    g.out_f32 = c.b.SPEC;
    g.ldo = SPEC_LD;
    e.gemm(g, c.st);
  }
  dim3 grid(cdiv(480 * c.Tmax, 1024), c.L.B);
  hift_istft_kernel<<<grid, ISTFT_THREADS, 0, c.st>>>(c.b.SPEC, SPEC_LD, c.sq, c.Tmax, wav, h->istft_tb, 0.99f, (long)c.L.rows_alloc[3]);
  JV_LAUNCHED();
}

}  // namespace jv

// =========================================================================================== C ABI


extern "C" {

int jv_hift_create(int device, int precision, jv_hift** out) {
  JV_API_BEGIN
  JV_REQUIRE(out != nullptr, JV_ERR_INVALID, "out is NULL");
  std::unique_ptr<jv_hift> h(new jv_hift());
  h->eng.init(device, precision);
  *out = h.release();
  JV_API_END
}

void jv_hift_destroy(jv_hift* h) { delete h; }

int jv_hift_set_weight(jv_hift* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "hift already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  h->store.set(key, data, shape, ndim);
  JV_API_END
}

int jv_hift_finalize(jv_hift* h) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  hift_finalize_impl(h);
  JV_API_END
}

size_t jv_hift_workspace_bytes(const jv_hift* h, int B, const int32_t* lens_host) {
  try {
    if (!h || B < 1 || !lens_host) return 0;
    HiftLayout L = hift_layout(B, lens_host);
    int tmax = 0;
    for (int b = 0; b < B; ++b) tmax = std::max(tmax, lens_host[b]);
    Arena ar(nullptr, 0);
    hift_carve(ar, h->eng, L, tmax);
    return ar.off + 256;
  } catch (const std::exception& e) {
    jv::set_last_error(e.what());
    return 0;
  }
}

int jv_hift_f0(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* mel, float* f0, void* ws, size_t ws_bytes,
               void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(mel && f0, JV_ERR_INVALID, "bad arguments");
  HCtx c;
  hift_setup(c, h, B, Tmax, lens_host, ws, ws_bytes, stream);
  run_f0(c, mel, f0);
  JV_API_END
}

int jv_hift_source(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* f0, const float* phase, const float* noise,
                   float* s, void* ws, size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(f0 && phase && noise && s, JV_ERR_INVALID, "bad arguments");
  HCtx c;
  hift_setup(c, h, B, Tmax, lens_host, ws, ws_bytes, stream);
  hift_phase_prefix_kernel<<<cdiv(B * 9, 64), 64, 0, c.st>>>(f0, Tmax, c.b.len, B, c.b.D, c.Tlong);
  JV_LAUNCHED();
  const long n = (long)B * 480 * Tmax;
  hift_source_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, c.st>>>(f0, c.b.D, phase, noise, h->src_w, h->src_b, c.b.len, B, Tmax, s,
                                                                    c.Tlong);
  JV_LAUNCHED();
  JV_API_END
}

int jv_hift_stft(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* s, float* out, void* ws, size_t ws_bytes,
                 void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(s && out, JV_ERR_INVALID, "bad arguments");
  HCtx c;
  hift_setup(c, h, B, Tmax, lens_host, ws, ws_bytes, stream);
  hift_frame_rows(c, 3);
  run_stft(c, s);
  const int Fmax = 120 * Tmax + 1;
  const long n = (long)B * 18 * Fmax;
  if (h->eng.is_bf16()) hift_unpack_stft_kernel<bf16><<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(out, (const bf16*)c.b.SST, c.sq, Fmax);
  else hift_unpack_stft_kernel<float><<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(out, (const float*)c.b.SST, c.sq, Fmax);
  JV_LAUNCHED();
  JV_API_END
}

int jv_hift_decode(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* mel, const float* s, float* wav, void* ws,
                   size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(mel && s && wav, JV_ERR_INVALID, "bad arguments");
  HCtx c;
  hift_setup(c, h, B, Tmax, lens_host, ws, ws_bytes, stream);
  run_decode(c, mel, s, wav);
  JV_API_END
}

}  // extern "C"
