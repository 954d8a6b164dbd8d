// Speech-token encoder on the GPU: what turns the prompt's speech tokens into `prompt_h`.
//   FlowEncoder.forward                 infer.py:66-82            (token embedding, encoder, 512 -> 80 projection)
//   UpsampleConformerEncoder.forward    jyutvoice/transformer/upsample_encoder.py:290-355
//     embed / up_embed                  transformer/subsampling.py:70-115 + embedding.py:201-298 (Linear, LayerNorm, * sqrt(512), rel-pos table)
//     pre_lookahead_layer, up_layer     upsample_encoder.py:37-137
//     6 + 4 ConformerEncoderLayer       transformer/encoder_layer.py:236-330 (pre-norm, no macaron, no cnn module)
//     RelPositionMultiHeadedAttention   transformer/attention.py:196-330
// Batched over ragged utterances; every utterance is computed as the reference's batch-1 call computes it (the zero rows
// between utterances are the convs' zero padding).  fp32 throughout: 3xTF32 tcgen05 GEMM-with-taps engine + the row kernels
// of flowenc_kernels.cuh.
#include <cmath>
#include <memory>
#include <algorithm>

#include "engine.cuh"
#include "weights.cuh"
#include "flowenc_kernels.cuh"

namespace jv {
struct FeLN {
  float *g = nullptr, *b = nullptr;
};
struct FeLayer {
  PackedW qkv, pos, out, w1, w2;
  FeLN n_mha, n_ff;
  float *bias_u = nullptr, *bias_v = nullptr;
};
struct FeEmbed {
  PackedW lin;
  FeLN ln;
};
}  // namespace jv

using namespace jv;

struct jv_flowenc {
  Engine eng;
  WeightStore store;
  DeviceAlloc mem;
  bool finalized = false, has_emb = false, has_proj = false;
  int vocab = 0;
  float* emb = nullptr;
  FeEmbed embed, up_embed;
  PackedW pre1, pre2, up_conv, proj;
  FeLayer LA[FE_LAYERS_A], LB[FE_LAYERS_B];
  FeLN after;
  float* pe = nullptr;  // [2 * FE_PE_MAX - 1, 512]: row r holds position FE_PE_MAX - 1 - r (embedding.py:236-253)
};

namespace jv {

static PackedW fe_pack(jv_flowenc* h, std::vector<float>&& w, const std::vector<float>& bias, int N, int K_tap, int n_taps) {
  PackedW p;
  p.N = p.N_pad = N;
  p.K_tap = K_tap;
  p.n_taps = n_taps;
  p.W = h->mem.upload_f32(w);
  h->mem.upload_tf32_split(w, &p.W_hi, &p.W_lo);
  p.bias = bias.empty() ? nullptr : h->mem.upload_f32(bias);
  return p;
}

static PackedW fe_linear(jv_flowenc* h, const std::string& name, int N, int K, bool bias = true) {
  const HostTensor& w = h->store.get(name + ".weight", {N, K});
  std::vector<float> b;
  if (bias) b = h->store.get(name + ".bias", {N}).data;
  return fe_pack(h, std::vector<float>(w.data), b, N, K, 1);
}

static PackedW fe_conv(jv_flowenc* h, const std::string& name, int Kw) {
  const HostTensor& w = h->store.get(name + ".weight", {FE_C, FE_C, Kw});
  const HostTensor& b = h->store.get(name + ".bias", {FE_C});
  std::vector<TapSrc> taps;
  for (int k = 0; k < Kw; ++k) taps.push_back({k, 0, FE_C});
  return fe_pack(h, pack_conv_taps(w.data.data(), FE_C, FE_C, Kw, taps, FE_C, FE_C), b.data, FE_C, FE_C, Kw);
}

static FeLN fe_lnw(jv_flowenc* h, const std::string& name) {
  FeLN l;
  l.g = h->mem.upload_f32(h->store.get(name + ".weight", {FE_C}).data);
  l.b = h->mem.upload_f32(h->store.get(name + ".bias", {FE_C}).data);
  return l;
}

static void fe_layer(jv_flowenc* h, FeLayer& L, const std::string& a) {
  std::vector<float> w, b;
  for (const char* c : {"q", "k", "v"}) {
    const HostTensor& wt = h->store.get(a + "self_attn.linear_" + c + ".weight", {FE_C, FE_C});
    const HostTensor& bt = h->store.get(a + "self_attn.linear_" + c + ".bias", {FE_C});
    w.insert(w.end(), wt.data.begin(), wt.data.end());
    b.insert(b.end(), bt.data.begin(), bt.data.end());
  }
  L.qkv = fe_pack(h, std::move(w), b, 3 * FE_C, FE_C, 1);
  L.pos = fe_linear(h, a + "self_attn.linear_pos", FE_C, FE_C, false);
  L.out = fe_linear(h, a + "self_attn.linear_out", FE_C, FE_C);
  L.bias_u = h->mem.upload_f32(h->store.get(a + "self_attn.pos_bias_u", {FE_HEADS, FE_DK}).data);
  L.bias_v = h->mem.upload_f32(h->store.get(a + "self_attn.pos_bias_v", {FE_HEADS, FE_DK}).data);
  L.w1 = fe_linear(h, a + "feed_forward.w_1", FE_FC, FE_C);
  L.w2 = fe_linear(h, a + "feed_forward.w_2", FE_C, FE_FC);
  L.n_mha = fe_lnw(h, a + "norm_mha");
  L.n_ff = fe_lnw(h, a + "norm_ff");
}

static void fe_finalize(jv_flowenc* h) {
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "flow-encoder handle already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  const std::string e = "encoder.";
  h->has_emb = h->store.has("input_embedding.weight");
  h->has_proj = h->store.has("encoder_proj.weight");
  if (h->has_emb) {
    auto it = h->store.t.find("input_embedding.weight");
    const HostTensor& t = it->second;
    JV_REQUIRE(t.shape.size() == 2 && t.shape[1] == FE_C, JV_ERR_INVALID, "input_embedding.weight must be [vocab, 512]");
    h->store.used.insert("input_embedding.weight");
    h->vocab = (int)t.shape[0];
    h->emb = h->mem.upload_f32(t.data);
  }
  for (int s = 0; s < 2; ++s) {
    FeEmbed& E = s ? h->up_embed : h->embed;
    const std::string n = e + (s ? "up_embed" : "embed") + ".out.";
    E.lin = fe_linear(h, n + "0", FE_C, FE_C);
    E.ln = fe_lnw(h, n + "1");
  }
  h->pre1 = fe_conv(h, e + "pre_lookahead_layer.conv1", 4);
  h->pre2 = fe_conv(h, e + "pre_lookahead_layer.conv2", 3);
  h->up_conv = fe_conv(h, e + "up_layer.conv", 5);
  for (int i = 0; i < FE_LAYERS_A; ++i) fe_layer(h, h->LA[i], e + "encoders." + std::to_string(i) + ".");
  for (int i = 0; i < FE_LAYERS_B; ++i) fe_layer(h, h->LB[i], e + "up_encoders." + std::to_string(i) + ".");
  h->after = fe_lnw(h, e + "after_norm");
  if (h->has_proj) h->proj = fe_linear(h, "encoder_proj", 80, FE_C);
  h->store.require_all_used();
  h->store.t.clear();
  // EspnetRelPositionalEncoding.extend_pe (embedding.py:236-253) for max_len 5000, fp32 like the reference's table:
  // pe[p, 2c] = sin(p * div_c), pe[p, 2c + 1] = cos(p * div_c), div_c = exp(2c * -(ln 10000 / 512)); rows by falling position
  {
    std::vector<float> pe((size_t)(2 * FE_PE_MAX - 1) * FE_C);
    std::vector<float> div(FE_C / 2);
    const float k = (float)(-(std::log(10000.0) / (double)FE_C));
    for (int c = 0; c < FE_C / 2; ++c) div[c] = std::exp((float)(2 * c) * k);
    for (int r = 0; r < 2 * FE_PE_MAX - 1; ++r) {
      const float pos = (float)(FE_PE_MAX - 1 - r);
      float* row = pe.data() + (size_t)r * FE_C;
      for (int c = 0; c < FE_C / 2; ++c) {
        const float a = pos * div[c];
        row[2 * c] = std::sin(a);
        row[2 * c + 1] = std::cos(a);
      }
    }
    h->pe = h->mem.upload_f32(pe);
  }
  JV_CUDA(cudaDeviceSynchronize());
  h->finalized = true;
}

struct FeLayout {
  int B = 0, M = 0, M_alloc = 0, Tlong = 0;
  std::vector<int> off, len, frame_row;
};
static FeLayout fe_layout(int B, int T, const int32_t* lens, int mult) {
  FeLayout L;
  L.B = B;
  L.off.resize(B + 1);
  L.len.resize(B);
  int off = 0;
  for (int b = 0; b < B; ++b) {
    JV_REQUIRE(lens[b] >= 1 && lens[b] <= T, JV_ERR_INVALID, "lens[%d] = %d outside [1, T = %d]", b, lens[b], T);
    L.len[b] = lens[b] * mult;
    L.off[b] = off;
    off += L.len[b] + FE_GAP;
    L.Tlong = std::max(L.Tlong, L.len[b]);
  }
  L.off[B] = off;
  L.M = off;
  L.M_alloc = round_up(off, 128);
  L.frame_row.assign(L.M_alloc, -1);
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < L.len[b]; ++t) L.frame_row[L.off[b] + t] = b;
  return L;
}

struct FeStage {  // device copies of one layout
  int *frame_row, *off, *len;
};
struct FeBuffers {
  FeStage a, b;
  float *X, *Y, *T1, *T2, *QKV, *ATT, *F, *PT, *H;
};
static FeBuffers fe_carve(Arena& ar, const FeLayout& A, const FeLayout& Bl) {
  FeBuffers b;
  const int B = A.B;
  const size_t M = (size_t)Bl.M_alloc;  // the upsampled layout is the larger one: every activation buffer is sized for it
  b.a.frame_row = ar.alloc<int>(A.M_alloc);
  b.a.off = ar.alloc<int>(B + 1);
  b.a.len = ar.alloc<int>(B);
  b.b.frame_row = ar.alloc<int>(Bl.M_alloc);
  b.b.off = ar.alloc<int>(B + 1);
  b.b.len = ar.alloc<int>(B);
  b.X = ar.alloc<float>(M * FE_C);
  b.Y = ar.alloc<float>(M * FE_C);
  b.T1 = ar.alloc<float>(M * FE_C);
  b.T2 = ar.alloc<float>(M * FE_C);
  b.QKV = ar.alloc<float>(M * 3 * FE_C);
  b.ATT = ar.alloc<float>(M * FE_C);
  b.F = ar.alloc<float>(M * FE_FC);
  b.PT = ar.alloc<float>((size_t)round_up(2 * Bl.Tlong - 1, 128) * FE_C);
  b.H = ar.alloc<float>(M * 80);
  return b;
}

struct FeCtx {
  jv_flowenc* h;
  FeLayout A, Bl;
  FeBuffers b;
  cudaStream_t st;
};

static void fe_upload(const FeLayout& L, const FeStage& s, cudaStream_t st) {
  JV_CUDA(cudaMemcpyAsync(s.frame_row, L.frame_row.data(), L.frame_row.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  JV_CUDA(cudaMemcpyAsync(s.off, L.off.data(), (L.B + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  JV_CUDA(cudaMemcpyAsync(s.len, L.len.data(), L.B * sizeof(int), cudaMemcpyHostToDevice, st));
}

// Conv1d / Linear over the packed rows as GEMM taps; tap k reads row m + first_shift + k
static GemmDesc fe_desc(const FeLayout& L, const PackedW& w, const float* A, int lda, int first_shift) {
  GemmDesc g = gemm_desc_default();
  g.A[0] = A;
  g.lda[0] = lda;
  g.a_rows[0] = L.M_alloc;
  g.n_taps = w.n_taps;
  g.K_tap = w.K_tap;
  for (int k = 0; k < w.n_taps; ++k) {
    g.tap_src[k] = 0;
    g.tap_shift[k] = first_shift + k;
  }
  g.W = w.W;
  g.W_hi = w.W_hi;
  g.W_lo = w.W_lo;
  g.M = L.M_alloc;
  g.N = w.N;
  g.bias = w.bias;
  g.o_rows = L.M_alloc;
  return g;
}

static void fe_ln(const FeCtx& c, const FeLayout& L, const FeStage& s, const float* x, const FeLN& ln, float eps, float scale, float* out) {
  fe_ln_kernel<<<cdiv(L.M_alloc * 32, 256), 256, 0, c.st>>>(x, ln.g, ln.b, eps, scale, s.frame_row, out, L.M_alloc);
  JV_LAUNCHED();
}

// LinearNoSubsampling.forward (subsampling.py:113-115): x = LayerNorm(Linear(x)) * sqrt(512); in T1 -> out X
static void fe_embed(const FeCtx& c, const FeLayout& L, const FeStage& s, const FeEmbed& E, const float* in, float* tmp, float* out) {
  GemmDesc g = fe_desc(L, E.lin, in, FE_C, 0);
  g.out_f32 = tmp;
  g.ldo = FE_C;
  c.h->eng.gemm(g, c.st);
  fe_ln(c, L, s, tmp, E.ln, 1e-5f, 22.627416997969522f, out);  // sqrt(512)
}

// ConformerEncoderLayer.forward (encoder_layer.py:268-330 with normalize_before, no macaron, no cnn module), x in place:
//   x += linear_out(rel_attention(LN_mha(x)));  x += w_2(swish(w_1(LN_ff(x))))
static void fe_run_layer(const FeCtx& c, const FeLayout& L, const FeStage& s, const FeLayer& Ly, int chunk, int att_smem) {
  Engine& e = c.h->eng;
  const FeBuffers& b = c.b;
  fe_ln(c, L, s, b.X, Ly.n_mha, 1e-12f, 1.0f, b.Y);
  GemmDesc g = fe_desc(L, Ly.qkv, b.Y, FE_C, 0);
  g.out_f32 = b.QKV;
  g.ldo = 3 * FE_C;
  e.gemm(g, c.st);
  {  // P = linear_pos(pos_emb) for positions Tlong-1 .. -(Tlong-1) (attention.py:307-309): rows of the resident table
    const int rows = 2 * L.Tlong - 1;
    GemmDesc p = gemm_desc_default();
    p.A[0] = c.h->pe + (size_t)(FE_PE_MAX - L.Tlong) * FE_C;
    p.lda[0] = FE_C;
    p.a_rows[0] = rows;
    p.n_taps = 1;
    p.K_tap = FE_C;
    p.W = Ly.pos.W;
    p.W_hi = Ly.pos.W_hi;
    p.W_lo = Ly.pos.W_lo;
    p.M = rows;
    p.N = FE_C;
    p.out_f32 = b.PT;
    p.ldo = FE_C;
    p.o_rows = rows;
    e.gemm(p, c.st);
  }
  fe_rel_attention_kernel<<<cdiv(L.M_alloc * FE_HEADS, 4), 128, att_smem, c.st>>>(b.QKV, b.PT, L.Tlong, Ly.bias_u, Ly.bias_v, b.ATT,
                                                                                  s.frame_row, s.off, s.len, L.M_alloc, L.Tlong, chunk);
  JV_LAUNCHED();
  g = fe_desc(L, Ly.out, b.ATT, FE_C, 0);
  g.frame_row = s.frame_row;
  g.resid = b.X;
  g.ldr = FE_C;
  g.out_f32 = b.X;
  g.ldo = FE_C;
  e.gemm(g, c.st);
  fe_ln(c, L, s, b.X, Ly.n_ff, 1e-12f, 1.0f, b.Y);
  g = fe_desc(L, Ly.w1, b.Y, FE_C, 0);
  g.act = ACT_SILU;
  g.out_f32 = b.F;
  g.ldo = FE_FC;
  e.gemm(g, c.st);
  g = fe_desc(L, Ly.w2, b.F, FE_FC, 0);
  g.frame_row = s.frame_row;
  g.resid = b.X;
  g.ldr = FE_C;
  g.out_f32 = b.X;
  g.ldo = FE_C;
  e.gemm(g, c.st);
}

static void fe_encode(FeCtx& c, int T, const long long* token, const float* xs, int chunk, float* out_hidden, float* out_h) {
  jv_flowenc* h = c.h;
  Engine& e = h->eng;
  const FeBuffers& b = c.b;
  const FeLayout &A = c.A, &U = c.Bl;
  JV_REQUIRE(2 * U.Tlong - 1 <= 2 * FE_PE_MAX - 1 && U.Tlong <= FE_PE_MAX, JV_ERR_INVALID,
             "utterances of more than %d tokens are not supported (rel-pos table of 5000 positions)", FE_PE_MAX / 2);
  static unsigned long long attr = 0;
  const int att_smem = 4 * U.Tlong * (int)sizeof(float);
  if (att_smem > 48 * 1024 && first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(fe_rel_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  JV_REQUIRE(att_smem <= 200 * 1024, JV_ERR_INVALID, "utterance too long for the attention kernel's score buffer");
  const unsigned gridA = (unsigned)(((long)A.M_alloc * (FE_C / 4) + 255) / 256);
  if (token) {
    JV_REQUIRE(h->has_emb, JV_ERR_STATE, "this handle holds no input_embedding (encoder-only weights): pass features, not tokens");
    fe_embed_kernel<<<gridA, 256, 0, c.st>>>(b.T2, b.a.frame_row, b.a.off, A.M_alloc, token, T, h->emb, h->vocab);
  } else {
    fe_pack_kernel<<<gridA, 256, 0, c.st>>>(b.T2, b.a.frame_row, b.a.off, A.M_alloc, xs, T);
  }
  JV_LAUNCHED();
  fe_embed(c, A, b.a, h->embed, b.T2, b.T1, b.Y);
  {  // PreLookaheadLayer (upsample_encoder.py:103-137): conv k4 over frames t .. t+3, leaky_relu, causal conv k3, + input
    GemmDesc g = fe_desc(A, h->pre1, b.Y, FE_C, 0);
    g.act = ACT_LRELU;
    g.act_param = 0.01f;
    g.frame_row = b.a.frame_row;  // gap rows must stay zero: they are conv2's left padding of the next utterance
    g.out_f32 = b.T1;
    g.ldo = FE_C;
    e.gemm(g, c.st);
    g = fe_desc(A, h->pre2, b.T1, FE_C, -2);
    g.frame_row = b.a.frame_row;
    g.resid = b.Y;
    g.ldr = FE_C;
    g.out_f32 = b.X;
    g.ldo = FE_C;
    e.gemm(g, c.st);
  }
  const int att_smem_a = 4 * A.Tlong * (int)sizeof(float);
  for (int i = 0; i < FE_LAYERS_A; ++i) fe_run_layer(c, A, b.a, h->LA[i], chunk, att_smem_a);
  {  // Upsample1D (upsample_encoder.py:62-78): nearest x2, 4 zero frames on the left, conv k5
    fe_repeat_kernel<<<(unsigned)(((long)U.M_alloc * (FE_C / 4) + 255) / 256), 256, 0, c.st>>>(b.T2, b.b.frame_row, b.b.off, U.M_alloc, b.X,
                                                                                               b.a.off);
    JV_LAUNCHED();
    GemmDesc g = fe_desc(U, h->up_conv, b.T2, FE_C, -4);
    g.frame_row = b.b.frame_row;
    g.out_f32 = b.T1;
    g.ldo = FE_C;
    e.gemm(g, c.st);
  }
  fe_embed(c, U, b.b, h->up_embed, b.T1, b.T2, b.X);
  for (int i = 0; i < FE_LAYERS_B; ++i) fe_run_layer(c, U, b.b, h->LB[i], 2 * chunk, att_smem);
  fe_ln(c, U, b.b, b.X, h->after, 1e-5f, 1.0f, b.Y);
  const int B = A.B, T2 = 2 * T;
  if (out_hidden) {
    fe_unpack_kernel<<<(unsigned)(((long)B * T2 * FE_C + 255) / 256), 256, 0, c.st>>>(out_hidden, b.Y, FE_C, b.b.off, b.b.len, B, T2);
    JV_LAUNCHED();
  }
  if (out_h) {
    JV_REQUIRE(h->has_proj, JV_ERR_STATE, "this handle holds no encoder_proj weights");
    GemmDesc g = fe_desc(U, h->proj, b.Y, FE_C, 0);
    g.frame_row = b.b.frame_row;
    g.out_f32 = b.H;
    g.ldo = 80;
    e.gemm(g, c.st);
    fe_unpack_kernel<<<(unsigned)(((long)B * T2 * 80 + 255) / 256), 256, 0, c.st>>>(out_h, b.H, 80, b.b.off, b.b.len, B, T2);
    JV_LAUNCHED();
  }
}

}  // namespace jv

// =========================================================================================== C ABI
extern "C" {

int jv_flowenc_create(int device, jv_flowenc** out) {
  JV_API_BEGIN
  JV_REQUIRE(out != nullptr, JV_ERR_INVALID, "out is NULL");
  std::unique_ptr<jv_flowenc> h(new jv_flowenc());
  h->eng.init(device, JV_PREC_FP32);
  *out = h.release();
  JV_API_END
}

void jv_flowenc_destroy(jv_flowenc* h) { delete h; }

int jv_flowenc_set_weight(jv_flowenc* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "flow-encoder handle already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  h->store.set(key, data, shape, ndim);
  JV_API_END
}

int jv_flowenc_finalize(jv_flowenc* h) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  fe_finalize(h);
  JV_API_END
}

size_t jv_flowenc_workspace_bytes(const jv_flowenc* h, int B, int T, const int32_t* lens_host) {
  try {
    if (!h || B < 1 || T < 1 || !lens_host) return 0;
    FeLayout A = fe_layout(B, T, lens_host, 1), U = fe_layout(B, T, lens_host, 2);
    Arena ar(nullptr, 0);
    fe_carve(ar, A, U);
    return ar.off + 256;
  } catch (const std::exception& e) {
    jv::set_last_error(e.what());
    return 0;
  }
}

int jv_flowenc_encode(jv_flowenc* h, int B, int T, const int32_t* lens_host, const int64_t* token, const float* xs, int chunk,
                      float* out_hidden, float* out_h, void* ws, size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "flow-encoder handle not finalised");
  JV_REQUIRE(B >= 1 && T >= 1 && lens_host && ws, JV_ERR_INVALID, "bad arguments");
  JV_REQUIRE((token != nullptr) != (xs != nullptr), JV_ERR_INVALID, "pass exactly one of token / xs");
  JV_REQUIRE(out_hidden || out_h, JV_ERR_INVALID, "no output requested");
  JV_REQUIRE(chunk >= 0, JV_ERR_INVALID, "chunk must be >= 0");
  JV_CUDA(cudaSetDevice(h->eng.device));
  FeCtx c;
  c.h = h;
  c.A = fe_layout(B, T, lens_host, 1);
  c.Bl = fe_layout(B, T, lens_host, 2);
  Arena ar(ws, ws_bytes);
  c.b = fe_carve(ar, c.A, c.Bl);
  c.st = (cudaStream_t)stream;
  fe_upload(c.A, c.b.a, c.st);
  fe_upload(c.Bl, c.b.b, c.st);
  JV_CUDA(cudaStreamSynchronize(c.st));  // the host vectors die with `c`
  fe_encode(c, T, (const long long*)token, xs, chunk, out_hidden, out_h);
  JV_API_END
}

}  // extern "C"
