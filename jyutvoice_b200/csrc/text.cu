// Text front of `synthesise` on the GPU, batched over ragged utterances:
//   TextEncoder.forward          jyutvoice/models/text_encoder.py:401-451
//   DurationPredictor.forward    jyutvoice/models/duration_predictor.py:48-60
//   durations -> lengths -> hard monotonic alignment -> mu_y gather   jyutvoice/models/jyutvoice_tts.py:184-203
// fp32 throughout (3xTF32 tcgen05 GEMM-with-taps engine + the small kernels of text_kernels.cuh): the durations pass through ceil(),
// so this part keeps the reference's arithmetic type in both precision modes.
#include <cmath>
#include <memory>
#include <algorithm>

#include "engine.cuh"
#include "weights.cuh"
#include "text_kernels.cuh"

namespace jv {
struct TeLN {
  float* g = nullptr;
  float* b = nullptr;
};
struct TeLayer {
  PackedW qkv, o, f1, f2;
  TeLN n1, n2;
};
}  // namespace jv

using namespace jv;

struct jv_text {
  Engine eng;
  WeightStore store;
  DeviceAlloc mem;
  bool finalized = false, has_enc = false, has_dp = false;
  int n_vocab = 0, n_lang = 0, n_tone = 0;
  float *emb = nullptr, *lang_emb = nullptr, *tone_emb = nullptr, *wp_emb = nullptr, *sp_emb = nullptr;
  PackedW pre_conv[3], pre_proj;
  TeLN pre_ln[3];
  TeLayer L[TE_LAYERS];
  PackedW proj;
  PackedW dp_c1, dp_c2, dp_proj, dp_cond;
  TeLN dp_n1, dp_n2;
  float* rope_cs = nullptr;  // [rope_T, 144] cos | sin
  int rope_T = 0;
};

namespace jv {

static PackedW te_pack(jv_text* h, std::vector<float>&& w, const std::vector<float>& bias, int N, int K_tap, int n_taps) {
  PackedW p;
  p.N = p.N_pad = N;
  p.K_tap = K_tap;
  p.n_taps = n_taps;
  p.W = h->mem.upload_f32(w);
  h->mem.upload_tf32_split(w, &p.W_hi, &p.W_lo);
  p.bias = h->mem.upload_f32(bias);
  return p;
}

// Conv1d [Cout, Cin, Kw] -> taps k = 0 .. Kw-1 over K_tap = Cin channels
static PackedW te_conv(jv_text* h, const std::string& name, int Cout, int Cin, int Kw) {
  const HostTensor& w = h->store.get(name + ".weight", {Cout, Cin, Kw});
  const HostTensor& b = h->store.get(name + ".bias", {Cout});
  std::vector<TapSrc> taps;
  for (int k = 0; k < Kw; ++k) taps.push_back({k, 0, Cin});
  return te_pack(h, pack_conv_taps(w.data.data(), Cout, Cin, Kw, taps, Cin, Cout), b.data, Cout, Cin, Kw);
}

static TeLN te_ln(jv_text* h, const std::string& name, int C) {
  TeLN l;
  l.g = h->mem.upload_f32(h->store.get(name + ".gamma", {C}).data);
  l.b = h->mem.upload_f32(h->store.get(name + ".beta", {C}).data);
  return l;
}

static float* te_table(jv_text* h, const std::string& name, int* rows) {
  auto it = h->store.t.find(name);
  JV_REQUIRE(it != h->store.t.end(), JV_ERR_STATE, "missing weight '%s'", name.c_str());
  const HostTensor& t = it->second;
  JV_REQUIRE(t.shape.size() == 2 && t.shape[1] == TE_C && t.shape[0] >= 1, JV_ERR_INVALID, "weight '%s' must be [n, 192]", name.c_str());
  h->store.used.insert(name);
  *rows = (int)t.shape[0];
  return h->mem.upload_f32(t.data);
}

static void te_finalize(jv_text* h) {
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "text handle already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  for (const auto& kv : h->store.t) {
    if (kv.first.rfind("encoder.", 0) == 0) h->has_enc = true;
    if (kv.first.rfind("dp.", 0) == 0) h->has_dp = true;
  }
  JV_REQUIRE(h->has_enc || h->has_dp, JV_ERR_STATE, "no weights set (keys start with 'encoder.' or 'dp.')");
  if (h->has_enc) {
    const std::string e = "encoder.";
    int four = 0;
    h->emb = te_table(h, e + "emb.weight", &h->n_vocab);
    h->lang_emb = te_table(h, e + "lang_emb.weight", &h->n_lang);
    h->tone_emb = te_table(h, e + "tone_emb.weight", &h->n_tone);
    h->wp_emb = te_table(h, e + "word_pos_emb.weight", &four);
    JV_REQUIRE(four == 4, JV_ERR_INVALID, "word_pos_emb must have 4 rows");
    h->sp_emb = te_table(h, e + "syllable_pos.weight", &four);
    JV_REQUIRE(four == 4, JV_ERR_INVALID, "syllable_pos must have 4 rows");
    for (int i = 0; i < 3; ++i) {
      h->pre_conv[i] = te_conv(h, e + "prenet.conv_layers." + std::to_string(i), TE_C, TE_C, 5);
      h->pre_ln[i] = te_ln(h, e + "prenet.norm_layers." + std::to_string(i), TE_C);
    }
    h->pre_proj = te_conv(h, e + "prenet.proj", TE_C, TE_C, 1);
    for (int i = 0; i < TE_LAYERS; ++i) {
      const std::string a = e + "encoder.attn_layers." + std::to_string(i);
      std::vector<float> w, b;
      for (const char* c : {"q", "k", "v"}) {
        const HostTensor& wt = h->store.get(a + ".conv_" + c + ".weight", {TE_H, TE_H, 1});
        const HostTensor& bt = h->store.get(a + ".conv_" + c + ".bias", {TE_H});
        w.insert(w.end(), wt.data.begin(), wt.data.end());
        b.insert(b.end(), bt.data.begin(), bt.data.end());
      }
      h->L[i].qkv = te_pack(h, std::move(w), b, 3 * TE_H, TE_H, 1);
      h->L[i].o = te_conv(h, a + ".conv_o", TE_H, TE_H, 1);
      h->L[i].n1 = te_ln(h, e + "encoder.norm_layers_1." + std::to_string(i), TE_H);
      h->L[i].f1 = te_conv(h, e + "encoder.ffn_layers." + std::to_string(i) + ".conv_1", TE_FC, TE_H, 3);
      h->L[i].f2 = te_conv(h, e + "encoder.ffn_layers." + std::to_string(i) + ".conv_2", TE_H, TE_FC, 3);
      h->L[i].n2 = te_ln(h, e + "encoder.norm_layers_2." + std::to_string(i), TE_H);
    }
    h->proj = te_conv(h, e + "proj", 80, TE_H, 1);
  }
  if (h->has_dp) {
    const std::string d = "dp.";
    h->dp_c1 = te_conv(h, d + "conv_1", TE_DP, TE_H, 3);
    h->dp_n1 = te_ln(h, d + "norm_1", TE_DP);
    h->dp_c2 = te_conv(h, d + "conv_2", TE_DP, TE_DP, 3);
    h->dp_n2 = te_ln(h, d + "norm_2", TE_DP);
    h->dp_proj = te_conv(h, d + "proj", 1, TE_DP, 1);
    h->dp_cond = te_conv(h, d + "cond", TE_H, TE_C, 1);
  }
  h->store.require_all_used();
  h->store.t.clear();
  JV_CUDA(cudaDeviceSynchronize());
  h->finalized = true;
}

// RoPE cache exactly as RotaryPositionalEmbeddings._build_cache (text_encoder.py:110-135), fp32 on the host
static void te_rope_cache(jv_text* h, int T) {
  if (T <= h->rope_T) return;
  const int Tn = std::max(T, 256);
  std::vector<float> cs((size_t)Tn * 144);
  for (int i = 0; i < 72; ++i) {
    const float theta = 1.0f / std::pow(10000.0f, (float)(2 * i) / 144.0f);
    for (int p = 0; p < Tn; ++p) {
      const float a = (float)p * theta;
      cs[(size_t)p * 144 + i] = std::cos(a);
      cs[(size_t)p * 144 + 72 + i] = std::sin(a);
    }
  }
  h->rope_cs = h->mem.upload_f32(cs);  // (the previous, shorter table stays allocated until the handle dies: a few hundred KB)
  h->rope_T = Tn;
}

struct TeLayout {
  int B = 0, M = 0, M_alloc = 0, Tlong = 0;
  std::vector<int> off, len, frame_row;
};
static TeLayout te_layout(int B, int Tx, const int32_t* lens) {
  TeLayout L;
  L.B = B;
  L.off.resize(B + 1);
  L.len.assign(lens, lens + B);
  int off = 0;
  for (int b = 0; b < B; ++b) {
    JV_REQUIRE(lens[b] >= 1 && lens[b] <= Tx, JV_ERR_INVALID, "x_lengths[%d] = %d outside [1, Tx = %d]", b, lens[b], Tx);
    L.off[b] = off;
    off += lens[b] + TE_GAP;
    L.Tlong = std::max(L.Tlong, lens[b]);
  }
  L.off[B] = off;
  L.M = off;
  L.M_alloc = round_up(off, 128);
  L.frame_row.assign(L.M_alloc, -1);
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < lens[b]; ++t) L.frame_row[L.off[b] + t] = b;
  return L;
}

struct TeBuffers {
  int *frame_row, *off, *len;
  float *X0, *T1, *T2, *XH, *QKV, *ATT, *Y, *F, *MU, *COND, *LOGW;
};
static TeBuffers te_carve(Arena& ar, int M_alloc, int B) {
  TeBuffers b;
  b.frame_row = ar.alloc<int>(M_alloc);
  b.off = ar.alloc<int>(B + 1);
  b.len = ar.alloc<int>(B);
  b.X0 = ar.alloc<float>((size_t)M_alloc * TE_C);
  b.T1 = ar.alloc<float>((size_t)M_alloc * TE_DP);   // prenet temporaries (192 wide) and the duration predictor's (256 wide)
  b.T2 = ar.alloc<float>((size_t)M_alloc * TE_DP);
  b.XH = ar.alloc<float>((size_t)M_alloc * TE_H);
  b.QKV = ar.alloc<float>((size_t)M_alloc * 3 * TE_H);
  b.ATT = ar.alloc<float>((size_t)M_alloc * TE_H);
  b.Y = ar.alloc<float>((size_t)M_alloc * TE_H);
  b.F = ar.alloc<float>((size_t)M_alloc * TE_FC);
  b.MU = ar.alloc<float>((size_t)M_alloc * 80);
  b.COND = ar.alloc<float>((size_t)round_up(B, 128) * TE_H);
  b.LOGW = ar.alloc<float>((size_t)M_alloc);
  return b;
}

struct TeCtx {
  jv_text* h;
  TeLayout L;
  TeBuffers b;
  cudaStream_t st;
};

static void te_setup(TeCtx& c, jv_text* h, int B, int Tx, const int32_t* lens, void* ws, size_t ws_bytes, void* stream) {
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "text handle not finalised");
  JV_REQUIRE(B >= 1 && Tx >= 1 && lens && ws, JV_ERR_INVALID, "bad arguments");
  JV_CUDA(cudaSetDevice(h->eng.device));
  c.h = h;
  c.L = te_layout(B, Tx, lens);
  Arena ar(ws, ws_bytes);
  c.b = te_carve(ar, c.L.M_alloc, B);
  c.st = (cudaStream_t)stream;
  JV_CUDA(cudaMemcpyAsync(c.b.frame_row, c.L.frame_row.data(), c.L.frame_row.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.off, c.L.off.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.len, c.L.len.data(), B * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaStreamSynchronize(c.st));  // the host vectors live in c.L, but keep the contract simple: nothing pending on return paths
}

// "same" Conv1d over the packed token rows as GEMM taps
static GemmDesc te_desc(const TeCtx& c, const PackedW& w, const float* A, int lda) {
  GemmDesc g = gemm_desc_default();
  g.A[0] = A;
  g.lda[0] = lda;
  g.a_rows[0] = c.L.M_alloc;
  g.n_taps = w.n_taps;
  g.K_tap = w.K_tap;
  for (int k = 0; k < w.n_taps; ++k) {
    g.tap_src[k] = 0;
    g.tap_shift[k] = k - (w.n_taps - 1) / 2;
  }
  g.W = w.W;
  g.W_hi = w.W_hi;
  g.W_lo = w.W_lo;
  g.M = c.L.M_alloc;
  g.N = w.N;
  g.bias = w.bias;
  g.o_rows = c.L.M_alloc;
  return g;
}

static void te_ln_launch(const TeCtx& c, const float* x, int ldx, const float* add, int ld_add, const TeLN& ln, int C, int relu, float* out,
                         int ldo) {
  te_ln_kernel<<<cdiv(c.L.M_alloc * 32, 256), 256, 0, c.st>>>(x, ldx, add, ld_add, ln.g, ln.b, C, relu, c.b.frame_row, out, ldo, c.L.M_alloc);
  JV_LAUNCHED();
}

static void te_encode(TeCtx& c, int Tx, const long long* x, const long long* lang, const long long* tone, const long long* wp,
                      const long long* sp, const float* spk, float* out_x, float* out_mu) {
  jv_text* h = c.h;
  Engine& e = h->eng;
  const int M = c.L.M_alloc;
  JV_REQUIRE(h->has_enc, JV_ERR_STATE, "this handle holds no TextEncoder weights");
  te_rope_cache(h, c.L.Tlong);
  te_embed_kernel<<<(unsigned)(((long)M * TE_C + 255) / 256), 256, 0, c.st>>>(c.b.X0, c.b.frame_row, c.b.off, M, x, tone, wp, sp, Tx, h->emb,
                                                                             h->tone_emb, h->wp_emb, h->sp_emb, h->n_vocab, h->n_tone);
  JV_LAUNCHED();
  // prenet (ConvReluNorm, text_encoder.py:76-83): 3 x [conv k5 (x * mask) -> LayerNorm -> ReLU], then x_org + proj(x), masked
  const float* in = c.b.X0;
  for (int i = 0; i < 3; ++i) {
    GemmDesc g = te_desc(c, h->pre_conv[i], in, TE_C);
    g.out_f32 = c.b.T1;
    g.ldo = TE_C;
    e.gemm(g, c.st);
    te_ln_launch(c, c.b.T1, TE_C, nullptr, 0, h->pre_ln[i], TE_C, 1, c.b.T2, TE_C);  // masked: it only feeds the next (x * mask) conv
    in = c.b.T2;
  }
  {
    GemmDesc g = te_desc(c, h->pre_proj, in, TE_C);
    g.frame_row = c.b.frame_row;
    g.resid = c.b.X0;
    g.ldr = TE_C;
    g.out_f32 = c.b.XH;  // phoneme channels 0 .. 191 of the encoder input
    g.ldo = TE_H;
    e.gemm(g, c.st);
  }
  te_concat_kernel<<<(unsigned)(((long)M * 2 * TE_C + 255) / 256), 256, 0, c.st>>>(c.b.XH, c.b.frame_row, c.b.off, M, spk, lang, Tx,
                                                                                  h->lang_emb, h->n_lang);
  JV_LAUNCHED();
  // Encoder (text_encoder.py:327-337).  Every stored x is masked, which is what each consumer sees in the reference.
  static unsigned long long attr = 0;
  const int att_smem = 4 * c.L.Tlong * (int)sizeof(float);
  if (att_smem > 48 * 1024 && first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(te_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  JV_REQUIRE(att_smem <= 200 * 1024, JV_ERR_INVALID, "utterances of more than 12800 tokens are not supported");
  for (int i = 0; i < TE_LAYERS; ++i) {
    const TeLayer& Ly = h->L[i];
    GemmDesc g = te_desc(c, Ly.qkv, c.b.XH, TE_H);
    g.out_f32 = c.b.QKV;
    g.ldo = 3 * TE_H;
    e.gemm(g, c.st);
    te_rope_kernel<<<(unsigned)(((long)M * 2 * TE_HEADS * 72 + 255) / 256), 256, 0, c.st>>>(c.b.QKV, c.b.frame_row, c.b.off, M, h->rope_cs);
    JV_LAUNCHED();
    te_attention_kernel<<<cdiv(M * TE_HEADS, 4), 128, att_smem, c.st>>>(c.b.QKV, c.b.ATT, c.b.frame_row, c.b.off, c.b.len, M, c.L.Tlong);
    JV_LAUNCHED();
    g = te_desc(c, Ly.o, c.b.ATT, TE_H);
    g.out_f32 = c.b.Y;
    g.ldo = TE_H;
    e.gemm(g, c.st);
    te_ln_launch(c, c.b.XH, TE_H, c.b.Y, TE_H, Ly.n1, TE_H, 0, c.b.XH, TE_H);  // x = LN(x + attn(x)), in place (one warp owns its row)
    g = te_desc(c, Ly.f1, c.b.XH, TE_H);
    g.act = ACT_LRELU;  // slope 0: ReLU
    g.act_param = 0.f;
    g.frame_row = c.b.frame_row;
    g.out_f32 = c.b.F;
    g.ldo = TE_FC;
    e.gemm(g, c.st);
    g = te_desc(c, Ly.f2, c.b.F, TE_FC);
    g.frame_row = c.b.frame_row;
    g.out_f32 = c.b.Y;
    g.ldo = TE_H;
    e.gemm(g, c.st);
    te_ln_launch(c, c.b.XH, TE_H, c.b.Y, TE_H, Ly.n2, TE_H, 0, c.b.XH, TE_H);
  }
  {
    GemmDesc g = te_desc(c, h->proj, c.b.XH, TE_H);
    g.frame_row = c.b.frame_row;
    g.out_f32 = c.b.MU;
    g.ldo = 80;
    e.gemm(g, c.st);
  }
  const int B = c.L.B;
  te_unpack_kernel<<<(unsigned)(((long)B * TE_H * Tx + 255) / 256), 256, 0, c.st>>>(out_x, c.b.XH, TE_H, TE_H, c.b.off, c.b.len, B, Tx);
  JV_LAUNCHED();
  te_unpack_kernel<<<(unsigned)(((long)B * 80 * Tx + 255) / 256), 256, 0, c.st>>>(out_mu, c.b.MU, 80, 80, c.b.off, c.b.len, B, Tx);
  JV_LAUNCHED();
}

static void te_predict(TeCtx& c, int Tx, const float* x, const float* g_spk, float* out_logw) {
  jv_text* h = c.h;
  Engine& e = h->eng;
  const int M = c.L.M_alloc, B = c.L.B;
  JV_REQUIRE(h->has_dp, JV_ERR_STATE, "this handle holds no DurationPredictor weights");
  {  // cond(g): one 576-vector per utterance (duration_predictor.py:50)
    GemmDesc g = gemm_desc_default();
    g.A[0] = g_spk;
    g.lda[0] = TE_C;
    g.a_rows[0] = B;
    g.n_taps = 1;
    g.K_tap = TE_C;
    g.W = h->dp_cond.W;
    g.W_hi = h->dp_cond.W_hi;
    g.W_lo = h->dp_cond.W_lo;
    g.M = B;
    g.N = TE_H;
    g.bias = h->dp_cond.bias;
    g.out_f32 = c.b.COND;
    g.ldo = TE_H;
    g.o_rows = B;
    e.gemm(g, c.st);
  }
  te_pack_cond_kernel<<<(unsigned)(((long)M * TE_H + 255) / 256), 256, 0, c.st>>>(c.b.XH, c.b.frame_row, c.b.off, M, x, c.b.COND, TE_H, Tx);
  JV_LAUNCHED();
  GemmDesc g = te_desc(c, h->dp_c1, c.b.XH, TE_H);  // conv_1 -> relu -> norm_1 (:51-53)
  g.act = ACT_LRELU;
  g.out_f32 = c.b.T1;
  g.ldo = TE_DP;
  e.gemm(g, c.st);
  te_ln_launch(c, c.b.T1, TE_DP, nullptr, 0, h->dp_n1, TE_DP, 0, c.b.T2, TE_DP);
  g = te_desc(c, h->dp_c2, c.b.T2, TE_DP);  // conv_2 -> relu -> norm_2 (:55-57)
  g.act = ACT_LRELU;
  g.out_f32 = c.b.T1;
  g.ldo = TE_DP;
  e.gemm(g, c.st);
  te_ln_launch(c, c.b.T1, TE_DP, nullptr, 0, h->dp_n2, TE_DP, 0, c.b.T2, TE_DP);
  g = te_desc(c, h->dp_proj, c.b.T2, TE_DP);  // proj, masked (:59-60)
  g.frame_row = c.b.frame_row;
  g.out_f32 = c.b.LOGW;
  g.ldo = 1;
  e.gemm(g, c.st);
  te_unpack_kernel<<<(unsigned)(((long)B * Tx + 255) / 256), 256, 0, c.st>>>(out_logw, c.b.LOGW, 1, 1, c.b.off, c.b.len, B, Tx);
  JV_LAUNCHED();
}

}  // namespace jv

// =========================================================================================== C ABI
extern "C" {

int jv_text_create(int device, jv_text** out) {
  JV_API_BEGIN
  JV_REQUIRE(out != nullptr, JV_ERR_INVALID, "out is NULL");
  std::unique_ptr<jv_text> h(new jv_text());
  h->eng.init(device, JV_PREC_FP32);
  *out = h.release();
  JV_API_END
}

void jv_text_destroy(jv_text* h) { delete h; }

int jv_text_set_weight(jv_text* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "text handle already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  h->store.set(key, data, shape, ndim);
  JV_API_END
}

int jv_text_finalize(jv_text* h) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  te_finalize(h);
  JV_API_END
}

size_t jv_text_workspace_bytes(const jv_text* h, int B, int Tx, const int32_t* x_lens_host) {
  try {
    if (!h || B < 1 || Tx < 1 || !x_lens_host) return 0;
    TeLayout L = te_layout(B, Tx, x_lens_host);
    Arena ar(nullptr, 0);
    te_carve(ar, L.M_alloc, B);
    return ar.off + 256;
  } catch (const std::exception& e) {
    jv::set_last_error(e.what());
    return 0;
  }
}

int jv_text_encode(jv_text* h, int B, int Tx, const int32_t* x_lens_host, const int64_t* x, const int64_t* lang, const int64_t* tone,
                   const int64_t* word_pos, const int64_t* syllable_pos, const float* spk_embed, float* out_x, float* out_mu, void* ws,
                   size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(x && lang && tone && word_pos && syllable_pos && spk_embed && out_x && out_mu, JV_ERR_INVALID, "bad arguments");
  TeCtx c;
  te_setup(c, h, B, Tx, x_lens_host, ws, ws_bytes, stream);
  te_encode(c, Tx, (const long long*)x, (const long long*)lang, (const long long*)tone, (const long long*)word_pos,
            (const long long*)syllable_pos, spk_embed, out_x, out_mu);
  JV_API_END
}

int jv_text_durations(jv_text* h, int B, int Tx, const int32_t* x_lens_host, const float* x, const float* spk_embed, float* out_logw,
                      void* ws, size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(x && spk_embed && out_logw, JV_ERR_INVALID, "bad arguments");
  TeCtx c;
  te_setup(c, h, B, Tx, x_lens_host, ws, ws_bytes, stream);
  te_predict(c, Tx, x, spk_embed, out_logw);
  JV_API_END
}

int jv_length_durations(int B, int Tx, const int32_t* x_lens_dev, const float* logw, float length_scale, float* cum, int64_t* y_lengths,
                        void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(B >= 1 && Tx >= 1 && x_lens_dev && logw && cum && y_lengths, JV_ERR_INVALID, "bad arguments");
  te_durations_kernel<<<cdiv(B, 64), 64, 0, (cudaStream_t)stream>>>(logw, x_lens_dev, B, Tx, length_scale, cum, (long long*)y_lengths);
  JV_LAUNCHED();
  JV_API_END
}

int jv_length_align(int B, int Tx, int Ty, const int32_t* x_lens_dev, const int64_t* y_lengths, const float* cum, const float* mu_x,
                    float* mu_y, int32_t* frame_token, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(B >= 1 && Tx >= 1 && Ty >= 1 && x_lens_dev && y_lengths && cum && mu_x && mu_y && frame_token, JV_ERR_INVALID, "bad arguments");
  te_align_kernel<<<(unsigned)(((long)B * Ty + 127) / 128), 128, 0, (cudaStream_t)stream>>>(cum, x_lens_dev, (const long long*)y_lengths, B, Tx,
                                                                                          Ty, mu_x, mu_y, frame_token);
  JV_LAUNCHED();
  JV_API_END
}

}  // extern "C"
