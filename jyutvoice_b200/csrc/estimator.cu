// CausalConditionalDecoder estimator + Euler/CFG solver on B200.
// Reference: jyutvoice/flow/decoder.py:798-1018 (estimator), jyutvoice/flow/flow_matching.py:215-265,356-401 (solver).
#include <cmath>
#include <memory>
#include <algorithm>

#include "engine.cuh"
#include "kernels.cuh"
#include "attention_tc.cuh"
#include "attention_tc2.cuh"
#include "attention_tc3.cuh"
#include "mlp_tc.cuh"
#include "weights.cuh"

namespace jv {

constexpr int N_MID = 12;
constexpr int N_TB = 4;
constexpr int N_RESNET = 14;  // down + 12 mid + up
constexpr int C = 256;

struct LNW {
  float* g = nullptr;
  float* b = nullptr;
};
struct ResnetW {
  PackedW c1, c2, res;
  LNW ln1, ln2;
};
struct TBlockW {
  LNW n1, n3;
  PackedW qkv, out, ff1, ff2;
};
struct GroupW {
  ResnetW rn;
  TBlockW tb[N_TB];
};

}  // namespace jv

using namespace jv;

struct jv_estimator {
  Engine eng;
  WeightStore store;
  DeviceAlloc mem;
  bool finalized = false;
  int stream_fmt = 0;       // bf16 mode: residual stream stored as 0 = fp16 (default), 1 = bf16, 2 = fp32 (jv_estimator_set_stream_format)
  bool stream16() const { return eng.is_bf16() && stream_fmt != 2; }   // the XB epilogue kernels
  bool stream_half() const { return eng.is_bf16() && stream_fmt == 0; }
  int* sat_dev = nullptr;   // device counter: rows of the fp16 stream that may have saturated (GemmDesc::sat_flag)
  int* sat_host = nullptr;  // pinned copy, refreshed asynchronously at the end of every forward / solve
  int chunk = 0;  // attention chunk mask of streaming=True (decoder.py:950-953); 0 = full context
  cudaStream_t cap_stream = nullptr, cap_stream2 = nullptr;  // private streams used only to capture an Euler step into a CUDA graph
  cudaStream_t aux_stream = nullptr;                         // second half-batch of a split solve (eager steps)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  GroupW groups[N_RESNET];  // 0 = down, 1..12 = mid, 13 = up
  PackedW down_conv, up_conv, final_conv, final_proj;
  LNW final_ln;
  // time embedding (fp32, tiny): linear_1 [1024,320], linear_2 [1024,1024], 14 x mlp.1 [256,1024]
  float *t_w1 = nullptr, *t_b1 = nullptr, *t_w2 = nullptr, *t_b2 = nullptr;
  float* mlp_w = nullptr;  // [14*256, 1024]
  float* mlp_b = nullptr;  // [14*256]
};

namespace jv {

static PackedW make_packed(jv_estimator* h, std::vector<float>&& w, const float* bias, int N, int K_tap, int n_taps) {
  PackedW p;
  p.N = N;
  p.N_pad = N;
  p.K_tap = K_tap;
  p.n_taps = n_taps;
  p.W = h->mem.upload_act(w, h->eng.is_bf16());
  if (!h->eng.is_bf16()) h->mem.upload_tf32_split(w, &p.W_hi, &p.W_lo);
  p.bias = bias ? h->mem.upload_f32(pad_vec(bias, N, N)) : nullptr;
  return p;
}

static PackedW pack_conv(jv_estimator* h, const std::string& name, int Cout, int Cin, int Kw, int n_src) {
  const HostTensor& w = h->store.get(name + ".weight", {Cout, Cin, Kw});
  const HostTensor& b = h->store.get(name + ".bias", {Cout});
  const int K_tap = Cin / n_src;
  std::vector<TapSrc> taps;
  for (int s = 0; s < n_src; ++s)
    for (int k = 0; k < Kw; ++k) taps.push_back({k, s * K_tap, K_tap});
  return make_packed(h, pack_conv_taps(w.data.data(), Cout, Cin, Kw, taps, K_tap, Cout), b.data.data(), Cout, K_tap, (int)taps.size());
}

static PackedW pack_linear(jv_estimator* h, const std::string& name, int N, int K, bool bias) {
  const HostTensor& w = h->store.get(name + ".weight", {N, K});
  const float* b = bias ? h->store.get(name + ".bias", {N}).data.data() : nullptr;
  std::vector<float> v(w.data);
  return make_packed(h, std::move(v), b, N, K, 1);
}

static LNW pack_ln(jv_estimator* h, const std::string& name) {
  LNW l;
  l.g = h->mem.upload_f32(h->store.get(name + ".weight", {C}).data);
  l.b = h->mem.upload_f32(h->store.get(name + ".bias", {C}).data);
  return l;
}

static void finalize_impl(jv_estimator* h) {
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "estimator already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  auto group_names = [](int g, std::string& rn, std::string& tb) {
    if (g == 0) { rn = "down_blocks.0.0"; tb = "down_blocks.0.1"; }
    else if (g == N_RESNET - 1) { rn = "up_blocks.0.0"; tb = "up_blocks.0.1"; }
    else { rn = "mid_blocks." + std::to_string(g - 1) + ".0"; tb = "mid_blocks." + std::to_string(g - 1) + ".1"; }
  };
  std::vector<float> mlp_w((size_t)N_RESNET * C * 1024), mlp_b((size_t)N_RESNET * C);
  for (int g = 0; g < N_RESNET; ++g) {
    std::string rn, tb;
    group_names(g, rn, tb);
    const int cin = g == 0 ? 320 : (g == N_RESNET - 1 ? 512 : 256);
    const int n_src = g == N_RESNET - 1 ? 2 : 1;
    GroupW& G = h->groups[g];
    G.rn.c1 = pack_conv(h, rn + ".block1.block.0", C, cin, 3, n_src);
    G.rn.ln1 = pack_ln(h, rn + ".block1.block.2");
    G.rn.c2 = pack_conv(h, rn + ".block2.block.0", C, C, 3, 1);
    G.rn.ln2 = pack_ln(h, rn + ".block2.block.2");
    G.rn.res = pack_conv(h, rn + ".res_conv", C, cin, 1, n_src);
    const HostTensor& mw = h->store.get(rn + ".mlp.1.weight", {C, 1024});
    const HostTensor& mb = h->store.get(rn + ".mlp.1.bias", {C});
    std::copy(mw.data.begin(), mw.data.end(), mlp_w.begin() + (size_t)g * C * 1024);
    std::copy(mb.data.begin(), mb.data.end(), mlp_b.begin() + (size_t)g * C);
    for (int j = 0; j < N_TB; ++j) {
      const std::string p = tb + "." + std::to_string(j);
      TBlockW& T = G.tb[j];
      T.n1 = pack_ln(h, p + ".norm1");
      T.n3 = pack_ln(h, p + ".norm3");
      const HostTensor& wq = h->store.get(p + ".attn1.to_q.weight", {512, C});
      const HostTensor& wk = h->store.get(p + ".attn1.to_k.weight", {512, C});
      const HostTensor& wv = h->store.get(p + ".attn1.to_v.weight", {512, C});
      std::vector<float> qkv;
      qkv.reserve(3 * 512 * C);
      qkv.insert(qkv.end(), wq.data.begin(), wq.data.end());
      qkv.insert(qkv.end(), wk.data.begin(), wk.data.end());
      qkv.insert(qkv.end(), wv.data.begin(), wv.data.end());
      T.qkv = make_packed(h, std::move(qkv), nullptr, 1536, C, 1);
      T.out = pack_linear(h, p + ".attn1.to_out.0", C, 512, true);
      T.ff1 = pack_linear(h, p + ".ff.net.0.proj", 1024, C, true);
      T.ff2 = pack_linear(h, p + ".ff.net.2", C, 1024, true);
    }
  }
  h->down_conv = pack_conv(h, "down_blocks.0.2", C, C, 3, 1);
  h->up_conv = pack_conv(h, "up_blocks.0.2", C, C, 3, 1);
  h->final_conv = pack_conv(h, "final_block.block.0", C, C, 3, 1);
  h->final_ln = pack_ln(h, "final_block.block.2");
  h->final_proj = pack_conv(h, "final_proj", 80, C, 1, 1);
  h->t_w1 = h->mem.upload_f32(h->store.get("time_mlp.linear_1.weight", {1024, 320}).data);
  h->t_b1 = h->mem.upload_f32(h->store.get("time_mlp.linear_1.bias", {1024}).data);
  h->t_w2 = h->mem.upload_f32(h->store.get("time_mlp.linear_2.weight", {1024, 1024}).data);
  h->t_b2 = h->mem.upload_f32(h->store.get("time_mlp.linear_2.bias", {1024}).data);
  h->mlp_w = h->mem.upload_f32(mlp_w);
  h->mlp_b = h->mem.upload_f32(mlp_b);
  h->store.require_all_used();  // names the first unexpected key (load_state_dict(strict=True) semantics)
  JV_REQUIRE(h->store.t.size() == 910, JV_ERR_STATE, "expected 910 estimator tensors, got %zu (unexpected keys present)",
             h->store.t.size());
  h->store.t.clear();
  h->sat_dev = (int*)h->mem.raw(sizeof(int));
  JV_CUDA(cudaMemset(h->sat_dev, 0, sizeof(int)));
  JV_CUDA(cudaMallocHost(&h->sat_host, sizeof(int)));
  *h->sat_host = 0;
  h->eng.sat_flag = h->sat_dev;
  JV_CUDA(cudaDeviceSynchronize());
  h->finalized = true;
}

// ------------------------------------------------------------------------------------------ workspace
struct EstLayout {
  int R = 0, M = 0, M_alloc = 0;
  std::vector<int> row_off, row_len, frame_row;
  std::vector<int> att_items;  // attention work items (row << 8 | head << 4 | query tile), ordered, non-empty tiles only
};

static EstLayout make_layout(int R, const int32_t* lens) {
  EstLayout L;
  L.R = R;
  L.row_off.resize(R + 1);
  L.row_len.assign(lens, lens + R);
  int off = 0;
  for (int r = 0; r < R; ++r) {
    JV_REQUIRE(lens[r] >= 1, JV_ERR_INVALID, "lens[%d] = %d must be >= 1", r, lens[r]);
    L.row_off[r] = off;
    off += lens[r] + EST_GAP;
  }
  L.row_off[R] = off;
  L.M = off;
  L.M_alloc = round_up(off, 128);
  L.frame_row.assign(L.M_alloc, -1);
  for (int r = 0; r < R; ++r)
    for (int t = 0; t < lens[r]; ++t) L.frame_row[L.row_off[r] + t] = r;
  for (int r = 0; r < R; ++r)
    for (int h = 0; h < 8; ++h)
      for (int qt = 0; qt * 128 < lens[r] && qt < 16; ++qt) L.att_items.push_back((r << 8) | (h << 4) | qt);
  return L;
}

struct EstBuffers {
  int *frame_row, *row_off, *row_len, *row_tidx;
  int* att_items;  // [n_att_items]
  float* temb;   // [nt, 14, 256]
  float* tsin;   // [nt, 320]
  float* th1;    // [nt, 1024]
  float* th2;    // [nt, 1024]
  void *A0, *XA, *XB, *SKIP, *H, *LNX, *QKV, *ATT, *FF;
  float *X, *Y, *RES, *V;
};

static EstBuffers carve(Arena& ar, const Engine& eng, int M_alloc, int R, int nt) {
  EstBuffers b;
  const size_t es = eng.act_size();
  b.frame_row = ar.alloc<int>(M_alloc);
  b.row_off = ar.alloc<int>(R + 1);
  b.row_len = ar.alloc<int>(R);
  b.row_tidx = ar.alloc<int>(R);
  b.att_items = ar.alloc<int>((size_t)(M_alloc / 128 + R) * 8);  // sum over rows of ceil(len / 128) <= M_alloc / 128 + R
  b.temb = ar.alloc<float>((size_t)nt * N_RESNET * C);
  b.tsin = ar.alloc<float>((size_t)nt * 320);
  b.th1 = ar.alloc<float>((size_t)nt * 1024);
  b.th2 = ar.alloc<float>((size_t)nt * 1024);
  b.A0 = ar.alloc<char>((size_t)M_alloc * 320 * es);
  b.XA = ar.alloc<char>((size_t)M_alloc * C * es);
  b.XB = ar.alloc<char>((size_t)M_alloc * C * es);
  b.SKIP = ar.alloc<char>((size_t)M_alloc * C * es);
  b.H = ar.alloc<char>((size_t)M_alloc * C * es);
  b.LNX = ar.alloc<char>((size_t)M_alloc * C * es);
  b.QKV = ar.alloc<char>((size_t)M_alloc * 1536 * es);
  b.ATT = ar.alloc<char>((size_t)M_alloc * 512 * es);
  b.FF = ar.alloc<char>((size_t)M_alloc * 1024 * es);
  b.X = ar.alloc<float>((size_t)M_alloc * C);
  b.Y = ar.alloc<float>((size_t)M_alloc * C);
  b.RES = ar.alloc<float>((size_t)M_alloc * C);
  b.V = ar.alloc<float>((size_t)M_alloc * 80);
  return b;
}

// ------------------------------------------------------------------------------------------ forward
struct FwdCtx {
  jv_estimator* h;
  EstBuffers b;
  int M, M_alloc, R, Tmax_len;  // Tmax_len: longest row (attention grid)
  int n_att_items;
  long valid_frames;            // sum of row lengths (algorithmic FLOP accounting)
  const float* temb_step;       // temb rows of the current step: [*, 14, 256]
  cudaStream_t st;
};

static GemmDesc conv_desc(const FwdCtx& c, const PackedW& w, const void* A0p, const void* A1p, int Kw) {
  GemmDesc g = gemm_desc_default();
  g.A[0] = A0p;
  g.A[1] = A1p;
  g.lda[0] = g.lda[1] = w.K_tap;
  g.a_rows[0] = g.a_rows[1] = c.M_alloc;
  g.n_taps = w.n_taps;
  g.K_tap = w.K_tap;
  const int n_src = A1p ? 2 : 1;
  for (int s = 0; s < n_src; ++s)
    for (int k = 0; k < Kw; ++k) {
      g.tap_src[s * Kw + k] = s;
      g.tap_shift[s * Kw + k] = k - (Kw - 1);  // causal: taps read t-2, t-1, t
    }
  g.W = w.W;
  g.W_hi = w.W_hi;
  g.W_lo = w.W_lo;
  g.M = c.M_alloc;
  g.N = w.N;
  g.bias = w.bias;
  g.frame_row = c.b.frame_row;
  g.o_rows = c.M_alloc;
  g.algo_flops = 2.0 * (double)c.valid_frames * w.N * w.n_taps * w.K_tap;
  return g;
}

static void run_attention(const FwdCtx& c) {
  if (c.h->eng.is_bf16()) {  // tcgen05 flash-style kernel
    launch_attention(c.h->eng.tmaps, c.b.QKV, c.b.ATT, c.b.row_off, c.b.row_len, c.b.att_items, c.n_att_items, c.M_alloc, c.R,
                     c.Tmax_len, c.h->chunk, c.h->eng.num_sms, c.st);
    return;
  }
  static unsigned long long attr = 0;
  if (first_use_on_device(attr))
    JV_CUDA(cudaFuncSetAttribute(attention_simt_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
  dim3 grid(cdiv(c.Tmax_len, 64), 8, c.R);
  attention_simt_kernel<float><<<grid, 256, ATT_SMEM_BYTES, c.st>>>((const float*)c.b.QKV, 1536, (float*)c.b.ATT, 512, c.b.row_off,
                                                                    c.b.row_len, 0.125f, c.h->chunk);
  JV_LAUNCHED();
}

static void set_ln1(GemmDesc& g, const LNW& ln, int act) {
  g.ln1_gamma = ln.g;
  g.ln1_beta = ln.b;
  g.act = act;
}
static void set_ln2(const FwdCtx& c, GemmDesc& g, const LNW& ln) {
  g.ln2_gamma = ln.g;
  g.ln2_beta = ln.b;
  g.out_ln = c.b.LNX;
  g.ldo3 = C;
}

// resnet (decoder.py:110-115) over conv input(s) in0 (+ in1 for the skip concat); result -> X (fp32 stream) and
// LNX = norm1(X) of the first transformer block.  Three GEMMs, every LayerNorm / Mish / mask / add fused.
static void run_resnet(const FwdCtx& c, const ResnetW& w, int layer, const void* in0, const void* in1, const LNW& next_ln) {
  Engine& e = c.h->eng;
  // h = (Mish(LN(conv1(x*m))) + mlp(t)) * m -> H
  GemmDesc g = conv_desc(c, w.c1, in0, in1, 3);
  set_ln1(g, w.ln1, ACT_MISH);
  g.add_row = c.temb_step + (size_t)layer * C;
  g.row_tidx = c.b.row_tidx;
  g.add_row_stride = N_RESNET * C;
  g.out_act = c.b.H; g.ldo2 = C;
  e.gemm(g, c.st);
  // res_conv(x*m) -> RES (bf16 mode: the residual stream X / RES is 16-bit unless the handle asks for fp32)
  const bool xb = c.h->stream16();
  g = conv_desc(c, w.res, in0, in1, 1);
  if (xb) { g.out_act = c.b.RES; g.ldo2 = C; }
  else { g.out_f32 = c.b.RES; g.ldo = C; }
  e.gemm(g, c.st);
  // x = Mish(LN(conv2(h))) * m + RES -> X ; LNX = norm1(x)
  g = conv_desc(c, w.c2, c.b.H, nullptr, 3);
  set_ln1(g, w.ln2, ACT_MISH);
  g.resid = c.b.RES; g.ldr = C;
  g.out_f32 = c.b.X; g.ldo = C;
  g.x_bf16 = xb;        // RES comes from a bf16-output GEMM; X is the 16-bit stream (fp16 unless the handle says bf16)
  g.x_out_half = c.h->stream_half();
  set_ln2(c, g, next_ln);
  e.gemm(g, c.st);
}

// The fused feed-forward kernel (mlp_tc.cuh, CTA-pair variant) is the default with the fp16 stream;
// JYUTVOICE_B200_MLP=0 selects the two-launch feed-forward (FF1 + GELU, FF2 + residual + norm)
static bool use_mlp_fused() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_MLP");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// BasicTransformerBlock (transformer.py:355-443).  On entry LNX = norm1(X).  `next_ln`: norm1 of the following
// block (fused into the FF2 epilogue) or null; `copy_to`: activation-typed copy of the block output (conv input).
static void run_tblock(const FwdCtx& c, const TBlockW& w, const LNW* next_ln, void* copy_to) {
  Engine& e = c.h->eng;
  GemmDesc g = conv_desc(c, w.qkv, c.b.LNX, nullptr, 1);
  g.out_act = c.b.QKV; g.ldo2 = 1536;
  e.gemm(g, c.st);
  run_attention(c);
  const bool xb = c.h->stream16();
  g = conv_desc(c, w.out, c.b.ATT, nullptr, 1);  // x += to_out(attn) ; LNX = norm3(x)
  g.resid = c.b.X; g.ldr = C;
  g.out_f32 = c.b.X; g.ldo = C;
  g.x_bf16 = xb;
  g.x_in_half = g.x_out_half = c.h->stream_half();
  set_ln2(c, g, w.n3);
  e.gemm(g, c.st);
  // FF1 + GELU + FF2 + residual (+ next norm1) in one kernel: the hidden never leaves the SM.  Not for a handful of tiles: a
  // tile streams the layer's 1 MB of weights through one SM pair's 5-slot ring (30 us at batch 1), where FF1 spreads its
  // weight tiles over several CTAs (tools/latency_probe.py: 2 tiles 37.6 vs 39.0 ms per utterance, 10 tiles 46.9 vs 45.3)
  if (c.h->stream_half() && use_mlp_fused() && (c.M_alloc > 4 * 128 || force_modes())) {
    const bool to_copy = copy_to != nullptr;
    launch_mlp_fused(e.tmaps, c.b.LNX, w.ff1.W, w.ff1.bias, w.ff2.W, w.ff2.bias, c.b.X, to_copy ? copy_to : c.b.X, to_copy ? 0 : 1,
                     next_ln ? next_ln->g : nullptr, next_ln ? next_ln->b : nullptr, next_ln ? c.b.LNX : nullptr, c.b.frame_row,
                     c.M_alloc, e.num_sms, 2.0 * (double)c.valid_frames * (2.0 * C * 1024), e.sat_flag, c.st, c.b.FF,
                     (size_t)c.M_alloc * 1024 * e.act_size());  // FF (the unfused path's hidden activation) is free: tail scratch
    return;
  }
  g = conv_desc(c, w.ff1, c.b.LNX, nullptr, 1);
  g.act = ACT_GELU;
  g.out_act = c.b.FF; g.ldo2 = 1024;
  e.gemm(g, c.st);
  g = conv_desc(c, w.ff2, c.b.FF, nullptr, 1);  // x += ff(norm3(x)) ; LNX = next norm1(x)
  g.resid = c.b.X; g.ldr = C;
  g.out_f32 = c.b.X; g.ldo = C;
  g.x_bf16 = xb;
  g.x_in_half = g.x_out_half = c.h->stream_half();
  if (next_ln) set_ln2(c, g, *next_ln);
  if (copy_to) {
    if (xb) { g.out_f32 = (float*)copy_to; g.x_out_half = 0; }  // bf16 copy for the next conv; X is not read again in this group
    else { g.out_act = copy_to; g.ldo2 = C; }
  }
  e.gemm(g, c.st);
}

// JYUTVOICE_B200_GRAPH=0: every Euler step is launched eagerly
static bool use_graph() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_GRAPH");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// A0 (packed [M,320]) -> V (fp32 [M,80], masked)
static void forward_packed(const FwdCtx& c) {
  jv_estimator* h = c.h;
  Engine& e = h->eng;
  e.scratch = c.b.Y;
  e.scratch_rows = c.M_alloc;
  for (int gi = 0; gi < N_RESNET; ++gi) {
    const GroupW& G = h->groups[gi];
    const void* in0 = gi == 0 ? c.b.A0 : (gi == 1 ? c.b.XB : c.b.XA);  // XB: output of the down CausalConv1d
    const void* in1 = gi == N_RESNET - 1 ? c.b.SKIP : nullptr;
    run_resnet(c, G.rn, gi, in0, in1, G.tb[0].n1);
    for (int j = 0; j < N_TB; ++j) {
      void* copy_to = nullptr;
      if (j == N_TB - 1) copy_to = gi == 0 ? c.b.SKIP : c.b.XA;
      run_tblock(c, G.tb[j], j + 1 < N_TB ? &G.tb[j + 1].n1 : nullptr, copy_to);
    }
    if (gi == 0) {  // x = CausalConv1d(x * mask) (decoder.py:968)
      GemmDesc g = conv_desc(c, h->down_conv, c.b.SKIP, nullptr, 3);
      g.out_act = c.b.XB; g.ldo2 = C;
      e.gemm(g, c.st);
    }
  }
  GemmDesc g = conv_desc(c, h->up_conv, c.b.XA, nullptr, 3);  // decoder.py:1015
  g.out_act = c.b.XB; g.ldo2 = C;
  e.gemm(g, c.st);
  g = conv_desc(c, h->final_conv, c.b.XB, nullptr, 3);  // final_block: Mish(LN(conv)) * m
  set_ln1(g, h->final_ln, ACT_MISH);
  g.out_act = c.b.H; g.ldo2 = C;
  e.gemm(g, c.st);
  g = conv_desc(c, h->final_proj, c.b.H, nullptr, 1);
  g.out_f32 = c.b.V; g.ldo = 80;
  e.gemm(g, c.st);
}

// sinusoidal embedding (decoder.py:21-30) on the host in double, rounded to fp32
static void host_sinusoid(const float* t, int nt, std::vector<float>& out) {
  out.resize((size_t)nt * 320);
  const float neg = (float)(-(std::log(10000.0) / 159.0));
  for (int i = 0; i < nt; ++i) {
    const float ts = 1000.0f * t[i];
    for (int k = 0; k < 160; ++k) {
      const float f = (float)std::exp((double)((float)k * neg));
      const float e = ts * f;
      out[(size_t)i * 320 + k] = (float)std::sin((double)e);
      out[(size_t)i * 320 + 160 + k] = (float)std::cos((double)e);
    }
  }
}

// temb[i, layer, :] = mlp_layer(Mish(time_mlp(sinusoid(t_i))))
static void run_time_embedding(const FwdCtx& c, const float* t_host, int nt) {
  jv_estimator* h = c.h;
  std::vector<float> e;
  host_sinusoid(t_host, nt, e);
  JV_CUDA(cudaMemcpyAsync(c.b.tsin, e.data(), e.size() * sizeof(float), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaStreamSynchronize(c.st));  // `e` is a stack-owned staging buffer
  gemv_rows_kernel<<<cdiv(nt * 1024 * 32, 256), 256, 0, c.st>>>(h->t_w1, h->t_b1, c.b.tsin, c.b.th1, nt, 1024, 320, ACT_NONE, ACT_SILU, 1024);
  JV_LAUNCHED();
  gemv_rows_kernel<<<cdiv(nt * 1024 * 32, 256), 256, 0, c.st>>>(h->t_w2, h->t_b2, c.b.th1, c.b.th2, nt, 1024, 1024, ACT_NONE, ACT_NONE, 1024);
  JV_LAUNCHED();
  gemv_rows_kernel<<<cdiv(nt * N_RESNET * C * 32, 256), 256, 0, c.st>>>(h->mlp_w, h->mlp_b, c.b.th2, c.b.temb, nt, N_RESNET * C, 1024,
                                                                        ACT_MISH, ACT_NONE, N_RESNET * C);
  JV_LAUNCHED();
}

static void upload_layout(const FwdCtx& c, const EstLayout& L, const std::vector<int>& tidx) {
  JV_CUDA(cudaMemcpyAsync(c.b.frame_row, L.frame_row.data(), L.frame_row.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.row_off, L.row_off.data(), L.row_off.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.row_len, L.row_len.data(), L.row_len.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.row_tidx, tidx.data(), tidx.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaMemcpyAsync(c.b.att_items, L.att_items.data(), L.att_items.size() * sizeof(int), cudaMemcpyHostToDevice, c.st));
  JV_CUDA(cudaStreamSynchronize(c.st));  // host vectors die with the caller's frame
}

template <typename TA>
static void launch_pack(const FwdCtx& c, const float* x, const float* mu, const float* spks, const float* cond, int Tmax, int cfg) {
  const long n = (long)c.M_alloc * 320;
  pack_input_kernel<TA><<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>((TA*)c.b.A0, c.b.frame_row, c.b.row_off, c.M_alloc, x, mu, spks,
                                                                       cond, Tmax, cfg);
  JV_LAUNCHED();
}

static void run_pack(const FwdCtx& c, const float* x, const float* mu, const float* spks, const float* cond, int Tmax, int cfg) {
  if (c.h->eng.is_bf16()) launch_pack<bf16>(c, x, mu, spks, cond, Tmax, cfg);
  else launch_pack<float>(c, x, mu, spks, cond, Tmax, cfg);
}

// JYUTVOICE_B200_SPLIT=1 (opt-in): a large solve runs as two concurrent half-batches.  Measured on B200 (batch 64 x 300,
// same box, alternating runs): 2396 / 2400 audio-s/s unsplit against 2306 / 2311 split — the kernels' fixed costs
// (prologue, resident weight loads, pipeline fill) double while the tail overlap recovers less than that.
static bool use_split() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("JYUTVOICE_B200_SPLIT");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// A solve over B utterances may run as two independent half-batches on two streams: utterances do not interact, and
// a grid of one persistent CTA per SM leaves most SMs idle while its last partial wave finishes (302 m-tiles on 148
// SMs = 2.04 waves: a third round on six SMs).  With two chains the other half's next kernel takes the freed SMs.
// Returns the first utterance of the second part (0 = no split).  Only when each half still fills the machine.
static int plan_split(int B, const int32_t* lens, int num_sms) {
  if (!use_split() || B < 2) return 0;
  long total = 0;
  for (int b = 0; b < B; ++b) total += 2L * (lens[b] + EST_GAP);
  if (total < 2L * num_sms * 128) return 0;
  long acc = 0;
  for (int b = 0; b < B - 1; ++b) {
    acc += 2L * (lens[b] + EST_GAP);
    if (2 * acc >= total) return b + 1;
  }
  return B - 1;
}

struct SolvePart {
  int b0 = 0, B = 0;
  FwdCtx c;
  int* lens_dev = nullptr;
  float* dts_dev = nullptr;
};

static void carve_part(SolvePart& pt, jv_estimator* h, Arena& ar, const int32_t* lens_host, int* tmax_len, EstLayout* Lout) {
  const int R = 2 * pt.B;
  std::vector<int32_t> l2(R);
  int tl = 0;
  for (int b = 0; b < pt.B; ++b) {
    l2[2 * b] = l2[2 * b + 1] = lens_host[pt.b0 + b];
    tl = std::max(tl, lens_host[pt.b0 + b]);
  }
  EstLayout L = make_layout(R, l2.data());
  pt.c.h = h;
  pt.c.b = carve(ar, h->eng, L.M_alloc, R, 64);
  pt.lens_dev = ar.alloc<int>(R);
  pt.dts_dev = ar.alloc<float>(64);
  pt.c.M = L.M; pt.c.M_alloc = L.M_alloc; pt.c.R = R; pt.c.Tmax_len = tl;
  pt.c.n_att_items = (int)L.att_items.size();
  pt.c.valid_frames = 0;
  for (int r = 0; r < R; ++r) pt.c.valid_frames += L.row_len[r];
  if (tmax_len) *tmax_len = tl;
  if (Lout) *Lout = std::move(L);
}

static size_t workspace_bytes(const jv_estimator* h, int R, const int32_t* lens, int nt) {
  EstLayout L = make_layout(R, lens);
  Arena ar(nullptr, 0);
  carve(ar, h->eng, L.M_alloc, R, nt);
  ar.alloc<int>(R);  // lens copy (solver)
  return ar.off + 256;
}

static size_t solve_workspace_bytes(jv_estimator* h, int B, const int32_t* lens_host) {
  Arena ar(nullptr, 0);
  const int bs = plan_split(B, lens_host, h->eng.num_sms);
  SolvePart parts[2];
  const int n_parts = bs ? 2 : 1;
  parts[0].b0 = 0; parts[0].B = bs ? bs : B;
  parts[1].b0 = bs; parts[1].B = B - bs;
  for (int i = 0; i < n_parts; ++i) carve_part(parts[i], h, ar, lens_host, nullptr, nullptr);
  size_t need = ar.off;
  if (n_parts == 2) {  // the unsplit plan (used while profiling) must fit as well
    Arena one(nullptr, 0);
    SolvePart whole;
    whole.b0 = 0; whole.B = B;
    carve_part(whole, h, one, lens_host, nullptr, nullptr);
    need = std::max(need, one.off);
  }
  return need + 256;
}

}  // namespace jv

// =========================================================================================== C ABI


extern "C" {

int jv_estimator_create(int device, int precision, jv_estimator** out) {
  JV_API_BEGIN
  JV_REQUIRE(out != nullptr, JV_ERR_INVALID, "out is NULL");
  std::unique_ptr<jv_estimator> h(new jv_estimator());
  h->eng.init(device, precision);
  *out = h.release();
  JV_API_END
}

void jv_estimator_destroy(jv_estimator* h) {
  if (h) {
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->cap_stream2) cudaStreamDestroy(h->cap_stream2);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->sat_host) cudaFreeHost(h->sat_host);
  }
  delete h;
}

int jv_estimator_set_chunk(jv_estimator* h, int chunk_size) {
  JV_API_BEGIN
  JV_REQUIRE(h && chunk_size >= 0, JV_ERR_INVALID, "bad arguments");
  h->chunk = chunk_size;
  JV_API_END
}

int jv_debug_attention_trace(void* buf) {  // development aid, not part of include/jyutvoice_b200.h
  attention_trace_buffer() = (long long*)buf;
  return JV_OK;
}

int jv_debug_gemm_trace(void* buf, int epi) {  // development aid, not part of include/jyutvoice_b200.h
  gemm_trace().buf = (long long*)buf;
  gemm_trace().epi = epi;
  return JV_OK;
}

int jv_estimator_set_stream_format(jv_estimator* h, int format) {
  JV_API_BEGIN
  JV_REQUIRE(h && format >= 0 && format <= 2, JV_ERR_INVALID, "format must be 0 (fp16), 1 (bf16) or 2 (fp32)");
  h->stream_fmt = format;
  JV_API_END
}

int jv_estimator_saturation_count(jv_estimator* h, int synchronize, int64_t* count) {
  JV_API_BEGIN
  JV_REQUIRE(h && h->finalized && count, JV_ERR_INVALID, "bad arguments");
  if (synchronize) {
    JV_CUDA(cudaSetDevice(h->eng.device));
    JV_CUDA(cudaDeviceSynchronize());
    JV_CUDA(cudaMemcpy(h->sat_host, h->sat_dev, sizeof(int), cudaMemcpyDeviceToHost));
  }
  *count = *h->sat_host;
  JV_API_END
}

int jv_estimator_time_embedding(jv_estimator* h, const float* t_host, int n, float* out, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "estimator not finalised");
  JV_REQUIRE(t_host && out && n >= 1 && n <= 4096, JV_ERR_INVALID, "bad arguments");
  JV_CUDA(cudaSetDevice(h->eng.device));
  DeviceAlloc tmp;  // test / inspection hook: scratch is allocated here, not taken from a workspace
  FwdCtx c;
  c.h = h;
  c.st = (cudaStream_t)stream;
  c.b.tsin = (float*)tmp.raw((size_t)n * 320 * sizeof(float));
  c.b.th1 = (float*)tmp.raw((size_t)n * 1024 * sizeof(float));
  c.b.th2 = (float*)tmp.raw((size_t)n * 1024 * sizeof(float));
  c.b.temb = out;
  run_time_embedding(c, t_host, n);
  JV_CUDA(cudaStreamSynchronize(c.st));  // scratch dies with this frame
  JV_API_END
}

int jv_estimator_set_weight(jv_estimator* h, const char* key, const float* data, const int64_t* shape, int ndim) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  JV_REQUIRE(!h->finalized, JV_ERR_STATE, "estimator already finalised");
  JV_CUDA(cudaSetDevice(h->eng.device));
  h->store.set(key, data, shape, ndim);
  JV_API_END
}

int jv_estimator_finalize(jv_estimator* h) {
  JV_API_BEGIN
  JV_REQUIRE(h != nullptr, JV_ERR_INVALID, "handle is NULL");
  finalize_impl(h);
  JV_API_END
}

size_t jv_cfm_workspace_bytes(const jv_estimator* h, int n_rows, const int32_t* lens_host) {
  try {
    if (!h || n_rows < 1 || !lens_host) return 0;
    return workspace_bytes(h, n_rows, lens_host, n_rows);
  } catch (const std::exception& e) {
    jv::set_last_error(e.what());
    return 0;
  }
}

size_t jv_cfm_solve_workspace_bytes(const jv_estimator* h, int B, const int32_t* lens_host) {
  try {
    if (!h || B < 1 || !lens_host) return 0;
    for (int b = 0; b < B; ++b)
      if (lens_host[b] < 1) return 0;
    return solve_workspace_bytes(const_cast<jv_estimator*>(h), B, lens_host);
  } catch (const std::exception& e) {
    jv::set_last_error(e.what());
    return 0;
  }
}

int jv_estimator_forward(jv_estimator* h, int R, int Tmax, const int32_t* lens_host, const float* x, const float* mu,
                         const float* t_host, const float* spks, const float* cond, float* out, void* ws, size_t ws_bytes,
                         void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "estimator not finalised");
  JV_REQUIRE(R >= 1 && Tmax >= 1 && lens_host && x && mu && t_host && out && ws, JV_ERR_INVALID, "bad arguments");
  JV_CUDA(cudaSetDevice(h->eng.device));
  EstLayout L = make_layout(R, lens_host);
  int tmax_len = 0;
  for (int r = 0; r < R; ++r) {
    JV_REQUIRE(lens_host[r] <= Tmax, JV_ERR_INVALID, "lens[%d] = %d exceeds Tmax = %d", r, lens_host[r], Tmax);
    tmax_len = std::max(tmax_len, lens_host[r]);
  }
  Arena ar(ws, ws_bytes);
  FwdCtx c;
  c.h = h;
  c.b = carve(ar, h->eng, L.M_alloc, R, R);
  int* lens_dev = ar.alloc<int>(R);
  c.M = L.M; c.M_alloc = L.M_alloc; c.R = R; c.Tmax_len = tmax_len;
  c.n_att_items = (int)L.att_items.size();
  c.valid_frames = 0;
  for (int r = 0; r < R; ++r) c.valid_frames += L.row_len[r];
  c.st = (cudaStream_t)stream;
  c.temb_step = c.b.temb;
  std::vector<int> tidx(R);
  for (int r = 0; r < R; ++r) tidx[r] = r;
  JV_CUDA(cudaMemcpyAsync(lens_dev, lens_host, R * sizeof(int), cudaMemcpyHostToDevice, c.st));
  upload_layout(c, L, tidx);
  run_time_embedding(c, t_host, R);
  run_pack(c, x, mu, spks, cond, Tmax, 0);
  forward_packed(c);
  const long n = (long)R * 80 * Tmax;
  unpack_output_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c.st>>>(out, c.b.V, 80, c.b.row_off, lens_dev, R, Tmax);
  JV_LAUNCHED();
  JV_CUDA(cudaMemcpyAsync(h->sat_host, h->sat_dev, sizeof(int), cudaMemcpyDeviceToHost, c.st));
  JV_API_END
}

int jv_cfm_solve(jv_estimator* h, int B, int Tmax, const int32_t* lens_host, const float* mu, const float* spks,
                 const float* cond, const float* noise, int64_t noise_stride, float temperature, int n_timesteps,
                 const float* t_span_host, float cfg_rate, float* out_mel, void* ws, size_t ws_bytes, void* stream) {
  JV_API_BEGIN
  JV_REQUIRE(h && h->finalized, JV_ERR_STATE, "estimator not finalised");
  JV_REQUIRE(B >= 1 && Tmax >= 1 && lens_host && mu && spks && noise && t_span_host && out_mel && ws, JV_ERR_INVALID, "bad arguments");
  JV_REQUIRE(n_timesteps >= 1 && n_timesteps <= 64, JV_ERR_INVALID, "n_timesteps = %d out of range [1, 64]", n_timesteps);
  JV_CUDA(cudaSetDevice(h->eng.device));
  for (int b = 0; b < B; ++b) {
    JV_REQUIRE(lens_host[b] >= 1 && lens_host[b] <= Tmax, JV_ERR_INVALID, "lens[%d] = %d outside [1, Tmax = %d]", b, lens_host[b], Tmax);
    JV_REQUIRE(lens_host[b] <= noise_stride, JV_ERR_INVALID, "lens[%d] = %d exceeds the noise bank (%lld frames)", b, lens_host[b],
               (long long)noise_stride);
  }
  const cudaStream_t user_st = (cudaStream_t)stream;
  // Euler bookkeeping exactly as flow_matching.py:230-263 (t and dt accumulate in fp32)
  std::vector<float> ts(n_timesteps), dts(n_timesteps);
  {
    float t = t_span_host[0];
    float dt = t_span_host[1] - t_span_host[0];
    for (int k = 1; k <= n_timesteps; ++k) {
      ts[k - 1] = t;
      dts[k - 1] = dt;
      t = t + dt;
      if (k < n_timesteps) dt = t_span_host[k + 1] - t;
    }
  }
  // one or two independent half-batches (plan_split)
  // (not while per-launch events are being recorded: a kernel's duration would include waiting for the other chain's SMs)
  const int bs = profile_state().on ? 0 : plan_split(B, lens_host, h->eng.num_sms);
  const int n_parts = bs ? 2 : 1;
  SolvePart parts[2];
  parts[0].b0 = 0; parts[0].B = bs ? bs : B;
  parts[1].b0 = bs; parts[1].B = B - bs;
  Arena ar(ws, ws_bytes);
  for (int i = 0; i < n_parts; ++i) {  // set-up of both parts on the caller's stream
    SolvePart& pt = parts[i];
    EstLayout L;
    carve_part(pt, h, ar, lens_host, nullptr, &L);
    pt.c.st = user_st;
    pt.c.temb_step = pt.c.b.temb;
    std::vector<int> tidx(pt.c.R, 0);
    JV_CUDA(cudaMemcpyAsync(pt.lens_dev, lens_host + pt.b0, pt.B * sizeof(int), cudaMemcpyHostToDevice, user_st));
    JV_CUDA(cudaMemcpyAsync(pt.dts_dev, dts.data(), n_timesteps * sizeof(float), cudaMemcpyHostToDevice, user_st));
    upload_layout(pt.c, L, tidx);                      // synchronises: the host vectors die with this frame
    run_time_embedding(pt.c, ts.data(), n_timesteps);  // all steps at once: temb depends on t only
  }
  auto part_ptr = [&](const float* base, const SolvePart& pt, long per_utt) { return base ? base + (long)pt.b0 * per_utt : nullptr; };
  // One Euler step of one part = pack, estimator forward (~330 launches), CFG + update, advance.  What differs between
  // steps (the time-embedding row and dt) is selected on the device through row_tidx (= step index for every row), so
  // the launch sequence is identical for all steps: step 0 runs eagerly, step 1 is captured into a CUDA graph (both
  // parts as two branches), steps 1 .. n-1 replay it.
  auto euler_step = [&](SolvePart& pt) {
    FwdCtx& c = pt.c;
    float* x = out_mel + (long)pt.b0 * 80 * Tmax;
    const long nx = (long)pt.B * 80 * Tmax;
    run_pack(c, x, part_ptr(mu, pt, 80L * Tmax), part_ptr(spks, pt, 80), part_ptr(cond, pt, 80L * Tmax), Tmax, 1);
    forward_packed(c);
    cfg_euler_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, c.st>>>(x, c.b.V, 80, c.b.row_off, pt.lens_dev, pt.B, Tmax, pt.dts_dev,
                                                                     c.b.row_tidx, cfg_rate);
    JV_LAUNCHED();
    step_advance_kernel<<<cdiv(c.R, 256), 256, 0, c.st>>>(c.b.row_tidx, c.R);
    JV_LAUNCHED();
  };
  // fork: part 1 runs on the handle's auxiliary stream
  if (n_parts == 2) {
    if (!h->aux_stream) JV_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    if (!h->ev_fork) JV_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    if (!h->ev_join) JV_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    JV_CUDA(cudaEventRecord(h->ev_fork, user_st));
    JV_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
    parts[1].c.st = h->aux_stream;
  }
  for (int i = 0; i < n_parts; ++i) {
    SolvePart& pt = parts[i];
    const long nx = (long)pt.B * 80 * Tmax;
    init_noise_kernel<<<(unsigned)((nx + 255) / 256), 256, 0, pt.c.st>>>(out_mel + (long)pt.b0 * 80 * Tmax, noise, noise_stride,
                                                                          pt.lens_dev, pt.B, Tmax, temperature);
    JV_LAUNCHED();
    euler_step(pt);
  }
  auto join = [&]() {
    if (n_parts == 2) {
      JV_CUDA(cudaEventRecord(h->ev_join, h->aux_stream));
      JV_CUDA(cudaStreamWaitEvent(user_st, h->ev_join, 0));
    }
  };
  int k = 1;
  if (n_timesteps >= 3 && use_graph() && !profile_state().on) {
    join();  // the graph is launched on the caller's stream only
    cudaGraphExec_t exec = nullptr;
    const uint64_t before = g_launch_count.load();
    bool captured = false;
    do {  // capture one step of every part: the parts are branches forked from / joined into the capture origin
      if (!h->cap_stream && cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking) != cudaSuccess) break;
      if (n_parts == 2 && !h->cap_stream2 && cudaStreamCreateWithFlags(&h->cap_stream2, cudaStreamNonBlocking) != cudaSuccess) break;
      if (cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) break;
      bool ok = true;
      try {
        if (n_parts == 2) {
          JV_CUDA(cudaEventRecord(h->ev_fork, h->cap_stream));
          JV_CUDA(cudaStreamWaitEvent(h->cap_stream2, h->ev_fork, 0));
        }
        for (int i = 0; i < n_parts; ++i) {
          parts[i].c.st = i == 0 ? h->cap_stream : h->cap_stream2;
          euler_step(parts[i]);
        }
        if (n_parts == 2) {
          JV_CUDA(cudaEventRecord(h->ev_join, h->cap_stream2));
          JV_CUDA(cudaStreamWaitEvent(h->cap_stream, h->ev_join, 0));
        }
      } catch (const std::exception&) {
        ok = false;
      }
      cudaGraph_t graph = nullptr;
      if (cudaStreamEndCapture(h->cap_stream, &graph) != cudaSuccess || !graph) ok = false;
      if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
      if (graph) cudaGraphDestroy(graph);
      captured = ok;
    } while (false);
    if (!captured) cudaGetLastError();
    const uint64_t nodes = g_launch_count.load() - before;
    g_launch_count.fetch_sub(nodes);  // a captured launch does not run; every replay runs `nodes` kernels
    parts[0].c.st = user_st;
    if (captured) {
      cudaError_t rc = cudaSuccess;
      for (; k < n_timesteps && rc == cudaSuccess; ++k) {
        rc = cudaGraphLaunch(exec, user_st);
        g_launch_count.fetch_add(nodes);
        g_graph_launches.fetch_add(1);
      }
      cudaGraphExecDestroy(exec);  // launches are asynchronous: the runtime defers the destruction until they have run
      JV_CUDA(rc);
      n_timesteps = k;  // nothing left for the eager loop; the parts are already joined
      if (n_parts == 2) parts[1].c.st = user_st;
    } else if (n_parts == 2) {  // eager fallback: fork again
      JV_CUDA(cudaEventRecord(h->ev_fork, user_st));
      JV_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
      parts[1].c.st = h->aux_stream;
    }
  }
  if (k < n_timesteps) {
    for (; k < n_timesteps; ++k)
      for (int i = 0; i < n_parts; ++i) euler_step(parts[i]);
    join();
  } else if (!(n_timesteps >= 3 && use_graph() && !profile_state().on)) {
    join();
  }
  JV_CUDA(cudaMemcpyAsync(h->sat_host, h->sat_dev, sizeof(int), cudaMemcpyDeviceToHost, user_st));
  JV_API_END
}

}  // extern "C"
