/*
 * jyutvoice_b200 — C-ABI of the B200-native CFM + HiFT hot path.
 *
 * The reference (indiejoseph/JyutVoice) is pure Python/PyTorch and has no FFI; the one seam it
 * offers for swapping the estimator is the raw-pointer engine contract in
 * jyutvoice/flow/flow_matching.py:267-297 (six device pointers x, mask, mu, t, spks, cond, result
 * written on the caller's CUDA stream).  This header is what a binding for that seam, and for
 * HiFTGenerator.inference/decode, calls.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every function returns 0 (JV_OK) or a negative JV_ERR_* code; jv_last_error() gives the text
 *     (thread-local).  No exception crosses the ABI.
 *   - "dev" pointers are device memory on the handle's device; "host" pointers are host memory.
 *   - all tensors are dense, row-major, float32 unless noted; lens are int32 on the host.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised, except
 *     where noted (set_weight / finalize synchronise the device).
 *   - the caller owns every buffer, including the workspace arena sized by *_workspace_bytes().
 *   - a handle is bound to one device and must not be used from two threads at once.
 *   - precision: JV_PREC_FP32 computes every contraction in fp32 FFMA (the <=1e-3 / >=60 dB mode);
 *     JV_PREC_BF16 runs contractions on tcgen05 tensor cores with bf16 operands, fp32 accumulation,
 *     fp32 normalisation statistics; the estimator's residual stream is stored in 16 bits between epilogues
 *     (jv_estimator_set_stream_format), HiFT's in fp32.
 */
#ifndef JYUTVOICE_B200_H
#define JYUTVOICE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JV_OK 0
#define JV_ERR_INVALID (-1)   /* bad argument (maps to ValueError) */
#define JV_ERR_CUDA (-2)      /* CUDA runtime / driver failure (RuntimeError) */
#define JV_ERR_STATE (-3)     /* handle not finalised, missing weights, workspace too small */

#define JV_PREC_FP32 0
#define JV_PREC_BF16 1

typedef struct jv_estimator jv_estimator;
typedef struct jv_hift jv_hift;

int jv_version(void);
const char* jv_last_error(void);
/* Number of kernel launches this library has enqueued in this process (bench.py's gpu_launches). */
uint64_t jv_launch_count(void);
/* CUDA-graph replays issued by jv_cfm_solve so far (one replay = one Euler step = ~330 kernels, all of them counted in
 * jv_launch_count).  Step 0 of a solve runs eagerly, step 1 is captured, steps 1 .. n-1 replay it; JYUTVOICE_B200_GRAPH=0
 * or a failed capture falls back to eager launches. */
uint64_t jv_graph_launch_count(void);
/* bf16-mode GEMMs whose shape the tcgen05 kernel cannot run and that were lowered to the FFMA engine instead (~50x
 * slower; reported once on stderr).  0 for the configs/base.yaml graphs. */
uint64_t jv_simt_fallback_count(void);

/* ------------------------------------------------------------------------------------------
 * Estimator = CausalConditionalDecoder (jyutvoice/flow/decoder.py:798-1018) with the
 * configs/base.yaml:88-99 hyper-parameters, plus the Euler/CFG solver around it
 * (ConditionalCFM.solve_euler, flow_matching.py:215-265).
 * ------------------------------------------------------------------------------------------ */
int jv_estimator_create(int device, int precision, jv_estimator** out);
void jv_estimator_destroy(jv_estimator* h);
/* Replaces nn.Module.load_state_dict for `decoder.estimator.*`: `key` is the reference's
 * state_dict key without the "estimator." prefix (e.g. "mid_blocks.3.1.0.attn1.to_q.weight");
 * `data` may be host or device fp32; `shape` is checked against the architecture. */
int jv_estimator_set_weight(jv_estimator* h, const char* key, const float* data, const int64_t* shape, int ndim);
/* Fails (JV_ERR_STATE, listing the first missing key) unless all 910 tensors were set. */
int jv_estimator_finalize(jv_estimator* h);

/* streaming=True of CausalConditionalDecoder.forward (flow/decoder.py:950-953, utils/mask.py:91-126): attention query t
 * sees keys < min(len, (t / chunk_size + 1) * chunk_size).  chunk_size = static_chunk_size (50) turns it on, 0 (the
 * default) restores full context.  State of the handle: applies to every later jv_estimator_forward / jv_cfm_solve. */
int jv_estimator_set_chunk(jv_estimator* h, int chunk_size);

/* bf16 mode stores the residual stream between GEMM epilogues as: format 0 = fp16 with saturating stores (default: 11
 * significand bits), 1 = bf16 (8 bits, fp32 range), 2 = fp32 (for weights whose activations exceed +-65504: a stream
 * that large also out-grows 16-bit resolution, so the fallback is fp32, at ~5 % of the throughput). */
int jv_estimator_set_stream_format(jv_estimator* h, int format);
/* Stores into the fp16 stream that hit the format's largest finite value (+-65504, saturating convert) since finalize,
 * counted per (row, column share) of a GEMM tile: 0 means nothing was clipped.  synchronize = 0 returns the value as of the last completed forward / solve (a pinned
 * host copy refreshed on the caller's stream); synchronize != 0 waits for the device first. */
int jv_estimator_saturation_count(jv_estimator* h, int synchronize, int64_t* count);
/* The time conditioning of n timesteps t (decoder.py:15-30 SinusoidalPosEmb, :127-171 TimestepEmbedding, :101-103 the
 * resnets' Mish -> Linear): out dev [n, 14, 256], resnet order down, mid 0..11, up.  Inspection / test hook (the
 * solver computes the same table once per solve); synchronises the stream. */
int jv_estimator_time_embedding(jv_estimator* h, const float* t_host, int n, float* out, void* stream);

/* Workspace for B utterances (R = 2B estimator rows with CFG) of the given lengths. */
size_t jv_cfm_workspace_bytes(const jv_estimator* h, int n_rows, const int32_t* lens_host);

/* One estimator evaluation, the reference's forward(x, mask, mu, t, spks, cond) (decoder.py:917):
 * x, mu, cond dev [R,80,Tmax]; spks dev [R,80]; t host [R]; mask is implied by lens (prefix masks,
 * which is all the reference ever builds: jyutvoice_tts.py:225,229).  cond / spks may be NULL
 * (treated as zeros).  out dev [R,80,Tmax]; frames >= lens[r] are written as 0. */
int jv_estimator_forward(jv_estimator* h, int R, int Tmax, const int32_t* lens_host,
                         const float* x, const float* mu, const float* t_host, const float* spks,
                         const float* cond, float* out, void* ws, size_t ws_bytes, void* stream);

/* CausalConditionalCFM.forward (flow_matching.py:356-401) for a ragged batch of B utterances:
 *   x0[b] = noise[:, :len_b] * temperature      (noise dev [80, noise_stride], the seed-0 bank)
 *   for k in 0..n-1: v = estimator(CFG pair); x += (t_span[k+1]-t_span[k]) * ((1+cfg)*v_c - cfg*v_u)
 * mu, cond dev [B,80,Tmax] (cond may be NULL); spks dev [B,80]; t_span host [n_timesteps+1];
 * out_mel dev [B,80,Tmax], frames >= len_b are 0.  Workspace from jv_cfm_workspace_bytes(h, 2B, lens
 * repeated per CFG pair) — or simply jv_cfm_solve_workspace_bytes(). */
size_t jv_cfm_solve_workspace_bytes(const jv_estimator* h, int B, const int32_t* lens_host);
int jv_cfm_solve(jv_estimator* h, int B, int Tmax, const int32_t* lens_host, const float* mu,
                 const float* spks, const float* cond, const float* noise, int64_t noise_stride,
                 float temperature, int n_timesteps, const float* t_span_host, float cfg_rate,
                 float* out_mel, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * HiFT vocoder = HiFTGenerator (jyutvoice/hifigan/generator.py:239-466) + ConvRNNF0Predictor
 * (f0_predictor.py:19-55), configs/base.yaml:26-48.
 * ------------------------------------------------------------------------------------------ */
int jv_hift_create(int device, int precision, jv_hift** out);
void jv_hift_destroy(jv_hift* h);
/* `key` is the reference's HiFTGenerator state_dict key (weight-norm pairs are folded here). */
int jv_hift_set_weight(jv_hift* h, const char* key, const float* data, const int64_t* shape, int ndim);
int jv_hift_finalize(jv_hift* h);

size_t jv_hift_workspace_bytes(const jv_hift* h, int B, const int32_t* lens_host);

/* f0 = |classifier(condnet(mel))| (f0_predictor.py:52-55). mel dev [B,80,Tmax] -> f0 dev [B,Tmax]. */
int jv_hift_f0(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* mel, float* f0,
               void* ws, size_t ws_bytes, void* stream);
/* Source module (generator.py:459-461,141-176,220-236): f0 dev [B,Tmax]; phase dev [B,9] (the
 * Uniform(-pi,pi) draw, entry 0 is ignored and treated as 0); noise dev [B,9,480*Tmax] (the
 * randn_like draw) -> s dev [B,480*Tmax].  The caller draws the RNG exactly as the reference does. */
int jv_hift_source(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* f0,
                   const float* phase, const float* noise, float* s, void* ws, size_t ws_bytes, void* stream);
/* HiFTGenerator._stft (generator.py:371-381): s dev [B,480*Tmax] -> out dev [B,18,120*Tmax+1] = [real(9) | imag(9)],
 * frames beyond 120*len_b + 1 are 0 (each utterance reflect-padded at its own length).  Inspection / test hook: decode
 * computes the same tensor internally (in bf16 mode it is rounded to bf16 there and here). */
int jv_hift_stft(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* s, float* out,
                 void* ws, size_t ws_bytes, void* stream);
/* HiFTGenerator.decode(x=mel, s) (generator.py:396-432): -> wav dev [B,480*Tmax]; samples beyond
 * 480*len_b are 0.  Each utterance is decoded with its own zero boundary, i.e. equals the
 * reference's unpadded batch-1 call. */
int jv_hift_decode(jv_hift* h, int B, int Tmax, const int32_t* lens_host, const float* mel,
                   const float* s, float* wav, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Text front of JyutVoiceTTS.synthesise (jyutvoice/models/jyutvoice_tts.py:175-203), batched over ragged utterances:
 * TextEncoder (models/text_encoder.py:340-451: embeddings, ConvReluNorm prenet, 6 RoPE attention layers, proj),
 * DurationPredictor (models/duration_predictor.py:26-60) and the length regulator.  fp32 in both precision modes (the
 * durations pass through ceil()).  One handle holds the encoder's weights, the predictor's, or both.
 * ------------------------------------------------------------------------------------------ */
typedef struct jv_text jv_text;
int jv_text_create(int device, jv_text** out);
void jv_text_destroy(jv_text* h);
/* `key` = "encoder." + TextEncoder.state_dict() key, or "dp." + DurationPredictor.state_dict() key (the names they have
 * inside JyutVoiceTTS.state_dict()).  Embedding tables may have any number of rows (n_vocab, n_lang, n_tone). */
int jv_text_set_weight(jv_text* h, const char* key, const float* data, const int64_t* shape, int ndim);
int jv_text_finalize(jv_text* h);
size_t jv_text_workspace_bytes(const jv_text* h, int B, int Tx, const int32_t* x_lens_host);
/* TextEncoder.forward(x, x_lengths, lang, tone, word_pos, syllable_pos, spk_embed) (text_encoder.py:401-451):
 * token streams dev int64 [B,Tx], spk_embed dev [B,192] -> out_x dev [B,576,Tx], out_mu dev [B,80,Tx]; positions beyond
 * x_lens[b] are 0 (the reference's `* x_mask`). */
int jv_text_encode(jv_text* h, int B, int Tx, const int32_t* x_lens_host, const int64_t* x, const int64_t* lang,
                   const int64_t* tone, const int64_t* word_pos, const int64_t* syllable_pos, const float* spk_embed,
                   float* out_x, float* out_mu, void* ws, size_t ws_bytes, void* stream);
/* DurationPredictor.forward(x, x_mask, g) (duration_predictor.py:48-60): x dev [B,576,Tx], spk_embed dev [B,192]
 * -> out_logw dev [B,1,Tx], 0 beyond x_lens[b]. */
int jv_text_durations(jv_text* h, int B, int Tx, const int32_t* x_lens_host, const float* x, const float* spk_embed,
                      float* out_logw, void* ws, size_t ws_bytes, void* stream);
/* Length regulation (jyutvoice_tts.py:184-203, utils/model.py:29-46), on the current device.  Step 1:
 * w_ceil = ceil(exp(logw) * mask) * length_scale, cum = cumsum(w_ceil) (dev [B,Tx]), y_lengths = max(sum, 1) truncated
 * (dev int64 [B]; the caller reads them to size the next call).  Step 2: frame t of utterance b copies the token i with
 * cum[i-1] <= t < cum[i]: mu_y dev [B,80,Ty] = the gather form of attn^T @ mu_x, frame_token dev int32 [B,Ty] = i or -1. */
int jv_length_durations(int B, int Tx, const int32_t* x_lens_dev, const float* logw, float length_scale, float* cum,
                        int64_t* y_lengths, void* stream);
int jv_length_align(int B, int Tx, int Ty, const int32_t* x_lens_dev, const int64_t* y_lengths, const float* cum,
                    const float* mu_x, float* mu_y, int32_t* frame_token, void* stream);

/* ------------------------------------------------------------------------------------------
 * Speech-token encoder that produces `prompt_h` for voice cloning, batched over ragged
 * utterances, every utterance computed as the reference's batch-1 call computes it:
 *   FlowEncoder.forward (infer.py:66-82): embedding(clamp(token, 0)) * mask -> encoder -> Linear(512, 80)
 *   UpsampleConformerEncoder.forward (jyutvoice/transformer/upsample_encoder.py:290-355) with the hyper-parameters of
 *   infer.py:44-60: 512 channels, 8 heads, 2048 FFN units, rel-pos attention (transformer/attention.py:196-330),
 *   pre-lookahead conv, 6 layers, x2 upsampling conv, 4 layers, after_norm.
 * Keys are those of flow_encoder.pt: `input_embedding.weight` (optional), `encoder.*`, `encoder_proj.*` (optional).  fp32.
 * ------------------------------------------------------------------------------------------ */
typedef struct jv_flowenc jv_flowenc;
int jv_flowenc_create(int device, jv_flowenc** out);
void jv_flowenc_destroy(jv_flowenc* h);
int jv_flowenc_set_weight(jv_flowenc* h, const char* key, const float* data, const int64_t* shape, int ndim);
int jv_flowenc_finalize(jv_flowenc* h);
size_t jv_flowenc_workspace_bytes(const jv_flowenc* h, int B, int T, const int32_t* lens_host);
/* Exactly one of `token` ([B, T] int64, device; needs input_embedding) and `xs` ([B, T, 512] fp32, device: the encoder called
 * on features, upsample_encoder.py:290).  chunk = 0: full context (streaming=False); chunk > 0: the static chunk mask of
 * streaming=True (utils/mask.py:161-200; the reference uses static_chunk_size = 25, doubled after the upsampling).
 * out_hidden [B, 2T, 512] (after_norm output) and / or out_h [B, 2T, 80] (needs encoder_proj); rows >= 2 * lens[b] are 0. */
int jv_flowenc_encode(jv_flowenc* h, int B, int T, const int32_t* lens_host, const int64_t* token, const float* xs, int chunk,
                      float* out_hidden, float* out_h, void* ws, size_t ws_bytes, void* stream);

/* Profiling of the dominant kernel (the tcgen05 GEMM): between begin and end every launch of it is
 * bracketed by CUDA events on its own stream.  end() synchronises the device and returns the summed
 * kernel time, the summed algorithmic FLOPs (valid frames only) and the launch count. */
int jv_profile_begin(void);
int jv_profile_end(double* kernel_ms, double* algo_flops, int64_t* launches);

/* Micro-benchmark hook: average milliseconds of one tcgen05 GEMM launch, C[M,N] = A[M,K] W[N,K]^T over `taps`
 * shifted taps (K per tap), timed with CUDA events over `iters` launches after 3 warm-ups.
 * mode bits: 1 = fp32 residual in + fp32 out, 2 = GELU, 4 = LayerNorm+Mish before the output, 8 = second LayerNorm output,
 * 16 = bf16 output.  Buffers are allocated and freed inside. */
int jv_bench_gemm(int M, int N, int K_tap, int taps, int mode, int iters, double* ms_out);

/* Test hook: C[M,N] = A[M,K] * W[N,K]^T (+bias) through the same GEMM engine the handles use
 * (precision selects fp32 FFMA or bf16 tcgen05).  A, W, bias, C: dev fp32; operands are rounded to
 * bf16 inside when precision == JV_PREC_BF16. */
int jv_test_gemm(int precision, int M, int N, int K, const float* A, const float* W, const float* bias,
                 float* C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JYUTVOICE_B200_H */
