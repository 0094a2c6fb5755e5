/*
 * s2s_b200.h -- C ABI of libs2s_b200.so, the B200 (sm_100a) engine behind
 * huggingface/speech-to-speech's STT / LLM / TTS handler slots.
 *
 * The reference has no FFI of its own (pure Python); the calls this library replaces are
 * the third-party model-runtime calls inside the three handler process() methods:
 *
 *   s2s_whisper_*   <- processor(...) + model.generate(...)            S/STT/whisper_stt_handler.py:83-87, 166-197, 243
 *   s2s_llama_*     <- pipeline("text-generation")(prompt, ...)        S/LLM/language_model.py:800-892
 *   s2s_tts_*       <- _resample_to_pipeline_sr/_to_int16/_stream      S/TTS/qwen3_tts_handler.py:612-613, 674-680, 695-749
 *   (S/ = /root/reference/src/speech_to_speech)
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success or a negative
 * S2S_ERR_* code (message via s2s_last_error(), thread-local).  Pointers suffixed _d are
 * device pointers, _h host pointers.  `stream` is a cudaStream_t passed as void*; device
 * work is enqueued on it and is asynchronous unless the call documents a host result.
 * The library owns its weights and workspaces (allocated at create/finalize); it never
 * frees caller memory.  A model handle may be used from one thread at a time; different
 * handles are independent.
 */
#ifndef S2S_B200_H
#define S2S_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2S_OK 0
#define S2S_ERR_INVALID -1   /* bad argument / shape / state */
#define S2S_ERR_CUDA -2      /* CUDA runtime or driver error */
#define S2S_ERR_NOT_FOUND -3 /* unknown tensor name */
#define S2S_ERR_UNSUPPORTED -4

/* element types for bind_tensor sources */
#define S2S_F32 0
#define S2S_F16 1
#define S2S_BF16 2

typedef struct s2s_ctx s2s_ctx;
typedef struct s2s_whisper s2s_whisper;
typedef struct s2s_llama s2s_llama;

/* ---- context ------------------------------------------------------------------------- */
int s2s_init(int device, s2s_ctx** out);
int s2s_destroy(s2s_ctx* ctx);
const char* s2s_last_error(void);
/* number of kernels this library launched on this thread's context since the last reset */
int64_t s2s_launch_count(s2s_ctx* ctx, int reset);

/* ---- Whisper (STT) ---------------------------------------------------------------------
 * Geometry mirrors transformers WhisperConfig.  compute_dtype: S2S_F16 (reference default
 * stt_torch_dtype=float16) or S2S_BF16 = storage type of weights and GEMM operands;
 * accumulation, LayerNorm, softmax and the residual stream are fp32.                      */
typedef struct {
  int32_t d_model, heads, enc_layers, dec_layers, ffn, n_mels, vocab;
  int32_t max_source_positions; /* 1500 */
  int32_t max_target_positions; /* 448  */
  int32_t compute_dtype;
  int32_t max_batch; /* concurrent utterances per call */
} s2s_whisper_config;

int s2s_whisper_create(s2s_ctx* ctx, const s2s_whisper_config* cfg, s2s_whisper** out);
int s2s_whisper_destroy(s2s_whisper* m);
/* Copy one parameter (transformers state-dict name, e.g. "model.encoder.conv1.weight") from
 * HOST memory into the library's arena, converting to compute_dtype and to the kernel layout. */
int s2s_whisper_bind_tensor(s2s_whisper* m, const char* name, const void* data_h, const int64_t* shape,
                            int32_t ndim, int32_t dtype);
/* Fill every parameter with seeded pseudo-random values on the device (bench only). */
int s2s_whisper_init_random(s2s_whisper* m, uint64_t seed);
int s2s_whisper_finalize(s2s_whisper* m);

/* decoding controls = what WhisperGenerationMixin.generate derives from generation_config */
typedef struct {
  const int32_t* prefix_h; /* forced decoder prompt, e.g. [sot, lang, transcribe, notimestamps] */
  int32_t n_prefix;
  int32_t max_new_tokens;
  int32_t eos_id;
  const int32_t* suppress_h; /* SuppressTokensLogitsProcessor */
  int32_t n_suppress;
  const int32_t* begin_suppress_h; /* SuppressTokensAtBeginLogitsProcessor */
  int32_t n_begin_suppress;
} s2s_whisper_decode_opts;

/* pcm_d: [B, pcm_stride] f32 mono 16 kHz, n_samples_h[b] valid samples each (rest ignored, padded
 * with zeros to 30 s like WhisperFeatureExtractor).  mel_out_d: optional [B, n_mels, 3000] f32. */
int s2s_whisper_logmel(s2s_whisper* m, const float* pcm_d, int64_t pcm_stride, const int32_t* n_samples_h,
                       int32_t B, float* mel_out_d, void* stream);
/* mel_in_d: optional [B, n_mels, 3000] f32 (NULL = use the features left by s2s_whisper_logmel).
 * enc_out_d: optional [B, 1500, d_model] f32 copy of the encoder output (after final LayerNorm).
 * Also computes the decoder cross-attention K/V for every decoder layer.                     */
int s2s_whisper_encode(s2s_whisper* m, const float* mel_in_d, int32_t B, float* enc_out_d, void* stream);
/* Greedy decode of the B utterances encoded last.  ids_out_d: [B, max_new_tokens] i32 (generated
 * tokens incl. EOS, then EOS padding), len_out_d: [B] i32.  forced_d (optional, [B, max_new_tokens]):
 * teacher-forced feedback tokens for parity tests.  logits_out_d (optional): [max_new_tokens, B, vocab]
 * f32 processed logits of every generation step (B <= s2s_whisper_max_decode_batch).  Up to 16 sessions share one
 * persistent launch (larger B is split); a single session uses the thread-block-cluster kernel.     */
int s2s_whisper_decode(s2s_whisper* m, const s2s_whisper_decode_opts* opts, int32_t B, int32_t* ids_out_d,
                       int32_t* len_out_d, const int32_t* forced_d, float* logits_out_d, void* stream);
/* One decoder step from <|sot|>, logits restricted to lang_ids -> lang_out_d[B] (detect_language). */
int s2s_whisper_detect_language(s2s_whisper* m, int32_t sot_id, const int32_t* lang_ids_h, int32_t n_lang,
                                int32_t B, int32_t* lang_out_d, void* stream);
/* Profiling aid: when trace_d != NULL the next decode launches record, for CTA 0 and the last CTA, the
 * %globaltimer (ns) at [phase begin, inputs staged, arrived at the grid barrier, next phase prepared, phase body end,
 * barrier exit] (the cluster kernel uses slots 2 and 3 for its intra-phase milestones) of the first
 * `capacity` phases into trace_d[2][capacity][6] (u64).  NULL disables tracing.                                          */
int s2s_whisper_set_trace(s2s_whisper* m, uint64_t* trace_d, int32_t capacity);
/* Sessions one persistent decode launch can carry for this geometry (<= 16; larger batches are split by the library).
 * The session batcher (speech_to_speech_b200/batcher.py) sizes its batches with it. */
int32_t s2s_whisper_max_decode_batch(s2s_whisper* m);
/* End-to-end with HOST buffers: H2D of pcm, log-mel, encode, greedy decode, D2H of ids; synchronous.
 * pcm_h: [B, pcm_stride] f32 (pinned or pageable), ids_out_h: [B, max_new_tokens], len_out_h: [B].  */
int s2s_whisper_transcribe(s2s_whisper* m, const s2s_whisper_decode_opts* opts, const float* pcm_h,
                           int64_t pcm_stride, const int32_t* n_samples_h, int32_t B, int32_t* ids_out_h,
                           int32_t* len_out_h, void* stream);

/* ---- generic GEMM (exposed for tests / roofline benches) --------------------------------
 * C[M,N] = A[M,K] * W[N,K]^T (+bias), tcgen05 + TMA; dtype S2S_F16 or S2S_BF16, fp32 accumulate.
 * out_dtype: S2S_F32 or same as dtype.  act: 0 none, 1 exact GELU.                          */
int s2s_gemm(s2s_ctx* ctx, const void* a_d, const void* w_d, const float* bias_d, void* c_d, int32_t M,
             int32_t N, int32_t K, int32_t dtype, int32_t out_dtype, int32_t act, void* stream);
/* O = softmax(Q K^T * scale [+causal]) V ; q/k/v/o row-major [T, heads*hd] slices with row strides
 * (elements); kv_heads <= heads (GQA).  hd in {64, 128}.                                     */
int s2s_attention(s2s_ctx* ctx, const void* q_d, const void* k_d, const void* v_d, void* o_d, int32_t B,
                  int32_t Tq, int32_t Tk, int32_t heads, int32_t kv_heads, int32_t hd, int64_t ldq,
                  int64_t ldk, int64_t ldv, int64_t ldo, float scale, int32_t causal, int32_t dtype,
                  void* stream);

/* ---- Llama-family LLM ------------------------------------------------------------------ */
typedef struct {
  int32_t d_model, layers, heads, kv_heads, head_dim, ffn, vocab;
  float rope_theta, rms_eps;
  int32_t compute_dtype;  /* S2S_BF16 (default) or S2S_F16 */
  int32_t max_sessions;   /* KV-cache slots */
  int32_t max_positions;  /* per session */
  int32_t max_prefill;    /* longest prompt chunk per call */
  int32_t qk_norm;        /* 1 = Qwen3-style per-head RMSNorm(head_dim) on q,k before RoPE (transformers
                             modeling_qwen3.py Qwen3Attention: the reference's default LLM family and the TTS talker); 0 = Llama */
  int32_t n_tables;       /* 0/1 = one embedding table + one output head; >1 = one pair per codebook
                             ("model.embed_tokens.<i>.weight", "lm_head.<i>.weight"), used by the TTS code predictor */
} s2s_llama_config;

int s2s_llama_create(s2s_ctx* ctx, const s2s_llama_config* cfg, s2s_llama** out);
int s2s_llama_destroy(s2s_llama* m);
int s2s_llama_bind_tensor(s2s_llama* m, const char* name, const void* data_h, const int64_t* shape,
                          int32_t ndim, int32_t dtype);
int s2s_llama_init_random(s2s_llama* m, uint64_t seed);
int s2s_llama_finalize(s2s_llama* m);
/* Reset a KV-cache slot (new response). */
int s2s_llama_session_reset(s2s_llama* m, int32_t slot);
/* Prefill `n` prompt tokens (host ids) into `slot`; logits_out_d optional [n, vocab] f32;
 * next_id_d optional [1] i32 = argmax of the last position.                                 */
int s2s_llama_prefill(s2s_llama* m, int32_t slot, const int32_t* ids_h, int32_t n, float* logits_out_d,
                      int32_t* next_id_d, void* stream);
/* Greedy decode for B (<= s2s_llama_max_decode_batch: 16, 4 for Llama-3-8B) sessions in one persistent launch.  slots_h[B]; first_ids_d[B] are the tokens to
 * feed first (the prefill argmax); ids_out_d [B, n_steps] receives the n_steps tokens generated AFTER them; eos stops
 * a row (eos_id < 0 disables); forced_d optional [B, n_steps] teacher-forced feedback; logits_out_d optional
 * [n_steps, B, vocab].                                                                              */
int s2s_llama_decode(s2s_llama* m, const int32_t* slots_h, int32_t B, const int32_t* first_ids_d,
                     int32_t n_steps, int32_t eos_id, int32_t* ids_out_d, int32_t* len_out_d,
                     const int32_t* forced_d, float* logits_out_d, void* stream);
/* Profiling aid (see s2s_whisper_set_trace): CTA 0's [phase begin, body end, barrier exit] stamps, trace_d[capacity][3]. */
int s2s_llama_set_trace(s2s_llama* m, uint64_t* trace_d, int32_t capacity);
/* Largest B accepted by s2s_llama_decode for this geometry (shared-memory budget of the decode kernel, <= 16). */
int32_t s2s_llama_max_decode_batch(s2s_llama* m);
/* End-to-end with HOST buffers: (chunked) prefill + greedy decode, synchronous.  ids_out_h[0] is the argmax of the
 * prompt's last position, n_steps tokens in total (generate() semantics of the reference's pipeline call). */
int s2s_llama_generate(s2s_llama* m, int32_t slot, const int32_t* prompt_h, int32_t n_prompt, int32_t n_steps,
                       int32_t eos_id, int32_t* ids_out_h, int32_t* len_out_h, void* stream);

/* ---- TTS post-processing (Qwen3TTSHandler._stream) -------------------------------------
 * wav24k_d f32[n] -> polyphase resample 24k->16k (scipy.signal.resample_poly(x, 2, 3) taps supplied by
 * the host mirror) -> clip(x*32768) -> int16.  out16k_d must hold ceil(n*2/3) samples.          */
int s2s_tts_postproc(s2s_ctx* ctx, const float* wav24k_d, int32_t n, const float* taps_d, int32_t n_taps,
                     int16_t* out16k_d, int32_t* n_out_h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2S_B200_H */
