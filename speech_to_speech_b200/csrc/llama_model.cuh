// llama_model.cuh -- the Llama-family model object (internal to libs2s_b200): shared by llama.cu (C ABI of the LLM slot)
// and qwen3tts.cu (Qwen3-TTS talker and code predictor are Qwen3-style decoders driven by embeddings).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "llama_decode.cuh"

enum LSlotKind { L_PLAIN = 0, L_ROPE_PERM = 1, L_INTERLEAVE = 2 };
struct LSlot {
  void* dst = nullptr;
  bool half = false;
  int64_t rows = 0, cols = 0;
  int kind = L_PLAIN;
  int hd = 0;          // ROPE_PERM
  int parity = 0;      // INTERLEAVE: 0 gate, 1 up
  bool bound = false;
  float rnd_scale = 0.02f, rnd_offset = 0.f;
};

struct s2s_llama {
  s2s_ctx* ctx = nullptr;
  s2s_llama_config cfg{};
  bool finalized = false;
  int debug_phases = 0;
  std::vector<void*> allocs;
  std::unordered_map<std::string, LSlot> slots;
  bool lm_head_bound = false;
  // weights
  void *embed = nullptr, *lm_head = nullptr;
  float* norm_f = nullptr;
  std::vector<LlamaDecLayer> layers_h;   // row-major weights (prefill GEMMs)
  std::vector<LlamaDecLayer> tiled_h;    // fragment-major copies streamed by the decode kernel
  LlamaDecLayer* layers_d = nullptr;     // device copy of tiled_h
  void* lm_head_t = nullptr;
  float2* rope = nullptr;
  // KV + sessions
  void* kv = nullptr;
  long long kv_slot_stride = 0, kv_layer_stride = 0, kv_which_stride = 0;
  std::vector<int> len;  // tokens in each slot
  // prefill workspace
  int* ids_d = nullptr;
  float *x = nullptr, *last_logits = nullptr, *batch_logits = nullptr;   // batch_logits: [MAX_DEC_B][vocab] (batched prefill)
  void *xn = nullptr, *qkv = nullptr, *attn = nullptr, *hbuf = nullptr, *vt = nullptr;
  size_t vt_elems = 0;
  // decode state
  float *dx = nullptr, *dq = nullptr, *dh = nullptr, *part = nullptr, *cand_val = nullptr;
  void* attn16 = nullptr;
  unsigned int* attn_cnt = nullptr;
  int *slot_d = nullptr, *pos_d = nullptr, *done = nullptr, *n_done = nullptr, *cand_idx = nullptr, *out_ids = nullptr,
      *out_len = nullptr, *next_id = nullptr;
  unsigned int* sync_counter = nullptr;
  int s_max = 0;
  unsigned long long* trace = nullptr;
  int trace_cap = 0;
  int n_tables = 1;                      // > 1: one embedding table and one output head per codebook (code predictor)
  size_t embed_table_elems = 0, head_table_elems = 0, head_t_table_elems = 0;
  float* qn = nullptr;                   // qk_norm: normalised q rows of the decode step [B][H*hd]
  float* kraw = nullptr;                 // qk_norm: raw k rows of the decode step [B][KV*hd]
};


constexpr int LLAMA_MAX_DEC_B = 16;  // upper bound; the shared-memory budget of the geometry may allow fewer

// Run the decoder stack over the n fp32 rows already in m->x (embeddings), appending to `slot`'s KV cache.
// logits_out_d optional [n, vocab]; next_id_d optional [1] = argmax of the last position (table 0);
// hidden_out_d optional [d] fp32 = the last position's residual stream BEFORE the final norm.
int llama_prefill_rows(s2s_llama* m, int slot, int n, float* logits_out_d, int32_t* next_id_d, float* hidden_out_d,
                       cudaStream_t st);
// Everything of LlamaDecParams that does not depend on the call (geometry, weights, workspace).
void llama_fill_dec_params(s2s_llama* m, LlamaDecParams& p);
int llama_max_decode_batch_of(const s2s_llama* m);
