// llama.cu -- Llama-family LLM model object behind the C ABI (placeholder until the decode path lands).
#include "common.cuh"
extern "C" {
#define NOT_YET(name) s2s_set_error(name ": not implemented yet"); return S2S_ERR_UNSUPPORTED
int s2s_llama_create(s2s_ctx*, const s2s_llama_config*, s2s_llama**) { NOT_YET("s2s_llama_create"); }
int s2s_llama_destroy(s2s_llama*) { return S2S_OK; }
int s2s_llama_bind_tensor(s2s_llama*, const char*, const void*, const int64_t*, int32_t, int32_t) { NOT_YET("s2s_llama_bind_tensor"); }
int s2s_llama_init_random(s2s_llama*, uint64_t) { NOT_YET("s2s_llama_init_random"); }
int s2s_llama_finalize(s2s_llama*) { NOT_YET("s2s_llama_finalize"); }
int s2s_llama_session_reset(s2s_llama*, int32_t) { NOT_YET("s2s_llama_session_reset"); }
int s2s_llama_prefill(s2s_llama*, int32_t, const int32_t*, int32_t, float*, int32_t*, void*) { NOT_YET("s2s_llama_prefill"); }
int s2s_llama_decode(s2s_llama*, const int32_t*, int32_t, const int32_t*, int32_t, int32_t, int32_t*, int32_t*,
                     const int32_t*, float*, void*) { NOT_YET("s2s_llama_decode"); }
int s2s_llama_generate(s2s_llama*, int32_t, const int32_t*, int32_t, int32_t, int32_t, int32_t*, int32_t*, void*) { NOT_YET("s2s_llama_generate"); }
}
