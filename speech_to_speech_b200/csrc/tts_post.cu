// tts_post.cu -- Qwen3TTSHandler._stream post-processing on the GPU (placeholder).
#include "common.cuh"
extern "C" int s2s_tts_postproc(s2s_ctx*, const float*, int32_t, const float*, int32_t, int16_t*, int32_t*, void*) {
  s2s_set_error("s2s_tts_postproc: not implemented yet");
  return S2S_ERR_UNSUPPORTED;
}
