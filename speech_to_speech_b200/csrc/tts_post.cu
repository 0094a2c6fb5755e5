// tts_post.cu -- Qwen3TTSHandler._stream post-processing on the GPU: 24 kHz f32 -> 16 kHz int16.
//
// Restates, fused in one pass over the chunk,
//   _resample_to_pipeline_sr : scipy.signal.resample_poly(x, up=2, down=3)   (S/TTS/qwen3_tts_handler.py:674-680)
//   _to_int16                : np.clip(x * 32768, -32768, 32767).astype(int16) (:612-613)
// resample_poly = upfirdn with h = 2 * firwin(61, 1/3, kaiser 5.0) cast to f32, 3 leading zero taps, output
// sliced from sample 11 (scipy/signal/_signaltools.py resample_poly).  The taps are designed by the host
// (scipy.signal.firwin) and passed in, so the filter is the reference's by construction:
//   y[m] = sum_j x[j] * h[3*(m + 11) - 2*j - 3]   over taps in [0, 61), j in [0, n)
// accumulated in fp32 in increasing j with separate multiply and add (the order/rounding of scipy's C loop).
// HBM-bound elementwise/FIR work: 4 B in + 2 B out per 1.5 input samples, one thread per output sample,
// coalesced int16 stores; the taps sit in shared memory.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
tts_post_kernel(const float* __restrict__ x, int n, const float* __restrict__ taps, int n_taps, int half_len,
                short* __restrict__ out, int n_out) {
  extern __shared__ float s_h[];
  for (int i = threadIdx.x; i < n_taps; i += blockDim.x) s_h[i] = taps[i];
  __syncthreads();
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  constexpr int UP = 2, DOWN = 3;
  const int n_pre_pad = DOWN - half_len % DOWN;
  const int n_pre_remove = (half_len + n_pre_pad) / DOWN;
  const int base = DOWN * (m + n_pre_remove) - n_pre_pad;  // tap index = base - UP * j
  int j_lo = (base - (n_taps - 1) + UP - 1) / UP;
  if (base - (n_taps - 1) < 0) j_lo = 0;
  int j_hi = base / UP;
  if (base < 0) j_hi = -1;
  j_hi = min(j_hi, n - 1);
  float acc = 0.f;
  for (int j = max(j_lo, 0); j <= j_hi; ++j) acc = __fadd_rn(acc, __fmul_rn(x[j], s_h[base - UP * j]));
  float v = __fmul_rn(acc, 32768.0f);
  v = fminf(fmaxf(v, -32768.0f), 32767.0f);
  out[m] = (short)(int)v;  // truncation toward zero == ndarray.astype(int16)
}

}  // namespace

extern "C" int s2s_tts_postproc(s2s_ctx* ctx, const float* wav24k_d, int32_t n, const float* taps_d, int32_t n_taps,
                                int16_t* out16k_d, int32_t* n_out_h, void* stream) {
  S2S_REQUIRE(ctx && wav24k_d && taps_d && out16k_d && n_out_h, "tts_postproc: null argument");
  S2S_REQUIRE(n >= 0 && n_taps >= 3 && (n_taps & 1) && n_taps <= 4096, "tts_postproc: bad sizes (n=%d taps=%d)", n, n_taps);
  const int n_out = (n * 2 + 2) / 3;  // ceil(n * up / down)
  *n_out_h = n_out;
  if (n_out == 0) return S2S_OK;
  S2S_CHECK_CUDA(cudaSetDevice(ctx->device));   // handler threads start on device 0 whatever GPU their session lives on
  tts_post_kernel<<<(n_out + 255) / 256, 256, n_taps * sizeof(float), (cudaStream_t)stream>>>(
      wav24k_d, n, taps_d, n_taps, (n_taps - 1) / 2, reinterpret_cast<short*>(out16k_d), n_out);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}
