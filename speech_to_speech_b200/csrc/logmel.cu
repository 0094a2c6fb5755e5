// logmel.cu -- Whisper log-mel front end on the GPU, batched over utterances.
//
// Restates WhisperFeatureExtractor._torch_extract_fbank_features (transformers
// models/whisper/feature_extraction_whisper.py:135-164) and the pad/truncate to 30 s (:296-303):
//   zero-pad to 480000, centred STFT (n_fft 400, hop 160, periodic hann, reflect pad), drop the last
//   frame, |.|^2, slaney mel filter bank, log10(clamp 1e-10), max(x, max - 8) per utterance, (x + 4) / 4.
//
// Kernel 1 (logmel_power): one CTA = 16 frames of one utterance; the 400-point DFT is evaluated
// directly in fp32 against a 400-entry twiddle table in shared memory (index k*n mod 400 is exact, so the
// only error is fp32 accumulation), every thread owns one frequency bin for all 16 frames; then the
// sparse mel projection + log10 and a block->global atomic max.  Frames that lie entirely in the
// zero padding (2/3 of a 10 s utterance) are written as the constant log10(1e-10) without any math.
// Kernel 2 (logmel_finalize): clamp to max-8, affine, write fp32 [B, n_mels, 3000] and the transposed
// 16-bit [B, 3002, n_mels] copy (rows 0 and 3001 stay zero) that conv1 consumes as a strided-window GEMM.
#include "common.cuh"
#include "kernels.cuh"

namespace {

constexpr int N_FFT = 400;
constexpr int HOP = 160;
constexpr int N_BINS = 201;
constexpr int N_SAMPLES = 480000;
constexpr int N_FRAMES = 3000;
constexpr int FPB = 16;  // frames per block

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_f32_kernel(float* p, float v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void __launch_bounds__(256)
logmel_power_kernel(const float* __restrict__ pcm, long long pcm_stride, const int* __restrict__ n_samples,
                    const float* __restrict__ hann, const float* __restrict__ twiddle /*[2][400] cos,sin*/,
                    const float* __restrict__ fb /*[n_mels][201]*/, const int2* __restrict__ fb_range /*[n_mels]*/,
                    int n_mels, float* __restrict__ mel_log /*[B][n_mels][3000]*/, float* __restrict__ mel_max) {
  __shared__ __align__(16) float s_frame[FPB][N_FFT];
  __shared__ float s_cos[N_FFT], s_sin[N_FFT];
  __shared__ float s_pow[FPB][N_BINS + 7];
  __shared__ float s_red[8];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * FPB;
  const int tid = threadIdx.x;
  const int n_valid = min(n_samples[b], N_SAMPLES);
  float* out = mel_log + (long long)b * n_mels * N_FRAMES;

  // first original sample index touched by frame f0 is f0*160 - 200; everything >= n_valid is zero
  if (f0 * HOP - N_FFT / 2 >= n_valid) {
    const float c = -10.0f;  // log10(1e-10)
    for (int i = tid; i < n_mels * FPB; i += 256) {
      const int m = i / FPB, f = f0 + (i % FPB);
      if (f < N_FRAMES) out[(long long)m * N_FRAMES + f] = c;
    }
    if (tid == 0) atomic_max_float(mel_max + b, c);
    return;
  }

  const float* x = pcm + (long long)b * pcm_stride;
  for (int i = tid; i < N_FFT; i += 256) { s_cos[i] = twiddle[i]; s_sin[i] = twiddle[N_FFT + i]; }
  for (int i = tid; i < FPB * N_FFT; i += 256) {
    const int f = i / N_FFT, n = i % N_FFT;
    int idx = (f0 + f) * HOP + n - N_FFT / 2;
    if (idx < 0) idx = -idx;                               // reflect (no edge repeat)
    if (idx >= N_SAMPLES) idx = 2 * N_SAMPLES - 2 - idx;
    const float v = (idx < n_valid) ? x[idx] : 0.f;
    s_frame[f][n] = v * hann[n];
  }
  __syncthreads();

  if (tid < N_BINS) {
    float re[FPB], im[FPB];
#pragma unroll
    for (int f = 0; f < FPB; ++f) re[f] = im[f] = 0.f;
    int idx = 0;
    for (int n = 0; n < N_FFT; n += 4) {
      float c[4], s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        c[j] = s_cos[idx]; s[j] = s_sin[idx];
        idx += tid; if (idx >= N_FFT) idx -= N_FFT;
      }
#pragma unroll
      for (int f = 0; f < FPB; ++f) {
        const float4 xv = *reinterpret_cast<const float4*>(&s_frame[f][n]);
        re[f] = fmaf(xv.x, c[0], re[f]); im[f] = fmaf(xv.x, s[0], im[f]);
        re[f] = fmaf(xv.y, c[1], re[f]); im[f] = fmaf(xv.y, s[1], im[f]);
        re[f] = fmaf(xv.z, c[2], re[f]); im[f] = fmaf(xv.z, s[2], im[f]);
        re[f] = fmaf(xv.w, c[3], re[f]); im[f] = fmaf(xv.w, s[3], im[f]);
      }
    }
#pragma unroll
    for (int f = 0; f < FPB; ++f) s_pow[f][tid] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();

  float lmax = -INFINITY;
  for (int i = tid; i < n_mels * FPB; i += 256) {
    const int m = i / FPB, f = i % FPB;
    const int2 rg = fb_range[m];
    const float* w = fb + m * N_BINS;
    float acc = 0.f;
    for (int k = rg.x; k < rg.y; ++k) acc = fmaf(w[k], s_pow[f][k], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    if (f0 + f < N_FRAMES) {
      out[(long long)m * N_FRAMES + f0 + f] = v;
      lmax = fmaxf(lmax, v);
    }
  }
  lmax = warp_max(lmax);
  if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, s_red[i]);
    atomic_max_float(mel_max + b, v);
  }
}

// grid (ceil(3000/32), B); normalise (optional) and emit both layouts
template <typename T>
__global__ void __launch_bounds__(256)
logmel_finalize_kernel(const float* src /*[B][n_mels][3000]*/, const float* __restrict__ mel_max,
                       int normalize, int n_mels, float* dst_f32 /*nullable, may alias src*/,
                       T* __restrict__ dst_t /*[B][3002][n_mels]*/) {
  extern __shared__ float s_tile[];  // [n_mels][33]
  const int b = blockIdx.y, f0 = blockIdx.x * 32;
  const float floor_v = normalize ? (mel_max[b] - 8.0f) : -INFINITY;
  const float* s = src + (long long)b * n_mels * N_FRAMES;
  for (int i = threadIdx.x; i < n_mels * 32; i += 256) {
    const int m = i >> 5, f = i & 31;
    float v = 0.f;
    if (f0 + f < N_FRAMES) {
      v = s[(long long)m * N_FRAMES + f0 + f];
      if (normalize) v = (fmaxf(v, floor_v) + 4.0f) * 0.25f;
      if (dst_f32) dst_f32[(long long)b * n_mels * N_FRAMES + (long long)m * N_FRAMES + f0 + f] = v;
    }
    s_tile[m * 33 + f] = v;
  }
  __syncthreads();
  T* d = dst_t + ((long long)b * (N_FRAMES + 2) + 1) * n_mels;
  for (int i = threadIdx.x; i < n_mels * 32; i += 256) {
    const int f = i / n_mels, m = i % n_mels;
    if (f0 + f < N_FRAMES) d[(long long)(f0 + f) * n_mels + m] = DT<T>::from_f(s_tile[m * 33 + f]);
  }
}

}  // namespace

int logmel_launch(const LogmelTables& tb, const float* pcm_d, long long pcm_stride, const int* n_samples_d, int B,
                  int n_mels, float* mel_f32, float* mel_max, void* mel_t, int dtype, cudaStream_t stream) {
  fill_f32_kernel<<<1, 256, 0, stream>>>(mel_max, -INFINITY, B);
  S2S_LAUNCH_CHECK();
  dim3 g1((N_FRAMES + FPB - 1) / FPB, B);
  logmel_power_kernel<<<g1, 256, 0, stream>>>(pcm_d, pcm_stride, n_samples_d, tb.hann, tb.twiddle, tb.fb, tb.fb_range,
                                             n_mels, mel_f32, mel_max);
  S2S_LAUNCH_CHECK();
  return logmel_finalize_launch(mel_f32, mel_max, 1, B, n_mels, mel_f32, mel_t, dtype, stream);
}

int logmel_finalize_launch(const float* src, const float* mel_max, int normalize, int B, int n_mels, float* dst_f32,
                           void* mel_t, int dtype, cudaStream_t stream) {
  dim3 g2((N_FRAMES + 31) / 32, B);
  const size_t sm = (size_t)n_mels * 33 * sizeof(float);
  if (dtype == S2S_F16)
    logmel_finalize_kernel<__half><<<g2, 256, sm, stream>>>(src, mel_max, normalize, n_mels, dst_f32, (__half*)mel_t);
  else
    logmel_finalize_kernel<__nv_bfloat16><<<g2, 256, sm, stream>>>(src, mel_max, normalize, n_mels, dst_f32,
                                                                  (__nv_bfloat16*)mel_t);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}
