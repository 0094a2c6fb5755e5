// codec_decode.cu -- TTS codec decoder: 16 codebook ids per 12.5 Hz frame -> 24 kHz waveform, fp32, time-major.
//
// Reference path: Qwen3TTSHandler._process_custom_voice -> model.generate_custom_voice_streaming(...) yields
// (audio_f32, 24000, timing) every `chunk_size` frames (S/TTS/qwen3_tts_handler.py:946-978); the arithmetic lives in
// the absent faster-qwen3-tts.  Structural oracle: transformers Qwen3OmniMoeCode2Wav
// (modeling_qwen3_omni_moe.py:3283-3790), restated in oracle/code2wav_ref.py:
//   mean of the frame's code embeddings -> 8 sliding-window RoPE transformer layers (LayerScale) -> RMSNorm
//   -> 2 x [transposed conv (k = stride) -> ConvNeXt block]                       (x4 in time)
//   -> causal conv 7 -> 4 x [SnakeBeta -> causal transposed conv (k = 2 stride) -> 3 dilated residual units]
//   -> SnakeBeta -> causal conv 7 -> clamp                                        (x480 in time: 1920 samples / frame)
// and chunked_decode(:3779-3790): decode `chunk` new frames behind `left_context` frames of history, drop the history.
//
// Everything convolutional is ONE contraction kernel over time-major [T, C] activations (conv1d_f32_kernel):
//   causal conv (k taps, dilation)   : rows t - (k-1-j) * dil, zero rows before the start
//   transposed conv, k = 2 s         : out[(t0-1) * s + r, o] = X[t0-1] . W[:, o, r + s] + X[t0] . W[:, o, r]  for t0 = 1..T-1
//                                      = a 2-tap contraction with N = s * C_out whose [T-1, s * C_out] result IS the
//                                      time-major [(T-1) * s, C_out] output (trim of k - s samples at both ends included)
//   transposed conv, k = s           : a 1-tap contraction with N = s * C_out
//   linear layers (transformer, ConvNeXt MLP): 1 tap.
// fp32 CUDA-core FMA with shared-memory tiles: the decoder is held to 1e-4-level waveform parity with the fp32 oracle;
// its cost is ~5 GFLOP per frame, far from the B200's limits at the session counts the talker sustains.
#include <math.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "codec_decode.cuh"
#include "kernels.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ contraction kernel
constexpr int CV_BM = 64, CV_BK = 16, CV_THREADS = 256;

template <int BN>
__global__ void __launch_bounds__(CV_THREADS) conv1d_f32_kernel(const ConvArgs a) {
  constexpr int TN = BN / 16;   // outputs per thread along n (4 or 2); 4 along m
  __shared__ float As[CV_BK][CV_BM + 4];
  __shared__ float Bs[CV_BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * CV_BM, n0 = blockIdx.x * BN, z = blockIdx.z;
  const float* X = a.x + (long long)z * a.x_bs;
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  // loaders: A tile 64 rows x 16 c (thread -> row tid / 4, 4 consecutive c), B tile 16 c x BN n
  const int a_row = tid >> 2, a_c = (tid & 3) * 4;
  for (int j = 0; j < a.k; ++j) {
    const int xr = m0 + a_row + a.x_row0 + j * a.dil;
    const bool row_ok = (m0 + a_row) < a.T_out && xr >= 0 && xr < a.T_in;
    const float* xrow = X + (long long)xr * a.ldx;
    const float* wj = a.w + (long long)j * a.C_in * a.N;
    for (int c0 = 0; c0 < a.C_in; c0 += CV_BK) {
      float av[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + a_c + i;
        av[i] = (row_ok && c < a.C_in) ? xrow[c] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) As[a_c + i][a_row] = av[i];
      for (int e = tid; e < CV_BK * BN; e += CV_THREADS) {
        const int kk = e / BN, nn = e % BN;
        const int c = c0 + kk, n = n0 + nn;
        Bs[kk][nn] = (c < a.C_in && n < a.N) ? __ldg(wj + (long long)c * a.N + n) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < CV_BK; ++kk) {
        float ar[4], br[TN];
#pragma unroll
        for (int i = 0; i < 4; ++i) ar[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int jn = 0; jn < TN; ++jn) br[jn] = Bs[kk][tx * TN + jn];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jn = 0; jn < TN; ++jn) acc[i][jn] = fmaf(ar[i], br[jn], acc[i][jn]);
      }
      __syncthreads();
    }
  }
  float* Y = a.y + (long long)z * a.y_bs;
  const float* R = a.resid ? a.resid + (long long)z * a.r_bs : nullptr;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = m0 + ty * 4 + i;
    if (t >= a.T_out) continue;
#pragma unroll
    for (int jn = 0; jn < TN; ++jn) {
      const int n = n0 + tx * TN + jn;
      if (n >= a.N) continue;
      float v = acc[i][jn];
      if (a.bias) v += __ldg(a.bias + (n % a.bias_mod));
      if (a.act == 1) v = gelu_erf(v);
      else if (a.act == 2) v = v / (1.0f + expf(-v));
      if (a.scale) v *= __ldg(a.scale + n);
      if (R) v += R[(long long)t * a.ldr + n];
      Y[(long long)t * a.ldy + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ small kernels
// x[t, :] = mean_q E[q * cb + codes[t][q], :]
__global__ void code_embed_mean_kernel(const int* __restrict__ codes, int Q, int cb, const float* __restrict__ E, int H,
                                       float* __restrict__ x) {
  const int t = blockIdx.x;
  const float inv = 1.0f / (float)Q;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float s = 0.f;
    for (int q = 0; q < Q; ++q) s += E[((long long)q * cb + codes[t * Q + q]) * H + i];
    x[(long long)t * H + i] = s * inv;
  }
}

// y = w * x * rsqrt(mean(x^2) + eps), one warp per row
__global__ void rmsnorm_rows_f32_kernel(const float* __restrict__ x, const float* __restrict__ w, float eps, int rows, int d,
                                        float* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (long long)row * d;
  float ss = 0.f;
  for (int i = lane; i < d; i += 32) ss = fmaf(xr[i], xr[i], ss);
  ss = warp_sum(ss);
  const float r = 1.0f / sqrtf(ss / (float)d + eps);
  for (int i = lane; i < d; i += 32) y[(long long)row * d + i] = w[i] * (xr[i] * r);
}

// RoPE (rotate_half pairs (j, j + hd/2)) on q and k in place; qkv [T, (H + 2 KV) * hd], position = row index
__global__ void rope_rows_f32_kernel(float* __restrict__ qkv, int T, int n_rot_heads, int hd, int ld, float theta) {
  const int t = blockIdx.x;
  const int half = hd >> 1;
  for (int i = threadIdx.x; i < n_rot_heads * half; i += blockDim.x) {
    const int h = i / half, j = i % half;
    const float inv = 1.0f / powf(theta, (float)(2 * j) / (float)hd);
    const float ang = (float)t * inv;
    const float c = cosf(ang), s = sinf(ang);
    float* p = qkv + (long long)t * ld + h * hd;
    const float a = p[j], b = p[j + half];
    p[j] = a * c - b * s;
    p[j + half] = b * c + a * s;
  }
}

// causal sliding-window attention, one warp per (query t, head h); keys t - W + 1 .. t; fp32 softmax
__global__ void swa_attention_f32_kernel(const float* __restrict__ qkv, int T, int H, int KV, int hd, int W, float* __restrict__ o) {
  extern __shared__ float sm[];   // per warp: q[hd] + p[W]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int item = blockIdx.x * nw + warp;
  if (item >= T * H) return;
  const int t = item / H, h = item % H, g = H / KV;
  const int ld = (H + 2 * KV) * hd;
  float* qs = sm + (size_t)warp * (hd + W);
  float* ps = qs + hd;
  const float* q = qkv + (long long)t * ld + h * hd;
  for (int i = lane; i < hd; i += 32) qs[i] = q[i];
  __syncwarp();
  const int lo = max(0, t - W + 1), n = t - lo + 1;
  const float scale = rsqrtf((float)hd);
  float mx = -INFINITY;
  for (int s = lane; s < n; s += 32) {
    const float* k = qkv + (long long)(lo + s) * ld + (H + h / g) * hd;
    float d = 0.f;
    for (int i = 0; i < hd; ++i) d = fmaf(qs[i], k[i], d);
    d *= scale;
    ps[s] = d;
    mx = fmaxf(mx, d);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int s = lane; s < n; s += 32) { const float e = expf(ps[s] - mx); ps[s] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.0f / sum;
  for (int i = lane; i < hd; i += 32) {
    float acc = 0.f;
    for (int s = 0; s < n; ++s) acc = fmaf(ps[s], qkv[(long long)(lo + s) * ld + (H + KV + h / g) * hd + i], acc);
    o[(long long)t * (H * hd) + h * hd + i] = acc * inv;
  }
}

// h = silu(g) * u ; gu [T, 2 * inter] as [gate | up]
__global__ void silu_mul_kernel(const float* __restrict__ gu, int inter, long long n, float* __restrict__ h) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long t = i / inter, c = i % inter;
  const float g = gu[t * 2 * inter + c], u = gu[t * 2 * inter + inter + c];
  h[i] = (g / (1.0f + expf(-g))) * u;
}

// SnakeBeta: y = x + ib[c] * sin(x * a[c])^2 with a = exp(alpha), ib = 1 / (exp(beta) + 1e-9) (precomputed at bind)
__global__ void snake_beta_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ ib, int C,
                                  long long n, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const float v = x[i], s = sinf(v * a[c]);
  y[i] = v + ib[c] * (s * s);
}

// ConvNeXt front: depthwise causal conv 7 + LayerNorm(eps) over channels; one CTA per time step
__global__ void dwconv7_ln_kernel(const float* __restrict__ x, int T, int C, const float* __restrict__ wd, const float* __restrict__ bd,
                                  const float* __restrict__ lw, const float* __restrict__ lb, float eps, float* __restrict__ y) {
  extern __shared__ float hs[];   // [C]
  __shared__ float red[2][32];
  const int t = blockIdx.x;
  float s1 = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v = bd[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int r = t - 6 + j;
      if (r >= 0) v = fmaf(x[(long long)r * C + c], wd[c * 7 + j], v);
    }
    hs[c] = v;
    s1 += v;
  }
  s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s1;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[0][w];
  const float mean = tot / (float)C;
  float s2 = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { const float dlt = hs[c] - mean; s2 = fmaf(dlt, dlt, s2); }
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) red[1][threadIdx.x >> 5] = s2;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) var += red[1][w];
  const float rstd = 1.0f / sqrtf(var / (float)C + eps);
  for (int c = threadIdx.x; c < C; c += blockDim.x) y[(long long)t * C + c] = (hs[c] - mean) * rstd * lw[c] + lb[c];
}

// wav[i] = clamp(x[(skip + i) * ld], -1, 1)
__global__ void clamp_out_kernel(const float* __restrict__ x, long long ld, int skip, int n, float* __restrict__ wav) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) wav[i] = fminf(1.0f, fmaxf(-1.0f, x[(long long)(skip + i) * ld]));
}

__global__ void exp_prep_kernel(float* a, float* b, int n) {   // alpha -> exp(alpha); beta -> 1 / (exp(beta) + 1e-9)
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = expf(a[i]); b[i] = 1.0f / (expf(b[i]) + 1e-9f); }
}

inline float src_f32(const void* data, int64_t i, int dtype) {
  if (dtype == S2S_F32) return reinterpret_cast<const float*>(data)[i];
  uint16_t u = reinterpret_cast<const uint16_t*>(data)[i];
  if (dtype == S2S_BF16) { uint32_t w = (uint32_t)u << 16; float f; memcpy(&f, &w, 4); return f; }
  __half h; memcpy(&h, &u, 2); return __half2float(h);
}

enum CSlotKind { C_PLAIN = 0, C_LINEAR = 1 /* [N,K] -> [K][N] */, C_CONV = 2 /* [O,C,k] -> [k][C][O] */,
                 C_TCONV = 3 /* [C,O,k] -> taps x [C][s*O] */ };
struct CSlot {
  float* dst = nullptr;
  int64_t n = 0;
  int kind = C_PLAIN;
  int d0 = 0, d1 = 0, d2 = 0, stride = 0;   // source dims
  bool bound = false;
  float rnd_scale = 0.02f, rnd_offset = 0.f;
};

}  // namespace

int conv1d_f32_launch(const ConvArgs& a, cudaStream_t st) {
  if (a.T_out <= 0 || a.N <= 0) return S2S_OK;
  const int batch = a.batch > 0 ? a.batch : 1;
  const int mt = (a.T_out + CV_BM - 1) / CV_BM;
  // few time tiles (transformer / first decoder stages): narrower n tiles put more SMs on the weight stream
  if ((long long)mt * ((a.N + 63) / 64) * batch < 148) {
    dim3 grid((a.N + 31) / 32, mt, batch);
    conv1d_f32_kernel<32><<<grid, CV_THREADS, 0, st>>>(a);
  } else {
    dim3 grid((a.N + 63) / 64, mt, batch);
    conv1d_f32_kernel<64><<<grid, CV_THREADS, 0, st>>>(a);
  }
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}

// ------------------------------------------------------------------------------------------------ the model
struct CodecLayer {   // transformer layer
  float *w_qkv, *w_o, *w_gu, *w_down, *n1, *n2, *ls_attn, *ls_mlp;
};
struct CodecConvNeXt { float *up_w, *up_b, *dw_w, *dw_b, *ln_w, *ln_b, *pw1_w, *pw1_b, *pw2_w, *pw2_b, *gamma; };
struct CodecResUnit { float *a1, *b1, *c1_w, *c1_b, *a2, *b2, *c2_w, *c2_b; };
struct CodecBlock { float *a0, *b0, *tc_w, *tc_b; CodecResUnit u[3]; int cin, cout, rate; };

struct s2s_codec {
  s2s_ctx* ctx = nullptr;
  s2s_codec_config cfg{};
  bool finalized = false;
  std::vector<void*> allocs;
  std::unordered_map<std::string, CSlot> slots;
  std::vector<std::pair<float*, float*>> snake_pairs;   // (alpha, beta) buffers and their length, transformed at finalize
  std::vector<int> snake_len;
  float* embed = nullptr;
  std::vector<CodecLayer> layers;
  float* norm_f = nullptr;
  std::vector<CodecConvNeXt> ups;
  float *d0_w = nullptr, *d0_b = nullptr;
  std::vector<CodecBlock> blocks;
  float *fa = nullptr, *fb = nullptr, *f_w = nullptr, *f_b = nullptr;
  // workspace
  float *x = nullptr, *xn = nullptr, *qkv = nullptr, *att = nullptr, *gu = nullptr, *hmid = nullptr;
  float *bufX = nullptr, *bufA = nullptr, *bufH = nullptr;
  size_t buf_elems = 0;
  int total_up = 1;
};

namespace {

int calloc_f(CodecDecoder* m, float** out, size_t n) {
  void* p = nullptr;
  S2S_CHECK_CUDA(cudaMalloc(&p, (n ? n : 4) * 4));
  S2S_CHECK_CUDA(cudaMemset(p, 0, (n ? n : 4) * 4));
  m->allocs.push_back(p);
  *out = reinterpret_cast<float*>(p);
  return S2S_OK;
}

int cslot(CodecDecoder* m, const std::string& name, float** out, int kind, int d0, int d1, int d2, int stride, float rs,
          float ro = 0.f) {
  CSlot s;
  s.kind = kind; s.d0 = d0; s.d1 = d1; s.d2 = d2; s.stride = stride; s.rnd_scale = rs; s.rnd_offset = ro;
  s.n = (int64_t)d0 * (d1 ? d1 : 1) * (d2 ? d2 : 1);
  S2S_CHECK(calloc_f(m, &s.dst, (size_t)s.n));
  *out = s.dst;
  m->slots[name] = s;
  return S2S_OK;
}

// sequence lengths along the decoder for T frames
void codec_lengths(const s2s_codec_config& c, int T, std::vector<int>& L) {
  int len = T;
  for (int i = 0; i < c.n_upsampling_ratios; ++i) len *= c.upsampling_ratios[i];
  L.clear();
  L.push_back(len);
  for (int i = 0; i < c.n_upsample_rates; ++i) { len = (len - 1) * c.upsample_rates[i]; L.push_back(len); }
}

int build(CodecDecoder* m) {
  const auto& c = m->cfg;
  const int H = c.hidden, hd = H / c.heads, qd = c.heads * hd, kvd = c.kv_heads * hd, I = c.inter;
  float* dummy;
  S2S_CHECK(cslot(m, "code_embedding.weight", &m->embed, C_PLAIN, c.codebook_size * c.quantizers, H, 0, 0, 1.0f));
  m->layers.resize(c.layers);
  const float sh = 1.0f / sqrtf((float)H);
  for (int l = 0; l < c.layers; ++l) {
    const std::string p = "pre_transformer.layers." + std::to_string(l) + ".";
    CodecLayer& L = m->layers[l];
    // q, k, v share one [H][qd + 2 kvd] matrix; gate and up one [H][2 I] matrix (column blocks)
    S2S_CHECK(calloc_f(m, &L.w_qkv, (size_t)H * (qd + 2 * kvd)));
    S2S_CHECK(calloc_f(m, &L.w_gu, (size_t)H * 2 * I));
    auto sub = [&](const std::string& name, float* base, int N_total, int col0, int N, int K, float rs) {
      CSlot s; s.dst = base; s.kind = C_LINEAR; s.d0 = N; s.d1 = K; s.d2 = N_total; s.stride = col0; s.n = (int64_t)N * K; s.rnd_scale = rs;
      m->slots[name] = s;
    };
    sub(p + "self_attn.q_proj.weight", L.w_qkv, qd + 2 * kvd, 0, qd, H, sh);
    sub(p + "self_attn.k_proj.weight", L.w_qkv, qd + 2 * kvd, qd, kvd, H, sh);
    sub(p + "self_attn.v_proj.weight", L.w_qkv, qd + 2 * kvd, qd + kvd, kvd, H, sh);
    sub(p + "mlp.gate_proj.weight", L.w_gu, 2 * I, 0, I, H, sh);
    sub(p + "mlp.up_proj.weight", L.w_gu, 2 * I, I, I, H, sh);
    S2S_CHECK(cslot(m, p + "self_attn.o_proj.weight", &L.w_o, C_LINEAR, H, qd, 0, 0, sh));
    S2S_CHECK(cslot(m, p + "mlp.down_proj.weight", &L.w_down, C_LINEAR, H, I, 0, 0, 1.0f / sqrtf((float)I)));
    S2S_CHECK(cslot(m, p + "input_layernorm.weight", &L.n1, C_PLAIN, H, 0, 0, 0, 0.1f, 1.0f));
    S2S_CHECK(cslot(m, p + "post_attention_layernorm.weight", &L.n2, C_PLAIN, H, 0, 0, 0, 0.1f, 1.0f));
    S2S_CHECK(cslot(m, p + "self_attn_layer_scale.scale", &L.ls_attn, C_PLAIN, H, 0, 0, 0, 0.1f, 0.5f));
    S2S_CHECK(cslot(m, p + "mlp_layer_scale.scale", &L.ls_mlp, C_PLAIN, H, 0, 0, 0, 0.1f, 0.5f));
  }
  S2S_CHECK(cslot(m, "pre_transformer.norm.weight", &m->norm_f, C_PLAIN, H, 0, 0, 0, 0.1f, 1.0f));
  m->ups.resize(c.n_upsampling_ratios);
  m->total_up = 1;
  for (int i = 0; i < c.n_upsampling_ratios; ++i) {
    const int f = c.upsampling_ratios[i];
    m->total_up *= f;
    const std::string p = "upsample." + std::to_string(i) + ".";
    CodecConvNeXt& U = m->ups[i];
    S2S_CHECK(cslot(m, p + "0.conv.weight", &U.up_w, C_TCONV, H, H, f, f, sh));
    S2S_CHECK(cslot(m, p + "0.conv.bias", &U.up_b, C_PLAIN, H, 0, 0, 0, 0.02f));
    S2S_CHECK(cslot(m, p + "1.dwconv.conv.weight", &U.dw_w, C_PLAIN, H, 7, 0, 0, 0.378f));
    S2S_CHECK(cslot(m, p + "1.dwconv.conv.bias", &U.dw_b, C_PLAIN, H, 0, 0, 0, 0.02f));
    S2S_CHECK(cslot(m, p + "1.norm.weight", &U.ln_w, C_PLAIN, H, 0, 0, 0, 0.1f, 1.0f));
    S2S_CHECK(cslot(m, p + "1.norm.bias", &U.ln_b, C_PLAIN, H, 0, 0, 0, 0.05f));
    S2S_CHECK(cslot(m, p + "1.pwconv1.weight", &U.pw1_w, C_LINEAR, 4 * H, H, 0, 0, sh));
    S2S_CHECK(cslot(m, p + "1.pwconv1.bias", &U.pw1_b, C_PLAIN, 4 * H, 0, 0, 0, 0.02f));
    S2S_CHECK(cslot(m, p + "1.pwconv2.weight", &U.pw2_w, C_LINEAR, H, 4 * H, 0, 0, 0.5f * sh));
    S2S_CHECK(cslot(m, p + "1.pwconv2.bias", &U.pw2_b, C_PLAIN, H, 0, 0, 0, 0.02f));
    S2S_CHECK(cslot(m, p + "1.gamma", &U.gamma, C_PLAIN, H, 0, 0, 0, 0.05f, 0.3f));
  }
  const int D = c.decoder_dim;
  S2S_CHECK(cslot(m, "decoder.0.conv.weight", &m->d0_w, C_CONV, D, H, 7, 0, 1.0f / sqrtf(7.0f * H)));
  S2S_CHECK(cslot(m, "decoder.0.conv.bias", &m->d0_b, C_PLAIN, D, 0, 0, 0, 0.02f));
  m->blocks.resize(c.n_upsample_rates);
  auto snake = [&](const std::string& pa, const std::string& pb, float** a, float** b, int n) -> int {
    S2S_CHECK(cslot(m, pa, a, C_PLAIN, n, 0, 0, 0, 0.2f));
    S2S_CHECK(cslot(m, pb, b, C_PLAIN, n, 0, 0, 0, 0.2f));
    m->snake_pairs.push_back({*a, *b});
    m->snake_len.push_back(n);
    return S2S_OK;
  };
  for (int i = 0; i < c.n_upsample_rates; ++i) {
    CodecBlock& B = m->blocks[i];
    B.cin = D >> i; B.cout = D >> (i + 1); B.rate = c.upsample_rates[i];
    m->total_up *= B.rate;
    const std::string p = "decoder." + std::to_string(i + 1) + ".block.";
    S2S_CHECK(snake(p + "0.alpha", p + "0.beta", &B.a0, &B.b0, B.cin));
    S2S_CHECK(cslot(m, p + "1.conv.weight", &B.tc_w, C_TCONV, B.cin, B.cout, 2 * B.rate, B.rate, 1.0f / sqrtf(2.0f * B.cin)));
    S2S_CHECK(cslot(m, p + "1.conv.bias", &B.tc_b, C_PLAIN, B.cout, 0, 0, 0, 0.02f));
    for (int u = 0; u < 3; ++u) {
      const std::string q = p + std::to_string(u + 2) + ".";
      CodecResUnit& R = B.u[u];
      S2S_CHECK(snake(q + "act1.alpha", q + "act1.beta", &R.a1, &R.b1, B.cout));
      S2S_CHECK(cslot(m, q + "conv1.conv.weight", &R.c1_w, C_CONV, B.cout, B.cout, 7, 0, 1.0f / sqrtf(7.0f * B.cout)));
      S2S_CHECK(cslot(m, q + "conv1.conv.bias", &R.c1_b, C_PLAIN, B.cout, 0, 0, 0, 0.02f));
      S2S_CHECK(snake(q + "act2.alpha", q + "act2.beta", &R.a2, &R.b2, B.cout));
      S2S_CHECK(cslot(m, q + "conv2.conv.weight", &R.c2_w, C_CONV, B.cout, B.cout, 1, 0, 1.0f / sqrtf((float)B.cout)));
      S2S_CHECK(cslot(m, q + "conv2.conv.bias", &R.c2_b, C_PLAIN, B.cout, 0, 0, 0, 0.02f));
    }
  }
  const int n = c.n_upsample_rates, cl = D >> n;
  S2S_CHECK(snake("decoder." + std::to_string(n + 1) + ".alpha", "decoder." + std::to_string(n + 1) + ".beta", &m->fa, &m->fb, cl));
  S2S_CHECK(cslot(m, "decoder." + std::to_string(n + 2) + ".conv.weight", &m->f_w, C_CONV, 1, cl, 7, 0, 0.02f / sqrtf(7.0f * cl)));
  S2S_CHECK(cslot(m, "decoder." + std::to_string(n + 2) + ".conv.bias", &m->f_b, C_PLAIN, 1, 0, 0, 0, 0.02f));
  (void)dummy;
  // workspace for max_frames
  const int T = c.max_frames;
  std::vector<int> L;
  codec_lengths(c, T, L);
  size_t mx = (size_t)L[0] * std::max(4 * H, D);
  for (int i = 0; i < c.n_upsample_rates; ++i) mx = std::max(mx, (size_t)L[i + 1] * (size_t)(D >> (i + 1)));
  mx = std::max(mx, (size_t)L[0] * (size_t)D);
  m->buf_elems = mx;
  S2S_CHECK(calloc_f(m, &m->bufX, mx));
  S2S_CHECK(calloc_f(m, &m->bufA, mx));
  S2S_CHECK(calloc_f(m, &m->bufH, mx));
  S2S_CHECK(calloc_f(m, &m->x, (size_t)T * H));
  S2S_CHECK(calloc_f(m, &m->xn, (size_t)T * H));
  S2S_CHECK(calloc_f(m, &m->qkv, (size_t)T * (qd + 2 * kvd)));
  S2S_CHECK(calloc_f(m, &m->att, (size_t)T * qd));
  S2S_CHECK(calloc_f(m, &m->gu, (size_t)T * 2 * I));
  S2S_CHECK(calloc_f(m, &m->hmid, (size_t)T * I));
  return S2S_OK;
}

ConvArgs linear_args(const float* x, int T, int K, const float* w, int N, float* y) {
  ConvArgs a{};
  a.x = x; a.ldx = K; a.T_in = T; a.x_row0 = 0; a.w = w; a.k = 1; a.dil = 1; a.C_in = K; a.N = N; a.bias = nullptr; a.bias_mod = N;
  a.y = y; a.ldy = N; a.T_out = T; a.batch = 1;
  return a;
}
ConvArgs causal_conv_args(const float* x, int T, int C_in, const float* w, const float* b, int k, int dil, int C_out, float* y) {
  ConvArgs a{};
  a.x = x; a.ldx = C_in; a.T_in = T; a.x_row0 = -(k - 1) * dil; a.w = w; a.k = k; a.dil = dil; a.C_in = C_in; a.N = C_out;
  a.bias = b; a.bias_mod = C_out; a.y = y; a.ldy = C_out; a.T_out = T; a.batch = 1;
  return a;
}

int snake_launch(const float* x, const float* a, const float* ib, int C, long long n, float* y, cudaStream_t st) {
  snake_beta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, a, ib, C, n, y);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}

}  // namespace

int codec_total_upsample(const CodecDecoder* m) { return m->total_up; }
int codec_samples_for(const CodecDecoder* m, int T) {
  std::vector<int> L;
  codec_lengths(m->cfg, T, L);
  return L.back();
}

int codec_create(s2s_ctx* ctx, const s2s_codec_config* cfg, CodecDecoder** out) {
  S2S_REQUIRE(ctx && cfg && out, "codec_create: null argument");
  S2S_REQUIRE(cfg->hidden % cfg->heads == 0 && cfg->heads % cfg->kv_heads == 0, "codec: bad head geometry");
  S2S_REQUIRE(cfg->n_upsample_rates >= 1 && cfg->n_upsample_rates <= 8 && cfg->n_upsampling_ratios >= 0 && cfg->n_upsampling_ratios <= 4,
              "codec: bad upsampling lists");
  S2S_REQUIRE(cfg->max_frames >= 1 && cfg->layers >= 0 && cfg->sliding_window >= 1, "codec: bad capacity");
  S2S_REQUIRE((cfg->decoder_dim >> cfg->n_upsample_rates) >= 1, "codec: decoder_dim too small for %d blocks", cfg->n_upsample_rates);
  S2S_CHECK_CUDA(cudaSetDevice(ctx->device));
  CodecDecoder* m = new s2s_codec();
  m->ctx = ctx;
  m->cfg = *cfg;
  const int r = build(m);
  if (r != S2S_OK) { codec_destroy(m); return r; }
  *out = m;
  return S2S_OK;
}

int codec_destroy(CodecDecoder* m) {
  if (!m) return S2S_OK;
  for (void* p : m->allocs) cudaFree(p);
  delete m;
  return S2S_OK;
}

int codec_bind_tensor(CodecDecoder* m, const char* name, const void* data_h, const int64_t* shape, int ndim, int dtype) {
  S2S_REQUIRE(m && name && data_h && shape, "codec bind_tensor: null argument");
  auto it = m->slots.find(name);
  if (it == m->slots.end()) { s2s_set_error("codec bind_tensor: unknown tensor '%s'", name); return S2S_ERR_NOT_FOUND; }
  CSlot& s = it->second;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  S2S_REQUIRE(n == s.n, "codec bind_tensor: '%s' has %lld elements, expected %lld", name, (long long)n, (long long)s.n);
  S2S_CHECK_CUDA(cudaSetDevice(m->ctx->device));
  if (s.kind == C_LINEAR && s.d2 > 0) {
    // column block [col0, col0 + N) of a shared [K][N_total] matrix: one strided 2-D copy
    const int N = s.d0, K = s.d1, Nt = s.d2, col0 = s.stride;
    std::vector<float> tmp((size_t)K * N);
    for (int nn = 0; nn < N; ++nn)
      for (int k = 0; k < K; ++k) tmp[(size_t)k * N + nn] = src_f32(data_h, (int64_t)nn * K + k, dtype);
    S2S_CHECK_CUDA(cudaMemcpy2D(s.dst + col0, (size_t)Nt * 4, tmp.data(), (size_t)N * 4, (size_t)N * 4, K, cudaMemcpyHostToDevice));
    s.bound = true;
    return S2S_OK;
  }
  std::vector<float> tmp((size_t)n);
  if (s.kind == C_LINEAR) {
    const int N = s.d0, K = s.d1;
    for (int nn = 0; nn < N; ++nn)
      for (int k = 0; k < K; ++k) tmp[(size_t)k * N + nn] = src_f32(data_h, (int64_t)nn * K + k, dtype);
  } else if (s.kind == C_CONV) {
    const int O = s.d0, C = s.d1, k = s.d2;
    for (int o = 0; o < O; ++o)
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < k; ++j) tmp[((size_t)j * C + c) * O + o] = src_f32(data_h, ((int64_t)o * C + c) * k + j, dtype);
  } else if (s.kind == C_TCONV) {
    // [C, O, k] with k = s or 2 s.  k = s: one tap, W'[c][r * O + o] = W[c][o][r].
    // k = 2 s: tap 0 multiplies X[t0 - 1] -> W[c][o][r + s]; tap 1 multiplies X[t0] -> W[c][o][r].
    const int C = s.d0, O = s.d1, k = s.d2, st = s.stride, taps = k / st;
    for (int c = 0; c < C; ++c)
      for (int o = 0; o < O; ++o)
        for (int r = 0; r < st; ++r) {
          if (taps == 1) tmp[(size_t)c * st * O + (size_t)r * O + o] = src_f32(data_h, ((int64_t)c * O + o) * k + r, dtype);
          else {
            tmp[((size_t)0 * C + c) * st * O + (size_t)r * O + o] = src_f32(data_h, ((int64_t)c * O + o) * k + r + st, dtype);
            tmp[((size_t)1 * C + c) * st * O + (size_t)r * O + o] = src_f32(data_h, ((int64_t)c * O + o) * k + r, dtype);
          }
        }
  } else {
    for (int64_t i = 0; i < n; ++i) tmp[(size_t)i] = src_f32(data_h, i, dtype);
  }
  S2S_CHECK_CUDA(cudaMemcpy(s.dst, tmp.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
  s.bound = true;
  return S2S_OK;
}

int codec_init_random(CodecDecoder* m, uint64_t seed) {
  S2S_REQUIRE(m, "codec init_random: null model");
  S2S_CHECK_CUDA(cudaSetDevice(m->ctx->device));
  for (auto& kv : m->slots) {
    CSlot& s = kv.second;
    uint64_t hsh = 1469598103934665603ull;
    for (char ch : kv.first) hsh = (hsh ^ (uint64_t)(unsigned char)ch) * 1099511628211ull;
    if (s.kind == C_LINEAR && s.d2 > 0) {
      // column block of a shared matrix: fill row by row
      std::vector<float> tmp((size_t)s.d0 * s.d1);
      uint64_t st = seed ^ hsh;
      for (auto& v : tmp) { st = st * 6364136223846793005ull + 1442695040888963407ull; v = s.rnd_scale * (((st >> 40) & 0xffff) / 32768.0f - 1.0f); }
      S2S_CHECK_CUDA(cudaMemcpy2D(s.dst + s.stride, (size_t)s.d2 * 4, tmp.data(), (size_t)s.d0 * 4, (size_t)s.d0 * 4, s.d1, cudaMemcpyHostToDevice));
    } else {
      S2S_CHECK(fill_random_launch(s.dst, s.n, S2S_F32, s.rnd_scale, s.rnd_offset, seed ^ hsh, 0));
    }
    s.bound = true;
  }
  S2S_CHECK_CUDA(cudaDeviceSynchronize());
  return S2S_OK;
}

int codec_finalize(CodecDecoder* m) {
  S2S_REQUIRE(m, "codec finalize: null model");
  for (auto& kv : m->slots)
    if (!kv.second.bound) { s2s_set_error("codec finalize: tensor '%s' was never bound", kv.first.c_str()); return S2S_ERR_INVALID; }
  S2S_CHECK_CUDA(cudaSetDevice(m->ctx->device));
  if (!m->finalized)   // SnakeBeta parameters -> (exp(alpha), 1 / (exp(beta) + 1e-9)), once
    for (size_t i = 0; i < m->snake_pairs.size(); ++i) {
      exp_prep_kernel<<<(m->snake_len[i] + 255) / 256, 256>>>(m->snake_pairs[i].first, m->snake_pairs[i].second, m->snake_len[i]);
      S2S_LAUNCH_CHECK();
    }
  S2S_CHECK_CUDA(cudaDeviceSynchronize());
  m->finalized = true;
  return S2S_OK;
}

int codec_decode(CodecDecoder* m, const int32_t* codes_d, int T, int ctx_frames, float* wav_out_d, int32_t* n_out_h,
                 float* hidden_out_d, cudaStream_t st) {
  S2S_REQUIRE(m && m->finalized && codes_d && wav_out_d, "codec decode: null argument / not finalized");
  const auto& c = m->cfg;
  S2S_REQUIRE(T >= 1 && T <= c.max_frames, "codec decode: T=%d outside [1,%d]", T, c.max_frames);
  S2S_REQUIRE(ctx_frames >= 0 && ctx_frames < T, "codec decode: context %d must be smaller than T=%d", ctx_frames, T);
  S2S_CHECK_CUDA(cudaSetDevice(m->ctx->device));
  const int H = c.hidden, hd = H / c.heads, qd = c.heads * hd, kvd = c.kv_heads * hd, I = c.inter, ldq = qd + 2 * kvd;
  code_embed_mean_kernel<<<T, 256, 0, st>>>(codes_d, c.quantizers, c.codebook_size, m->embed, H, m->x);
  S2S_LAUNCH_CHECK();
  for (int l = 0; l < c.layers; ++l) {
    const CodecLayer& L = m->layers[l];
    rmsnorm_rows_f32_kernel<<<(T + 7) / 8, 256, 0, st>>>(m->x, L.n1, c.rms_eps, T, H, m->xn);
    S2S_LAUNCH_CHECK();
    S2S_CHECK(conv1d_f32_launch(linear_args(m->xn, T, H, L.w_qkv, ldq, m->qkv), st));
    rope_rows_f32_kernel<<<T, 256, 0, st>>>(m->qkv, T, c.heads + c.kv_heads, hd, ldq, c.rope_theta);
    S2S_LAUNCH_CHECK();
    {
      const int nw = 8;
      const size_t sm = (size_t)nw * (hd + c.sliding_window) * 4;
      swa_attention_f32_kernel<<<(T * c.heads + nw - 1) / nw, nw * 32, sm, st>>>(m->qkv, T, c.heads, c.kv_heads, hd, c.sliding_window, m->att);
      S2S_LAUNCH_CHECK();
    }
    {
      ConvArgs a = linear_args(m->att, T, qd, L.w_o, H, m->x);
      a.scale = L.ls_attn; a.resid = m->x; a.ldr = H;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
    rmsnorm_rows_f32_kernel<<<(T + 7) / 8, 256, 0, st>>>(m->x, L.n2, c.rms_eps, T, H, m->xn);
    S2S_LAUNCH_CHECK();
    S2S_CHECK(conv1d_f32_launch(linear_args(m->xn, T, H, L.w_gu, 2 * I, m->gu), st));
    silu_mul_kernel<<<(unsigned)(((long long)T * I + 255) / 256), 256, 0, st>>>(m->gu, I, (long long)T * I, m->hmid);
    S2S_LAUNCH_CHECK();
    {
      ConvArgs a = linear_args(m->hmid, T, I, L.w_down, H, m->x);
      a.scale = L.ls_mlp; a.resid = m->x; a.ldr = H;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
  }
  rmsnorm_rows_f32_kernel<<<(T + 7) / 8, 256, 0, st>>>(m->x, m->norm_f, c.rms_eps, T, H, m->bufX);
  S2S_LAUNCH_CHECK();
  if (hidden_out_d) S2S_CHECK_CUDA(cudaMemcpyAsync(hidden_out_d, m->bufX, (size_t)T * H * 4, cudaMemcpyDeviceToDevice, st));
  // upsampling stages: transposed conv (k = stride) + ConvNeXt
  float *X = m->bufX, *A = m->bufA, *Hb = m->bufH;
  int len = T;
  for (int i = 0; i < c.n_upsampling_ratios; ++i) {
    const CodecConvNeXt& U = m->ups[i];
    const int f = c.upsampling_ratios[i];
    {
      ConvArgs a = linear_args(X, len, H, U.up_w, f * H, A);
      a.bias = U.up_b; a.bias_mod = H;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
    len *= f;   // A is [len, H]
    dwconv7_ln_kernel<<<len, 256, (size_t)H * 4, st>>>(A, len, H, U.dw_w, U.dw_b, U.ln_w, U.ln_b, 1e-6f, Hb);
    S2S_LAUNCH_CHECK();
    {
      ConvArgs a = linear_args(Hb, len, H, U.pw1_w, 4 * H, X);
      a.bias = U.pw1_b; a.act = 1;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
    {
      ConvArgs a = linear_args(X, len, 4 * H, U.pw2_w, H, A);
      a.bias = U.pw2_b; a.scale = U.gamma; a.resid = A; a.ldr = H;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
    std::swap(X, A);   // X = stage output [len, H]
  }
  const int D = c.decoder_dim;
  S2S_CHECK(conv1d_f32_launch(causal_conv_args(X, len, H, m->d0_w, m->d0_b, 7, 1, D, A), st));
  std::swap(X, A);
  for (int i = 0; i < c.n_upsample_rates; ++i) {
    const CodecBlock& B = m->blocks[i];
    S2S_CHECK(snake_launch(X, B.a0, B.b0, B.cin, (long long)len * B.cin, A, st));
    {
      ConvArgs a{};
      a.x = A; a.ldx = B.cin; a.T_in = len; a.x_row0 = 0; a.w = B.tc_w; a.k = 2; a.dil = 1; a.C_in = B.cin; a.N = B.rate * B.cout;
      a.bias = B.tc_b; a.bias_mod = B.cout; a.y = X; a.ldy = (long long)B.rate * B.cout; a.T_out = len - 1; a.batch = 1;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
    len = (len - 1) * B.rate;   // X is [len, cout]
    static const int dil[3] = {1, 3, 9};
    for (int u = 0; u < 3; ++u) {
      const CodecResUnit& R = B.u[u];
      S2S_CHECK(snake_launch(X, R.a1, R.b1, B.cout, (long long)len * B.cout, A, st));
      S2S_CHECK(conv1d_f32_launch(causal_conv_args(A, len, B.cout, R.c1_w, R.c1_b, 7, dil[u], B.cout, Hb), st));
      S2S_CHECK(snake_launch(Hb, R.a2, R.b2, B.cout, (long long)len * B.cout, A, st));
      ConvArgs a = causal_conv_args(A, len, B.cout, R.c2_w, R.c2_b, 1, 1, B.cout, X);
      a.resid = X; a.ldr = B.cout;
      S2S_CHECK(conv1d_f32_launch(a, st));
    }
  }
  const int cl = D >> c.n_upsample_rates;
  S2S_CHECK(snake_launch(X, m->fa, m->fb, cl, (long long)len * cl, A, st));
  S2S_CHECK(conv1d_f32_launch(causal_conv_args(A, len, cl, m->f_w, m->f_b, 7, 1, 1, Hb), st));
  const int skip = ctx_frames * m->total_up;
  const int n_out = len - skip;
  S2S_REQUIRE(n_out > 0, "codec decode: nothing left after dropping %d context frames", ctx_frames);
  clamp_out_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(Hb, 1, skip, n_out, wav_out_d);
  S2S_LAUNCH_CHECK();
  if (n_out_h) *n_out_h = n_out;
  return S2S_OK;
}

extern "C" {
int s2s_codec_create(s2s_ctx* ctx, const s2s_codec_config* cfg, s2s_codec** out) { return codec_create(ctx, cfg, out); }
int s2s_codec_destroy(s2s_codec* m) { return codec_destroy(m); }
int s2s_codec_bind_tensor(s2s_codec* m, const char* name, const void* data_h, const int64_t* shape, int32_t ndim, int32_t dtype) {
  return codec_bind_tensor(m, name, data_h, shape, ndim, dtype);
}
int s2s_codec_init_random(s2s_codec* m, uint64_t seed) { return codec_init_random(m, seed); }
int s2s_codec_finalize(s2s_codec* m) { return codec_finalize(m); }
int s2s_codec_decode(s2s_codec* m, const int32_t* codes_d, int32_t T, int32_t ctx_frames, float* wav_out_d, int32_t* n_out_h,
                     float* hidden_out_d, void* stream) {
  return codec_decode(m, codes_d, T, ctx_frames, wav_out_d, n_out_h, hidden_out_d, (cudaStream_t)stream);
}
int32_t s2s_codec_samples(s2s_codec* m, int32_t T) { return m ? codec_samples_for(m, T) : 0; }
int32_t s2s_codec_total_upsample(s2s_codec* m) { return m ? codec_total_upsample(m) : 0; }
}  // extern "C"
