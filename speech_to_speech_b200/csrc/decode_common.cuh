// decode_common.cuh -- building blocks of the persistent single-token decode kernels
// (Whisper decoder, Llama-family decoder).  All of them are HBM-bound weight/KV streaming loops:
// 16-byte coalesced loads through the read-only no-L1-allocate path, fp32 accumulation, warp-shuffle
// reductions.  One CTA per SM, DEC_THREADS threads; phases are separated by a grid-wide barrier.
#pragma once
#include "common.cuh"

constexpr int DEC_THREADS = 512;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int ATT_CHUNK = 64;   // keys per attention work item
constexpr int PART_STRIDE = 2;  // partial record = [m, l, o[hd]]

// ---- grid barrier (all CTAs co-resident: cooperative launch) ------------------------------
// Monotonic counter: barrier #e completes when counter == e * gridDim.x.  Bounded spin -> trap.
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    const unsigned int target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 28)) __trap();
    } while (v < target);
    __threadfence();  // gpu-scope fence: drops stale L1 lines before the CTA reads other CTAs' results
  }
  __syncthreads();
}

// ---- normalise B rows of x (fp32 [B, d]) into shared memory xs[B][d] -----------------------
// bias != null: LayerNorm ; bias == null: RMSNorm.  One warp per row.
__device__ __forceinline__ void norm_rows_to_smem(const float* x, const float* w, const float* bias, float eps, int B,
                                                  int d, float* xs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += DEC_WARPS) {
    const float* xr = x + (long long)b * d;
    float mean = 0.f;
    if (bias) {
      float s = 0.f;
      for (int i = lane; i < d; i += 32) s += __ldcg(xr + i);
      mean = warp_sum(s) / (float)d;
    }
    float ss = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float a = __ldcg(xr + i) - mean;
      ss += a * a;
    }
    const float rstd = rsqrtf(warp_sum(ss) / (float)d + eps);
    for (int i = lane; i < d; i += 32) {
      float y = (__ldcg(xr + i) - mean) * rstd * w[i];
      if (bias) y += bias[i];
      xs[b * d + i] = y;
    }
  }
}

__device__ __forceinline__ void copy_rows_to_smem(const float* x, int n, float* xs) {
  for (int i = threadIdx.x * 4; i < n; i += DEC_THREADS * 4)
    *reinterpret_cast<float4*>(xs + i) = __ldcg(reinterpret_cast<const float4*>(x + i));
}

// ---- skinny GEMV: out[b][row] = sum_k W[row][k] * xs[b][k], rows distributed over every warp of the grid
// Each warp takes R consecutive rows so that R * (K/256) independent 16-byte loads are in flight per lane.
template <typename T, int NB, int R, typename Epi>
__device__ __forceinline__ void gemv_rows(const T* __restrict__ W, int N, int K, const float* xs, int B, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * DEC_WARPS + warp;
  const int GW = gridDim.x * DEC_WARPS;
  for (int row0 = gw * R; row0 < N; row0 += GW * R) {
    float acc[R][NB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
      uint4 wv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int row = min(row0 + r, N - 1);
        wv[r] = ld_stream16(W + (long long)row * K + k);
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (b < B) {
          const float4 x0 = *reinterpret_cast<const float4*>(xs + b * K + k);
          const float4 x1 = *reinterpret_cast<const float4*>(xs + b * K + k + 4);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float2 w0 = DT<T>::to_f2(wv[r].x), w1 = DT<T>::to_f2(wv[r].y);
            const float2 w2 = DT<T>::to_f2(wv[r].z), w3 = DT<T>::to_f2(wv[r].w);
            float a = acc[r][b];
            a = fmaf(w0.x, x0.x, a); a = fmaf(w0.y, x0.y, a);
            a = fmaf(w1.x, x0.z, a); a = fmaf(w1.y, x0.w, a);
            a = fmaf(w2.x, x1.x, a); a = fmaf(w2.y, x1.y, a);
            a = fmaf(w3.x, x1.z, a); a = fmaf(w3.y, x1.w, a);
            acc[r][b] = a;
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);
      if (row0 + r < N) epi(row0 + r, acc[r], lane);
    }
  }
}

// ---- attention over one chunk of <= 64 keys for one (batch, head): warp-level -----------------
// 8 lanes cover one key row (HD 16-bit values, HD/8 per lane... for HD=64: 16 B per lane; HD=128: 2 x 16 B),
// 4 keys per warp step, 16 steps.  Scores are kept in registers so that all K loads (then all V loads)
// are independent.  q must already carry the softmax scale.  Writes [m, l, o[HD]] (unnormalised).
template <typename T, int HD>
__device__ __forceinline__ void attend_chunk(const float* q /*global fp32 [HD]*/, const T* K, const T* V,
                                             long long ldk, long long ldv, int n_keys, float* part) {
  constexpr int PER = HD / 8;      // elements per lane
  constexpr int NV = PER / 8;      // 16-byte vectors per lane
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  float qr[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) qr[i] = __ldcg(q + j * PER + i);
  float s[ATT_CHUNK / 4];
#pragma unroll
  for (int i = 0; i < ATT_CHUNK / 4; ++i) {
    const int key = g + 4 * i;
    float dot = 0.f;
    if (key < n_keys) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint4 kv = __ldcg(reinterpret_cast<const uint4*>(K + (long long)key * ldk + j * PER + v * 8));
        const float2 a = DT<T>::to_f2(kv.x), b = DT<T>::to_f2(kv.y), c = DT<T>::to_f2(kv.z), d = DT<T>::to_f2(kv.w);
        dot = fmaf(a.x, qr[v * 8 + 0], dot); dot = fmaf(a.y, qr[v * 8 + 1], dot);
        dot = fmaf(b.x, qr[v * 8 + 2], dot); dot = fmaf(b.y, qr[v * 8 + 3], dot);
        dot = fmaf(c.x, qr[v * 8 + 4], dot); dot = fmaf(c.y, qr[v * 8 + 5], dot);
        dot = fmaf(d.x, qr[v * 8 + 6], dot); dot = fmaf(d.y, qr[v * 8 + 7], dot);
      }
    }
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    s[i] = (key < n_keys) ? dot : -INFINITY;
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < ATT_CHUNK / 4; ++i) m = fmaxf(m, s[i]);
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
  float l = 0.f;
  float o[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) o[i] = 0.f;
#pragma unroll
  for (int i = 0; i < ATT_CHUNK / 4; ++i) {
    const int key = g + 4 * i;
    if (key < n_keys) {
      const float p = __expf(s[i] - m);
      l += p;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const uint4 vv = __ldcg(reinterpret_cast<const uint4*>(V + (long long)key * ldv + j * PER + v * 8));
        const float2 a = DT<T>::to_f2(vv.x), b = DT<T>::to_f2(vv.y), c = DT<T>::to_f2(vv.z), d = DT<T>::to_f2(vv.w);
        o[v * 8 + 0] = fmaf(p, a.x, o[v * 8 + 0]); o[v * 8 + 1] = fmaf(p, a.y, o[v * 8 + 1]);
        o[v * 8 + 2] = fmaf(p, b.x, o[v * 8 + 2]); o[v * 8 + 3] = fmaf(p, b.y, o[v * 8 + 3]);
        o[v * 8 + 4] = fmaf(p, c.x, o[v * 8 + 4]); o[v * 8 + 5] = fmaf(p, c.y, o[v * 8 + 5]);
        o[v * 8 + 6] = fmaf(p, d.x, o[v * 8 + 6]); o[v * 8 + 7] = fmaf(p, d.y, o[v * 8 + 7]);
      }
    }
  }
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < PER; ++i) part[PART_STRIDE + j * PER + i] = o[i];
  }
  if (lane == 0) { part[0] = m; part[1] = l; }
}

// ---- merge attention partials into shared memory xs[b][h*HD + dd] (normalised) ---------------
// part layout: [B][H][s_max][2 + HD]; n_chunks valid records per (b, h).
template <int HD>
__device__ __forceinline__ void combine_partials_to_smem(const float* part, int B, int H, int s_max, int n_chunks,
                                                         float* xs) {
  const int D = H * HD;
  for (int e = threadIdx.x; e < B * D; e += DEC_THREADS) {
    const int b = e / D, r = e % D, h = r / HD, dd = r % HD;
    const float* pp = part + ((long long)(b * H + h) * s_max) * (PART_STRIDE + HD);
    float M = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) M = fmaxf(M, __ldcg(pp + c * (PART_STRIDE + HD)));
    float num = 0.f, den = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      const float* rec = pp + c * (PART_STRIDE + HD);
      const float wgt = __expf(__ldcg(rec) - M);
      den = fmaf(__ldcg(rec + 1), wgt, den);
      num = fmaf(__ldcg(rec + PART_STRIDE + dd), wgt, num);
    }
    xs[e] = num / den;
  }
}
