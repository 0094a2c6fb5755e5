// attention.cu -- fused softmax(Q K^T) V for the multi-token paths (Whisper encoder self-attention
// 1500x1500 non-causal, LLM prefill causal GQA).  Never materialises the score matrix.
//
// v1 layout: one CTA = 64 query rows of one head (4 warps x 16 rows), K/V streamed in 64-key tiles through
// a double-buffered cp.async pipeline into XOR-swizzled shared memory, QK^T and PV on the warp-level
// tensor-core path (mma.sync m16n8k16, fp32 accumulate), online softmax in the exp2 domain with fp32
// running max / sum.  Reference ops: transformers modeling_whisper.py:215-239 (eager_attention_forward),
// modeling_llama.py:187-222.  (A tcgen05/TMEM version is the planned replacement; see DESIGN.md.)
#include "common.cuh"

namespace {

struct AttnParams {
  const void* q; const void* k; const void* v; void* o;
  int B, Tq, Tk, heads, kv_heads;
  long long ldq, ldk, ldv, ldo;
  float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <typename T>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int ABM = 64;  // query rows per CTA
constexpr int ABN = 64;  // keys per tile

// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][HD] 16-bit tile, XOR-swizzled so that the
// 8 row addresses of one ldmatrix 8x8 fall in distinct bank groups.
template <int HD> __device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return (uint32_t)(row * (HD * 2) + (((chunk & ~7) | ((chunk ^ row) & 7)) << 4));
}

template <typename T, int HD>
__device__ __forceinline__ void load_tile(uint8_t* dst, const T* src, long long ld, int row0, int nrows_valid) {
  constexpr int CH = HD / 8;  // 16-byte chunks per row
  for (int i = threadIdx.x; i < ABN * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    const bool ok = (row0 + r) < nrows_valid;
    const T* g = src + (long long)(ok ? (row0 + r) : 0) * ld + c * 8;
    cp_async16(dst + swz<HD>(r, c), g, ok);
  }
}

template <typename T, int HD, bool CAUSAL>
__global__ void __launch_bounds__(128) flash_attn_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int TILE = ABN * HD * 2;
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE;      // 2 buffers
  uint8_t* sV = smem + 3 * TILE;  // 2 buffers

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ABM, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.heads / p.kv_heads);
  const T* Q = reinterpret_cast<const T*>(p.q) + (long long)b * p.Tq * p.ldq + (long long)h * HD;
  const T* K = reinterpret_cast<const T*>(p.k) + (long long)b * p.Tk * p.ldk + (long long)kvh * HD;
  const T* V = reinterpret_cast<const T*>(p.v) + (long long)b * p.Tk * p.ldv + (long long)kvh * HD;
  const int causal_off = p.Tk - p.Tq;  // query i sees keys <= i + causal_off

  int n_tiles = (p.Tk + ABN - 1) / ABN;
  if (CAUSAL) {
    const int last_key = min(p.Tk - 1, q0 + ABM - 1 + causal_off);
    n_tiles = min(n_tiles, last_key / ABN + 1);
  }

  load_tile<T, HD>(sQ, Q, p.ldq, q0, p.Tq);
  load_tile<T, HD>(sK, K, p.ldk, 0, p.Tk);
  load_tile<T, HD>(sV, V, p.ldv, 0, p.Tk);
  cp_async_commit();

  uint32_t qf[HD / 16][4];
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int r_lo = q0 + warp * 16 + (lane >> 2);  // query row of c0,c1 ; +8 for c2,c3

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      load_tile<T, HD>(sK + (buf ^ 1) * TILE, K, p.ldk, (t + 1) * ABN, p.Tk);
      load_tile<T, HD>(sV + (buf ^ 1) * TILE, V, p.ldv, (t + 1) * ABN, p.Tk);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = ks * 2 + (lane >> 4);
        ldsm_x4(smem_u32(sQ) + swz<HD>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T : 16 x 64 per warp ----
    float s[ABN / 8][4];
#pragma unroll
    for (int i = 0; i < ABN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    const uint32_t kbase = smem_u32(sK + buf * TILE);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
#pragma unroll
      for (int nb2 = 0; nb2 < ABN / 16; ++nb2) {
        uint32_t b0, b1, b2, b3;
        const int row = nb2 * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(kbase + swz<HD>(row, chunk), b0, b1, b2, b3);
        mma16816<T>(s[nb2 * 2], qf[ks], b0, b1);
        mma16816<T>(s[nb2 * 2 + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask + online softmax ----
    const int kv0 = t * ABN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < ABN / 8; ++nb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = kv0 + nb * 8 + (lane & 3) * 2 + (j & 1);
        const int qrow = r_lo + (j >> 1) * 8;
        bool ok = key < p.Tk;
        if (CAUSAL) ok = ok && (key <= qrow + causal_off);
        const float val = ok ? s[nb][j] * p.scale_log2 : -INFINITY;
        s[nb][j] = val;
        mx[j >> 1] = fmaxf(mx[j >> 1], val);
      }
    }
    float corr[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r]);
      const float m_use = (m_new[r] == -INFINITY) ? 0.f : m_new[r];
      corr[r] = exp2f(m_run[r] - m_use);  // m_run = -inf -> 0
      m_run[r] = m_new[r];
      m_new[r] = m_use;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[ABN / 16][4];
#pragma unroll
    for (int nb = 0; nb < ABN / 8; ++nb) {
      const float p0 = exp2f(s[nb][0] - m_new[0]);
      const float p1 = exp2f(s[nb][1] - m_new[0]);
      const float p2 = exp2f(s[nb][2] - m_new[1]);
      const float p3 = exp2f(s[nb][3] - m_new[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      // C fragment of n-blocks (2kk, 2kk+1) -> A fragment of k-step kk
      pf[nb >> 1][(nb & 1) * 2 + 0] = DT<T>::pack2(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = DT<T>::pack2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + rs[r];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V ----
    const uint32_t vbase = smem_u32(sV + buf * TILE);
#pragma unroll
    for (int kk = 0; kk < ABN / 16; ++kk) {
#pragma unroll
      for (int nb2 = 0; nb2 < HD / 16; ++nb2) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;  // key
        const int chunk = nb2 * 2 + (lane >> 4);                        // head-dim chunk
        ldsm_x4_t(vbase + swz<HD>(row, chunk), b0, b1, b2, b3);
        mma16816<T>(o[nb2 * 2], pf[kk], b0, b1);
        mma16816<T>(o[nb2 * 2 + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();
  }

  // ---- finalize ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  T* O = reinterpret_cast<T*>(p.o) + (long long)b * p.Tq * p.ldo + (long long)h * HD;
#pragma unroll
  for (int nb = 0; nb < HD / 8; ++nb) {
    const int col = nb * 8 + (lane & 3) * 2;
    if (r_lo < p.Tq)
      *reinterpret_cast<uint32_t*>(O + (long long)r_lo * p.ldo + col) = DT<T>::pack2(o[nb][0] * inv0, o[nb][1] * inv0);
    if (r_lo + 8 < p.Tq)
      *reinterpret_cast<uint32_t*>(O + (long long)(r_lo + 8) * p.ldo + col) = DT<T>::pack2(o[nb][2] * inv1, o[nb][3] * inv1);
  }
}

template <typename T, int HD, bool CAUSAL>
int launch_attn(const AttnParams& p, cudaStream_t stream) {
  constexpr int SMEM = 5 * ABN * HD * 2;
  static bool attr_set = false;
  if (!attr_set && SMEM > 48 * 1024) {
    S2S_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<T, HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    attr_set = true;
  }
  dim3 grid((p.Tq + ABM - 1) / ABM, p.heads, p.B);
  flash_attn_kernel<T, HD, CAUSAL><<<grid, 128, SMEM, stream>>>(p);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}

template <typename T>
int dispatch(const AttnParams& p, int hd, int causal, cudaStream_t stream) {
  if (hd == 64) return causal ? launch_attn<T, 64, true>(p, stream) : launch_attn<T, 64, false>(p, stream);
  if (hd == 128) return causal ? launch_attn<T, 128, true>(p, stream) : launch_attn<T, 128, false>(p, stream);
  s2s_set_error("attention: head_dim %d unsupported (64 or 128)", hd);
  return S2S_ERR_UNSUPPORTED;
}

}  // namespace

int attention_launch(const void* q, const void* k, const void* v, void* o, int B, int Tq, int Tk, int heads,
                     int kv_heads, int hd, long long ldq, long long ldk, long long ldv, long long ldo, float scale,
                     int causal, int dtype, cudaStream_t stream) {
  S2S_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0, "attention: bad shape");
  S2S_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: strides must be multiples of 8");
  AttnParams p;
  p.q = q; p.k = k; p.v = v; p.o = o; p.B = B; p.Tq = Tq; p.Tk = Tk; p.heads = heads; p.kv_heads = kv_heads;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  if (dtype == S2S_F16) return dispatch<__half>(p, hd, causal, stream);
  if (dtype == S2S_BF16) return dispatch<__nv_bfloat16>(p, hd, causal, stream);
  s2s_set_error("attention: unsupported dtype %d", dtype);
  return S2S_ERR_UNSUPPORTED;
}
