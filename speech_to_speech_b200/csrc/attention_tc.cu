// attention_tc.cu -- fused softmax(Q K^T) V on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// One CTA = 128 query rows of one head.  Per 64-key tile:
//   S = Q K^T        tcgen05.mma  M=128 N=64  K=hd     (Q, K tiles in shared memory via TMA, S in TMEM cols [0,64))
//   softmax          4 warps, one query row per thread: tcgen05.ld S -> registers, mask, online max/sum (exp2, fp32),
//                    P (16-bit) written to shared memory in the K-major 128-byte-swizzled operand layout,
//                    O (TMEM cols [64, 64+hd)) rescaled in place with tcgen05.ld / tcgen05.st
//   O += P V         tcgen05.mma  M=128 N=hd  K=64      (V^T tile [hd, 64 keys] in shared memory via TMA)
// Warp roles (192 threads): warps 0-3 softmax/epilogue (TMEM lane group = warp id), warp 4 TMA producer,
// warp 5 TMEM allocator + MMA issuer.  K/V tiles are double buffered; S(t+1) is issued while softmax(t) runs.
// V is consumed as V^T (K-major B operand), produced once per call by a small transpose kernel into a
// caller-owned scratch [B, kv_heads, hd, Tpad].
// Reference ops: transformers modeling_whisper.py:215-239 (eager_attention_forward, encoder: non-causal 1500x1500),
// modeling_llama.py:187-222 (causal GQA, prefill).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

namespace {

constexpr int QM = 128;  // queries per CTA
constexpr int KN = 64;   // keys per tile
constexpr int ATC_THREADS = 192;

struct AtcParams {
  void* o;
  long long ldo;
  int B, Tq, Tk, heads, kv_heads;
  float scale_log2;
  int causal_off;  // query i sees keys <= i + causal_off
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int HD> struct AtcSmem {
  static constexpr int Q_BYTES = QM * HD * 2;
  static constexpr int K_BYTES = KN * HD * 2;   // per stage
  static constexpr int V_BYTES = HD * KN * 2;   // per stage
  static constexpr int P_BYTES = QM * KN * 2;
  static constexpr int OFF_K = Q_BYTES;
  static constexpr int OFF_V = OFF_K + 2 * K_BYTES;
  static constexpr int OFF_P = OFF_V + 2 * V_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

template <typename T, int HD, bool CAUSAL>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_vt, const AtcParams p) {
  using L = AtcSmem<HD>;
  constexpr int KK = HD / 64;  // 64-element (128-byte) swizzle atoms along the head dimension
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + L::OFF_K;
  uint8_t* sV = smem + L::OFF_V;
  uint8_t* sP = smem + L::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_ready = bars + 7;
  uint64_t* o_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * QM, h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.heads / p.kv_heads);

  int n_tiles = (p.Tk + KN - 1) / KN;
  if (CAUSAL) {
    const int last_key = min(p.Tk - 1, q0 + QM - 1 + p.causal_off);
    n_tiles = min(n_tiles, last_key / KN + 1);
    if (n_tiles < 1) n_tiles = 1;
  }

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_vt);
    mbar_init(q_full, 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_ready, 128);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // columns [0, 64)
  const uint32_t tmem_O = tmem_base + 64;   // columns [64, 64 + HD)

  if (warp == 4) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      mbar_expect_tx(q_full, L::Q_BYTES);
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) tma_load_2d(sQ + kk * (QM * 128), &map_q, q_full, h * HD + kk * 64, b * p.Tq + q0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t & 1;
        mbar_wait(&kv_empty[s], ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[s], L::K_BYTES + L::V_BYTES);
#pragma unroll
        for (int kk = 0; kk < KK; ++kk)
          tma_load_2d(sK + s * L::K_BYTES + kk * (KN * 128), &map_k, &kv_full[s], kvh * HD + kk * 64, b * p.Tk + t * KN);
        tma_load_3d(sV + s * L::V_BYTES, &map_vt, &kv_full[s], t * KN, 0, b * p.kv_heads + kvh);
      }
    }
  } else if (warp == 5) {
    // ============================== MMA issuer (one thread) ==============================
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(DT<T>::umma_fmt, QM, KN);
      constexpr uint32_t idesc_o = umma_idesc_f16(DT<T>::umma_fmt, QM, HD);
      auto issue_s = [&](int t) {
        const int s = t & 1;
        mbar_wait(&kv_full[s], (t >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const uint64_t ad = umma_smem_desc_sw128(smem_u32(sQ + kk * (QM * 128)));
          const uint64_t bd = umma_smem_desc_sw128(smem_u32(sK + s * L::K_BYTES + kk * (KN * 128)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(tmem_S, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_s, (kk | k) != 0);
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t & 1;
        if (t + 1 < n_tiles) {
          mbar_wait(s_free, t & 1);  // softmax(t) holds S(t) in registers: the S columns may be overwritten
          issue_s(t + 1);
        }
        mbar_wait(p_ready, t & 1);   // P(t) is in shared memory, O has been rescaled
        tc_fence_after();
        const uint64_t ad = umma_smem_desc_sw128(smem_u32(sP));
        const uint64_t bd = umma_smem_desc_sw128(smem_u32(sV + s * L::V_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_O, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc_o, (t | k) != 0);
        umma_commit(o_done);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ============================== softmax + epilogue: one query row per thread ==============================
    const int row = warp * 32 + lane;          // TMEM lane == query row inside the tile
    const int q = q0 + row;                    // query index inside the sequence
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    float m = -INFINITY, l = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld_32x32(tmem_S + lane_sel + 0, sr[0]);
      tmem_ld_32x32(tmem_S + lane_sel + 32, sr[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);
      float mx = -INFINITY;
      const int kbase = t * KN;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int key = kbase + hh * 32 + j;
          bool ok = key < p.Tk;
          if (CAUSAL) ok = ok && (key <= q + p.causal_off);
          const float v = ok ? __uint_as_float(sr[hh][j]) * p.scale_log2 : -INFINITY;
          sr[hh][j] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      const float m_new = fmaxf(m, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float corr = exp2f(m - m_use);  // m = -inf -> 0
      float rs = 0.f;
      uint32_t pk[32];  // 64 probabilities packed 2 x 16-bit
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = exp2f(__uint_as_float(sr[hh][j]) - m_use);
          const float p1 = exp2f(__uint_as_float(sr[hh][j + 1]) - m_use);
          rs += p0 + p1;
          pk[hh * 16 + (j >> 1)] = DT<T>::pack2(p0, p1);
        }
      l = l * corr + rs;
      m = m_new;
      if (t > 0) {
        mbar_wait(o_done, (t - 1) & 1);  // PV(t-1) finished: P buffer free, O stable
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t orr[32];
          tmem_ld_32x32(tmem_O + lane_sel + c * 32, orr);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * corr);
          tmem_st_32x32(tmem_O + lane_sel + c * 32, orr);
        }
        tmem_st_wait();
      }
      // P row -> shared memory, K-major 128-byte swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r & 7))
      uint8_t* prow = sP + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(prow + ((c ^ (row & 7)) << 4)) = make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> 16-bit global ----
    mbar_wait(o_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv = l > 0.f ? 1.f / l : 0.f;
    T* orow = reinterpret_cast<T*>(p.o) + ((long long)b * p.Tq + q) * p.ldo + (long long)h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t orr[32];
      tmem_ld_32x32(tmem_O + lane_sel + c * 32, orr);
      tmem_ld_wait();
      if (q < p.Tq) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 w;
          w.x = DT<T>::pack2(__uint_as_float(orr[j]) * inv, __uint_as_float(orr[j + 1]) * inv);
          w.y = DT<T>::pack2(__uint_as_float(orr[j + 2]) * inv, __uint_as_float(orr[j + 3]) * inv);
          w.z = DT<T>::pack2(__uint_as_float(orr[j + 4]) * inv, __uint_as_float(orr[j + 5]) * inv);
          w.w = DT<T>::pack2(__uint_as_float(orr[j + 6]) * inv, __uint_as_float(orr[j + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c * 32 + j) = w;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// v [B*Tk rows, ld] (offset to the V columns applied by the caller) -> vt [B, kv_heads, hd, Tpad]
template <typename T>
__global__ void __launch_bounds__(256) transpose_v_kernel(const T* __restrict__ v, long long ld, int Tk, int cols, int hd,
                                                          int Tpad, T* __restrict__ vt) {
  __shared__ T tile[64][66];
  const int t0 = blockIdx.x * 64, c0 = blockIdx.y * 64, b = blockIdx.z;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    T val = DT<T>::from_f(0.f);
    if (t0 + r < Tk && c0 + c < cols) val = v[((long long)b * Tk + t0 + r) * ld + c0 + c];
    tile[r][c] = val;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, r = i & 63;  // r fastest: coalesced along keys
    const int col = c0 + c;
    if (col < cols && t0 + r < Tpad) {
      const int head = col / hd, dd = col % hd;
      vt[(((long long)b * (cols / hd) + head) * hd + dd) * Tpad + t0 + r] = tile[r][c];
    }
  }
}

template <typename T, int HD, bool CAUSAL>
int launch_atc(s2s_ctx* ctx, const void* q, const void* k, const void* v, void* o, int B, int Tq, int Tk, int heads,
               int kv_heads, long long ldq, long long ldk, long long ldv, long long ldo, float scale, void* vt, int Tpad,
               cudaStream_t stream) {
  using L = AtcSmem<HD>;
  {
    dim3 g((Tpad + 63) / 64, (kv_heads * HD + 63) / 64, B);
    transpose_v_kernel<T><<<g, 256, 0, stream>>>(reinterpret_cast<const T*>(v), ldv, Tk, kv_heads * HD, HD, Tpad,
                                                 reinterpret_cast<T*>(vt));
    S2S_LAUNCH_CHECK();
  }
  CUtensorMap mq, mk, mv;
  {
    cuuint64_t dims[2] = {(cuuint64_t)heads * HD, (cuuint64_t)B * Tq};
    cuuint64_t strides[1] = {(cuuint64_t)ldq * 2};
    cuuint32_t box[2] = {64, QM};
    S2S_CHECK(tma_encode_map(ctx, &mq, DT<T>::tma, 2, q, dims, strides, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)kv_heads * HD, (cuuint64_t)B * Tk};
    cuuint64_t strides[1] = {(cuuint64_t)ldk * 2};
    cuuint32_t box[2] = {64, KN};
    S2S_CHECK(tma_encode_map(ctx, &mk, DT<T>::tma, 2, k, dims, strides, box));
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Tpad, (cuuint64_t)HD, (cuuint64_t)B * kv_heads};
    cuuint64_t strides[2] = {(cuuint64_t)Tpad * 2, (cuuint64_t)HD * Tpad * 2};
    cuuint32_t box[3] = {64, HD, 1};
    S2S_CHECK(tma_encode_map(ctx, &mv, DT<T>::tma, 3, vt, dims, strides, box));
  }
  AtcParams p;
  p.o = o; p.ldo = ldo; p.B = B; p.Tq = Tq; p.Tk = Tk; p.heads = heads; p.kv_heads = kv_heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal_off = Tk - Tq;
  static bool attr_set = false;
  if (!attr_set) {
    S2S_CHECK_CUDA(cudaFuncSetAttribute(attn_tc_kernel<T, HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    attr_set = true;
  }
  dim3 grid((Tq + QM - 1) / QM, heads, B);
  attn_tc_kernel<T, HD, CAUSAL><<<grid, ATC_THREADS, L::TOTAL, stream>>>(mq, mk, mv, p);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}

template <typename T>
int dispatch_atc(s2s_ctx* ctx, const void* q, const void* k, const void* v, void* o, int B, int Tq, int Tk, int heads,
                 int kv_heads, int hd, long long ldq, long long ldk, long long ldv, long long ldo, float scale, int causal,
                 void* vt, int Tpad, cudaStream_t st) {
#define ATC_GO(HD_, C_) return launch_atc<T, HD_, C_>(ctx, q, k, v, o, B, Tq, Tk, heads, kv_heads, ldq, ldk, ldv, ldo, scale, vt, Tpad, st)
  if (hd == 64) { if (causal) ATC_GO(64, true); else ATC_GO(64, false); }
  if (hd == 128) { if (causal) ATC_GO(128, true); else ATC_GO(128, false); }
#undef ATC_GO
  s2s_set_error("attention_tc: head_dim %d unsupported (64 or 128)", hd);
  return S2S_ERR_UNSUPPORTED;
}

}  // namespace

size_t attention_tc_scratch_elems(int B, int Tk, int kv_heads, int hd) {
  const int Tpad = (Tk + 63) / 64 * 64;
  return (size_t)B * kv_heads * hd * Tpad;
}

int attention_tc_launch(s2s_ctx* ctx, const void* q, const void* k, const void* v, void* o, int B, int Tq, int Tk,
                        int heads, int kv_heads, int hd, long long ldq, long long ldk, long long ldv, long long ldo,
                        float scale, int causal, int dtype, void* vt_scratch, size_t vt_elems, cudaStream_t stream) {
  S2S_REQUIRE(B > 0 && Tq > 0 && Tk > 0 && heads > 0 && kv_heads > 0 && heads % kv_heads == 0, "attention_tc: bad shape");
  S2S_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldo % 8 == 0, "attention_tc: row strides must be multiples of 8 elements");
  S2S_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention_tc: q/k/o must be 16-byte aligned");
  const int Tpad = (Tk + 63) / 64 * 64;
  S2S_REQUIRE(vt_scratch && vt_elems >= attention_tc_scratch_elems(B, Tk, kv_heads, hd), "attention_tc: V^T scratch too small");
  if (dtype == S2S_F16)
    return dispatch_atc<__half>(ctx, q, k, v, o, B, Tq, Tk, heads, kv_heads, hd, ldq, ldk, ldv, ldo, scale, causal, vt_scratch, Tpad, stream);
  if (dtype == S2S_BF16)
    return dispatch_atc<__nv_bfloat16>(ctx, q, k, v, o, B, Tq, Tk, heads, kv_heads, hd, ldq, ldk, ldv, ldo, scale, causal, vt_scratch, Tpad, stream);
  s2s_set_error("attention_tc: unsupported dtype %d", dtype);
  return S2S_ERR_UNSUPPORTED;
}
