// gemm_tc.cu -- C = A * W^T on Blackwell 5th-gen tensor cores.
//
//   * operands: 16-bit (fp16 / bf16), K-major, staged global->shared by TMA (cp.async.bulk.tensor,
//     128-byte swizzle) through a STAGES-deep mbarrier ring;
//   * math: tcgen05.mma.cta_group::1.kind::f16, M=128 x N=BN x K=16 per instruction, issued by ONE
//     thread; the fp32 accumulator tile lives in TMEM (BN columns x 128 lanes);
//   * epilogue: 4 warps read TMEM with tcgen05.ld (32 lanes x 32 columns per instruction) and fuse
//     bias, exact GELU, fp32 residual add (in-place residual stream or broadcast positional table)
//     and the 16-bit / fp32 stores.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM lane group = warp_id % 4).
//
// Used for every GEMM-shaped op with M >= 64 on the hot path: Whisper conv1/conv2 (as strided-window
// GEMMs, no im2col), QKV / out-proj / fc1 / fc2 of the encoder, the cross-attention K/V projection and
// the LLM prefill projections.  (Reference ops: transformers modeling_whisper.py:284-357,380-414,619-625;
// modeling_llama.py:171-289.)
#include "gemm_tc.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 x 16-bit = 128 B = one swizzle atom
constexpr int UMMA_K = 16;
constexpr int NTHREADS = 192;

template <int BN> struct TileCfg {
  static constexpr int STAGES = (BN == 64) ? 8 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = (BN < 32) ? 32 : BN;
};

struct EpiParams {
  const float* bias;
  int act;
  void* out_h;
  long long ldo_h;
  float* out_f;
  long long ldo_f;
  const float* resid;
  long long ld_resid;
  int resid_mode;
  long long out_batch_rows;
  long long out_row_offset;
  int head_major_rows;  // > 0: out_h is [row / hmr][N / 64][hmr][64] (64-column blocks contiguous per row range)
  int M, N, K;
};

template <typename T, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
               const EpiParams ep) {
  using Cfg = TileCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment is required by the 128-byte swizzle pattern (TMA and UMMA must agree)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m_tile = blockIdx.y, batch = blockIdx.z;
  const int num_kb = (ep.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
        tma_load_3d(sa, &map_a, &full_bar[s], kb * BK, m_tile * BM, batch);
        tma_load_2d(sb, &map_w, &full_bar[s], kb * BK, n_tile * BN);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(DT<T>::umma_fmt, BM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::STAGES;
        const uint32_t ph = (kb / Cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
        const uint32_t sb = sa + Cfg::A_BYTES;
        const uint64_t adesc = umma_smem_desc_sw128(sa);
        const uint64_t bdesc = umma_smem_desc_sw128(sb);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advancing K inside the 128-byte swizzle atom = +32 B on the start address (>>4 => +2)
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(&empty_bar[s]);  // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);    // accumulator complete
    }
  } else {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int lg = warp & 3;  // TMEM lane group this warp may access
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int m = m_tile * BM + lg * 32 + lane;
    const bool valid = m < ep.M;
    const long long orow = (long long)batch * ep.out_batch_rows + ep.out_row_offset + m;
    const long long rrow = (ep.resid_mode == 2) ? (long long)m : orow;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
      const int n0 = n_tile * BN + c * 32;
      if (valid) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(ep.bias + n0 + j);
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          }
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (ep.act == 2) {
          // SwiGLU over interleaved (gate, up) columns: out[n/2] = silu(v[2i]) * v[2i+1]; 16-bit output, ld = N/2
          T* dst = reinterpret_cast<T*>(ep.out_h) + orow * ep.ldo_h + (n0 >> 1);
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float gte = v[2 * j];
            o[j] = (gte / (1.0f + __expf(-gte))) * v[2 * j + 1];
          }
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 q;
            q.x = DT<T>::pack2(o[j], o[j + 1]);
            q.y = DT<T>::pack2(o[j + 2], o[j + 3]);
            q.z = DT<T>::pack2(o[j + 4], o[j + 5]);
            q.w = DT<T>::pack2(o[j + 6], o[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = q;
          }
          continue;
        }
        if (ep.out_f) {
          float* dst = ep.out_f + orow * ep.ldo_f + n0;
          if (ep.resid) {
            const float* rs = ep.resid + rrow * ep.ld_resid + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 q = *reinterpret_cast<const float4*>(rs + j);
              v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (ep.out_h) {
          T* dst = reinterpret_cast<T*>(ep.out_h) + orow * ep.ldo_h + n0;
          if (ep.head_major_rows > 0) {
            const long long ob = orow / ep.head_major_rows, ot = orow - ob * ep.head_major_rows;
            dst = reinterpret_cast<T*>(ep.out_h) + ((ob * (ep.N >> 6) + (n0 >> 6)) * ep.head_major_rows + ot) * 64 + (n0 & 63);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q;
            q.x = DT<T>::pack2(v[j], v[j + 1]);
            q.y = DT<T>::pack2(v[j + 2], v[j + 3]);
            q.z = DT<T>::pack2(v[j + 4], v[j + 5]);
            q.w = DT<T>::pack2(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(dst + j) = q;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int tma_encode_map(s2s_ctx* ctx, CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base,
                   const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled);
  if (!fn) {
    s2s_set_error("cuTensorMapEncodeTiled unavailable");
    return S2S_ERR_CUDA;
  }
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    s2s_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu] strides [%llu,%llu]", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)strides_bytes[0],
                  (unsigned long long)(rank > 2 ? strides_bytes[1] : 0));
    return S2S_ERR_CUDA;
  }
  return S2S_OK;
}

namespace {

template <typename T, int BN>
int launch_impl(s2s_ctx* ctx, const GemmProblem& p, cudaStream_t stream) {
  using Cfg = TileCfg<BN>;
  CUtensorMap map_a, map_w;
  {
    cuuint64_t dims[3] = {(cuuint64_t)p.K, (cuuint64_t)p.M, (cuuint64_t)p.batch};
    cuuint64_t strides[2] = {(cuuint64_t)p.a_row_stride * 2, (cuuint64_t)p.a_batch_stride * 2};
    if (p.batch == 1) strides[1] = (cuuint64_t)p.a_row_stride * 2 * (cuuint64_t)p.M;  // unused but must be valid
    cuuint32_t box[3] = {BK, BM, 1};
    S2S_CHECK(tma_encode_map(ctx, &map_a, DT<T>::tma, 3, p.a, dims, strides, box));
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
    cuuint64_t strides[1] = {(cuuint64_t)p.ldw * 2};
    cuuint32_t box[2] = {BK, BN};
    S2S_CHECK(tma_encode_map(ctx, &map_w, DT<T>::tma, 2, p.w, dims, strides, box));
  }
  EpiParams ep;
  ep.bias = p.bias; ep.act = p.act; ep.out_h = p.out_h; ep.ldo_h = p.ldo_h; ep.out_f = p.out_f; ep.ldo_f = p.ldo_f;
  ep.resid = p.resid; ep.ld_resid = p.ld_resid; ep.resid_mode = p.resid_mode;
  ep.out_batch_rows = p.out_batch_rows; ep.out_row_offset = p.out_row_offset; ep.head_major_rows = p.head_major_rows;
  ep.M = p.M; ep.N = p.N; ep.K = p.K;

  static bool attr_set = false;  // per (T,BN) instantiation
  if (!attr_set) {
    S2S_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<T, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  dim3 grid(p.N / BN, (p.M + BM - 1) / BM, p.batch);
  gemm_tc_kernel<T, BN><<<grid, NTHREADS, Cfg::SMEM, stream>>>(map_a, map_w, ep);
  S2S_LAUNCH_CHECK();
  return S2S_OK;
}

}  // namespace

int gemm_tc_launch(s2s_ctx* ctx, const GemmProblem& p, int dtype, cudaStream_t stream) {
  S2S_REQUIRE(p.N % 64 == 0, "gemm: N=%d must be a multiple of 64", p.N);
  S2S_REQUIRE(p.K % 8 == 0 && p.K > 0, "gemm: K=%d must be a positive multiple of 8", p.K);
  S2S_REQUIRE(p.a_row_stride % 8 == 0 && p.ldw % 8 == 0, "gemm: row strides must be multiples of 8 elements");
  S2S_REQUIRE((reinterpret_cast<uintptr_t>(p.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  S2S_REQUIRE(p.M > 0 && p.batch > 0, "gemm: empty problem");
  // Pick the N tile so the grid covers the 148 SMs when the problem allows it.
  const long long tiles128 = (long long)(p.N / 128) * ((p.M + BM - 1) / BM) * p.batch;
  const bool use128 = (p.N % 128 == 0) && tiles128 >= ctx->num_sms;
  if (dtype == S2S_F16) {
    return use128 ? launch_impl<__half, 128>(ctx, p, stream) : launch_impl<__half, 64>(ctx, p, stream);
  } else if (dtype == S2S_BF16) {
    return use128 ? launch_impl<__nv_bfloat16, 128>(ctx, p, stream) : launch_impl<__nv_bfloat16, 64>(ctx, p, stream);
  }
  s2s_set_error("gemm: unsupported dtype %d", dtype);
  return S2S_ERR_UNSUPPORTED;
}
