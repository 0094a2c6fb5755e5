// decode_common.cuh -- building blocks of the persistent single-token decode kernels
// (Whisper decoder, Llama-family decoder).
//
// A decode step at small batch is a chain of tiny matrix-vector phases; per phase every SM touches only a few
// KB, so the phase time is (HBM/L2 round trips on its critical path) + (grid barrier), not bytes / bandwidth.
// Everything here is therefore written for memory-level parallelism: all loads of a phase are independent and
// issued back to back (explicit register staging, compile-time unrolling) so that each phase costs ONE round
// trip; work items are interleaved across CTAs so all 148 SMs pull from HBM; buffers produced by other CTAs
// inside the launch are read with ld.global.cg (L2) so a stale L1 line can never be observed; weights use the
// read-only no-L1-allocate path.  Reductions are warp shuffles; accumulation is fp32.
#pragma once
#include "common.cuh"

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;
constexpr int ATT_CHUNK = 64;   // keys per attention work item
constexpr int PART_PAD = 4;     // partial record = [o[hd], m, l, pad, pad] (16-byte aligned records)

// ---- grid barrier (all CTAs co-resident: cooperative launch) ------------------------------
// Monotonic counter: barrier #e completes when counter == e * gridDim.x.  Bounded spin -> trap.
// Split-phase: grid_arrive() publishes this CTA's results; everything that does not depend on other CTAs (argument
// setup, weight-ring issue, norm-weight loads for the NEXT phase) runs between arrive and wait, i.e. off the
// critical path; grid_wait() then blocks until every CTA has arrived.
__device__ __forceinline__ void grid_arrive(unsigned int* counter, unsigned int& epoch) {
  __syncthreads();  // CTA-scope: every thread's phase results happen-before thread 0's release below
  if (threadIdx.x == 0) {
    epoch += 1;
    // release-add at gpu scope (cumulative over the bar.sync above); no MEMBAR.SC, no L1 flush: all cross-CTA
    // data is read with ld.global.cg / .nc, never through L1.
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned int* counter, unsigned int epoch_thread0) {
  if (threadIdx.x == 0) {
    const unsigned int target = epoch_thread0 * gridDim.x;
    unsigned int v, spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 28)) __trap();
    } while (v < target);
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_sync(unsigned int* counter, unsigned int& epoch) {
  grid_arrive(counter, epoch);
  grid_wait(counter, epoch);
}

// Warp w of CTA c takes work items  w * gridDim.x + c, then + total_warps ...: consecutive items land on
// different SMs, so a phase with few items still spreads over the whole chip.
__device__ __forceinline__ int dec_first_item() { return (threadIdx.x >> 5) * gridDim.x + blockIdx.x; }
__device__ __forceinline__ int dec_item_stride() { return gridDim.x * DEC_WARPS; }

// Warm L2 with the weight rows this warp will stream in the NEXT phase (issued before the grid barrier, so the
// HBM fetch overlaps the barrier latency).  One 128-byte line per lane per instruction.
template <typename T, int R>
__device__ __forceinline__ void prefetch_rows_l2(const T* W, int N, int K) {
  const int lane = threadIdx.x & 31;
  const int row_bytes = K * (int)sizeof(T);
#pragma unroll 1
  for (int row0 = dec_first_item() * R; row0 < N; row0 += dec_item_stride() * R) {
    const int nrows = min(R, N - row0);
    const char* base = reinterpret_cast<const char*>(W + (long long)row0 * K);
    const int total = nrows * row_bytes;
#pragma unroll 1
    for (int o = lane * 128; o < total; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
  }
}

// Same idea for one attention work item: n_keys rows of `row_bytes` bytes at stride ld_bytes (K and V adjacent).
__device__ __forceinline__ void prefetch_strided_l2(const void* base, long long ld_bytes, int n_rows, int row_bytes) {
  const int lane = threadIdx.x & 31;
  const int lines_per_row = (row_bytes + 127) >> 7;
#pragma unroll 1
  for (int i = lane; i < n_rows * lines_per_row; i += 32) {
    const char* a = reinterpret_cast<const char*>(base) + (long long)(i / lines_per_row) * ld_bytes + (i % lines_per_row) * 128;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
  }
}

// ---- stage B rows of x (fp32 [B, d], produced by other CTAs) into shared memory, optionally normalised ----
// mode 0: plain copy; 1: LayerNorm (w, bias); 2: RMSNorm (w).  One-pass statistics (sum, sum of squares) from shared memory.
// The norm weights are fetched together with x (one round trip) into wb[2*d].  s_red: >= DEC_WARPS floats.
// Ends with a __syncthreads().
// Norm weights of a phase (w | bias -> wb[2*d]); callable ahead of time between grid_arrive and grid_wait.
static __device__ __forceinline__ void stage_norm_weights(const float* w, const float* bias, int d, float* wb) {
#pragma unroll 1
  for (int i = threadIdx.x * 4; i < d; i += DEC_THREADS * 4) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(w + i));
    float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bb = __ldg(reinterpret_cast<const float4*>(bias + i));
    *reinterpret_cast<float4*>(wb + i) = g;
    *reinterpret_cast<float4*>(wb + d + i) = bb;
  }
}

static __device__ __noinline__ void stage_rows(const float* x, int B, int d, float* xs, int mode, const float* w,
                                           const float* bias, float eps, float* s_red, float* wb, int wb_ready = 0) {
  const int n = B * d;
  // x rows: two independent 16-byte loads per thread per trip (one trip for B*d <= 2048)
#pragma unroll 1
  for (int i = threadIdx.x * 4; i < n; i += DEC_THREADS * 8) {
    const int j = i + DEC_THREADS * 4;
    const float4 v0 = __ldcg(reinterpret_cast<const float4*>(x + i));
    float4 v1;
    if (j < n) v1 = __ldcg(reinterpret_cast<const float4*>(x + j));
    *reinterpret_cast<float4*>(xs + i) = v0;
    if (j < n) *reinterpret_cast<float4*>(xs + j) = v1;
  }
  if (mode != 0 && !wb_ready) stage_norm_weights(w, mode == 1 ? bias : nullptr, d, wb);
  __syncthreads();
  if (mode == 0) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int wpr = DEC_WARPS;  // warps per row: largest power of two with wpr * B <= DEC_WARPS (min 1)
  while (wpr > 1 && wpr * B > DEC_WARPS) wpr >>= 1;
  const int rows_per_iter = DEC_WARPS / wpr;
#pragma unroll 1
  for (int r0 = 0; r0 < B; r0 += rows_per_iter) {
    const int row = r0 + warp / wpr, sub = warp % wpr, grp = (warp / wpr) * wpr;
    const bool valid = row < B;
    float* xr = xs + row * d;
    // one pass: sum and sum of squares together (one block reduction); var = E[x^2] - mean^2 in fp32 is accurate
    // to ~1e-6 * (1 + mean^2/var), far below the stated tolerances for residual-stream statistics
    float s1 = 0.f, s2 = 0.f;
    if (valid) {
#pragma unroll 2
      for (int i = sub * 32 + lane; i < d; i += wpr * 32) {
        const float v = xr[i];
        s1 += v;
        s2 = fmaf(v, v, s2);
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) { s_red[warp] = s1; s_red[DEC_WARPS + warp] = s2; }
    __syncthreads();
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < wpr; ++k) { t1 += s_red[grp + k]; t2 += s_red[DEC_WARPS + grp + k]; }
    const float mean = (mode == 1) ? t1 / (float)d : 0.f;
    const float var = fmaxf(t2 - (float)d * mean * mean, 0.f);
    const float rstd = rsqrtf(var / (float)d + eps);
    if (valid) {
#pragma unroll 2
      for (int i = sub * 32 + lane; i < d; i += wpr * 32) {
        float y = (xr[i] - mean) * rstd * wb[i];
        y += wb[d + i];
        xr[i] = y;
      }
    }
    __syncthreads();
  }
}

// ---- skinny GEMV: out[b][row] = sum_k W[row][k] * xs[b][k] ---------------------------------------
// ONE non-inlined routine serves every projection of every layer (runtime shapes and a runtime epilogue
// mode): the decode step executes each phase once per layer, so per-phase inlined copies (245 KB of SASS)
// turned the kernel into an instruction-cache streaming problem; sharing the routine keeps the hot loop resident.
// A warp owns GV_R consecutive rows and keeps GV_R * GV_U independent 16-byte loads in flight per lane; the
// bias / residual / mask reads are issued BEFORE the weight loads; the NEXT row group of the warp is L2-prefetched
// while the current one is reduced.  Lane b (< B) applies the epilogue for batch row b.
constexpr int GV_R = 2, GV_PF = 3;
enum GemvEpi { EPI_STORE = 0, EPI_GELU = 1, EPI_RESID = 2, EPI_QKV = 3, EPI_LOGITS = 4,
               EPI_QKV_ROPE = 5 /* row pairs */, EPI_SWIGLU = 6 /* row pairs */ };
struct GemvArgs {
  const void* W; int N, K;
  const float* bias;           // [N] or null
  int mode;
  float* out; int ldo;         // STORE / GELU / RESID (in-place residual stream) / QKV (q rows)
  // QKV: rows [d, 2d) -> k cache, [2d, 3d) -> v cache (16-bit), element (b, c) at kv0 + which*kv_which + b*kv_batch + c
  void* kv0; long long kv_which, kv_batch; int d;
  // LOGITS
  const unsigned char* suppress; int first_step; float* logits_out; long long logits_ld;  // logits_out + b*logits_ld + row
  // QKV_ROPE (Llama family): rows are stored so that rotate_half partners (j, j + hd/2) are adjacent (2j, 2j+1).
  // rows [0, q_rows) -> out (fp32 q, rotated); [q_rows, q_rows + k_rows) -> K cache (rotated); rest -> V cache.
  // cache element (b, c): kv0 + which*kv_which + slot[b]*kv_slot + pos[b]*kv_ld + c ; rope[pos][j] = (cos, sin)
  const int* pos; const int* slot; long long kv_slot; int kv_ld; const float2* rope; int hd, q_rows, k_rows;
  float q_scale;               // softmax scale folded into the stored q (head_dim^-0.5)
};

// epilogue of one reduced row for batch row `lane`
template <typename T>
__device__ __forceinline__ void gemv_epilogue(const GemvArgs& a, int mode, int row, int lane, float v, float bias,
                                              float resid, int sup, float& best_v, int& best_i) {
  v += bias;
  if (mode == EPI_STORE) {
    a.out[lane * a.ldo + row] = v;
  } else if (mode == EPI_GELU) {
    a.out[lane * a.ldo + row] = gelu_erf(v);
  } else if (mode == EPI_RESID) {
    a.out[lane * a.ldo + row] = resid + v;
  } else if (mode == EPI_QKV) {
    if (row < a.d) {
      a.out[lane * a.ldo + row] = v;
    } else {
      const int which = (row < 2 * a.d) ? 0 : 1;
      reinterpret_cast<T*>(a.kv0)[which * a.kv_which + lane * a.kv_batch + (row - (which + 1) * a.d)] = DT<T>::from_f(v);
    }
  } else {  // EPI_LOGITS
    if ((sup & 1) || (a.first_step && (sup & 2))) v = -INFINITY;
    if (a.logits_out) a.logits_out[lane * a.logits_ld + row] = v;
    if (v > best_v || (v == best_v && row < best_i)) { best_v = v; best_i = row; }
  }
}

// epilogue of one reduced ROW PAIR (rows row0, row0 + 1) for batch row `lane`
template <typename T>
__device__ __forceinline__ void gemv_pair_epilogue(const GemvArgs& a, int mode, int row0, int lane, float v0, float v1) {
  if (mode == EPI_SWIGLU) {
    a.out[lane * a.ldo + (row0 >> 1)] = (v0 / (1.0f + __expf(-v0))) * v1;
    return;
  }
  // EPI_QKV_ROPE
  const int kv_end = a.q_rows + a.k_rows;
  if (row0 < kv_end) {
    const int p = __ldcg(a.pos + lane);
    const float2 cs = a.rope[(long long)p * (a.hd >> 1) + ((row0 % a.hd) >> 1)];
    const float r0 = v0 * cs.x - v1 * cs.y, r1 = v1 * cs.x + v0 * cs.y;
    v0 = r0; v1 = r1;
  }
  if (row0 < a.q_rows) {
    *reinterpret_cast<float2*>(a.out + lane * a.ldo + row0) = make_float2(v0 * a.q_scale, v1 * a.q_scale);
  } else {
    const int which = (row0 < kv_end) ? 0 : 1;
    const int c = row0 - (which ? kv_end : a.q_rows);
    T* dst = reinterpret_cast<T*>(a.kv0) + which * a.kv_which + (long long)a.slot[lane] * a.kv_slot +
             (long long)__ldcg(a.pos + lane) * a.kv_ld + c;
    *reinterpret_cast<uint32_t*>(dst) = DT<T>::pack2(v0, v1);
  }
}

// ---- weight streaming through a per-warp shared-memory ring filled by the bulk-copy engine ---------------
// Each warp owns `ring_slots` slots of GV_R rows x GV_CH 16-bit weights.  Lane 0 issues cp.async.bulk
// (global -> shared, completion on the slot's mbarrier); the warp then multiplies from shared memory.  Bytes in
// flight are bounded by shared memory (128 KB per SM), not by registers, and the copies are not droppable hints.
// The warp is its own producer and consumer, so no cross-warp synchronisation is needed: slot reuse is ordered
// by program order + __syncwarp + fence.proxy.async.
constexpr int GV_CH = 1024;                       // weights per row chunk (2 KB); the code shifts by 10 for /GV_CH
constexpr int GV_SLOT_BYTES = GV_R * GV_CH * 2;   // one slot: GV_R row chunks

struct GemvRing {
  uint32_t base_s;       // this warp's ring: 32-bit shared-space address (16-byte aligned)
  uint32_t bars_s;       // this warp's mbarriers [slots], shared-space address
  int slots;
  unsigned int slot;     // next slot to consume and its phase parity; persist across phases (all lanes identical)
  unsigned int parity;
  // units of the NEXT gemv already issued into the ring by gemv_prefetch() (weights do not depend on the barrier)
  int pre_valid, pre_pg, pre_pc;
  const void* pre_W;
};

__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds16(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ float4 lds16f(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t ok, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 26)) __trap();
  } while (!ok);
}

// producer side of one unit = chunk c of the GV_R rows starting at row0 (lane 0 only)
template <typename T>
__device__ __forceinline__ void gemv_issue_unit(const T* __restrict__ W, int N, int K, int row0, int c, uint32_t dst,
                                                uint32_t bar) {
  const int nrows = min(GV_R, N - row0);
  const int len = min(GV_CH, K - c * GV_CH);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic-proxy reads of the slot are done
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(nrows * len * 2)) : "memory");
  const T* src = W + (long long)row0 * K + c * GV_CH;
  bulk_g2s(dst, src, (uint32_t)(len * 2), bar);
  if (nrows > 1) bulk_g2s(dst + GV_CH * 2, src + K, (uint32_t)(len * 2), bar);
}

// Issue the first `slots` units of a gemv into the (empty) ring.  Called right after the previous gemv finished,
// i.e. BEFORE the grid barrier and the input staging of the phase that will consume them: the weight stream of the
// next projection is already landing in shared memory while the chip synchronises.
template <typename T>
__device__ __forceinline__ void gemv_prefetch(const GemvArgs& a, GemvRing& ring) {
  const int lane = threadIdx.x & 31;
  const T* __restrict__ W = reinterpret_cast<const T*>(a.W);
  const int N = a.N, K = a.K;
  const int first = dec_first_item(), istride = dec_item_stride();
  ring.pre_valid = 1; ring.pre_pg = 0; ring.pre_pc = 0; ring.pre_W = a.W;
  if (first * GV_R >= N) return;
  int n_groups = 0;
  for (int gi = first; gi * GV_R < N; gi += istride) ++n_groups;
  const int cpr = (K + GV_CH - 1) >> 10;
  if (lane == 0) {
    int pg = 0, pc = 0, n_ahead = 0;
    unsigned int pslot = ring.slot;
#pragma unroll 1
    while (n_ahead < ring.slots && pg < n_groups) {
      gemv_issue_unit<T>(W, N, K, (first + pg * istride) * GV_R, pc, ring.base_s + pslot * GV_SLOT_BYTES, ring.bars_s + pslot * 8);
      if (++pc == cpr) { pc = 0; ++pg; }
      if (++pslot == (unsigned)ring.slots) pslot = 0;
      ++n_ahead;
    }
    ring.pre_pg = pg; ring.pre_pc = pc;
  }
}

// Wait for prefetched units that will never be consumed (early exit) so no bulk copy is in flight at CTA exit.
template <typename T>
__device__ __forceinline__ void gemv_drain(const GemvArgs& a, GemvRing& ring) {
  if (!ring.pre_valid) return;
  const int first = dec_first_item(), istride = dec_item_stride();
  ring.pre_valid = 0;
  if (first * GV_R >= a.N) return;
  const int n_groups = ((a.N + GV_R - 1) / GV_R - first + istride - 1) / istride;
  const int n_units = min(ring.slots, n_groups * ((a.K + GV_CH - 1) / GV_CH));
  unsigned int cslot = ring.slot, cpar = ring.parity;
  for (int u = 0; u < n_units; ++u) {
    mbar_wait_s(ring.bars_s + cslot * 8, cpar);
    if (++cslot == (unsigned)ring.slots) { cslot = 0; cpar ^= 1u; }
  }
  ring.slot = cslot; ring.parity = cpar;
}

// The unit stream of a warp: row groups first + j*istride (j = 0..n_groups-1), each split into cpr chunks.
// Producer cursor (pg, pc) runs `slots` units ahead of the consumer cursor (g, c); no divisions in the loop.
template <typename T, int NB>
__device__ __noinline__ void gemv_generic(const GemvArgs& a, uint32_t xs_s, int B, float& best_v, int& best_i,
                                          GemvRing& ring) {
  static_assert(GV_R == 2, "two rows per group");
  const int lane = threadIdx.x & 31;
  const T* __restrict__ W = reinterpret_cast<const T*>(a.W);
  const int N = a.N, K = a.K, mode = a.mode;
  const int first = dec_first_item(), istride = dec_item_stride();
  if (first * GV_R >= N) { ring.pre_valid = 0; return; }
  int n_groups = 0;  // row groups of this warp (counted, not divided: 1-3 for the small projections)
  for (int gi = first; gi * GV_R < N; gi += istride) ++n_groups;
  const int cpr = (K + GV_CH - 1) >> 10;
  const int slots = ring.slots;
  // ---- producer prologue (normally already done by gemv_prefetch before the barrier) ----
  if (ring.pre_valid && ring.pre_W != a.W) __trap();  // prefetch bookkeeping bug: the ring holds another matrix
  if (!ring.pre_valid) gemv_prefetch<T>(a, ring);
  int pg = ring.pre_pg, pc = ring.pre_pc;  // lane 0's cursor
  ring.pre_valid = 0;
  unsigned int cslot = ring.slot, cpar = ring.parity;
  const uint32_t lane_off = lane * 16;
#pragma unroll 1
  for (int g = 0; g < n_groups; ++g) {
    const int row0 = (first + g * istride) * GV_R;
    const int r1 = min(row0 + 1, N - 1);
    const float bias0 = a.bias ? __ldg(a.bias + row0) : 0.f;
    const float bias1 = a.bias ? __ldg(a.bias + r1) : 0.f;
    float resid0 = 0.f, resid1 = 0.f;
    int sup0 = 0, sup1 = 0;
    if (mode == EPI_RESID) {
      if (lane < B) { resid0 = __ldcg(a.out + lane * a.ldo + row0); resid1 = __ldcg(a.out + lane * a.ldo + r1); }
    } else if (mode == EPI_LOGITS && a.suppress) {
      sup0 = __ldg(a.suppress + row0); sup1 = __ldg(a.suppress + r1);
    }
    float acc0[NB], acc1[NB], acc2[NB], acc3[NB];  // rows 0/1 x even/odd element pairs: four independent chains
#pragma unroll
    for (int b = 0; b < NB; ++b) acc0[b] = acc1[b] = acc2[b] = acc3[b] = 0.f;
#pragma unroll 1
    for (int c = 0; c < cpr; ++c) {
      mbar_wait_s(ring.bars_s + cslot * 8, cpar);
      const uint32_t src = ring.base_s + cslot * GV_SLOT_BYTES + lane_off;
      const int len = min(GV_CH, K - c * GV_CH);
      const uint32_t xk = xs_s + (uint32_t)(c * GV_CH + lane * 8) * 4;
#pragma unroll 2
      for (int e = lane * 8; e < len; e += 256) {
        const uint32_t eo = (uint32_t)(e - lane * 8);
        const uint4 w0 = lds16(src + eo * 2), w1 = lds16(src + GV_CH * 2 + eo * 2);
        const float2 a0 = DT<T>::to_f2(w0.x), a1 = DT<T>::to_f2(w0.y), a2 = DT<T>::to_f2(w0.z), a3 = DT<T>::to_f2(w0.w);
        const float2 c0 = DT<T>::to_f2(w1.x), c1 = DT<T>::to_f2(w1.y), c2 = DT<T>::to_f2(w1.z), c3 = DT<T>::to_f2(w1.w);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          if (b < B) {
            const float4 x0 = lds16f(xk + (uint32_t)(b * K) * 4 + eo * 4), x1 = lds16f(xk + (uint32_t)(b * K) * 4 + eo * 4 + 16);
            acc0[b] = fmaf(a0.x, x0.x, acc0[b]); acc1[b] = fmaf(a0.y, x0.y, acc1[b]);
            acc0[b] = fmaf(a1.x, x0.z, acc0[b]); acc1[b] = fmaf(a1.y, x0.w, acc1[b]);
            acc0[b] = fmaf(a2.x, x1.x, acc0[b]); acc1[b] = fmaf(a2.y, x1.y, acc1[b]);
            acc0[b] = fmaf(a3.x, x1.z, acc0[b]); acc1[b] = fmaf(a3.y, x1.w, acc1[b]);
            acc2[b] = fmaf(c0.x, x0.x, acc2[b]); acc3[b] = fmaf(c0.y, x0.y, acc3[b]);
            acc2[b] = fmaf(c1.x, x0.z, acc2[b]); acc3[b] = fmaf(c1.y, x0.w, acc3[b]);
            acc2[b] = fmaf(c2.x, x1.x, acc2[b]); acc3[b] = fmaf(c2.y, x1.y, acc3[b]);
            acc2[b] = fmaf(c3.x, x1.z, acc2[b]); acc3[b] = fmaf(c3.y, x1.w, acc3[b]);
          }
        }
      }
      __syncwarp();
      if (lane == 0 && pg < n_groups) {  // refill the slot just drained
        gemv_issue_unit<T>(W, N, K, (first + pg * istride) * GV_R, pc, ring.base_s + cslot * GV_SLOT_BYTES, ring.bars_s + cslot * 8);
        if (++pc == cpr) { pc = 0; ++pg; }
      }
      if (++cslot == (unsigned)slots) { cslot = 0; cpar ^= 1u; }
    }
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float s0 = warp_sum(acc0[b] + acc1[b]), s1 = warp_sum(acc2[b] + acc3[b]);
      if (b == 0 || lane == b) { v0 = s0; v1 = s1; }
    }
    if (lane < B) {
      if (mode >= EPI_QKV_ROPE) {
        gemv_pair_epilogue<T>(a, mode, row0, lane, v0 + bias0, v1 + bias1);
      } else {
        gemv_epilogue<T>(a, mode, row0, lane, v0, bias0, resid0, sup0, best_v, best_i);
        if (row0 + 1 < N) gemv_epilogue<T>(a, mode, row0 + 1, lane, v1, bias1, resid1, sup1, best_v, best_i);
      }
    }
  }
  ring.slot = cslot;
  ring.parity = cpar;
}

// ---- attention over one chunk of <= 64 keys for one (batch, head): warp-level -----------------
// 8 lanes cover one key row (HD/8 values per lane), 4 keys per warp step.  Keys are processed in two halves of
// 32: the 8 K vectors and 8 V vectors of a half are loaded into registers before any arithmetic, so a half
// costs one memory round trip.  q must already carry the softmax scale.  Writes [o[HD], m, l] (unnormalised).
template <typename T, int HD>
__device__ __noinline__ void attend_chunk(const float* q /*global fp32 [HD]*/, const T* K, const T* V,
                                             long long ldk, long long ldv, int n_keys, float* part) {
  constexpr int PER = HD / 8;      // elements per lane
  constexpr int NV = PER / 8;      // 16-byte vectors per lane
  constexpr int KH = 8;            // keys per lane slot per half
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, j = lane & 7;
  float qr[PER];
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(q + j * PER + i));
    qr[i] = t.x; qr[i + 1] = t.y; qr[i + 2] = t.z; qr[i + 3] = t.w;
  }
  float m = -INFINITY, l = 0.f;
  float o[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) o[i] = 0.f;

#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    const int kbase = half * 32;
    if (kbase >= n_keys) break;
    uint4 kr[KH][NV], vr[KH][NV];
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const int key = kbase + g + 4 * i;
      if (key < n_keys) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          kr[i][v] = __ldcg(reinterpret_cast<const uint4*>(K + (long long)key * ldk + j * PER + v * 8));
          vr[i][v] = __ldcg(reinterpret_cast<const uint4*>(V + (long long)key * ldv + j * PER + v * 8));
        }
      }
    }
    float s[KH];
    float mh = -INFINITY;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const int key = kbase + g + 4 * i;
      float dot = 0.f;
      if (key < n_keys) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float2 a = DT<T>::to_f2(kr[i][v].x), b = DT<T>::to_f2(kr[i][v].y);
          const float2 c = DT<T>::to_f2(kr[i][v].z), d = DT<T>::to_f2(kr[i][v].w);
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
      mh = fmaxf(mh, s[i]);
    }
    mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, 8));
    mh = fmaxf(mh, __shfl_xor_sync(0xffffffffu, mh, 16));
    const float m_new = fmaxf(m, mh);          // finite: the half holds at least one key
    const float corr = __expf(m - m_new);      // m = -inf on the first half -> 0
    l *= corr;
#pragma unroll
    for (int i = 0; i < PER; ++i) o[i] *= corr;
    m = m_new;
#pragma unroll
    for (int i = 0; i < KH; ++i) {
      const int key = kbase + g + 4 * i;
      if (key < n_keys) {
        const float p = __expf(s[i] - m);
        l += p;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float2 a = DT<T>::to_f2(vr[i][v].x), b = DT<T>::to_f2(vr[i][v].y);
          const float2 c = DT<T>::to_f2(vr[i][v].z), d = DT<T>::to_f2(vr[i][v].w);
          o[v * 8 + 0] = fmaf(p, a.x, o[v * 8 + 0]); o[v * 8 + 1] = fmaf(p, a.y, o[v * 8 + 1]);
          o[v * 8 + 2] = fmaf(p, b.x, o[v * 8 + 2]); o[v * 8 + 3] = fmaf(p, b.y, o[v * 8 + 3]);
          o[v * 8 + 4] = fmaf(p, c.x, o[v * 8 + 4]); o[v * 8 + 5] = fmaf(p, c.y, o[v * 8 + 5]);
          o[v * 8 + 6] = fmaf(p, d.x, o[v * 8 + 6]); o[v * 8 + 7] = fmaf(p, d.y, o[v * 8 + 7]);
        }
      }
    }
  }
  // merge the 4 key slots (lanes differing in bits 3,4): l and o are per-slot partial sums under the common m
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 8);
    o[i] += __shfl_xor_sync(0xffffffffu, o[i], 16);
  }
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < PER; i += 4)
      *reinterpret_cast<float4*>(part + j * PER + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
  }
  if (lane == 0) { part[HD] = m; part[HD + 1] = l; }
}

// ---- merge attention partials into shared memory xs[b][h*HD + dd] (normalised) ---------------
// part layout: [B][H][s_max][HD + 4]; n_chunks valid records per (b, h).  One warp per (b, h): lane c owns record
// c's (m, l); the o vectors of up to CG records are fetched with independent loads issued together with the
// (m, l) load, so a pass over <= CG records costs a single L2 round trip.
template <int HD, int CG>
__device__ __noinline__ void combine_partials_to_smem(const float* part, int B, int H, int s_max, int n_chunks,
                                                         float* xs) {
  static_assert(CG <= 32, "one lane per record");
  constexpr int REC = HD + PART_PAD;
  constexpr int DPL = HD / 32;  // dims per lane
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
  for (int bh = warp; bh < B * H; bh += DEC_WARPS) {
    const float* pp = part + (long long)bh * s_max * REC;
    float M = -INFINITY, den = 0.f;
    float num[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) num[i] = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < n_chunks; c0 += CG) {
      const int cnt = min(CG, n_chunks - c0);
      float mc = -INFINITY, lc = 0.f;
      if (lane < cnt) {
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(pp + (long long)(c0 + lane) * REC + HD));
        mc = ml.x; lc = ml.y;
      }
      float ov[CG][DPL];
#pragma unroll
      for (int u = 0; u < CG; ++u)
#pragma unroll
        for (int i = 0; i < DPL; ++i)
          ov[u][i] = (u < cnt) ? __ldcg(pp + (long long)(c0 + u) * REC + lane + 32 * i) : 0.f;
      const float M_new = fmaxf(M, warp_max(mc));
      const float corr = __expf(M - M_new);
      const float wc = (lane < cnt) ? __expf(mc - M_new) : 0.f;
      den = den * corr + warp_sum(lc * wc);
#pragma unroll
      for (int i = 0; i < DPL; ++i) num[i] *= corr;
      M = M_new;
#pragma unroll
      for (int u = 0; u < CG; ++u) {
        const float wu = __shfl_sync(0xffffffffu, wc, u);
#pragma unroll
        for (int i = 0; i < DPL; ++i) num[i] = fmaf(ov[u][i], wu, num[i]);
      }
    }
    const float inv = 1.f / den;
    const int b = bh / H, h = bh % H;
#pragma unroll
    for (int i = 0; i < DPL; ++i) xs[b * (H * HD) + h * HD + lane + 32 * i] = num[i] * inv;
  }
}
