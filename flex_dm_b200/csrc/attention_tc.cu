// Masked multi-head self-attention core on the 5th-generation tensor cores (reference: architecture/transformer.py:60-76).
//
// One "unit" = one (document, head): Q, K, V tiles of [S <= 128 elements][32] fp32 cut straight out of the [T, 3D] QKV
// activation by 3-D TMA maps ([B][S][3D]: rows past the document end are out of bounds, so they load as zeros and are
// clipped on store -- no padding copies, no cross-document reads).  Per unit:
//   S = Q K^T          tcgen05.mma kind::tf32, both operands K-major in shared memory, accumulator [128 x 128] in TMEM
//   P = softmax(S/sqrt(dh)) over the document's valid keys: one thread per query row reads its row from TMEM (no
//       cross-thread reduction at all), writes the unnormalised probabilities (rounded to TF32) back IN PLACE
//   O = P V            tcgen05.mma with the A operand read from TMEM and V (MN-major) from shared memory; only the
//       ceil(n/8) key slabs that hold valid keys are issued
//   O / rowsum -> swizzled staging -> TMA store into the head's 32 columns of the [T, D] output; lse for the backward.
// Persistent CTAs loop over units; a 3-slot TMA ring, two score buffers and two output buffers in TMEM let the loads
// and S = QK^T of unit i+1 run under the softmax of unit i, and the epilogue of unit i-1 run while P V of unit i is in
// flight.  Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = softmax + epilogue.
#include "gemm.cuh"
#include "kernels.cuh"
#include "tc.cuh"

namespace mfp {

constexpr int kAtRows = 128;                    // queries / keys per unit tile
constexpr int kAtTileBytes = kAtRows * kDh * 4;  // 16 KB: [128][32] fp32, 128 B rows
constexpr int kAtThreads = 256;
constexpr float kAtScale = 0.17677669529663687f;  // 1/sqrt(32), transformer.py:62-63
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ------------------------------------------------------------------------------------------------- forward
constexpr int kFwdSlots = 3;
constexpr int kAtFwdThreads = 384;  // warps 0-3: TMA / MMA / TMEM allocator / idle; warps 4-7 and 8-11: two softmax + epilogue groups
struct AttnFwdSmem {
  static constexpr int kSlotBytes = 3 * kAtTileBytes;                 // Q, K, V
  static constexpr int kStageOff = kFwdSlots * kSlotBytes;            // 8 warps x 2 x [32][32] fp32 output staging
  static constexpr int kBarOff = kStageOff + 8 * 2 * 4096;
  static constexpr int kNumBars = 2 * kFwdSlots + 6;                  // full, empty, s_full[2], p_ready[2], o_full[2]
  static constexpr int kTotal = kBarOff + 8 * kNumBars + 16 + 1024;
};

__global__ void __launch_bounds__(kAtFwdThreads, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQK, const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                        const int* __restrict__ length, int B, int S, float* __restrict__ lse) {
  using L = AttnFwdSmem;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + L::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kFwdSlots + s); };
  auto sfull_bar = [&](int a) { return bar_base + 8u * (2 * kFwdSlots + a); };
  auto pready_bar = [&](int a) { return bar_base + 8u * (2 * kFwdSlots + 2 + a); };
  auto ofull_bar = [&](int a) { return bar_base + 8u * (2 * kFwdSlots + 4 + a); };
  const uint32_t tmem_slot = bar_base + 8u * L::kNumBars;
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(base_ptr + L::kBarOff + 8 * L::kNumBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = B * kH;
  const int n_local = (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kFwdSlots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(sfull_bar(a), 1); mbar_init(pready_bar(a), 4); mbar_init(ofull_bar(a), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tensormap(&tmQK);
    prefetch_tensormap(&tmV);
    prefetch_tensormap(&tmO);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();  // one resident CTA per SM for the whole kernel: the next kernel's CTAs only queue up behind it
  pdl_wait();               // everything above touched no global memory
  // TMEM columns: scores / probabilities [0,128) and [128,256); outputs [256,288) and [288,320)

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int u = blockIdx.x + i * gridDim.x, b = u / kH, h = u % kH;
        const int slot = i % kFwdSlots;
        mbar_wait(empty_bar(slot), ((i / kFwdSlots) & 1) ^ 1u);
        const uint32_t sq = base + slot * L::kSlotBytes;
        mbar_expect_tx(full_bar(slot), L::kSlotBytes);
        tma_load_3d(sq, &tmQK, h * kDh, 0, b, full_bar(slot));
        tma_load_3d(sq + kAtTileBytes, &tmQK, kD + h * kDh, 0, b, full_bar(slot));
        tma_load_3d(sq + 2 * kAtTileBytes, &tmV, 2 * kD + h * kDh, 0, b, full_bar(slot));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptors: c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13), a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kAtRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kDh >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      // shared-memory descriptors built once (this thread's serial instruction stream is on the critical path): only the address
      // field (bits [0,14), units of 16 B) changes with the slot, the operand tile and the k-step
      const uint64_t dq0 = make_smem_desc(base, 16, 1024, 2);                                   // Q (K-major); K is one tile further
      const uint64_t dv0 = make_smem_desc(base + 2 * kAtTileBytes, kAtTileBytes, 512, 1);       // V (MN-major)
      constexpr uint64_t kTileStep = (uint64_t)(kAtTileBytes >> 4), kSlotStep = (uint64_t)(L::kSlotBytes >> 4);
      for (int i = 0; i <= n_local; ++i) {
        if (i < n_local) {
          const int slot = i % kFwdSlots;
          mbar_wait(full_bar(slot), (i / kFwdSlots) & 1);
          tcgen05_fence_after();
          const uint64_t dq = dq0 + (uint64_t)slot * kSlotStep, dk = dq + kTileStep;
          const uint32_t sbuf = tmem_base + (uint32_t)((i & 1) * 128);
#pragma unroll
          for (int kk = 0; kk < kDh / 8; ++kk) umma_tf32(sbuf, dq + (uint64_t)(kk * 2), dk + (uint64_t)(kk * 2), idesc_s, kk > 0 ? 1u : 0u);
          tcgen05_commit(sfull_bar(i & 1));
        }
        if (i >= 1) {
          const int j = i - 1;
          const int u = blockIdx.x + j * gridDim.x, b = u / kH;
          const int n = min(S, __ldg(length + b) + 1);
          const int nk = (n + 7) >> 3;
          mbar_wait(pready_bar(j & 1), (j >> 1) & 1);
          tcgen05_fence_after();
          const uint64_t dv = dv0 + (uint64_t)(j % kFwdSlots) * kSlotStep;
          const uint32_t pbuf = tmem_base + (uint32_t)((j & 1) * 128);
          const uint32_t obuf = tmem_base + 256u + (uint32_t)((j & 1) * 32);
          for (int kk = 0; kk < nk; ++kk) umma_tf32_ts(obuf, pbuf + (uint32_t)(kk * 8), dv + (uint64_t)(kk * 64), idesc_o, kk > 0 ? 1u : 0u);
          tcgen05_commit(ofull_bar(j & 1));
          tcgen05_commit(empty_bar(j % kFwdSlots));
        }
      }
    }
  } else if (warp >= 4) {
    // Two softmax groups (warps 4-7 and 8-11) take the units alternately, one per score / output buffer in TMEM: a group is
    // one warp per scheduler and latency-bound on its TMEM round trips, so two of them in flight nearly double the rate.
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stage = base + L::kStageOff + (g * 4 + q) * 8192;
    uint8_t* stage_ptr = base_ptr + L::kStageOff + (g * 4 + q) * 8192;
    const uint32_t sw = (uint32_t)(lane & 7);
    const float c2 = kAtScale * kLog2e;
    const uint32_t sbuf = lane_base + (uint32_t)(g * 128);
    const uint32_t obuf = lane_base + 256u + (uint32_t)(g * 32);
    int ob = 0;
    for (int i = g; i < n_local; i += 2) {
      const int u = blockIdx.x + i * gridDim.x, b = u / kH, h = u % kH;
      const int n = min(S, __ldg(length + b) + 1);
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      mbar_wait(sfull_bar(g), ph);
      tcgen05_fence_after();
      const int nch = (n + 31) >> 5;
      // pass 1: row maximum over the valid keys (the load of chunk c+1 flies under the reduction of chunk c)
      float m = -INFINITY;
      {
        uint32_t ra[32], rb[32];
        auto reduce = [&](const int c, uint32_t (&r)[32]) {
          if (c * 32 + 32 <= n) {  // whole chunk valid (always, for full-length documents): no per-key predicate
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < n) m = fmaxf(m, __uint_as_float(r[j]));
          }
        };
        tmem_ld32_issue(sbuf, ra);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld32_wait(ra);
          if (c + 1 < nch) tmem_ld32_issue(sbuf + (uint32_t)((c + 1) * 32), rb);
          reduce(c, ra);
          if (c + 1 < nch) {
            tmem_ld32_wait(rb);
            if (c + 2 < nch) tmem_ld32_issue(sbuf + (uint32_t)((c + 2) * 32), ra);
            reduce(c + 1, rb);
          }
        }
      }
      // pass 2: unnormalised probabilities, rounded to TF32, written back in place
      float l = 0.f;
      const float mc = m * c2;
      {
        uint32_t ra[32], rb[32];
        auto expo = [&](const int c, uint32_t (&r)[32]) {
          if (c * 32 + 32 <= n) {
            float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;  // four independent sums: the adds do not chain behind the MUFU latency
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float p0 = fast_exp2(fmaf(__uint_as_float(r[j]), c2, -mc)), p1 = fast_exp2(fmaf(__uint_as_float(r[j + 1]), c2, -mc));
              const float p2 = fast_exp2(fmaf(__uint_as_float(r[j + 2]), c2, -mc)), p3 = fast_exp2(fmaf(__uint_as_float(r[j + 3]), c2, -mc));
              l0 += p0; l1 += p1; l2 += p2; l3 += p3;
              r[j] = to_tf32(p0); r[j + 1] = to_tf32(p1); r[j + 2] = to_tf32(p2); r[j + 3] = to_tf32(p3);
            }
            l += (l0 + l1) + (l2 + l3);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float p = (c * 32 + j < n) ? fast_exp2(fmaf(__uint_as_float(r[j]), c2, -mc)) : 0.f;
              l += p;
              r[j] = to_tf32(p);
            }
          }
          tmem_st32(sbuf + (uint32_t)(c * 32), r);
        };
        tmem_ld32_issue(sbuf, ra);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld32_wait(ra);
          if (c + 1 < nch) tmem_ld32_issue(sbuf + (uint32_t)((c + 1) * 32), rb);
          expo(c, ra);
          if (c + 1 < nch) {
            tmem_ld32_wait(rb);
            if (c + 2 < nch) tmem_ld32_issue(sbuf + (uint32_t)((c + 2) * 32), ra);
            expo(c + 1, rb);
          }
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pready_bar(g));
      const float inv_l = 1.0f / l;
      if (row < S) lse[((size_t)b * kH + h) * S + row] = m * kAtScale + __logf(l);
      // epilogue of the same unit: O / rowsum -> swizzled staging -> TMA store (the other group's softmax runs meanwhile)
      mbar_wait(ofull_bar(g), ph);
      tcgen05_fence_after();
      uint32_t r[32];
      tmem_ld32(obuf, r);
      if (lane == 0) tma_wait_group_read<1>();
      __syncwarp();
      uint8_t* outp = stage_ptr + ob * 4096 + lane * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(outp + ((k ^ sw) << 4)) =
            make_float4(__uint_as_float(r[4 * k]) * inv_l, __uint_as_float(r[4 * k + 1]) * inv_l, __uint_as_float(r[4 * k + 2]) * inv_l,
                        __uint_as_float(r[4 * k + 3]) * inv_l);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmO, stage + ob * 4096, h * kDh, q * 32, b);
        tma_commit_group();
      }
      ob ^= 1;
    }
    if (lane == 0) tma_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------- backward
// Works in the TRANSPOSED domain (keys on the 128 TMEM lanes) so that both probability-shaped operands of the
// gradient GEMMs are already where tcgen05 wants an A operand -- in tensor memory, one key per lane:
//   S^T  = K Q^T,   dP^T = V dO^T                (both operands K-major in shared memory)
//   P^T  = exp(S^T/sqrt(dh) - lse_i),  dS^T = P^T o (dP^T - D_i)   thread j owns key j; lse_i, D_i = dO_i.O_i from smem
//   dV   = P^T dO,  dK = dS^T Q                   (A = TMEM in place of S^T / dP^T, B = MN-major shared tiles)
//   dQ   = dS K                                   (A = dS^T written once to shared memory in the MN-major layout)
// K-major and MN-major tf32 operands need different 128-byte swizzles, so Q, K and dO are fetched twice by TMA (the
// second fetch hits L2).  Two independent single-slot rings: the K-major set is released as soon as S^T / dP^T are
// done, the MN-major set after the three gradient GEMMs, so the next unit's loads fly under this unit's math.
struct AttnBwdSmem {
  static constexpr int kKmajOff = 0;                       // K, Q, V, dO, O   (K-major, SWIZZLE_128B; O only feeds D_i = dO_i . O_i)
  static constexpr int kMnOff = 5 * kAtTileBytes;          // dO, Q, K      (MN-major, SWIZZLE_128B_ATOM_32B)
  static constexpr int kDsOff = 8 * kAtTileBytes;          // dS^T as the MN-major A operand of dQ: 4 chunks x [128 keys][32 q]
  static constexpr int kStageOff = kDsOff + 4 * kAtTileBytes;  // 4 warps x 2 x [32][32] fp32
  static constexpr int kVecOff = kStageOff + 4 * 2 * 4096;     // lse[128], D[128]
  static constexpr int kBarOff = kVecOff + 1024;
  static constexpr int kNumBars = 7;                       // kfull, kempty, mnfull, mnempty, s_full, p_ready, o_full
  static constexpr int kTotal = kBarOff + 8 * kNumBars + 16 + 1024;
};

constexpr int kAtBwdThreads = 384;  // warps 0-3: TMA / MMA / TMEM allocator / idle; warps 4-7 and 8-11: two compute groups
__global__ void __launch_bounds__(kAtBwdThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQkvK, const __grid_constant__ CUtensorMap tmQkvMN, const __grid_constant__ CUtensorMap tmDoK,
                        const __grid_constant__ CUtensorMap tmDoMN, const __grid_constant__ CUtensorMap tmDqkv, const __grid_constant__ CUtensorMap tmOutK,
                        const float* __restrict__ lse, const int* __restrict__ length, int B, int S, unsigned long long* __restrict__ trace) {
  using L = AttnBwdSmem;
  // FLEXDM_ATTN_TRACE: wait / work cycles per role, trace[blockIdx.x * 16 + k] (see the launcher for the slots)
  unsigned long long tw[6] = {0, 0, 0, 0, 0, 0};
  const long long t_begin = trace ? clock64() : 0;
  auto timed_wait = [&](uint32_t bar, uint32_t parity, int slot) {
    if (trace) { const long long t0 = clock64(); mbar_wait(bar, parity); tw[slot] += (unsigned long long)(clock64() - t0); }
    else mbar_wait(bar, parity);
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + L::kBarOff;
  const uint32_t kfull = bar_base, kempty = bar_base + 8, mnfull = bar_base + 16, mnempty = bar_base + 24, sfull = bar_base + 32,
                 pready = bar_base + 40, ofull = bar_base + 48;
  const uint32_t tmem_slot = bar_base + 8u * L::kNumBars;
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(base_ptr + L::kBarOff + 8 * L::kNumBars);
  float* lse_s = reinterpret_cast<float*>(base_ptr + L::kVecOff);
  float* d_s = lse_s + kAtRows;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int units = B * kH;
  const int n_local = (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    mbar_init(kfull, 1); mbar_init(kempty, 5); mbar_init(mnfull, 1); mbar_init(mnempty, 1);  // kempty: MMA commit + the 4 warps that read dO
    mbar_init(sfull, 1); mbar_init(pready, 8); mbar_init(ofull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    prefetch_tensormap(&tmQkvK); prefetch_tensormap(&tmQkvMN); prefetch_tensormap(&tmDoK); prefetch_tensormap(&tmDoMN); prefetch_tensormap(&tmDqkv); prefetch_tensormap(&tmOutK);
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();  // one resident CTA per SM for the whole kernel: the next kernel's CTAs only queue up behind it
  pdl_wait();               // everything above touched no global memory
  // TMEM columns: S^T / P^T [0,128), dP^T / dS^T [128,256), dV [256,288), dK [288,320), dQ [320,352)

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int u = blockIdx.x + i * gridDim.x, b = u / kH, h = u % kH;
        const uint32_t ph = (uint32_t)(i & 1);
        mbar_wait(kempty, ph ^ 1u);
        mbar_expect_tx(kfull, 5 * kAtTileBytes);
        const uint32_t sk = base + L::kKmajOff;
        tma_load_3d(sk, &tmQkvK, kD + h * kDh, 0, b, kfull);                          // K
        tma_load_3d(sk + kAtTileBytes, &tmQkvK, h * kDh, 0, b, kfull);                // Q
        tma_load_3d(sk + 2 * kAtTileBytes, &tmQkvK, 2 * kD + h * kDh, 0, b, kfull);   // V
        tma_load_3d(sk + 3 * kAtTileBytes, &tmDoK, h * kDh, 0, b, kfull);             // dO
        tma_load_3d(sk + 4 * kAtTileBytes, &tmOutK, h * kDh, 0, b, kfull);            // O
        mbar_wait(mnempty, ph ^ 1u);
        mbar_expect_tx(mnfull, 3 * kAtTileBytes);
        const uint32_t sm = base + L::kMnOff;
        tma_load_3d(sm, &tmDoMN, h * kDh, 0, b, mnfull);                              // dO
        tma_load_3d(sm + kAtTileBytes, &tmQkvMN, h * kDh, 0, b, mnfull);              // Q
        tma_load_3d(sm + 2 * kAtTileBytes, &tmQkvMN, kD + h * kDh, 0, b, mnfull);     // K
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kAtRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_g = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(kDh >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // B MN-major
      const uint32_t idesc_q = idesc_g | (1u << 15);                                                                                       // A MN-major too
      // descriptors built once; a k-step is one add on the address field (units of 16 B)
      const uint32_t sk = base + L::kKmajOff, sq = sk + kAtTileBytes, sv = sk + 2 * kAtTileBytes, sdo = sk + 3 * kAtTileBytes;
      const uint32_t mdo = base + L::kMnOff, mq = mdo + kAtTileBytes, mk = mdo + 2 * kAtTileBytes;
      const uint64_t dk_k = make_smem_desc(sk, 16, 1024, 2), dq_k = make_smem_desc(sq, 16, 1024, 2), dv_k = make_smem_desc(sv, 16, 1024, 2),
                     ddo_k = make_smem_desc(sdo, 16, 1024, 2);
      const uint64_t ddo_m = make_smem_desc(mdo, kAtTileBytes, 512, 1), dq_m = make_smem_desc(mq, kAtTileBytes, 512, 1),
                     dk_m = make_smem_desc(mk, kAtTileBytes, 512, 1), dds_m = make_smem_desc(base + L::kDsOff, kAtTileBytes, 512, 1);
      for (int i = 0; i < n_local; ++i) {
        const int u = blockIdx.x + i * gridDim.x, b = u / kH;
        const int n = min(S, __ldg(length + b) + 1);
        const int nk = (n + 7) >> 3;
        const uint32_t ph = (uint32_t)(i & 1);
        timed_wait(kfull, ph, 0);
        tcgen05_fence_after();
#pragma unroll
        for (int kk = 0; kk < kDh / 8; ++kk) umma_tf32(tmem_base, dk_k + (uint64_t)(kk * 2), dq_k + (uint64_t)(kk * 2), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < kDh / 8; ++kk) umma_tf32(tmem_base + 128u, dv_k + (uint64_t)(kk * 2), ddo_k + (uint64_t)(kk * 2), idesc_s, kk > 0 ? 1u : 0u);
        tcgen05_commit(sfull);
        tcgen05_commit(kempty);
        timed_wait(pready, ph, 1);
        timed_wait(mnfull, ph, 2);
        tcgen05_fence_after();
        for (int kk = 0; kk < nk; ++kk)  // dV = P^T dO
          umma_tf32_ts(tmem_base + 256u, tmem_base + (uint32_t)(kk * 8), ddo_m + (uint64_t)(kk * 64), idesc_g, kk > 0 ? 1u : 0u);
        for (int kk = 0; kk < nk; ++kk)  // dK = dS^T Q
          umma_tf32_ts(tmem_base + 288u, tmem_base + 128u + (uint32_t)(kk * 8), dq_m + (uint64_t)(kk * 64), idesc_g, kk > 0 ? 1u : 0u);
        for (int kk = 0; kk < nk; ++kk)  // dQ = dS K
          umma_tf32(tmem_base + 320u, dds_m + (uint64_t)(kk * 64), dk_m + (uint64_t)(kk * 64), idesc_q, kk > 0 ? 1u : 0u);
        tcgen05_commit(ofull);
        tcgen05_commit(mnempty);
      }
      if (trace) { trace[blockIdx.x * 16 + 0] = tw[0]; trace[blockIdx.x * 16 + 1] = tw[1]; trace[blockIdx.x * 16 + 2] = tw[2]; }
    }
  } else if (warp >= 4) {
    // Two compute groups of four warps (one warp per TMEM lane quarter each).  Both own all 128 keys (lanes); group g takes the
    // query chunks 2g, 2g+1 of P^T / dS^T.  Afterwards group 0 stores dV and dK while group 1 prepares lse / D of the next unit
    // and stores dQ.  Named barriers 1 and 2 (256 threads) order the writes and reads of lse_s / d_s between the groups.
    const int g = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;  // key index in softmax / dS; query index in the D preparation; output row in the epilogue
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t stage = base + L::kStageOff + (g * 4 + q) * 4096;  // one [32][32] fp32 staging chunk per warp
    uint8_t* stage_ptr = base_ptr + L::kStageOff + (g * 4 + q) * 4096;
    uint8_t* ds_ptr = base_ptr + L::kDsOff;
    const uint8_t* do_ptr = base_ptr + L::kKmajOff + 3 * kAtTileBytes;
    const uint8_t* o_ptr = base_ptr + L::kKmajOff + 4 * kAtTileBytes;
    const uint32_t sw = (uint32_t)(lane & 7);
    const float c2 = kAtScale * kLog2e;

    // lse of a unit for this thread's query row: fetched one unit ahead, so that its latency is off the path
    auto load_lse = [&](int i) {
      const int u = blockIdx.x + i * gridDim.x, b = u / kH, h = u % kH;
      return row < S ? __ldg(lse + ((size_t)b * kH + h) * S + row) : INFINITY;
    };
    float lse_next = (g == 1 && n_local > 0) ? load_lse(0) : 0.f;
    // lse_i and D_i = dO_i . O_i of unit i -> shared memory (group 1: one query row per thread).  O comes with the K-major tiles by TMA
    // (plain fp32 map): per-thread global loads of the O rows put a DRAM round trip (~3 800 cycles per unit, FLEXDM_ATTN_TRACE) on the
    // path between two units' compute phases -- longer than the gradient MMAs it was meant to hide under.
    auto prepare = [&](int i) {
      mbar_wait(kfull, (uint32_t)(i & 1));
      float dsum = 0.f;
      const float l = lse_next;
      if (row < S) {  // rows past the document are zero-filled by the TMA unit
        const uint8_t* op = o_ptr + row * 128;
        const uint8_t* dp = do_ptr + row * 128;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 o = *reinterpret_cast<const float4*>(op + ((k ^ sw) << 4));
          const float4 gv = *reinterpret_cast<const float4*>(dp + ((k ^ sw) << 4));
          dsum += o.x * gv.x + o.y * gv.y + o.z * gv.z + o.w * gv.w;
        }
      }
      if (i + 1 < n_local) lse_next = load_lse(i + 1);
      lse_s[row] = l * kLog2e;
      d_s[row] = dsum;
      __syncwarp();
      if (lane == 0) mbar_arrive(kempty);  // this warp no longer reads the K-major dO tile
    };
    auto store_tile = [&](int i, int t) {  // t = 0 dV, 1 dK, 2 dQ: TMEM -> registers -> swizzled staging -> TMA store
      const int u = blockIdx.x + i * gridDim.x, b = u / kH, h = u % kH;
      uint32_t r[32];
      tmem_ld32(lane_base + 256u + (uint32_t)(t * 32), r);
      const float sc = (t == 0) ? 1.0f : kAtScale;
      if (lane == 0) tma_wait_group_read<0>();  // the staging chunk has been read out by the previous store
      __syncwarp();
      uint8_t* outp = stage_ptr + lane * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        *reinterpret_cast<float4*>(outp + ((k ^ sw) << 4)) = make_float4(__uint_as_float(r[4 * k]) * sc, __uint_as_float(r[4 * k + 1]) * sc,
                                                                         __uint_as_float(r[4 * k + 2]) * sc, __uint_as_float(r[4 * k + 3]) * sc);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&tmDqkv, stage, (2 - t) * kD + h * kDh, q * 32, b);
        tma_commit_group();
      }
    };

    if (n_local > 0 && g == 1) prepare(0);
    for (int i = 0; i < n_local; ++i) {
      const int u = blockIdx.x + i * gridDim.x, b = u / kH;
      const int n = min(S, __ldg(length + b) + 1);
      const bool key_valid = row < n;
      long long tc0 = trace ? clock64() : 0;
      asm volatile("bar.sync 1, 256;" ::: "memory");  // lse_s / d_s of this unit are in place
      if (trace) { const long long t1 = clock64(); tw[3] += (unsigned long long)(t1 - tc0); tc0 = t1; }
      timed_wait(sfull, (uint32_t)(i & 1), 0);
      tcgen05_fence_after();
      if (trace) tc0 = clock64();
#pragma unroll 1
      for (int c = 2 * g; c < 2 * g + 2; ++c) {  // 32 queries at a time
        uint32_t rs[32], rd[32];
        tmem_ld32_issue(lane_base + (uint32_t)(c * 32), rs);  // both loads in flight before the single wait
        tmem_ld32_issue(lane_base + 128u + (uint32_t)(c * 32), rd);
        tmem_ld32_wait(rs);
        tmem_ld32_wait(rd);
        uint8_t* dsp = ds_ptr + c * kAtTileBytes + row * 128;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int qi = c * 32 + j;
          const float p = key_valid ? fast_exp2(fmaf(__uint_as_float(rs[j]), c2, -lse_s[qi])) : 0.f;
          const float ds = p * (__uint_as_float(rd[j]) - d_s[qi]);
          rs[j] = to_tf32(p);
          rd[j] = to_tf32(ds);
        }
        tmem_st32(lane_base + (uint32_t)(c * 32), rs);
        tmem_st32(lane_base + 128u + (uint32_t)(c * 32), rd);
#pragma unroll
        for (int f = 0; f < 8; ++f)
          *reinterpret_cast<uint4*>(dsp + ((((f >> 1) ^ (row & 3)) << 5) | ((f & 1) << 4))) = make_uint4(rd[4 * f], rd[4 * f + 1], rd[4 * f + 2], rd[4 * f + 3]);
      }
      tmem_st_wait();
      fence_proxy_async_smem();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pready);
      if (trace) { const long long t1 = clock64(); tw[1] += (unsigned long long)(t1 - tc0); tc0 = t1; }  // P / dS computation
      asm volatile("bar.sync 2, 256;" ::: "memory");  // everyone is done with lse_s / d_s before the next unit overwrites them
      if (g == 1 && i + 1 < n_local) prepare(i + 1);
      if (trace) { const long long t1 = clock64(); tw[4] += (unsigned long long)(t1 - tc0); tc0 = t1; }  // barrier 2 + (group 1) prepare
      timed_wait(ofull, (uint32_t)(i & 1), 2);
      tcgen05_fence_after();
      if (trace) tc0 = clock64();
      if (g == 0) { store_tile(i, 0); store_tile(i, 1); }
      else store_tile(i, 2);
      if (trace) tw[5] += (unsigned long long)(clock64() - tc0);  // stores
    }
    if (trace && lane == 0 && q == 0) {
      unsigned long long* t = trace + blockIdx.x * 16 + 3 + g * 6;
      t[0] = tw[0]; t[1] = tw[1]; t[2] = tw[2]; t[3] = tw[3]; t[4] = tw[4]; t[5] = tw[5];
      if (g == 0) trace[blockIdx.x * 16 + 15] = (unsigned long long)(clock64() - t_begin);
    }
    if (lane == 0) tma_wait_group_read<0>();
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

int launch_attention_fwd_tc(TensorMapCache* maps, const float* qkv, const int* length, int B, int S, float* out, float* lse, cudaStream_t st) {
  tensor_map_cache_trim(maps);
  if (S > kAtRows) { set_error("attention (tcgen05): S = %d exceeds the %d-row unit tile", S, kAtRows); return MFP_ERR_UNSUPPORTED; }
  static bool attr_set = false;
  if (!attr_set) {
    MFP_CUDA_OK(cudaFuncSetAttribute(attention_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnFwdSmem::kTotal));
    attr_set = true;
  }
  const uint64_t qdims[3] = {(uint64_t)(3 * kD), (uint64_t)S, (uint64_t)B}, qstr[2] = {(uint64_t)(3 * kD), (uint64_t)S * 3 * kD};
  const uint32_t qbox[3] = {(uint32_t)kDh, (uint32_t)kAtRows, 1};
  const uint64_t odims[3] = {(uint64_t)kD, (uint64_t)S, (uint64_t)B}, ostr[2] = {(uint64_t)kD, (uint64_t)S * kD};
  const uint32_t obox[3] = {(uint32_t)kDh, 32, 1};
  const CUtensorMap* mqk = tensor_map_get(maps, qkv, 3, qdims, qstr, qbox, kMapOperandK);
  const CUtensorMap* mv = tensor_map_get(maps, qkv, 3, qdims, qstr, qbox, kMapOperandMN);
  const CUtensorMap* mo = tensor_map_get(maps, out, 3, odims, ostr, obox, kMapEpilogue);
  if (!mqk || !mv || !mo) return MFP_ERR_CUDA;
  const int units = B * kH;
  const int grid = units < sm_count() ? units : sm_count();
  MFP_CUDA_OK(launch_pdl(attention_fwd_tc_kernel, grid, kAtFwdThreads, AttnFwdSmem::kTotal, st, *mqk, *mv, *mo, length, B, S, lse));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_attention_bwd_tc(TensorMapCache* maps, const float* qkv, const float* out, const float* lse, const float* dout, const int* length, int B, int S,
                            float* dqkv, cudaStream_t st) {
  tensor_map_cache_trim(maps);
  if (S > kAtRows) { set_error("attention backward (tcgen05): S = %d exceeds the %d-row unit tile", S, kAtRows); return MFP_ERR_UNSUPPORTED; }
  static bool attr_set = false;
  if (!attr_set) {
    MFP_CUDA_OK(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdSmem::kTotal));
    attr_set = true;
  }
  const uint64_t qdims[3] = {(uint64_t)(3 * kD), (uint64_t)S, (uint64_t)B}, qstr[2] = {(uint64_t)(3 * kD), (uint64_t)S * 3 * kD};
  const uint64_t odims[3] = {(uint64_t)kD, (uint64_t)S, (uint64_t)B}, ostr[2] = {(uint64_t)kD, (uint64_t)S * kD};
  const uint32_t tbox[3] = {(uint32_t)kDh, (uint32_t)kAtRows, 1};
  const uint32_t sbox[3] = {(uint32_t)kDh, 32, 1};
  const CUtensorMap* mqk = tensor_map_get(maps, qkv, 3, qdims, qstr, tbox, kMapOperandK);
  const CUtensorMap* mqm = tensor_map_get(maps, qkv, 3, qdims, qstr, tbox, kMapOperandMN);
  const CUtensorMap* mdk = tensor_map_get(maps, dout, 3, odims, ostr, tbox, kMapOperandK);
  const CUtensorMap* mdm = tensor_map_get(maps, dout, 3, odims, ostr, tbox, kMapOperandMN);
  const CUtensorMap* mg = tensor_map_get(maps, dqkv, 3, qdims, qstr, sbox, kMapEpilogue);
  const CUtensorMap* mo = tensor_map_get(maps, out, 3, odims, ostr, tbox, kMapEpilogue);  // O rows as plain fp32 (not rounded to TF32), same swizzle as the K-major tiles
  if (!mqk || !mqm || !mdk || !mdm || !mg || !mo) return MFP_ERR_CUDA;
  const int units = B * kH;
  const int grid = units < sm_count() ? units : sm_count();
  static const bool trace_on = getenv("FLEXDM_ATTN_TRACE") != nullptr;  // debugging aid: per-role wait / work cycles of every launch
  static unsigned long long* trace = nullptr;
  if (trace_on && !trace) MFP_CUDA_OK(cudaMalloc(&trace, 148 * 16 * sizeof(unsigned long long)));
  if (trace_on) MFP_CUDA_OK(cudaMemsetAsync(trace, 0, 148 * 16 * sizeof(unsigned long long), st));
  MFP_CUDA_OK(launch_pdl(attention_bwd_tc_kernel, grid, kAtBwdThreads, AttnBwdSmem::kTotal, st, *mqk, *mqm, *mdk, *mdm, *mg, *mo, lse, length, B, S,
                         trace_on ? trace : nullptr));
  MFP_CUDA_OK(cudaGetLastError());
  if (trace_on) {
    unsigned long long hb[148 * 16];
    MFP_CUDA_OK(cudaStreamSynchronize(st));
    MFP_CUDA_OK(cudaMemcpy(hb, trace, sizeof(hb), cudaMemcpyDeviceToHost));
    double a[16] = {};
    for (int b = 0; b < grid && b < 148; ++b)
      for (int k = 0; k < 16; ++k) a[k] += (double)hb[b * 16 + k] / grid;
    fprintf(stderr, "attention bwd trace units/cta=%.2f | mma: wait kfull %.0f pready %.0f mnfull %.0f | group0: sfull %.0f compute %.0f ofull %.0f bar1 %.0f bar2 %.0f store %.0f | "
            "group1: sfull %.0f compute %.0f ofull %.0f bar1 %.0f bar2+prepare %.0f store %.0f | total %.0f cycles\n", (double)units / grid, a[0], a[1], a[2], a[3], a[4],
            a[5], a[6], a[7], a[8], a[9], a[10], a[11], a[12], a[13], a[14], a[15]);
  }
  return MFP_OK;
}

}  // namespace mfp

