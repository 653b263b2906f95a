// TF32 tcgen05/TMA GEMM + a SIMT bring-up kernel with the same epilogue.  See gemm.cuh.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>

#include "gemm.cuh"
#include "tc.cuh"

namespace mfp {

// ------------------------------------------------------------------------------------------------- epilogue
__device__ __forceinline__ void epilogue_store4(float (&v)[4], int row, int col, int N, const GemmEpilogue& ep, bool lead_split) {
  // col is a multiple of 4; N is a multiple of 4 (checked on the host).  bias/residual are added by split 0 only.
  if (col >= N) return;
  if (ep.bias && lead_split) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (ep.relu) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (ep.relu_src) {
    const float4 s = *reinterpret_cast<const float4*>(ep.relu_src + (size_t)row * ep.ld_relu + col);
    v[0] = s.x > 0.0f ? v[0] : 0.0f; v[1] = s.y > 0.0f ? v[1] : 0.0f;
    v[2] = s.z > 0.0f ? v[2] : 0.0f; v[3] = s.w > 0.0f ? v[3] : 0.0f;
  }
  if (ep.relu_bits) {
    const uint32_t nib = ep.relu_bits[(size_t)(col >> 5) * ep.relu_bits_ld + row] >> (col & 31);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = ((nib >> j) & 1u) ? v[j] : 0.0f;
  }
  if (ep.drop_enabled) dropout4(v, ((uint32_t)row + ep.drop_row0) * (uint32_t)N + (uint32_t)col, ep.drop_rate, ep.drop_seed, ep.drop_step, ep.drop_site);
  if (ep.rowflag && ep.rowflag[row]) { v[0] = v[1] = v[2] = v[3] = 0.0f; }
  if (ep.residual && lead_split) {
    const float4 r = *reinterpret_cast<const float4*>(ep.residual + (size_t)row * ep.ldr + col);
    v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
  }
  float* dst = ep.out + (size_t)row * ep.ldo + col;
  if (ep.atomic) {
    atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
  } else {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ------------------------------------------------------------------------------------------------- tcgen05 kernel
// Persistent, warp-specialised: one CTA per SM loops over output tiles (n fastest, so CTAs running together share
// A tiles through L2).  Roles (8 warps):
//   warp 0      TMA producer of the A half of every stage of the kStages-deep shared-memory ring (mbarrier full/empty)
//   warp 3      TMA producer of the B half (one issuing thread gets a box out only every ~600 cycles: two issuers, one box each)
//   warp 1      MMA issuer: one thread issues tcgen05.mma kind::tf32 into one of two TMEM accumulators
//   warp 2      TMEM allocator; optional column-sum role (bias gradients of wgrad GEMMs: sums the MN-major B tiles while they
//               sit in shared memory, so dY is never re-read from HBM)
//   warps 4-11  epilogue, two warps per TMEM lane quarter: TMEM -> registers -> fused ops -> swizzled shared staging -> TMA store (or
//               TMA reduce-add for split-K); the residual / ReLU-mask operand arrives by TMA as well, one 32x32 chunk ahead.
// The epilogue of tile i overlaps the main loop of tile i+1 through the double-buffered accumulator.
constexpr int kBM = 128;        // UMMA M (one TMEM lane per row)
constexpr int kBK = 32;         // 32 fp32 = 128 B = one swizzle row
constexpr int kUmmaK = 8;       // tf32: 32 B of K per instruction
constexpr int kGemmThreads = 384;   // 4 role warps (TMA A, MMA, TMEM / column sums, TMA B) + 8 epilogue warps
constexpr int kEpiWarp0 = 4;    // first epilogue warp (warp & 3 = TMEM lane quarter)
constexpr int kChunkBytes = 32 * 32 * 4;  // one 32-row x 32-column fp32 staging chunk

struct GemmTune {
  uint32_t mn_lbo, mn_sbo, k_lbo, k_sbo;
  uint32_t k_layout;  // descriptor layout type of K-major operands: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (experiment)
};

struct GemmTiles {
  int tiles_m, tiles_n, splits, kb_per_split;
  int passes;     // 1, or 3 for the compensated 3xTF32 mode (see the kernel)
  int slab_rows;  // > 0 (deterministic split-K): split s stores its tiles, without reduction, at row offset s * slab_rows of the scratch output
  int det_colsum; // column-sum partials are stored per (split, m-tile) slot instead of atomically added
};

// AUX: the epilogue stages a residual / ReLU-mask operand (one more 4 KB chunk per epilogue warp).  Without it the 32 KB saved
// buy a fourth operand stage at BN = 256: the ring is latency-bound (a slot is refilled only after its MMAs retire), so the
// bytes in flight set the fill rate.
// CG = 2: a CTA PAIR works on a 256 x BN tile (tcgen05.mma.cta_group::2): each CTA stages its own 128 rows of A and only its half of
// the B tile (BN / 2 rows), so a k-block costs 32 KB of L2 -> shared traffic per SM instead of 48 KB, and the same shared memory holds
// six stages instead of four.
template <int BN, int CG>
struct GemmSmem {
  static constexpr int kStages = (BN <= 128 || CG == 2) ? 6 : 4;
  static constexpr int kEpiChunks = 1;                           // per epilogue warp: one 4 KB staging chunk (transposes aux in, results out)
  static constexpr int kABytes = kBM * kBK * 4;
  static constexpr int kBRows = BN / CG;                          // rows of the B tile this CTA stages
  static constexpr int kBBytes = kBRows * kBK * 4;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiOff = kStages * kStageBytes;          // 8 warps x {out, aux} chunks
  static constexpr int kBarOff = kEpiOff + 8 * kEpiChunks * kChunkBytes;
  static constexpr int kNumBars = 2 * kStages + 4;               // full, empty, tmem full[2]/empty[2]
  static constexpr int kTotal = kBarOff + 8 * kNumBars + 16 + 1024;  // + TMEM slot + alignment slack
};

// Epilogue variants are compiled in (EPI bit mask), so each instantiation carries only the code it runs: the epilogue
// warps are issue-bound (one warp per scheduler), every dead branch in their loop costs throughput.
enum : int { kEpiBias = 1, kEpiRelu = 2, kEpiResidual = 4, kEpiReluMask = 8, kEpiDropout = 16, kEpiRowflag = 32, kEpiLayerNorm = 64,
             kEpiReluBits = 128,   // ReLU gates read as bits (one word per row and 32-column chunk) instead of the aux operand
             kEpiBitsOut = 256 };  // ... and written by the forward GEMM that applies the ReLU

// One GEMM of a launch.  A launch carries up to kMaxGroup INDEPENDENT problems whose tiles form one tile space (problem 0's tiles first):
// a weight gradient and the input gradient that hangs off the same dY run as one launch -- every CTA takes its one long split-K tile of
// the weight gradient and then its share of the input-gradient tiles, so the reduce-add epilogue of the first overlaps the main loop of
// the second through the double-buffered accumulator, and one launch boundary (drain, prologue, ramp) disappears.
constexpr int kMaxGroup = 3;
struct GemmProblem {
  CUtensorMap tmA, tmB, tmOut, tmALo, tmBLo;
  int M, N, K, a_mn, b_mn;
  GemmTiles tl;
  GemmEpilogue ep;
  float* colsum;
  int tile0;     // first index of this problem's tiles in the launch's tile space
  int aux_kind;  // 0, or the kernel's aux_mode (1 residual, 2 ReLU mask) when this problem uses the aux operand
  int a_hint, b_hint;  // L2 eviction priority of the operand loads (tc.cuh l2_policy): 1 = another tile of this launch reads it again, 2 = last use
};
struct GemmGroup {
  int n, total_tiles, any_colsum, pad;
  GemmProblem p[kMaxGroup];
};

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32_tcgen05(const __grid_constant__ GemmGroup grp, GemmTune tune, unsigned long long* __restrict__ trace) {
  constexpr int aux_mode = (EPI & kEpiResidual) ? 1 : ((EPI & kEpiReluMask) ? 2 : 0);
  using L = GemmSmem<BN, CG>;
  constexpr int kTileM = kBM * CG;  // rows of a tile of the launch's tile space (one CTA, or a CTA pair)
  // CTA pair (CG == 2, launched as clusters of two): rank 0 is the leader -- it alone issues the MMAs, and the operand-ring "full"
  // barriers and the accumulator "empty" barriers that its MMA thread waits on live in ITS shared memory: both CTAs' TMA loads complete
  // on the leader's full barrier, the peer's epilogue warps arrive remotely on the leader's tempty barrier; tcgen05.commit multicasts
  // the "slot free" and "accumulator ready" arrivals to both CTAs.
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int unit = (int)blockIdx.x / CG, units = (int)gridDim.x / CG;
  constexpr int kStages = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_base = base + L::kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * L::kNumBars;
  const uint32_t* tmem_slot_ptr = reinterpret_cast<const uint32_t*>(base_ptr + L::kBarOff + 8 * L::kNumBars);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = grp.total_tiles;
  const bool any_colsum = grp.any_colsum != 0;
  // which problem a tile of the launch's tile space belongs to
  auto problem_of = [&](int tile) {
    int pi = 0;
    if (grp.n > 1 && tile >= grp.p[1].tile0) pi = 1;
    if (grp.n > 2 && tile >= grp.p[2].tile0) pi = 2;
    return pi;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      // the A producer and the B producer each arrive with their own byte count.  Pair: the leader's barrier collects both CTAs' loads;
      // the peer's own full barrier only serves its column-sum warp, which the leader's column-sum warp notifies once the stage has landed
      mbar_init(full_bar(s), rank == 0 ? 2 : 1);
      mbar_init(empty_bar(s), any_colsum ? 2 : 1);  // MMA commit (+ the column-sum warp, which then attends every tile of the launch)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8 * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < grp.n; ++i) {
      prefetch_tensormap(&grp.p[i].tmA);
      prefetch_tensormap(&grp.p[i].tmB);
      prefetch_tensormap(&grp.p[i].tmOut);
      if (grp.p[i].tl.passes == 3) { prefetch_tensormap(&grp.p[i].tmALo); prefetch_tensormap(&grp.p[i].tmBLo); }
    }
  }
  if (warp == 2) {
    if constexpr (CG == 2) {  // both CTAs of the pair, same warp
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();  // the peer's barriers exist before anything arrives on them remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();  // one resident CTA per SM for the whole kernel: the next kernel's CTAs only queue up behind it
  pdl_wait();               // everything above (barriers, TMEM, tensor-map prefetch) touched no global memory
  // FLEXDM_GEMM_TRACE: per-role wait cycles (which stage of the pipeline starves which), written to trace[blockIdx.x * 8 + k]
  const long long t_start = trace ? clock64() : 0;
  unsigned long long w0 = 0, w1 = 0, w2 = 0;
  auto wait_t = [&](uint32_t bar, uint32_t parity, unsigned long long& acc) {
    if (trace) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += (unsigned long long)(clock64() - t0);
    } else {
      mbar_wait(bar, parity);
    }
  };
  // 3xTF32 (tl.passes == 3): the K loop runs three times over the operands -- (A, B), (A_lo, B), (A, B_lo), x_lo = x - tf32(x) -- into the
  // same accumulator: a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32-accurate products on the TF32 tensor cores (only the producers and the
  // column-sum role know; k-block kbt of the tripled range is block kbt % kb_single of pass kbt / kb_single)

  if (warp == 0 || warp == 3) {
    // ===== TMA producers: warp 0 feeds the A half of every stage, warp 3 the B half =====
    // One thread can issue a tensor box only every ~600 cycles whatever its size (tools/microbench/rowrate.cu: 16 KB boxes from one
    // thread top out at 55 GB/s per SM, from four threads at 135 GB/s); with one producer issuing both operands a k-block took
    // ~1100 cycles against 512 of tensor time.  Two issuing warps, one box each per k-block.
    // MN-major operands whose MN extent is whole 32-column blocks come in ONE 3-D box per stage ([block][k][32]: the same
    // shared-memory image as separate 32-column boxes; a_mn / b_mn == 2).
    if (lane == 0) {
      const bool is_a = (warp == 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += units) {
        const GemmProblem& P = grp.p[problem_of(tile)];
        const GemmTiles tl = P.tl;
        const int a_mn = P.a_mn, b_mn = P.b_mn;
        const int kb_single = (P.K + kBK - 1) / kBK, num_kb = kb_single * tl.passes;
        const int tiles_mn = tl.tiles_m * tl.tiles_n;
        const int lt = tile - P.tile0;
        const int split = lt / tiles_mn, r = lt - split * tiles_mn;
        // this CTA's rows of A and its rows of the B tile (pair: the second CTA takes the second half of both)
        const int m0 = (r / tl.tiles_n) * kTileM + (int)rank * kBM, n0 = (r % tl.tiles_n) * BN + (int)rank * L::kBRows;
        const int kb0 = split * tl.kb_per_split, kb1 = min(num_kb, kb0 + tl.kb_per_split);
        const uint64_t pol = l2_policy(is_a ? P.a_hint : P.b_hint);
        for (int kbt = kb0; kbt < kb1; ++kbt) {
          const int pass = kbt / kb_single, kb = kbt - pass * kb_single;
          const CUtensorMap* mapA = (pass == 1) ? &P.tmALo : &P.tmA;
          const CUtensorMap* mapB = (pass == 2) ? &P.tmBLo : &P.tmB;
          wait_t(empty_bar(stage), phase ^ 1u, w0);
          const uint32_t sa = base + stage * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
          // pair: the leader's producers announce both CTAs' bytes on the leader's barrier; the peer's only load (their bytes may land
          // before the announcement: the transaction count is signed, and the phase cannot complete before the leader's two arrivals)
          const uint32_t fbar = (CG == 2) ? mapa_shared(full_bar(stage), 0u) : full_bar(stage);
          if (is_a) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), CG * L::kABytes);
            if (a_mn == 0) {
              tma_ld2<CG>(sa, mapA, kb * kBK, m0, fbar, pol);
            } else if (a_mn == 2) {
              tma_ld3<CG>(sa, mapA, 0, kb * kBK, m0 >> 5, fbar, pol);
            } else {
#pragma unroll
              for (int j = 0; j < kBM / 32; ++j) tma_ld2<CG>(sa + j * (kBK * 128), mapA, m0 + 32 * j, kb * kBK, fbar, pol);
            }
          } else {
            if (rank == 0) mbar_expect_tx(full_bar(stage), CG * L::kBBytes);
            if (b_mn == 0) {
              tma_ld2<CG>(sb, mapB, kb * kBK, n0, fbar, pol);
            } else if (b_mn == 2) {
              tma_ld3<CG>(sb, mapB, 0, kb * kBK, n0 >> 5, fbar, pol);
            } else {
#pragma unroll
              for (int j = 0; j < L::kBRows / 32; ++j) tma_ld2<CG>(sb + j * (kBK * 128), mapB, n0 + 32 * j, kb * kBK, fbar, pol);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (trace && is_a) { trace[blockIdx.x * 8 + 0] = w0; trace[blockIdx.x * 8 + 7] = (unsigned long long)(clock64() - t_start); }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && rank == 0) {
      // Shared-memory descriptors: built per tile for stage 0 / k-step 0; the address field (bits [0,14), units of 16 B) is all that
      // changes inside a tile, so a stage or k-step is one 64-bit add.  This thread's instruction stream is serial and sits on the
      // critical path (measured: ~1000 cycles per k-block against 512 of tensor time), so nothing is recomputed inside the k loop.
      // K-major: 8 fp32 of K = 32 B inside the 128 B swizzle row (SWIZZLE_128B, 8-row atoms 1024 B apart).
      // MN-major: 8 K-rows = two 4-row 512 B atoms of SWIZZLE_128B_BASE32B (SBO), 32-wide MN chunks kBK*128 B apart (LBO).
      constexpr uint64_t kStageStep = (uint64_t)(L::kStageBytes >> 4);
      int stage = 0, as = 0;
      uint32_t phase = 0, aphase = 0;
      for (int tile = unit; tile < num_tiles; tile += units) {
        const GemmProblem& P = grp.p[problem_of(tile)];
        const GemmTiles tl = P.tl;
        const int a_mn = P.a_mn, b_mn = P.b_mn;
        // instruction descriptor: c=F32 [4,6), a=TF32 [7,10), b=TF32 [10,13), a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
        const uint64_t da0 = a_mn ? make_smem_desc(base, tune.mn_lbo, tune.mn_sbo, 1) : make_smem_desc(base, tune.k_lbo, tune.k_sbo, tune.k_layout);
        const uint64_t db0 = b_mn ? make_smem_desc(base + L::kABytes, tune.mn_lbo, tune.mn_sbo, 1) : make_smem_desc(base + L::kABytes, tune.k_lbo, tune.k_sbo, tune.k_layout);
        const uint64_t a_step = a_mn ? (1024u >> 4) : (32u >> 4), b_step = b_mn ? (1024u >> 4) : (32u >> 4);
        const int num_kb = ((P.K + kBK - 1) / kBK) * tl.passes;
        const int tiles_mn = tl.tiles_m * tl.tiles_n;
        const int split = (tile - P.tile0) / tiles_mn;
        const int kb0 = split * tl.kb_per_split, kb1 = min(num_kb, kb0 + tl.kb_per_split);
        uint64_t da = da0 + (uint64_t)stage * kStageStep, db = db0 + (uint64_t)stage * kStageStep;
        wait_t(tempty_bar(as), aphase ^ 1u, w1);  // epilogue has drained this accumulator
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * BN);
        uint32_t accumulate = 0u;
        for (int kb = kb0; kb < kb1; ++kb) {
          wait_t(full_bar(stage), phase, w0);
          tcgen05_fence_after();
#pragma unroll
          for (int kk = 0; kk < kBK / kUmmaK; ++kk) {
            umma_tf32_cg<CG>(tacc, da + kk * a_step, db + kk * b_step, idesc, accumulate);
            accumulate = 1u;
          }
          tcgen05_commit_cg<CG>(empty_bar(stage));  // frees the smem slot (in both CTAs of a pair) once these MMAs retire
          da += kStageStep; db += kStageStep;
          if (++stage == kStages) { stage = 0; phase ^= 1u; da = da0; db = db0; }
        }
        tcgen05_commit_cg<CG>(tfull_bar(as));  // accumulator complete (both CTAs' epilogues)
        as ^= 1;
        if (as == 0) aphase ^= 1u;
      }
      if (trace) { trace[blockIdx.x * 8 + 1] = w0; trace[blockIdx.x * 8 + 2] = w1; }
    }
  } else if (warp == 2) {
    // ===== column sums of the MN-major B tiles (bias gradient) =====
    // Every CTA of an n-tile sees the same B tiles; the 32 K-rows of each are shared out over the m-tiles (tiles_m = 2 or 4
    // for the Dense layers: 16 or 8 rows per CTA; otherwise m-tile 0 takes them all), so that no CTA's ring is held up by
    // this role.  Thread t owns two groups of 4 consecutive columns, 128 columns apart: chunk (t>>3) + 4i (32 columns, 4 KB
    // apart), 32-byte atom (t>>1)&3, half t&1; a quarter-warp reads one contiguous 128-byte row per LDS.128 (conflict-free under
    // the 32-byte-atom swizzle).  In a launch that mixes problems with and without a column sum this warp attends every tile (the
    // ring's empty barriers count it) and only sums where one is asked for.
    if (any_colsum) {
      const int chunk0 = lane >> 3, atom = (lane >> 1) & 3, half = lane & 1;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += units) {
        const GemmProblem& P = grp.p[problem_of(tile)];
        const GemmTiles tl = P.tl;
        float* colsum = P.colsum;
        const int N = P.N;
        const int kb_single = (P.K + kBK - 1) / kBK, num_kb = kb_single * tl.passes;
        const int tiles_mn = tl.tiles_m * tl.tiles_n;
        const bool share = (32 % tl.tiles_m) == 0;
        const int rows_per = share ? 32 / tl.tiles_m : 32;
        const int lt = tile - P.tile0;
        const int split = lt / tiles_mn, r = lt - split * tiles_mn;
        const int mt = r / tl.tiles_n;
        const bool active = colsum != nullptr && (share || mt == 0);
        const int k_first = share ? mt * rows_per : 0;
        const int n0 = (r % tl.tiles_n) * BN + (int)rank * L::kBRows;  // first column of the B rows this CTA stages
        const int kb0 = split * tl.kb_per_split, kb1 = min(num_kb, kb0 + tl.kb_per_split);
        constexpr int kColGroups = (L::kBRows + 127) / 128;
        float4 acc[kColGroups];
#pragma unroll
        for (int i = 0; i < kColGroups; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if constexpr (CG == 2) {  // the stage has landed in both CTAs: tell the peer's column-sum warp (its own full barrier sees no loads)
            if (rank == 0 && lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(full_bar(stage), 1u));
          }
          if (active && kb / kb_single != 1) {  // 3xTF32: pass 1 shows the B tiles a second time (hi from pass 0 + lo from pass 2 = the fp32 sum)
            const uint8_t* cb = base_ptr + stage * L::kStageBytes + L::kABytes + chunk0 * (kBK * 128) + half * 16;
#pragma unroll 4
            for (int k = k_first; k < k_first + rows_per; ++k) {
#pragma unroll
              for (int i = 0; i < kColGroups; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(cb + i * 4 * (kBK * 128) + k * 128 + ((atom ^ (k & 3)) << 5));
                acc[i].x += v.x; acc[i].y += v.y; acc[i].z += v.z; acc[i].w += v.w;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (colsum != nullptr && tl.det_colsum) {  // slot (split, m-tile): summed in slot order by splitk_reduce_kernel (CTAs without a share store zeros)
#pragma unroll
          for (int i = 0; i < kColGroups; ++i) {
            const int col = n0 + (chunk0 + 4 * i) * 32 + atom * 8 + half * 4;
            if (col < N && (chunk0 + 4 * i) * 32 < L::kBRows) *reinterpret_cast<float4*>(colsum + (size_t)(split * tl.tiles_m + mt) * N + col) = acc[i];
          }
        } else if (active) {
#pragma unroll
          for (int i = 0; i < kColGroups; ++i) {
            const int col = n0 + (chunk0 + 4 * i) * 32 + atom * 8 + half * 4;
            if (col < N && (chunk0 + 4 * i) * 32 < L::kBRows) {  // N is a multiple of 4
              atomicAdd(colsum + col, acc[i].x); atomicAdd(colsum + col + 1, acc[i].y);
              atomicAdd(colsum + col + 2, acc[i].z); atomicAdd(colsum + col + 3, acc[i].w);
            }
          }
        }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fused ops -> swizzled staging (transpose) -> coalesced global stores =====
    // Eight warps, two per TMEM lane quarter (a warp may only touch lanes 32 (warp % 4) ..): warp (q, h) takes the 32-column chunks
    // c = h, h + 2, ... of rows 32 q .. 32 q + 31.
    // Data path: NOT the TMA unit.  The SM's one TMA engine serves its requests in order, and the operand ring keeps up to four 48 KB
    // loads queued in it: a 4 KB store (or aux load) issued behind them waited ~2 us for its turn -- the source-level profile of the
    // TMA-store epilogue had 25-40 % of all samples on cp.async.bulk.wait_group.read / the aux mbarrier, and a plain store epilogue took
    // 5 800 cycles per 128 x 256 tile.  So results go accumulator registers (lane = row) -> swizzled 4 KB staging chunk -> registers in
    // the transposed mapping (8 lanes per 128-byte row segment) -> st.global.v4, full lines per quarter-warp; the residual / ReLU-mask
    // operand comes in by ld.global.v4 in the same mapping, one chunk ahead (registers), and through the same staging chunk.
    // Only split-K accumulation (weight gradients: a few tiles) still uses TMA reduce-add.
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int h = (warp - kEpiWarp0) >> 2;
    const int ew = warp - kEpiWarp0;
    const uint32_t epi = base + L::kEpiOff + ew * kChunkBytes;
    uint8_t* epi_ptr = base_ptr + L::kEpiOff + ew * kChunkBytes;
    const uint32_t sw = (uint32_t)(lane & 7);
    const int tr = lane >> 3, ts = lane & 7;  // transposed mapping: instruction i covers rows 4 i + tr, 16-byte segment ts
    uint8_t* own_row = epi_ptr + lane * 128;
    int as = 0;
    uint32_t aphase = 0;
    // geometry of a tile of the launch's tile space (for this warp: rows of its quarter, chunks of its parity)
    struct TileGeo { const GemmProblem* P; int split, m0, n0, nchunks; bool use_aux; };
    auto geo_of = [&](int tile) {
      TileGeo g;
      g.P = &grp.p[problem_of(tile)];
      const GemmTiles& tl = g.P->tl;
      const int tiles_mn = tl.tiles_m * tl.tiles_n;
      const int lt = tile - g.P->tile0;
      g.split = lt / tiles_mn;
      const int r = lt - g.split * tiles_mn;
      g.m0 = (r / tl.tiles_n) * kTileM + (int)rank * kBM;  // this CTA's 128 rows (= its TMEM lanes) of the tile
      g.n0 = (r % tl.tiles_n) * BN;
      g.nchunks = min(BN / 32, (g.P->N - g.n0 + 31) / 32);
      g.use_aux = aux_mode != 0 && g.P->aux_kind != 0 && (aux_mode == 2 || g.split == 0);  // the residual is added by split 0 only
      return g;
    };
    float4 an[8];  // aux operand of the NEXT chunk this warp will process, transposed mapping
    auto fetch_aux = [&](const TileGeo& g, int c) {
      const float* aux_src = aux_mode == 1 ? g.P->ep.residual : g.P->ep.relu_src;
      const int aux_ld = aux_mode == 1 ? g.P->ep.ldr : g.P->ep.ld_relu;
      const int grow0 = g.m0 + q * 32 + tr, gcol = g.n0 + c * 32 + ts * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int grow = grow0 + 4 * i;
        // plain (coherent) loads: the residual may be the output buffer itself (in-place x += ...); every element is read before the one warp that owns it writes it
        an[i] = (grow < g.P->M && gcol < g.P->N) ? *reinterpret_cast<const float4*>(aux_src + (size_t)grow * aux_ld + gcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    if constexpr (aux_mode != 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) an[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (unit < num_tiles) {
        const TileGeo g0 = geo_of(unit);
        if (g0.use_aux && h < g0.nchunks) fetch_aux(g0, h);
      }
    }
    // ReLU gates as bits (kEpiReluBits): lane = row, so a chunk's gates are ONE word per thread ([chunk][row] layout: a 128-byte line per
    // warp); fetched one chunk ahead like the aux operand.  A problem of the launch without gate words gets all ones.
    uint32_t gate_next = 0xffffffffu;
    auto fetch_gates = [&](const TileGeo& g, int c) {
      const int grow = g.m0 + q * 32 + lane;
      const uint32_t* bits = g.P->ep.relu_bits;
      return (bits != nullptr && grow < g.P->M) ? __ldg(bits + (size_t)((g.n0 >> 5) + c) * g.P->ep.relu_bits_ld + grow) : 0xffffffffu;
    };
    if constexpr (EPI & kEpiReluBits) {
      if (unit < num_tiles) {
        const TileGeo g0 = geo_of(unit);
        if (h < g0.nchunks) gate_next = fetch_gates(g0, h);
      }
    }
    for (int tile = unit; tile < num_tiles; tile += units) {
      const TileGeo g = geo_of(tile);
      const GemmEpilogue& ep = g.P->ep;
      const GemmTiles& tl = g.P->tl;
      const int M = g.P->M, N = g.P->N;
      const int split = g.split, m0 = g.m0, n0 = g.n0, nchunks = g.nchunks;
      const bool use_aux = g.use_aux;
      const int row0 = m0 + q * 32;
      const int row = row0 + lane;
      const int out_row0 = row0 + split * tl.slab_rows;
      const bool lead_split = (split == 0);
      const uint32_t drop_thr = dropout_threshold(ep.drop_rate);
      const float drop_scale = 1.0f / (1.0f - ep.drop_rate);
      bool flagged = false;
      if constexpr (EPI & kEpiRowflag) flagged = ep.rowflag != nullptr && row < M && ep.rowflag[row];
      // The tile's bias row (BN columns) is fetched once, 8 columns per lane, while the accumulator is still being computed,
      // and handed out by warp shuffles: a global load per chunk would put its L2 latency on the epilogue's critical path.
      float bt[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if constexpr (EPI & kEpiBias) {
        if (ep.bias != nullptr && lead_split && lane < BN / 8) {
          const int bc = n0 + lane * 8;
          if (bc < N) { const float4 t = __ldg(reinterpret_cast<const float4*>(ep.bias + bc)); bt[0] = t.x; bt[1] = t.y; bt[2] = t.z; bt[3] = t.w; }
          if (bc + 4 < N) { const float4 t = __ldg(reinterpret_cast<const float4*>(ep.bias + bc + 4)); bt[4] = t.x; bt[5] = t.y; bt[6] = t.z; bt[7] = t.w; }
        }
      }
      wait_t(tfull_bar(as), aphase, w0);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);
      auto release_acc = [&]() {  // this warp has its last chunk of the accumulator in registers
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(as), 0u));  // the leader's MMA thread waits for both CTAs' epilogues
          else mbar_arrive(tempty_bar(as));
        }
      };
      if (h >= nchunks) release_acc();  // a one-chunk tail tile: the odd warp has nothing to read
      float ln_s1 = 0.f, ln_s2 = 0.f;
      for (int c = h; c < nchunks; c += 2) {
        const int col0 = n0 + c * 32;
        uint32_t rr[32];
        if (trace) { const long long t0 = clock64(); tmem_ld32(tacc + (uint32_t)(c * 32), rr); w2 += (unsigned long long)(clock64() - t0); }
        else tmem_ld32(tacc + (uint32_t)(c * 32), rr);
        if constexpr (!(EPI & kEpiLayerNorm)) {
          if (c + 2 >= nchunks) release_acc();
        }
        if (lane == 0) {  // the staging chunk has been read out by an earlier tile's reduce-add (a no-op when none is pending)
          if (trace) { const long long t0 = clock64(); tma_wait_group_read<0>(); w1 += (unsigned long long)(clock64() - t0); }
          else tma_wait_group_read<0>();
        }
        __syncwarp();
        if constexpr (aux_mode != 0) {
          // this chunk's aux operand (fetched one chunk ago): registers (transposed mapping) -> staging -> own-row reads below; then
          // start fetching the next chunk's (this tile's chunk c + 2, or the first chunk of this CTA's next tile)
          if (use_aux) {
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(epi_ptr + (4 * i + tr) * 128 + ((ts ^ ((4 * i + tr) & 7)) << 4)) = an[i];
          }
          __syncwarp();
          if (c + 2 < nchunks) {
            if (use_aux) fetch_aux(g, c + 2);
          } else {
            const int nt = tile + units;
            if (nt < num_tiles) {
              const TileGeo gn = geo_of(nt);
              if (gn.use_aux && h < gn.nchunks) fetch_aux(gn, h);
            }
          }
        }
        uint32_t gates = 0xffffffffu, gates_out = 0u;
        if constexpr (EPI & kEpiReluBits) {
          gates = gate_next;
          if (c + 2 < nchunks) {
            gate_next = fetch_gates(g, c + 2);
          } else {
            const int nt = tile + units;
            gate_next = 0xffffffffu;
            if (nt < num_tiles) {
              const TileGeo gn = geo_of(nt);
              if (h < gn.nchunks) gate_next = fetch_gates(gn, h);
            }
          }
        }
        U4 dr = {0u, 0u, 0u, 0u};
        float4 av[8];
        if constexpr (aux_mode != 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) av[j] = use_aux ? *reinterpret_cast<const float4*>(own_row + ((j ^ sw) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          __syncwarp();  // every lane has its aux row: the staging chunk may take the results
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[4] = {__uint_as_float(rr[4 * j]), __uint_as_float(rr[4 * j + 1]), __uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])};
          if constexpr (EPI & kEpiBias) {  // bias of columns col0 + 4j .. +3 sits in lane 4c + j/2, registers 4(j&1) .. +3
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] += __shfl_sync(0xffffffffu, bt[(j & 1) * 4 + e], c * 4 + (j >> 1));
          }
          if constexpr (EPI & kEpiRelu) {
            if (ep.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) v[e] = fmaxf(v[e], 0.0f);
            }
          }
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (aux_mode != 0) a = av[j];
          if constexpr (aux_mode == 2) {
            if (use_aux) {
              v[0] = a.x > 0.0f ? v[0] : 0.0f; v[1] = a.y > 0.0f ? v[1] : 0.0f;
              v[2] = a.z > 0.0f ? v[2] : 0.0f; v[3] = a.w > 0.0f ? v[3] : 0.0f;
            }
          }
          if constexpr (EPI & kEpiBitsOut) {  // after bias + ReLU: the gate is "the stored value is positive", what relu_src > 0 tests
#pragma unroll
            for (int e = 0; e < 4; ++e) gates_out |= (v[e] > 0.0f ? 1u : 0u) << (4 * j + e);
          }
          if constexpr (EPI & kEpiReluBits) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = ((gates >> (4 * j + e)) & 1u) ? v[e] : 0.0f;
          }
          if constexpr (EPI & kEpiDropout) {  // one Philox block per eight columns (j even computes it, j odd uses its second half)
            if (ep.drop_enabled) {
              if ((j & 1) == 0) dr = philox4x32_10((((uint32_t)row + ep.drop_row0) * (uint32_t)N + (uint32_t)(col0 + 4 * j)) >> 3, ep.drop_site, 0u, 0u, ep.drop_seed, ep.drop_step);
              dropout_apply4(v, (j & 1) ? dr.z : dr.x, (j & 1) ? dr.w : dr.y, drop_thr, drop_scale);
            }
          }
          if constexpr (EPI & kEpiRowflag) {
            if (flagged) { v[0] = v[1] = v[2] = v[3] = 0.0f; }
          }
          if constexpr (aux_mode == 1) { v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; }
          *reinterpret_cast<float4*>(own_row + ((j ^ sw) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
          if constexpr (EPI & kEpiLayerNorm) {  // row statistics + the finished values back into the accumulator for the second pass
            ln_s1 += (v[0] + v[1]) + (v[2] + v[3]);
            ln_s2 += (v[0] * v[0] + v[1] * v[1]) + (v[2] * v[2] + v[3] * v[3]);
            rr[4 * j] = __float_as_uint(v[0]); rr[4 * j + 1] = __float_as_uint(v[1]); rr[4 * j + 2] = __float_as_uint(v[2]); rr[4 * j + 3] = __float_as_uint(v[3]);
          }
        }
        if constexpr (EPI & kEpiLayerNorm) tmem_st32(tacc + (uint32_t)(c * 32), rr);
        if constexpr (EPI & kEpiBitsOut) {
          if (ep.relu_bits_out != nullptr && row < M) ep.relu_bits_out[(size_t)(col0 >> 5) * ep.relu_bits_ld + row] = gates_out;
        }
        if (ep.atomic) {
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&g.P->tmOut, epi, col0, row0);
            tma_commit_group();
          }
        } else {
          __syncwarp();
          const int gcol = col0 + ts * 4;
          float* orow = ep.out + (size_t)(out_row0 + tr) * ep.ldo + gcol;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 o = *reinterpret_cast<const float4*>(epi_ptr + (4 * i + tr) * 128 + ((ts ^ ((4 * i + tr) & 7)) << 4));
            if (row0 + 4 * i + tr < M && gcol < N) *reinterpret_cast<float4*>(orow + (size_t)(4 * i) * ep.ldo) = o;
          }
        }
      }
      if constexpr (EPI & kEpiLayerNorm) {
        // ---- fused LayerNorm of the rows just produced (N == BN: this tile holds whole rows; the two warps of a lane quarter hold the even
        // and the odd chunks of the same 32 rows).  Pass 1 above left the values in the accumulator and per-lane partial sums; the partner
        // warp's come through the staging chunks; pass 2 re-reads the values and writes the normalised rows.
        tmem_st_wait();
        __syncwarp();
        *reinterpret_cast<float2*>(epi_ptr + lane * 8) = make_float2(ln_s1, ln_s2);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // both warps of the quarter have published their sums
        const float2 other = *reinterpret_cast<const float2*>(base_ptr + L::kEpiOff + (ew ^ 4) * kChunkBytes + lane * 8);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");  // ... and read the partner's, before the staging chunks are reused
        const float mean = (ln_s1 + other.x) * (1.0f / BN);
        const float var = fmaxf((ln_s2 + other.y) * (1.0f / BN) - mean * mean, 0.0f);
        const float rstd = rsqrtf(var + kLnEps);
        if (h == 0 && row < M) { ep.ln_mean[row] = mean; ep.ln_rstd[row] = rstd; }
        for (int c = h; c < nchunks; c += 2) {
          const int col0 = n0 + c * 32;
          uint32_t rr[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), rr);
          if (c + 2 >= nchunks) release_acc();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + col0) + j), b4 = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + col0) + j);
            *reinterpret_cast<float4*>(own_row + ((j ^ sw) << 4)) =
                make_float4((__uint_as_float(rr[4 * j]) - mean) * rstd * g4.x + b4.x, (__uint_as_float(rr[4 * j + 1]) - mean) * rstd * g4.y + b4.y,
                            (__uint_as_float(rr[4 * j + 2]) - mean) * rstd * g4.z + b4.z, (__uint_as_float(rr[4 * j + 3]) - mean) * rstd * g4.w + b4.w);
          }
          __syncwarp();
          const int gcol = col0 + ts * 4;
          float* orow = ep.ln_out + (size_t)(row0 + tr) * ep.ln_ldo + gcol;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 o = *reinterpret_cast<const float4*>(epi_ptr + (4 * i + tr) * 128 + ((ts ^ ((4 * i + tr) & 7)) << 4));
            if (row0 + 4 * i + tr < M) *reinterpret_cast<float4*>(orow + (size_t)(4 * i) * ep.ln_ldo) = o;
          }
          __syncwarp();
        }
      }
      as ^= 1;
      if (as == 0) aphase ^= 1u;
    }
    if (lane == 0) tma_wait_group_read<0>();
    if (trace && warp == kEpiWarp0 && lane == 0) {
      trace[blockIdx.x * 8 + 3] = w0; trace[blockIdx.x * 8 + 4] = w1; trace[blockIdx.x * 8 + 5] = w2;
      trace[blockIdx.x * 8 + 6] = (unsigned long long)(clock64() - t_start);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync();  // neither CTA leaves (or frees its TMEM) while the pair's MMAs / remote arrivals may still touch it
  if (warp == 2) {
    if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------- SIMT bring-up kernel
__global__ void __launch_bounds__(256) gemm_simt(const float* __restrict__ A, int a_mn, int lda, const float* __restrict__ B, int b_mn, int ldb,
                                                 int M, int N, int K, int k_per_split, GemmEpilogue ep) {
  pdl_wait();
  __shared__ float As[16][65];
  __shared__ float Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int k_begin = blockIdx.z * k_per_split, k_end = min(K, k_begin + k_per_split);
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int i = threadIdx.x; i < 16 * 64; i += 256) {
      const int kk = i & 15, mm = i >> 4;
      const int k = k0 + kk;
      float a = 0.f, b = 0.f;
      if (k < k_end && m0 + mm < M) a = a_mn ? A[(size_t)k * lda + m0 + mm] : A[(size_t)(m0 + mm) * lda + k];
      if (k < k_end && n0 + mm < N) b = b_mn ? B[(size_t)k * ldb + n0 + mm] : B[(size_t)(n0 + mm) * ldb + k];
      As[kk][mm] = a;
      Bs[kk][mm] = b;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    epilogue_store4(v, row, n0 + tx * 4, N, ep, blockIdx.z == 0);
  }
}

// ------------------------------------------------------------------------------------------------- deterministic split-K
// out[m, n] = sum over splits s (ascending) of part[s][m][n]; colsum[n] += sum over slots (ascending) of colpart[slot][n].
// One thread per float4 of the output; the column sums are taken by the first CTAs.  Fixed summation order: bit-identical
// results from run to run, whatever order the GEMM's CTAs finished in.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t slab_floats, int M, int N, float* __restrict__ out,
                                                            int ldo, const float* __restrict__ colpart, int slots, float* __restrict__ colsum) {
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // float4 index into [M, N]
  const size_t n4 = (size_t)M * N / 4;
  if (part && i < n4) {
    float4 acc = reinterpret_cast<const float4*>(part)[i];
    for (int s = 1; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(part + (size_t)s * slab_floats)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const size_t e = i * 4, m = e / N, n = e - m * N;
    *reinterpret_cast<float4*>(out + m * ldo + n) = acc;
  }
  if (colpart && i < (size_t)N / 4) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < slots; ++s) {
      const float4 v = reinterpret_cast<const float4*>(colpart + (size_t)s * N)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* dst = reinterpret_cast<float4*>(colsum) + i;
    const float4 o = *dst;
    *dst = make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w);
  }
}

// ------------------------------------------------------------------------------------------------- 3xTF32 operand split
// lo = x - tf32(x), tf32() = round to nearest even on 10 mantissa bits: exactly what the TFLOAT32 tensor maps deliver as the "hi" part
// (measured: tests/test_gpu_parity.py::test_tf32_operand_rounding_of_the_product_path), so hi + lo == x exactly in fp32.
__global__ void __launch_bounds__(256) split_tf32_lo_kernel(const float* __restrict__ x, int rows, int cols4, int ld4, float* __restrict__ lo) {
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * cols4) return;
  const size_t r = i / cols4, c = i - r * cols4;
  const float4 v = reinterpret_cast<const float4*>(x)[r * ld4 + c];
  auto low = [](float f) {
    uint32_t b = __float_as_uint(f);
    b = (b + 0xFFFu + ((b >> 13) & 1u)) & ~0x1FFFu;
    return f - __uint_as_float(b);
  };
  reinterpret_cast<float4*>(lo)[r * ld4 + c] = make_float4(low(v.x), low(v.y), low(v.z), low(v.w));
}

int launch_split_tf32_lo(const float* x, int rows, int cols, int ld, float* lo, cudaStream_t stream) {
  if ((cols % 4) || (ld % 4)) { set_error("split_tf32_lo: cols and pitch must be multiples of 4"); return MFP_ERR_ARG; }
  const size_t n4 = (size_t)rows * (cols / 4);
  MFP_CUDA_OK(launch_pdl(split_tf32_lo_kernel, (unsigned)((n4 + 255) / 256), 256, 0, stream, x, rows, cols / 4, ld / 4, lo));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static bool k_atom32() { static const bool v = getenv("FLEXDM_K_ATOM32") != nullptr; return v; }  // experiment: one smem image for both majors

struct MapKey {
  const void* ptr;
  uint64_t dims[3], strides[2];
  uint32_t box[3], rank, kind;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    for (int i = 0; i < 3; ++i) { mix(k.dims[i]); mix(k.box[i]); }
    mix(k.strides[0]); mix(k.strides[1]); mix(k.rank); mix(k.kind);
    return h;
  }
};

class TensorMapCache {
 public:
  // fp32 tensor of rank 2 or 3, dims[0] innermost (contiguous), strides (in floats) of dims 1.. ; box per dim.
  //   kMapOperandK : MMA operand, K-major, SWIZZLE_128B, values rounded to TF32 (RN) by the TMA unit
  //   kMapOperandMN: MMA operand, MN-major, SWIZZLE_128B_ATOM_32B (the only MN-major layout for 32-bit types), TF32
  //   kMapEpilogue : epilogue staging chunks (store / reduce-add / residual load), SWIZZLE_128B, plain fp32
  const CUtensorMap* get(const float* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box, MapKind kind) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    key.ptr = ptr; key.rank = (uint32_t)rank; key.kind = (uint32_t)kind;
    for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
    auto it = maps_.find(key);
    if (it != maps_.end()) return &it->second;

    EncodeTiledFn enc = get_encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)"); return nullptr; }
    if (reinterpret_cast<uintptr_t>(ptr) & 15) { set_error("TMA operand must be 16-byte aligned"); return nullptr; }
    CUtensorMap m;
    cuuint64_t gdims[3] = {1, 1, 1}, gstr[2] = {0, 0};
    cuuint32_t gbox[3] = {1, 1, 1}, estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gdims[i] = dims[i]; gbox[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) {
      if (strides[i] % 4) { set_error("TMA operand pitch must be a multiple of 4 floats"); return nullptr; }
      gstr[i] = strides[i] * sizeof(float);
    }
    static const bool plain_f32 = getenv("FLEXDM_TMA_F32") != nullptr;  // default: round operands to TF32 (RN) in the TMA unit
    const CUtensorMapDataType dt = (kind == kMapEpilogue || plain_f32) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
    const CUtensorMapSwizzle sw = (kind == kMapOperandMN || (kind == kMapOperandK && k_atom32())) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUresult r = enc(&m, dt, (cuuint32_t)rank, const_cast<float*>(ptr), gdims, gstr, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed: %d (rank=%d dims=%llu,%llu,%llu)", (int)r, rank, (unsigned long long)gdims[0], (unsigned long long)gdims[1],
                (unsigned long long)gdims[2]);
      return nullptr;
    }
    auto res = maps_.emplace(key, m);
    return &res.first->second;
  }
  // Descriptors are copied into kernel parameters at launch, so none outlives a launcher call: when a caller keeps feeding fresh
  // buffers (Model.call on user tensors, debug entry points) the cache is restarted -- by the launchers, before their first get()
  // -- instead of growing without bound.
  void trim() {
    if (maps_.size() >= 4096) maps_.clear();
  }
  // 2-D [outer][inner] with row pitch ld (floats); box = [box_outer][box_inner]
  const CUtensorMap* get(const float* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer, MapKind kind) {
    const uint64_t dims[2] = {inner, outer}, strides[1] = {ld};
    const uint32_t box[2] = {box_inner, box_outer};
    return get(ptr, 2, dims, strides, box, kind);
  }

 private:
  std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps_;
};

const CUtensorMap* tensor_map_get(TensorMapCache* cache, const float* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                                  MapKind kind) {
  return cache->get(ptr, rank, dims, strides, box, kind);
}

void tensor_map_cache_trim(TensorMapCache* cache) { cache->trim(); }
TensorMapCache* tensor_map_cache_create() { return new TensorMapCache(); }
void tensor_map_cache_destroy(TensorMapCache* c) { delete c; }

bool pdl_enabled() {
  static const bool on = [] { const char* s = getenv("FLEXDM_PDL"); return !(s && s[0] == '0'); }();
  return on;
}

static uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* s = getenv(name);
  return s ? (uint32_t)strtoul(s, nullptr, 0) : dflt;
}

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static int epi_bits(const GemmEpilogue& ep) {
  return (ep.bias ? kEpiBias : 0) | (ep.relu ? kEpiRelu : 0) | (ep.residual ? kEpiResidual : 0) | (ep.relu_src ? kEpiReluMask : 0) |
         (ep.drop_enabled ? kEpiDropout : 0) | (ep.rowflag ? kEpiRowflag : 0) | (ep.ln_out ? kEpiLayerNorm : 0) |
         (ep.relu_bits ? kEpiReluBits : 0) | (ep.relu_bits_out ? kEpiBitsOut : 0);
}

// What has to follow a problem's GEMM in deterministic mode: the fixed-order sum of its split-K / column-sum partials.
struct DetReduce {
  bool split = false;
  float* col = nullptr;
  int splits = 0, tiles_m = 0, slab_rows = 0;
};

// Tensor maps, tiling and epilogue of one problem of a launch (tile width BN is the launch's).
template <int BN, int CG>
static int prepare_problem(TensorMapCache* cache, const GemmCall& c, int epi_launch, GemmProblem* P, DetReduce* red) {
  constexpr int kBRows = BN / CG;     // rows of the B tile one CTA stages (a CTA pair splits the tile's columns between its two CTAs)
  constexpr int kTileM = kBM * CG;    // rows of a tile of the tile space
  if (c.ep.residual && c.ep.relu_src) { set_error("gemm: residual and relu_src cannot be combined"); return MFP_ERR_ARG; }
  if (c.ep.relu_bits && c.ep.relu_src) { set_error("gemm: relu_bits and relu_src are two forms of the same operand"); return MFP_ERR_ARG; }
  if ((c.ep.relu_bits || c.ep.relu_bits_out) && (c.ep.relu_bits_ld < c.M || (c.ep.relu_bits_ld % 32) || c.splits > 1)) {
    set_error("gemm: ReLU gate words need relu_bits_ld >= M, a multiple of 32, and no split-K");
    return MFP_ERR_ARG;
  }
  if (c.ep.relu_bits_out && !c.ep.relu) { set_error("gemm: relu_bits_out belongs to a ReLU epilogue"); return MFP_ERR_ARG; }
  if (c.colsum && !c.b.mn_major) { set_error("gemm: the fused column sum needs an MN-major B operand"); return MFP_ERR_ARG; }
  if (c.ep.ln_out && (c.N != BN || c.splits > 1 || !c.ep.ln_gamma || !c.ep.ln_beta || !c.ep.ln_mean || !c.ep.ln_rstd || (c.ep.ln_ldo % 4))) {
    set_error("gemm: the fused LayerNorm needs N == %d (whole rows in one tile), no split-K, and gamma / beta / mean / rstd", BN);
    return MFP_ERR_ARG;
  }
  // MN-major operand [K][MN] with pitch ld: 2-D boxes of 32 columns x kBK rows, or -- when MN is whole 32-column blocks -- a 3-D
  // view [MN / 32][K][32] whose box (32, kBK, tile / 32) brings a whole stage in one instruction (mode 2 in the kernel)
  auto mn_map = [&](const GemmOperand& o, int mn, int tile_mn, int* mode) -> const CUtensorMap* {
    if (mn % 32 == 0 && o.ld % 32 == 0) {
      const uint64_t d3[3] = {32, (uint64_t)c.K, (uint64_t)mn / 32}, s3[2] = {(uint64_t)o.ld, 32};
      const uint32_t b3[3] = {32, (uint32_t)kBK, (uint32_t)(tile_mn / 32)};
      *mode = 2;
      return cache->get(o.ptr, 3, d3, s3, b3, kMapOperandMN);
    }
    *mode = 1;
    return cache->get(o.ptr, mn, c.K, o.ld, 32, kBK, kMapOperandMN);
  };
  int a_mode = 0, b_mode = 0;
  const CUtensorMap* ma = c.a.mn_major ? mn_map(c.a, c.M, kBM, &a_mode) : cache->get(c.a.ptr, c.K, c.M, c.a.ld, kBK, kBM, kMapOperandK);
  const CUtensorMap* mb = c.b.mn_major ? mn_map(c.b, c.N, kBRows, &b_mode) : cache->get(c.b.ptr, c.K, c.N, c.b.ld, kBK, kBRows, kMapOperandK);
  const CUtensorMap* mo = cache->get(c.ep.out, c.N, c.M, c.ep.ldo, 32, 32, kMapEpilogue);  // split-K reduce-add target
  // 3xTF32: the low parts x - tf32(x) of both operands, same geometry (pitch = the operand's own)
  const CUtensorMap *mal = ma, *mbl = mb;
  if (c.a_lo && c.b_lo) {
    GemmOperand alo = c.a, blo = c.b;
    alo.ptr = c.a_lo; blo.ptr = c.b_lo;
    int unused_mode = 0;
    mal = c.a.mn_major ? mn_map(alo, c.M, kBM, &unused_mode) : cache->get(alo.ptr, c.K, c.M, alo.ld, kBK, kBM, kMapOperandK);
    mbl = c.b.mn_major ? mn_map(blo, c.N, kBRows, &unused_mode) : cache->get(blo.ptr, c.K, c.N, blo.ld, kBK, kBRows, kMapOperandK);
  }
  if (!ma || !mb || !mo || !mal || !mbl) return MFP_ERR_CUDA;
  GemmTiles tl;
  tl.passes = (c.a_lo && c.b_lo) ? 3 : 1;
  const int num_kb = tl.passes * ((c.K + kBK - 1) / kBK);
  int splits = c.splits < 1 ? 1 : c.splits;
  if (splits > num_kb) splits = num_kb;
  tl.kb_per_split = (num_kb + splits - 1) / splits;
  tl.splits = (num_kb + tl.kb_per_split - 1) / tl.kb_per_split;  // no empty split
  tl.tiles_m = (c.M + kTileM - 1) / kTileM;
  tl.tiles_n = (c.N + BN - 1) / BN;
  GemmEpilogue ep = c.ep;
  tl.slab_rows = 0;
  tl.det_colsum = 0;
  float* colsum = c.colsum;
  const bool det = c.det_ws != nullptr;
  red->split = det && tl.splits > 1;
  red->col = nullptr;
  if (red->split || (det && c.colsum)) {
    tl.slab_rows = red->split ? tl.tiles_m * kTileM : 0;
    const size_t part_floats = red->split ? (size_t)tl.splits * tl.slab_rows * c.N : 0;
    const size_t col_floats = c.colsum ? (size_t)tl.splits * tl.tiles_m * c.N : 0;
    if (part_floats + col_floats > c.det_ws_floats) { set_error("gemm: deterministic scratch too small (%zu > %zu floats)", part_floats + col_floats, c.det_ws_floats); return MFP_ERR_ARG; }
    if (red->split) {
      if (c.ep.bias || c.ep.residual || c.ep.relu || c.ep.relu_src || c.ep.drop_enabled || c.ep.rowflag) { set_error("gemm: deterministic split-K takes a plain epilogue"); return MFP_ERR_ARG; }
      ep.out = c.det_ws;  // split s stores its tiles at row offset s * slab_rows of the scratch block
      ep.ldo = c.N;
    }
    if (c.colsum) { red->col = c.det_ws + part_floats; colsum = red->col; tl.det_colsum = 1; }
  } else if (tl.splits > 1) {
    ep.atomic = 1;
  }
  red->splits = tl.splits; red->tiles_m = tl.tiles_m; red->slab_rows = tl.slab_rows;
  P->tmA = *ma; P->tmB = *mb; P->tmOut = *mo; P->tmALo = *mal; P->tmBLo = *mbl;
  P->M = c.M; P->N = c.N; P->K = c.K; P->a_mn = a_mode; P->b_mn = b_mode;
  P->tl = tl;
  P->ep = ep;
  P->colsum = colsum;
  const int launch_aux = (epi_launch & kEpiResidual) ? 1 : ((epi_launch & kEpiReluMask) ? 2 : 0);
  P->aux_kind = (c.ep.residual || c.ep.relu_src) ? launch_aux : 0;
  return MFP_OK;
}

template <int BN, int EPI, int CG>
static int launch_tcgen05(TensorMapCache* cache, const GemmCall* calls, int n, cudaStream_t stream) {
  using L = GemmSmem<BN, CG>;
  static_assert(L::kTotal <= 227 * 1024, "GEMM shared memory exceeds the 227 KB per-CTA limit");
  static_assert(sizeof(GemmGroup) <= 4000, "kernel parameters must stay under 4 KB");
  static bool attr_set = false;
  if (!attr_set) {
    MFP_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_tcgen05<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  cache->trim();
  static const GemmTune tune = {env_u32("FLEXDM_MN_LBO", kBK * 128), env_u32("FLEXDM_MN_SBO", 512), env_u32("FLEXDM_K_LBO", 16), env_u32("FLEXDM_K_SBO", 1024),
                                k_atom32() ? 1u : 2u};
  GemmGroup grp;
  memset(&grp, 0, sizeof(grp));
  DetReduce red[kMaxGroup];
  grp.n = n;
  int tiles = 0;
  for (int i = 0; i < n; ++i) {
    MFP_TRY((prepare_problem<BN, CG>(cache, calls[i], EPI, &grp.p[i], &red[i])));
    grp.p[i].tile0 = tiles;
    tiles += grp.p[i].tl.tiles_m * grp.p[i].tl.tiles_n * grp.p[i].tl.splits;
    if (grp.p[i].colsum) grp.any_colsum = 1;
  }
  // A weight gradient next to the input gradient off the same dY: dY is read twice (as the MN-major B operand of the first, as the K-major
  // A operand of the second), and the FFN's hidden activation as well (A operand of dW2, ReLU mask of the input gradient).  The first
  // reads ask L2 to keep the lines, the last ones to drop them first (FLEXDM_L2_HINTS=0: off).
  static const bool l2_hints = [] { const char* e = getenv("FLEXDM_L2_HINTS"); return !(e && e[0] == '0'); }();
  if (l2_hints && n == 2 && grp.p[0].tl.splits > 1) {
    if (calls[1].a.ptr == calls[0].b.ptr) { grp.p[0].b_hint = 1; grp.p[1].a_hint = 2; }
    if (calls[1].ep.relu_src == calls[0].a.ptr) grp.p[0].a_hint = 1;
  }
  grp.total_tiles = tiles;
  const int units = num_sms() / CG;  // CTAs, or CTA pairs (clusters of two: tcgen05 cta_group::2)
  const int grid = (tiles < units ? tiles : units) * CG;
  static const bool trace_on = getenv("FLEXDM_GEMM_TRACE") != nullptr;  // debugging aid: prints per-role wait cycles of every launch
  static unsigned long long* trace = nullptr;
  if (trace_on && !trace) { MFP_CUDA_OK(cudaMalloc(&trace, 148 * 8 * sizeof(unsigned long long))); }
  if (trace_on) MFP_CUDA_OK(cudaMemsetAsync(trace, 0, 148 * 8 * sizeof(unsigned long long), stream));
  MFP_CUDA_OK(launch_pdl_cluster(gemm_tf32_tcgen05<BN, EPI, CG>, grid, kGemmThreads, L::kTotal, stream, CG, grp, tune, trace_on ? trace : nullptr));
  if (trace_on) {
    unsigned long long hbuf[148 * 8];
    MFP_CUDA_OK(cudaStreamSynchronize(stream));
    MFP_CUDA_OK(cudaMemcpy(hbuf, trace, sizeof(hbuf), cudaMemcpyDeviceToHost));
    double acc[8] = {};
    for (int b = 0; b < grid && b < 148; ++b)
      for (int k = 0; k < 8; ++k) acc[k] += (double)hbuf[b * 8 + k] / grid;
    const GemmCall& c = calls[0];
    fprintf(stderr, "gemm trace cg=%d problems=%d first: M=%d N=%d K=%d a_mn=%d b_mn=%d epi=%d splits=%d tiles/cta=%.2f | producer: wait_empty %.0f of %.0f | mma: wait_full %.0f wait_tempty %.0f | "
            "epilogue(w4): wait_tfull %.0f wait_store %.0f wait_tmem_ld %.0f of %.0f cycles\n", CG, n, c.M, c.N, c.K, c.a.mn_major, c.b.mn_major, EPI, grp.p[0].tl.splits,
            (double)tiles / grid, acc[0], acc[7], acc[1], acc[2], acc[3], acc[4], acc[5], acc[6]);
  }
  MFP_CUDA_OK(cudaGetLastError());
  for (int i = 0; i < n; ++i) {
    if (!red[i].split && !red[i].col) continue;
    const GemmCall& c = calls[i];
    const size_t n4 = red[i].split ? (size_t)c.M * c.N / 4 : (size_t)c.N / 4;
    MFP_CUDA_OK(launch_pdl(splitk_reduce_kernel, (unsigned)((n4 + 255) / 256), 256, 0, stream, red[i].split ? (const float*)c.det_ws : (const float*)nullptr, red[i].splits,
                           (size_t)red[i].slab_rows * c.N, c.M, c.N, c.ep.out, c.ep.ldo, (const float*)red[i].col, red[i].splits * red[i].tiles_m, c.colsum));
    MFP_CUDA_OK(cudaGetLastError());
  }
  return MFP_OK;
}

static int check_call(const GemmCall& c) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) { set_error("gemm: empty problem %dx%dx%d", c.M, c.N, c.K); return MFP_ERR_ARG; }
  if ((c.N % 4) || (c.ep.ldo % 4)) { set_error("gemm: N and ldo must be multiples of 4 (N=%d ldo=%d)", c.N, c.ep.ldo); return MFP_ERR_ARG; }
  return MFP_OK;
}

// One launch for up to kMaxGroup independent problems (tcgen05 path); `epi` = union of their epilogue bits.
static int launch_tcgen05_any(TensorMapCache* cache, const GemmCall* calls, int n, int epi, bool wide, bool pair, cudaStream_t stream) {
#define MFP_GEMM_CASE(E)                                                                                \
  case (E):                                                                                             \
    return !wide ? launch_tcgen05<128, (E), 1>(cache, calls, n, stream)                                 \
                 : (pair ? launch_tcgen05<256, (E), 2>(cache, calls, n, stream) : launch_tcgen05<256, (E), 1>(cache, calls, n, stream));
  switch (epi) {
    MFP_GEMM_CASE(0)                                           // dgrad / wgrad
    MFP_GEMM_CASE(kEpiBias)                                    // QKV, heads
    MFP_GEMM_CASE(kEpiBias | kEpiRelu)                         // FFN 1
    MFP_GEMM_CASE(kEpiBias | kEpiResidual)                     // attention output / FFN 2, eval
    MFP_GEMM_CASE(kEpiBias | kEpiResidual | kEpiDropout)       // attention output / FFN 2, training
    MFP_GEMM_CASE(kEpiBias | kEpiResidual | kEpiLayerNorm)                 // ... with the following LayerNorm fused, eval
    MFP_GEMM_CASE(kEpiBias | kEpiResidual | kEpiDropout | kEpiLayerNorm)   // ... training
    MFP_GEMM_CASE(kEpiResidual | kEpiRowflag)                  // encoder Dense of a numerical field
    MFP_GEMM_CASE(kEpiResidual | kEpiRowflag | kEpiLayerNorm)  // ... the last one, with the first block's LayerNorm 1 fused
    MFP_GEMM_CASE(kEpiReluMask)                                // dgrad through the FFN ReLU (alone or next to a weight gradient)
    MFP_GEMM_CASE(kEpiReluBits)                                // ... with the gates as bits (no aux operand)
    MFP_GEMM_CASE(kEpiBias | kEpiRelu | kEpiBitsOut)           // FFN 1 of a training step: also writes the gate words
    MFP_GEMM_CASE(kEpiResidual)                                // dgrad + the gradient of the skip path (post-LayerNorm block)
    default:
      set_error("gemm: epilogue combination 0x%x is not instantiated", epi);
      return MFP_ERR_UNSUPPORTED;
  }
#undef MFP_GEMM_CASE
}

// CTA pairs (256-row tiles) for problems of at least one full pair tile; FLEXDM_GEMM_PAIR=0 keeps every launch on single CTAs
static bool pair_mode(int M) {
  static const bool on = [] { const char* e = getenv("FLEXDM_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  return on && M >= 2 * kBM;
}

// Split-K factor of a weight-gradient GEMM (K = tokens): the largest that keeps its tiles within one wave of the persistent kernel
// (one tile per CTA, or per CTA pair), with at least four k-blocks per split.
int gemm_wgrad_splits(int M, int N, int K) {
  const int bn = (N <= 128) ? 128 : 256;
  const bool pair = bn == 256 && pair_mode(M);
  const int tile_m = pair ? 2 * kBM : kBM;
  const int tiles = ((M + tile_m - 1) / tile_m) * ((N + bn - 1) / bn);
  int s = (num_sms() / (pair ? 2 : 1)) / tiles;
  const int kb = (K + kBK - 1) / kBK;
  if (s > kb / 4) s = kb / 4;
  return s < 1 ? 1 : s;
}

int launch_gemm(TensorMapCache* cache, const GemmCall& c, int impl, cudaStream_t stream) {
  MFP_TRY(check_call(c));
  if (impl == 1) {
    if (c.colsum) { set_error("gemm: the SIMT bring-up kernel has no fused column sum"); return MFP_ERR_ARG; }
    if (c.ep.relu_bits_out) { set_error("gemm: the SIMT bring-up kernel does not write ReLU gate words"); return MFP_ERR_ARG; }
    int splits = (c.splits < 1 || c.det_ws) ? 1 : c.splits;  // deterministic mode: no atomic accumulation over splits
    int k_per_split = ((c.K + splits - 1) / splits + 15) / 16 * 16;
    splits = (c.K + k_per_split - 1) / k_per_split;
    GemmEpilogue ep = c.ep;
    if (splits > 1) ep.atomic = 1;
    dim3 grid((c.M + 63) / 64, (c.N + 63) / 64, splits);
    MFP_CUDA_OK(launch_pdl(gemm_simt, grid, 256, 0, stream, c.a.ptr, c.a.mn_major, c.a.ld, c.b.ptr, c.b.mn_major, c.b.ld, c.M, c.N, c.K, k_per_split, ep));
    MFP_CUDA_OK(cudaGetLastError());
    return MFP_OK;
  }
  // Tile width: 256 columns, except for narrow outputs and for small problems that would leave most SMs without a tile (a 64-document
  // shard has 64 row tiles: at N = 256 that is 64 CTAs of 148) -- those take 128-column tiles, twice as many CTAs.
  static const int small_tiles = [] { const char* e = getenv("FLEXDM_SMALL_TILES"); return e ? atoi(e) : 100; }();
  const int tiles256 = ((c.M + kBM - 1) / kBM) * ((c.N + 255) / 256) * (c.splits < 1 ? 1 : c.splits);
  const bool wide = c.N > 128 && !(tiles256 < small_tiles && c.N % 128 == 0 && !c.ep.ln_out);
  return launch_tcgen05_any(cache, &c, 1, epi_bits(c.ep), wide, wide && pair_mode(c.M), stream);
}

int launch_gemm_group(TensorMapCache* cache, const GemmCall* calls, int n, cudaStream_t stream) {
  if (n < 1 || n > kMaxGroup) { set_error("gemm group: 1..%d problems", kMaxGroup); return MFP_ERR_ARG; }
  int epi = 0;
  bool wide = false, pair = true;
  for (int i = 0; i < n; ++i) {
    MFP_TRY(check_call(calls[i]));
    pair = pair && pair_mode(calls[i].M);
    const int e = epi_bits(calls[i].ep);
    // the kernel's epilogue variant is compiled in: problems of one launch may differ only in whether they use the aux operand / bias /
    // ReLU / dropout / row flags of that variant (all guarded at run time), not in the KIND of aux operand
    if ((e & kEpiResidual) && (epi & kEpiReluMask) || (e & kEpiReluMask) && (epi & kEpiResidual)) {
      set_error("gemm group: residual and ReLU-mask epilogues cannot share a launch");
      return MFP_ERR_ARG;
    }
    epi |= e;
    wide = wide || calls[i].N > 128;
  }
  return launch_tcgen05_any(cache, calls, n, epi, wide, wide && pair, stream);
}

}  // namespace mfp
