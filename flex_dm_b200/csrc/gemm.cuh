// TF32 GEMM of the MFP engine: D[M,N] (+)= A[M,K] . B[N,K]^T with a fused epilogue.
// Persistent tcgen05.mma kernel (kind::tf32, cta_group::1): operands staged by TMA (SWIZZLE_128B) through an mbarrier
// ring, two TMEM accumulators so the epilogue of one tile overlaps the main loop of the next, epilogue through
// swizzled shared staging and TMA store / reduce-add.  See gemm.cu for the warp roles.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mfp {

// Operand view. K-major: memory is [rows = M|N][cols = K] row-major with pitch ld (floats).
//               MN-major: memory is [rows = K][cols = M|N] row-major with pitch ld.
struct GemmOperand {
  const float* ptr;
  int mn_major;
  int ld;
};

// Fused epilogue, applied per element in this order:
//   v = acc; v += bias[col]; v = relu(v); v *= (relu_src[row,col] > 0); v = dropout(v);
//   if (rowflag[row]) v = 0; v += residual[row,col]; out[row,col] = v   (or atomicAdd when atomic != 0)
// ReLU gates as bits (tcgen05 path): a forward GEMM with relu_bits_out also writes, per row and 32-column chunk, the word whose bit k says
// out[row, 32 chunk + k] > 0 -- layout [N / 32][relu_bits_ld] uint32, rows contiguous, so a warp of 32 rows moves one 128-byte line --
// and the input-gradient GEMM through that ReLU takes relu_bits instead of relu_src: 1/32 of the bytes, no aux operand in its epilogue.
struct GemmEpilogue {
  float* out;
  int ldo;
  const float* bias;
  const float* residual;
  int ldr;
  const float* relu_src;
  int ld_relu;
  const unsigned char* rowflag;
  int relu;
  int atomic;
  int drop_enabled;
  float drop_rate;
  uint32_t drop_seed, drop_step, drop_site;
  uint32_t drop_row0;  // global index of row 0 (data-parallel shards): dropout counters are formed from global rows
  // Fused LayerNorm of the output rows (N == 256: one tile holds whole rows): besides out = v, the epilogue writes
  // ln_out = (v - mean) * rstd * ln_gamma + ln_beta and the row statistics the LayerNorm backward needs (eps = 1e-3, biased variance:
  // transformer.py:216,222 via Keras LayerNormalization).  ln_out == nullptr: off.
  float* ln_out;
  int ln_ldo;
  const float* ln_gamma;
  const float* ln_beta;
  float* ln_mean;
  float* ln_rstd;
  uint32_t* relu_bits_out;    // forward: gate words of the rows / chunks this GEMM writes (nullptr: off)
  const uint32_t* relu_bits;  // backward: v *= bit(row, col) (instead of relu_src)
  int relu_bits_ld;           // rows per chunk plane of either (>= M, multiple of 32)
};

inline GemmEpilogue make_epilogue(float* out, int ldo) {
  GemmEpilogue e{};
  e.out = out;
  e.ldo = ldo;
  return e;
}

struct GemmCall {
  GemmOperand a, b;
  int M, N, K;
  int splits;  // split-K factor (>1 forces atomic accumulation into a zeroed/pre-filled output)
  GemmEpilogue ep;
  float* colsum;  // optional [N]: += column sums of the (MN-major) B operand over K, i.e. the bias gradient of a wgrad GEMM
  // Deterministic reductions (mfp_set_deterministic): split-K partial tiles and the column-sum partials of every CTA go to this scratch
  // block with plain stores and are summed in a fixed order by a second kernel, instead of TMA reduce-add / atomicAdd in arrival order.
  // 3xTF32 (mfp_set_gemm_impl(h, 2)): x_lo = x - tf32(x) of both operands, same layout and pitch as a / b (split_tf32_lo); both or neither
  const float* a_lo;
  const float* b_lo;
  float* det_ws;        // nullptr = arrival-order accumulation (the fast default)
  size_t det_ws_floats;
};

constexpr size_t kDetWsFloats = (size_t)148 * 128 * 256 + (size_t)148 * 2048;  // one wave of 128 x 256 partial tiles + column-sum partials

// TMA descriptors, cached per (pointer, geometry): building one costs a driver call.
enum MapKind : uint32_t { kMapOperandK = 0, kMapOperandMN = 1, kMapEpilogue = 2 };
class TensorMapCache;
TensorMapCache* tensor_map_cache_create();
void tensor_map_cache_destroy(TensorMapCache*);
void tensor_map_cache_trim(TensorMapCache*);  // call at the start of a launcher, before its first tensor_map_get (pointers from earlier calls die)
// fp32 tensor of rank 2 or 3: dims[0] innermost (contiguous), strides (floats) of dims 1.., box per dim; see gemm.cu for the kinds
const CUtensorMap* tensor_map_get(TensorMapCache* cache, const float* ptr, int rank, const uint64_t* dims, const uint64_t* strides, const uint32_t* box,
                                  MapKind kind);

// impl: 0 = tcgen05 (TF32, or 3xTF32 when call.a_lo / b_lo are given), 1 = SIMT bring-up kernel.  Returns MFP_OK or an error (message via set_error).
int launch_gemm(TensorMapCache* cache, const GemmCall& call, int impl, cudaStream_t stream);
// Up to three INDEPENDENT problems in one tcgen05 launch (a weight gradient next to the input gradient off the same dY, the encoder's
// weight gradients): their tiles form one tile space, so epilogues overlap the next problem's main loop and launch boundaries disappear.
int launch_gemm_group(TensorMapCache* cache, const GemmCall* calls, int n, cudaStream_t stream);
// split-K factor the engine asks for a weight gradient [M, N] over K tokens (one wave of the persistent kernel's tiles)
int gemm_wgrad_splits(int M, int N, int K);
// lo[i] = x[i] - tf32_rne(x[i]) over a [rows, cols] matrix of pitch ld (lo has the same pitch): the compensation operand of the 3xTF32 mode
int launch_split_tf32_lo(const float* x, int rows, int cols, int ld, float* lo, cudaStream_t stream);

// keep-mask/scale of the engine's dropout sites (shared by the GEMM epilogue and the backward pass).
// RNG contract: element e of the flattened [T, D] activation takes the 16-bit half (e & 7) of philox(counter = (e >> 3, site, 0, 0),
// key = (seed, step)) -- words x, y, z, w hold halves (0,1), (2,3), (4,5), (6,7), low half first -- and is kept iff that half is
// >= round(rate * 65536) (rate 0.1: 6554, i.e. P(drop) = 0.100006).  One Philox block serves eight elements: inside the GEMM
// epilogues the generator was 70 % of the instructions of the dropout variants (and those epilogues are issue-bound).
__device__ __forceinline__ uint32_t dropout_threshold(float rate) { return (uint32_t)(rate * 65536.0f + 0.5f); }
__device__ __forceinline__ void dropout_apply4(float (&v)[4], uint32_t wa, uint32_t wb, uint32_t thr, float scale) {
  v[0] = ((wa & 0xffffu) >= thr) ? v[0] * scale : 0.0f;
  v[1] = ((wa >> 16) >= thr) ? v[1] * scale : 0.0f;
  v[2] = ((wb & 0xffffu) >= thr) ? v[2] * scale : 0.0f;
  v[3] = ((wb >> 16) >= thr) ? v[3] * scale : 0.0f;
}
__device__ __forceinline__ void dropout8(float (&lo)[4], float (&hi)[4], uint32_t e0, float rate, uint32_t seed, uint32_t step, uint32_t site) {
  // e0 = index of lo[0] in the flattened [T, D] activation, multiple of 8; hi = the next four elements
  const U4 r = philox4x32_10(e0 >> 3, site, 0u, 0u, seed, step);
  const uint32_t thr = dropout_threshold(rate);
  const float scale = 1.0f / (1.0f - rate);
  dropout_apply4(lo, r.x, r.y, thr, scale);
  dropout_apply4(hi, r.z, r.w, thr, scale);
}
__device__ __forceinline__ void dropout4(float (&v)[4], uint32_t e0, float rate, uint32_t seed, uint32_t step, uint32_t site) {
  // e0 multiple of 4: one half of the eight-element block (sites outside the hot loops)
  const U4 r = philox4x32_10(e0 >> 3, site, 0u, 0u, seed, step);
  dropout_apply4(v, (e0 & 4u) ? r.z : r.x, (e0 & 4u) ? r.w : r.y, dropout_threshold(rate), 1.0f / (1.0f - rate));
}

}  // namespace mfp
