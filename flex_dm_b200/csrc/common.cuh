// Shared device/host definitions of the MFP engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flexdm_mfp.h"

namespace mfp {

constexpr int kD = 256;        // latent_dim of every BASELINE config (args.py:29-33)
constexpr int kH = 8;          // transformer.py:43
constexpr int kDh = 32;        // kD / kH
constexpr int kF = 512;        // FFN hidden = 2*emb_size (transformer.py:164)
constexpr int kMaxFields = MFP_MAX_FIELDS;
constexpr int kMaxLookups = 96;

// masking.py:8-15
constexpr float kMaskValue = 10.0f;
constexpr float kNullValue = 0.0f;
constexpr float kMaskProb = 0.15f;
constexpr float kChangeProb = 0.9f;                      // 1 - UNCHANGE_PROB, evaluated in double like Python
constexpr float kThresh = (float)(0.1 / (1.0 - 0.1));    // REPLACE_PROB / CHANGE_PROB
// Keras defaults
constexpr float kLnEps = 1e-3f;
constexpr float kCeEps = 1e-7f;
constexpr float kAdamB1 = 0.9f, kAdamB2 = 0.999f, kAdamEps = 1e-7f;

// RNG contract; mirrored by oracle/philox.py
constexpr uint32_t kStreamRandomU = 0;
constexpr uint32_t kStreamRandomCat = 1;
constexpr uint32_t kStreamRandomNum = 16;
constexpr uint32_t kFieldElem = 1000;
constexpr uint32_t kFieldTask = 1001;
constexpr uint32_t kFieldShuffle = 1002;   // shuffle_inputs: per-element sort key
constexpr uint32_t kSitePosDropout = 1999;  // Dropout of the PositionEmbedding (input_dtype != "set")
constexpr uint32_t kSiteDropout = 2000;

struct FieldDev {
  int kind;        // 0 categorical, 1 numerical
  int C;           // sub-targets (categorical) or vector width (numerical)
  int input_dim;   // categorical vocabulary size
  int logit_off;   // first column of this field in the logits matrix [T, LW]
  int logit_w;     // C*input_dim or C
  int task_id;     // task index of the attribute group holding this field (spec.py:364-377)
  int has_cond;    // loss_condition present (crello-spec.yml:88-121)
  int num_slot;    // index among numerical fields, -1 otherwise
  int grow_off;    // first row of this field in the encoder's gradient-row space (encoder.cu, embed_onehot_kernel)
  unsigned long long cond_mask;  // bit i: elements of type i carry this field
  long long table_off;   // params offset: embedding table (categorical) or 2-row special table (numerical)
  long long kernel_off;  // numerical Dense kernel [C, D]
  long long bias_off;    // numerical Dense bias [D]
};

struct Schema {
  int F;           // sequence fields, get_valid_input_columns order
  int type_field;  // index of "type"
  int LW;          // logits row pitch: heads start at multiples of 4 columns, the row is padded to 32 columns (128 B) so that every
                   // row of the [T, LW] logits / gradient matrices is sector- and TMA-box-aligned
  int LWu;         // columns in use (sum of the heads' 4-padded widths); [LWu, LW) is zero padding
  int n_num;       // numerical fields
  int sort_field[5];  // indices of type,left,top,width,height (tensor_utils.py:11)
  int R, Rp;          // gradient rows of the encoder (tables + {<MASK>, <UNUSED>, bias} per numerical field); Rp = R padded to 4
  int n_lookups;      // embedding rows summed per element: one per categorical sub-target + one per numerical field
  unsigned char lk_field[kMaxLookups], lk_sub[kMaxLookups];
  FieldDev f[kMaxFields];
};

struct BatchPtrs {
  const int* length;             // [B] zero-based
  const void* cols[kMaxFields];  // int32 [T,C] or float [T,C]
  // Packed numerical columns (mfp_set_packed_rows): rowmap[f] != null means cols[f] is float [n_rows, C] holding only the elements that
  // carry the field (valid position and type gate), and rowmap[f][t] is element t's row or -1 -- a missing row reads as <UNUSED>, which
  // is what filter_padding (masking.py:24-53) turns those elements into anyway.
  const int* rowmap[kMaxFields];
};

struct MaskPtrs {
  const unsigned char* m[kMaxFields];  // [T] each
};

// ------------------------------------------------------------------------------------------------- programmatic dependent launch
// Every kernel of the step is launched with the programmatic-stream-serialization attribute: its CTAs may be scheduled while
// the previous kernel of the stream is still draining, which hides the launch latency between the ~100 short kernels of a
// step.  pdl_wait() (griddepcontrol.wait) blocks until the previous kernel has completed and its writes are visible; it is the
// first statement of every kernel, or follows the barrier / TMEM set-up of the persistent ones.  FLEXDM_PDL=0 turns it off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ... as clusters of `cluster_x` CTAs (grid.x a multiple of it); cluster_x == 1: plain launch_pdl
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)na;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ------------------------------------------------------------------------------------------------- Philox4x32-10
struct U4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }  // 2^-24
__host__ __device__ __forceinline__ uint32_t mulhi_range(uint32_t x, uint32_t n) { return (uint32_t)(((uint64_t)x * n) >> 32); }

// ------------------------------------------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
#define MFP_CUDA_OK(expr)                                                                         \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      mfp::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      return MFP_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)
#define MFP_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != MFP_OK) return _r; \
  } while (0)

}  // namespace mfp
