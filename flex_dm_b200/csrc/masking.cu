// Input corruption of the MFP train step: preprocess_for_train / preprocess_for_test
// (reference: src/mfp/mfp/models/mfp.py:72-138, models/masking.py:24-155,227-269).
// Every field of an element is produced in one launch, and only the variant the document's task id selects is generated (the reference
// materialises all variants, then tf.where-selects).
#include "kernels.cuh"

namespace mfp {

__global__ void sample_tasks_kernel(TaskSet allowed, int B, uint32_t seed, uint32_t step, uint32_t doc0, int* __restrict__ tasks) {
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const U4 r = philox4x32_10((uint32_t)b + doc0, kFieldTask, 0u, 0u, seed, step);
  tasks[b] = allowed.ids[mulhi_range(r.x, (uint32_t)allowed.n)];
}

__device__ __forceinline__ float2 box_muller(uint32_t xa, uint32_t xb) {
  const float u1 = ((float)(xa >> 8) + 1.0f) * 5.9604644775390625e-08f;
  const float u2 = u01(xb);
  const float r = sqrtf(-2.0f * logf(u1));
  const float t = 6.283185307179586f * u2;
  return make_float2(r * cosf(t), r * sinf(t));
}

// mode 0: train (tasks != null), mode 1: test (test_masks given)
// A CTA takes kMcElems consecutive elements.  Pass 0: one thread per element gathers what every field of it needs (validity, type, task,
// elem_masking's pick).  Pass 1: one thread per (field, element), elements fastest -- a warp works on ONE field of 32 consecutive
// elements, so the field's descriptor and kind are warp-uniform and the categorical loads / stores are coalesced; it writes the masks and
// the categorical columns and leaves (action, source row) of the numerical fields in shared memory.  Pass 2: one warp per numerical row
// (2 KB for the 512-float embeddings): copy / <MASK> / noise / <UNUSED>, and the encoder's by-value special-token flag of what was written.
// (The first version ran one warp per element through all ten fields: 1 400 warp instructions per element, two thirds of them control
// flow executed for one to three active lanes -- issue-bound at 72 us for 190 MB.)
constexpr int kMcElems = 64;
constexpr int kMcThreads = 256;
constexpr int kMcMaxNum = 4;  // numerical fields per spec (crello: 2)
__global__ void __launch_bounds__(kMcThreads, 4) mask_corrupt_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs in,
                                                                    const int* __restrict__ tasks, const __grid_constant__ MaskPtrs test_masks, int mode, int B,
                                                                    int S, uint32_t seed, uint32_t step, uint32_t doc0, const __grid_constant__ ModifiedPtrs out,
                                                                    unsigned char* __restrict__ flags) {
  pdl_wait();
  __shared__ int s_type[kMcElems], s_task[kMcElems], s_row[kMcMaxNum][kMcElems];
  __shared__ unsigned char s_valid[kMcElems], s_pick[kMcElems], s_act[kMcMaxNum][kMcElems];
  const int T = B * S;
  const int e0 = blockIdx.x * kMcElems;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- pass 0: per element
  if (tid < kMcElems) {
    const int t = e0 + tid;
    int type_val = 0, task = -1;
    bool valid = false, pick = false;
    if (t < T) {
      const int b = t / S, s = t - b * S;
      const int n_valid = in.length[b] + 1;  // mask.py:28-29
      valid = s < n_valid;
      type_val = reinterpret_cast<const int*>(in.cols[sc.type_field])[(size_t)t * sc.f[sc.type_field].C];
      task = (mode == 0) ? tasks[b] : -1;
      if (task == 1) {  // masking.py:98-113
        const U4 r = philox4x32_10((uint32_t)b + doc0, kFieldElem, 0u, 0u, seed, step);
        pick = s == (int)(u01(r.x) * (float)n_valid);
      }
    }
    s_type[tid] = type_val; s_task[tid] = task; s_valid[tid] = valid ? 1 : 0; s_pick[tid] = pick ? 1 : 0;
  }
  __syncthreads();
  // ---- pass 1: per (field, element)
  for (int idx = tid; idx < sc.F * kMcElems; idx += kMcThreads) {
    const int f = idx / kMcElems, e = idx - f * kMcElems;
    const int t = e0 + e;
    if (t >= T) continue;
    const FieldDev& fd = sc.f[f];
    const uint32_t gt = (uint32_t)t + doc0 * (uint32_t)S;  // global element index: the Philox counters of a sharded batch are the single-process ones
    const bool valid = s_valid[e] != 0;
    const int task = s_task[e];
    // filter_padding (masking.py:24-53)
    bool unused = !valid || (fd.has_cond && !((fd.cond_mask >> s_type[e]) & 1ull));
    int src_row = t;
    if (fd.kind == 1 && in.rowmap[f]) {  // packed column: element t's row, or none (<UNUSED>)
      const int pr = __ldg(in.rowmap[f] + t);
      unused = unused || pr < 0;
      src_row = pr < 0 ? 0 : pr;
    }
    int action = 0;  // 0 keep filtered, 1 <MASK>, 2 random token
    bool mfp = false;
    if (mode == 1) {
      mfp = test_masks.m[f][t] != 0;
      action = mfp ? 1 : 0;
    } else if (task == 0) {  // random_masking (masking.py:227-269): three uniforms per (element, field)
      const U4 rf = philox4x32_10(gt, (uint32_t)f, kStreamRandomU, 0u, seed, step);
      mfp = valid && (u01(rf.x) < kMaskProb);
      const bool chg = mfp && (u01(rf.y) < kChangeProb);
      if (chg) action = (u01(rf.z) >= kThresh) ? 1 : 2;
    } else if (task == 1) {  // elem_masking (masking.py:136-155)
      mfp = s_pick[e] != 0;
      action = mfp ? 1 : 0;
    } else {  // feat_masking of one attribute group (masking.py:116-133)
      mfp = valid && (fd.task_id == task);
      action = mfp ? 1 : 0;
    }
    if (mode == 0) out.masks[f][t] = mfp ? 1 : 0;
    if (fd.kind == 0) {
      const int* src = reinterpret_cast<const int*>(in.cols[f]) + (size_t)t * fd.C;
      int* dst = reinterpret_cast<int*>(out.cols[f]) + (size_t)t * fd.C;
      for (int c = 0; c < fd.C; ++c) {
        int v = unused ? fd.input_dim + 1 : src[c];
        if (action == 1) v = fd.input_dim;
        if (action == 2) {
          const U4 r = philox4x32_10(gt, (uint32_t)f, kStreamRandomCat + (uint32_t)c, 0u, seed, step);
          v = (int)mulhi_range(r.x, (uint32_t)fd.input_dim);
        }
        dst[c] = v;
      }
    } else {
      s_act[fd.num_slot][e] = (unsigned char)(action | (unused ? 4 : 0));
      s_row[fd.num_slot][e] = src_row;
    }
  }
  if (sc.n_num == 0) return;
  __syncthreads();
  // ---- pass 2: one warp per numerical row
  for (int r = warp; r < sc.n_num * kMcElems; r += kMcThreads / 32) {
    const int slot = r / kMcElems, e = r - slot * kMcElems;
    const int t = e0 + e;
    if (t >= T) continue;
    int f = 0;
    for (int k = 0; k < sc.F; ++k)
      if (sc.f[k].kind == 1 && sc.f[k].num_slot == slot) f = k;
    const FieldDev& fd = sc.f[f];
    const uint32_t gt = (uint32_t)t + doc0 * (uint32_t)S;
    const int action = s_act[slot][e] & 3;
    const bool unused = (s_act[slot][e] & 4) != 0;
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in.cols[f]) + (size_t)s_row[slot][e] * fd.C);
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out.cols[f]) + (size_t)t * fd.C);
    bool all_mask = true, all_null = true;  // the encoder's by-value special-token test (row_flags_kernel), on what is written
    int q_begin = lane;
    if (action == 0 && !unused) {
      // plain copy (the common case): four 16-byte loads in flight per lane before the stores
      for (; q_begin + 96 < fd.C / 4; q_begin += 128) {
        const float4 v0 = src[q_begin], v1 = src[q_begin + 32], v2 = src[q_begin + 64], v3 = src[q_begin + 96];
        dst[q_begin] = v0; dst[q_begin + 32] = v1; dst[q_begin + 64] = v2; dst[q_begin + 96] = v3;
        const float4 vs[4] = {v0, v1, v2, v3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          all_mask = all_mask && vs[k].x == kMaskValue && vs[k].y == kMaskValue && vs[k].z == kMaskValue && vs[k].w == kMaskValue;
          all_null = all_null && vs[k].x == kNullValue && vs[k].y == kNullValue && vs[k].z == kNullValue && vs[k].w == kNullValue;
        }
      }
    }
    for (int q = q_begin; q < fd.C / 4; q += 32) {
      float4 v;
      if (action == 1) {
        v = make_float4(kMaskValue, kMaskValue, kMaskValue, kMaskValue);
      } else if (action == 2) {
        const U4 rr = philox4x32_10(gt, (uint32_t)f, kStreamRandomNum + (uint32_t)q, 0u, seed, step);
        const float2 z0 = box_muller(rr.x, rr.y), z1 = box_muller(rr.z, rr.w);
        v = make_float4(z0.x * 0.1f, z0.y * 0.1f, z1.x * 0.1f, z1.y * 0.1f);  // stddev 0.1, masking.py:91
      } else if (unused) {
        v = make_float4(kNullValue, kNullValue, kNullValue, kNullValue);
      } else {
        v = src[q];
      }
      dst[q] = v;
      all_mask = all_mask && v.x == kMaskValue && v.y == kMaskValue && v.z == kMaskValue && v.w == kMaskValue;
      all_null = all_null && v.x == kNullValue && v.y == kNullValue && v.z == kNullValue && v.w == kNullValue;
    }
    if (flags) {
      all_mask = __all_sync(0xffffffffu, all_mask);
      all_null = __all_sync(0xffffffffu, all_null);
      if (lane == 0) flags[(size_t)fd.num_slot * B * S + t] = all_null ? 2 : (all_mask ? 1 : 0);
    }
  }
}

// Encoder special-token detection by value (encoder.py:165-166): 1 = all == MASK_VALUE, 2 = all == NULL_VALUE.
__global__ void __launch_bounds__(256) row_flags_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs mod, int T,
                                                        unsigned char* __restrict__ flags) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  for (int f = 0; f < sc.F; ++f) {
    if (sc.f[f].kind != 1) continue;
    const int C = sc.f[f].C;
    const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(mod.cols[f]) + (size_t)t * C);
    bool all_mask = true, all_null = true;
    for (int q = lane; q < C / 4; q += 32) {
      const float4 v = src[q];
      all_mask = all_mask && v.x == kMaskValue && v.y == kMaskValue && v.z == kMaskValue && v.w == kMaskValue;
      all_null = all_null && v.x == kNullValue && v.y == kNullValue && v.z == kNullValue && v.w == kNullValue;
    }
    all_mask = __all_sync(0xffffffffu, all_mask);
    all_null = __all_sync(0xffffffffu, all_null);
    if (lane == 0) flags[(size_t)sc.f[f].num_slot * T + t] = all_null ? 2 : (all_mask ? 1 : 0);  // is_unused is applied last (encoder.py:174-175)
  }
}

// shuffle_inputs (tensor_utils.py:47-76): grid = documents.  perm[b][r] = source position of output position r: the valid positions
// ordered by their Philox key (ties by position), padding in place.
__global__ void __launch_bounds__(128) shuffle_perm_kernel(const int* __restrict__ length, int S, uint32_t seed, uint32_t step, uint32_t doc0, int* __restrict__ perm) {
  pdl_wait();
  extern __shared__ uint32_t skeys[];  // [S]
  const int b = blockIdx.x;
  const int n = min(S, length[b] + 1);
  for (int s = threadIdx.x; s < n; s += blockDim.x) skeys[s] = philox4x32_10((uint32_t)(b * S + s) + doc0 * (uint32_t)S, kFieldShuffle, 0u, 0u, seed, step).x;
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    if (s >= n) { perm[b * S + s] = s; continue; }
    const uint32_t k = skeys[s];
    int r = 0;
    for (int j = 0; j < n; ++j) r += (skeys[j] < k) || (skeys[j] == k && j < s);
    perm[b * S + r] = s;
  }
}

// sort_inputs (tensor_utils.py:14-44, --input_dtype sorted_set): elements ordered by the base-100 digits of (type, left, top, width, height)[..., 0],
// padding last (key + 100^5), stable.  grid = documents.
__global__ void __launch_bounds__(128) sort_perm_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs in, int S, int* __restrict__ perm) {
  pdl_wait();
  extern __shared__ long long lkeys[];  // [S]
  const int b = blockIdx.x;
  const int n = in.length[b] + 1;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const size_t t = (size_t)b * S + s;
    long long k = 0;
    for (int i = 0; i < 5; ++i) k = k * 100 + reinterpret_cast<const int*>(in.cols[sc.sort_field[i]])[t * sc.f[sc.sort_field[i]].C];
    lkeys[s] = k + ((s >= n) ? 10000000000LL : 0LL);
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long k = lkeys[s];
    int r = 0;
    for (int j = 0; j < S; ++j) r += (lkeys[j] < k) || (lkeys[j] == k && j < s);
    perm[b * S + r] = s;
  }
}

// one warp per output element: every sequence column gathered through the permutation
__global__ void __launch_bounds__(256) gather_columns_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs in,
                                                             const int* __restrict__ perm, int B, int S, const __grid_constant__ ModifiedPtrs out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= B * S) return;
  const int b = t / S;
  const size_t src_t = (size_t)b * S + perm[t];
  for (int f = 0; f < sc.F; ++f) {
    const FieldDev& fd = sc.f[f];
    if (fd.kind == 0) {
      if (lane < fd.C) reinterpret_cast<int*>(out.cols[f])[(size_t)t * fd.C + lane] = reinterpret_cast<const int*>(in.cols[f])[src_t * fd.C + lane];
    } else {
      const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in.cols[f]) + src_t * fd.C);
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(out.cols[f]) + (size_t)t * fd.C);
      for (int q = lane; q < fd.C / 4; q += 32) dst[q] = src[q];
    }
  }
}

int launch_shuffle_inputs(const Schema& sc, const BatchPtrs& in, int B, int S, uint32_t seed, uint32_t step, int* perm, const ModifiedPtrs& out,
                          cudaStream_t st, bool sorted, uint32_t doc0) {
  if (sorted) MFP_CUDA_OK(launch_pdl(sort_perm_kernel, B, 128, S * sizeof(long long), st, sc, in, S, perm));
  else MFP_CUDA_OK(launch_pdl(shuffle_perm_kernel, B, 128, S * sizeof(uint32_t), st, in.length, S, seed, step, doc0, perm));
  MFP_CUDA_OK(cudaGetLastError());
  const int T = B * S;
  MFP_CUDA_OK(launch_pdl(gather_columns_kernel, (T + 7) / 8, 256, 0, st, sc, in, perm, B, S, out));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_sample_tasks(const TaskSet& allowed, int B, uint32_t seed, uint32_t step, int* tasks, cudaStream_t st, uint32_t doc0) {
  MFP_CUDA_OK(launch_pdl(sample_tasks_kernel, (B + 127) / 128, 128, 0, st, allowed, B, seed, step, doc0, tasks));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_mask_corrupt(const Schema& sc, const BatchPtrs& in, const int* tasks, const MaskPtrs* test_masks, int B, int S, uint32_t seed,
                        uint32_t step, const ModifiedPtrs& out, cudaStream_t st, unsigned char* flags, uint32_t doc0) {
  MaskPtrs tm{};
  if (test_masks) tm = *test_masks;
  const int T = B * S;
  if (sc.n_num > kMcMaxNum) { set_error("mask_corrupt: at most %d numerical fields", kMcMaxNum); return MFP_ERR_UNSUPPORTED; }
  MFP_CUDA_OK(launch_pdl(mask_corrupt_kernel, (T + kMcElems - 1) / kMcElems, kMcThreads, 0, st, sc, in, tasks, tm, test_masks ? 1 : 0, B, S, seed, step, doc0, out, flags));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_row_flags(const Schema& sc, const BatchPtrs& mod, int T, unsigned char* flags, cudaStream_t st) {
  if (sc.n_num == 0) return MFP_OK;
  MFP_CUDA_OK(launch_pdl(row_flags_kernel, (T + 7) / 8, 256, 0, st, sc, mod, T, flags));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
