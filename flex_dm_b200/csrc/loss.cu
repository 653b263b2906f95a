// LossLayer of the MFP step (reference: models/metrics.py:36-57,173-299; sort branch tensor_utils.py:14-44) fused
// with its gradient: one warp per element position computes, for every field, the masked / type-gated /
// length-gated loss, score numerator and denominator, and writes d(loss)/d(logits) in the same pass.
// Reductions are deterministic: per-(field, position) partials, then a fixed-order tree sum.
#include "kernels.cuh"

namespace mfp {

constexpr int kMaxVPerLane = 8;  // categorical vocabulary <= 256 per sub-target

// ------------------------------------------------------------------------------------------------- rico sort branch
// key = base-100 digits of (type,left,top,width,height)[...,0] (+100^5 when padded); stable rank by counting.
__global__ void __launch_bounds__(128) sort_indices_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs targets,
                                                           const float* __restrict__ logits, const unsigned char* __restrict__ sort_flag,
                                                           const int* __restrict__ sort_tasks, int pos_task_id, int S,
                                                           int* __restrict__ idx_true, int* __restrict__ idx_pred) {
  pdl_wait();
  extern __shared__ long long keys[];  // [2][S]
  const int b = blockIdx.x;
  const bool sorted = sort_flag ? (sort_flag[b] != 0) : (sort_tasks && sort_tasks[b] == pos_task_id);
  if (!sorted) {
    for (int s = threadIdx.x; s < S; s += blockDim.x) { idx_true[b * S + s] = s; idx_pred[b * S + s] = s; }
    return;
  }
  const int n = targets.length[b] + 1;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const size_t t = (size_t)b * S + s;
    long long kt = 0, kp = 0;
    for (int i = 0; i < 5; ++i) {
      const FieldDev& fd = sc.f[sc.sort_field[i]];
      kt = kt * 100 + reinterpret_cast<const int*>(targets.cols[sc.sort_field[i]])[t * fd.C];
      const float* lg = logits + t * sc.LW + fd.logit_off;  // sub-target 0
      int arg = 0;
      float best = lg[0];
      for (int v = 1; v < fd.input_dim; ++v)
        if (lg[v] > best) { best = lg[v]; arg = v; }
      kp = kp * 100 + arg;
    }
    const long long pad = (s >= n) ? 10000000000LL : 0LL;  // 100^5
    keys[s] = kt + pad;
    keys[S + s] = kp + pad;
  }
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long kt = keys[s], kp = keys[S + s];
    int rt = 0, rp = 0;
    for (int j = 0; j < S; ++j) {
      const long long a = keys[j], c = keys[S + j];
      rt += (a < kt) || (a == kt && j < s);
      rp += (c < kp) || (c == kp && j < s);
    }
    idx_true[b * S + rt] = s;
    idx_pred[b * S + rp] = s;
  }
}

// ------------------------------------------------------------------------------------------------- fused loss + gradient
// Masked categorical sub-target of one element (one warp): softmax cross-entropy on clipped probabilities (Keras),
// accuracy, and -- when drow is given -- the gradient.  NK = vocabulary entries per lane (ceil(V / 32)): the loops are unrolled for the
// vocabulary at hand instead of for the largest one (crello's are 6-64 wide: one or two entries per lane, not eight; the kernel is
// issue-bound).  Skipped entries would only have added 0.0f / compared -inf: the results are bit-identical.
template <int NK>
__device__ __forceinline__ void categorical_loss(const float* __restrict__ x, int V, int y, int lane, float inv_batch, float* __restrict__ dx, float& loss,
                                           float& score, float& den) {
  constexpr int kMaxVPerLane = NK;
  float xv[kMaxVPerLane];
  float mx = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < kMaxVPerLane; ++k) {
    const int v = lane + 32 * k;
    xv[k] = (v < V) ? x[v] : -INFINITY;
    if (xv[k] > mx) { mx = xv[k]; arg = v; }
  }
  // warp argmax, first index on ties (tf.argmax)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVPerLane; ++k) {
    xv[k] = (lane + 32 * k < V) ? expf(xv[k] - mx) : 0.f;
    sum += xv[k];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  // A6: p clipped to [eps, 1-eps], log, softmax-CE on the logs: -log c_y + log sum_j c_j
  float Z = 0.f, cy = 0.f, py = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVPerLane; ++k) {
    const int v = lane + 32 * k;
    const float p = xv[k] * inv;
    xv[k] = p;
    if (v < V) {
      const float cl = fminf(fmaxf(p, kCeEps), 1.0f - kCeEps);
      Z += cl;
      if (v == y) { cy = cl; py = p; }
    }
  }
  Z = warp_sum(Z);
  cy = warp_sum(cy);
  py = warp_sum(py);
  loss += -logf(cy) + logf(Z);
  score += (arg == y) ? 1.f : 0.f;
  den += 1.f;
  if (dx) {
    // g_v = dL/dp_v = m_v (-[v==y]/c_y + 1/Z), m_v = 1 inside the clip range; dx_v = p_v (g_v - sum_j g_j p_j)
    const float invZ = 1.0f / Z;
    float dot = 0.f;
    float g[kMaxVPerLane];
#pragma unroll
    for (int k = 0; k < kMaxVPerLane; ++k) {
      const int v = lane + 32 * k;
      const float p = xv[k];
      const bool inside = (v < V) && p >= kCeEps && p <= 1.0f - kCeEps;
      g[k] = inside ? (invZ - ((v == y) ? 1.0f / cy : 0.f)) : 0.f;
      dot += g[k] * p;
    }
    dot = warp_sum(dot);
#pragma unroll
    for (int k = 0; k < kMaxVPerLane; ++k) {
      const int v = lane + 32 * k;
      if (v < V) dx[v] = xv[k] * (g[k] - dot) * inv_batch;
    }
  }
  (void)py;
}

__global__ void __launch_bounds__(256) loss_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs targets,
                                                   const __grid_constant__ MaskPtrs masks, const float* __restrict__ logits, int use_sort, int B, int S,
                                                   float inv_batch, float* __restrict__ dlogits, const __grid_constant__ LossBuffers buf) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int T = B * S;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int b = t / S, s = t - b * S;
  const size_t tt = use_sort ? (size_t)b * S + buf.idx_true[t] : (size_t)t;  // row of the (sorted) target
  const size_t tp = use_sort ? (size_t)b * S + buf.idx_pred[t] : (size_t)t;  // row of the (sorted) prediction
  const bool valid = s <= targets.length[b];
  const int type_true = reinterpret_cast<const int*>(targets.cols[sc.type_field])[tt * sc.f[sc.type_field].C];
  const float* lrow = logits + tp * sc.LW;
  float* drow = dlogits ? dlogits + tp * sc.LW : nullptr;
  if (drow) {
    // The gradient row is zero except under the (15 % or so of) fields this element is masked in: clear it once with 16-byte
    // stores (LW is a multiple of 32 floats), then only the active fields write their entries.
    float4* d4 = reinterpret_cast<float4*>(drow);
    for (int c = lane; c < sc.LW / 4; c += 32) d4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
  }
  // the element's per-field weights in one go: lane f loads field f's mask byte (independent loads; one per field inside the loop below
  // put ten dependent L2 latencies on every warp's critical path)
  bool my_w = false;
  if (lane < sc.F) {
    const FieldDev& fl = sc.f[lane];
    // metrics.py:251-267: mfp mask (unsorted position) x type gate (sorted target) x seq mask
    my_w = valid && masks.m[lane][t] && (!fl.has_cond || ((fl.cond_mask >> type_true) & 1ull));
  }
  const unsigned wbits = __ballot_sync(0xffffffffu, my_w);
  float my_loss = 0.f, my_score = 0.f, my_den = 0.f;
  for (int f = 0; f < sc.F; ++f) {
    const FieldDev& fd = sc.f[f];
    const bool w = (wbits >> f) & 1u;
    float loss = 0.f, score = 0.f, den = 0.f;
    if (!w) {
      // nothing to add: the row was cleared above
    } else if (fd.kind == 0) {
      const int V = fd.input_dim;
      for (int c = 0; c < fd.C; ++c) {
        const int y = reinterpret_cast<const int*>(targets.cols[f])[tt * fd.C + c];
        const float* x = lrow + fd.logit_off + c * V;
        float* dx = drow ? drow + fd.logit_off + c * V : nullptr;
        if (V <= 32) categorical_loss<1>(x, V, y, lane, inv_batch, dx, loss, score, den);
        else if (V <= 64) categorical_loss<2>(x, V, y, lane, inv_batch, dx, loss, score, den);
        else if (V <= 128) categorical_loss<4>(x, V, y, lane, inv_batch, dx, loss, score, den);
        else categorical_loss<kMaxVPerLane>(x, V, y, lane, inv_batch, dx, loss, score, den);
      }
    } else {
      // metrics.py:52-57,246-248: sum_d (yhat - y)^2; score = 0.5 cos + 0.5 (l2_normalize eps 1e-12)
      // packed target column (mfp_set_packed_rows): w implies a valid element that carries the field, i.e. a row exists
      const size_t yrow = targets.rowmap[f] ? (size_t)max(__ldg(targets.rowmap[f] + tt), 0) : tt;
      const float* y = reinterpret_cast<const float*>(targets.cols[f]) + yrow * fd.C;
      const float* x = lrow + fd.logit_off;
      float sq = 0.f, yy = 0.f, xx = 0.f, xy = 0.f;
      // 16-byte accesses (logit_off and C are multiples of 4; rows are 128-byte aligned), four independent load pairs in flight per lane
      const float4* x4 = reinterpret_cast<const float4*>(x);
      const float4* y4 = reinterpret_cast<const float4*>(y);
      float4* d4 = drow ? reinterpret_cast<float4*>(drow + fd.logit_off) : nullptr;
      const float g2 = 2.0f * inv_batch;
      for (int c0 = 0; c0 < fd.C / 4; c0 += 128) {
        float4 xv4[4], yv4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + lane + 32 * u;
          xv4[u] = c < fd.C / 4 ? x4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          yv4[u] = c < fd.C / 4 ? y4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = c0 + lane + 32 * u;
          const float4 a = xv4[u], b = yv4[u];
          const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
          sq += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
          yy += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
          xx += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
          xy += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
          if (d4 && c < fd.C / 4) d4[c] = make_float4(g2 * d.x, g2 * d.y, g2 * d.z, g2 * d.w);
        }
      }
      sq = warp_sum(sq); yy = warp_sum(yy); xx = warp_sum(xx); xy = warp_sum(xy);
      loss = sq;
      score = 0.5f * xy * rsqrtf(fmaxf(yy, 1e-12f)) * rsqrtf(fmaxf(xx, 1e-12f)) + 0.5f;
      den = 1.f;
    }
    if (lane == f) { my_loss = loss; my_score = score; my_den = den; }  // warp-uniform values: lane f keeps field f's
  }
  if (lane < sc.F) {  // one store per quantity for all fields of the element (was three single-lane stores per field)
    buf.part[((size_t)0 * sc.F + lane) * T + t] = my_loss;
    buf.part[((size_t)1 * sc.F + lane) * T + t] = my_score;
    buf.part[((size_t)2 * sc.F + lane) * T + t] = my_den;
  }
}

// fixed-order reduction of one (quantity, field) row of partials; grid = (F, 3)
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ part, int F, int T, float inv_batch, float* __restrict__ metrics) {
  pdl_wait();
  __shared__ float red[1024];
  const int f = blockIdx.x, k = blockIdx.y;
  const float* p = part + ((size_t)k * F + f) * T;
  float acc = 0.f;
  for (int i = threadIdx.x; i < T; i += 1024) acc += p[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) metrics[f * 3 + k] = (k == 0) ? red[0] * inv_batch : red[0];  // loss: reduce_mean over the batch (metrics.py:277)
}

__global__ void loss_total_kernel(int F, float* __restrict__ metrics) {
  pdl_wait();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float total = 0.f;
    for (int f = 0; f < F; ++f) total += metrics[f * 3];  // metrics.py:291-297
    metrics[3 * F] = total;
  }
}

// merge_inputs_and_prediction (mfp.py:46-69) for one field
__global__ void __launch_bounds__(256) merge_prediction_kernel(const __grid_constant__ Schema sc, int f, const void* __restrict__ input_col,
                                                               const unsigned char* __restrict__ mask, const float* __restrict__ logits, int T,
                                                               float* __restrict__ out) {
  pdl_wait();
  const FieldDev& fd = sc.f[f];
  const size_t total = (size_t)T * fd.logit_w;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t t = i / fd.logit_w;
  const int c = (int)(i - t * fd.logit_w);
  const float pred = logits[t * sc.LW + fd.logit_off + c];
  float gt;
  if (fd.kind == 0) {
    const int sub = c / fd.input_dim, v = c - sub * fd.input_dim;
    gt = (reinterpret_cast<const int*>(input_col)[t * fd.C + sub] == v) ? 1.f : 0.f;
  } else {
    gt = reinterpret_cast<const float*>(input_col)[t * fd.C + c];
  }
  out[i] = mask[t] ? pred : gt;
}

int launch_sort_indices(const Schema& sc, const BatchPtrs& targets, const float* logits, const unsigned char* sort_flag, const int* sort_tasks,
                        int pos_task_id, int B, int S, const LossBuffers& buf, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(sort_indices_kernel, B, 128, 2 * S * sizeof(long long), st, sc, targets, logits, sort_flag, sort_tasks, pos_task_id, S, buf.idx_true, buf.idx_pred));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_loss(const Schema& sc, const BatchPtrs& targets, const MaskPtrs& masks, const float* logits, int use_sort, int B, int S, float inv_batch,
                float* dlogits, const LossBuffers& buf, float* metrics_out, cudaStream_t st) {
  const int T = B * S;
  MFP_CUDA_OK(launch_pdl(loss_kernel, (T + 7) / 8, 256, 0, st, sc, targets, masks, logits, use_sort, B, S, inv_batch, dlogits, buf));
  MFP_CUDA_OK(cudaGetLastError());
  MFP_CUDA_OK(launch_pdl(loss_reduce_kernel, dim3(sc.F, 3), 1024, 0, st, buf.part, sc.F, T, inv_batch, metrics_out));
  MFP_CUDA_OK(cudaGetLastError());
  MFP_CUDA_OK(launch_pdl(loss_total_kernel, 1, 32, 0, st, sc.F, metrics_out));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_merge_prediction(const Schema& sc, int field, const void* input_col, const unsigned char* mask, const float* logits, int T, float* out,
                            cudaStream_t st) {
  const size_t total = (size_t)T * sc.f[field].logit_w;
  MFP_CUDA_OK(launch_pdl(merge_prediction_kernel, (unsigned)((total + 255) / 256), 256, 0, st, sc, field, input_col, mask, logits, T, out));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
