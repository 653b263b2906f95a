// Adam(learning_rate, clipnorm=1.0) + Keras L2 regularisers (reference: train.py:71-77, architecture/utils.py:8-22;
// semantics SURVEY.md Appendix A3-A5): g <- g + 2*l2*w on regularised variables, per-VARIABLE clip_by_norm, then
// TF-form Adam (epsilon outside the bias-corrected sqrt).  Variables are strided 2-D views of the flat buffers.
#include "kernels.cuh"

namespace mfp {

constexpr int kNormChunks = 16;  // CTAs per variable in the norm pass; partials are summed in a fixed order (deterministic)

// grid = (V, kNormChunks): partial sums of (g + 2 l2 w)^2 and w^2 over one slice of one variable
__global__ void __launch_bounds__(256) var_norms_kernel(const VarDev* __restrict__ vars, const float* __restrict__ params,
                                                        const float* __restrict__ grads, float l2, int V, float* __restrict__ part) {
  pdl_wait();
  __shared__ float red[2][8];
  const VarDev v = vars[blockIdx.x];
  const int n = v.rows * v.cols;
  const float k = (v.l2 && l2 > 0.f) ? 2.0f * l2 : 0.f;
  const int per = (n + kNormChunks - 1) / kNormChunks;
  const int i0 = blockIdx.y * per, i1 = min(n, i0 + per);
  float sg = 0.f, sw = 0.f;
  if (((v.cols | v.ld) & 3) == 0 && (v.off & 3) == 0) {  // 16-byte path: whole float4s of a row (all Dense kernels, tables, biases of width % 4 == 0)
    const int c4 = v.cols >> 2, n4 = v.rows * c4;
    const int per4 = (n4 + kNormChunks - 1) / kNormChunks;
    const int j0 = blockIdx.y * per4, j1 = min(n4, j0 + per4);
    for (int j = j0 + threadIdx.x; j < j1; j += 256) {
      const int r = j / c4, c = (j - r * c4) << 2;
      const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
      const float4 w = *reinterpret_cast<const float4*>(params + idx);
      float4 g = grads ? *reinterpret_cast<const float4*>(grads + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
      g.x += k * w.x; g.y += k * w.y; g.z += k * w.z; g.w += k * w.w;
      sg += g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w;
      sw += w.x * w.x + w.y * w.y + w.z * w.z + w.w * w.w;
    }
  } else {
    for (int i = i0 + threadIdx.x; i < i1; i += 256) {
      const int r = i / v.cols, c = i - r * v.cols;
      const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
      const float w = params[idx];
      const float g = (grads ? grads[idx] : 0.f) + k * w;
      sg += g * g;
      sw += w * w;
    }
  }
  sg = warp_sum(sg);
  sw = warp_sum(sw);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sg; red[1][threadIdx.x >> 5] = sw; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; }
    part[(blockIdx.x * kNormChunks + blockIdx.y) * 2] = a;
    part[(blockIdx.x * kNormChunks + blockIdx.y) * 2 + 1] = (v.l2 && l2 > 0.f) ? b : 0.f;
  }
}

// one warp: l2 * sum over regularised variables of sum w^2 (fixed order)
__global__ void l2_loss_kernel(const float* __restrict__ part, int V, float l2, float* __restrict__ out) {
  pdl_wait();
  float s = 0.f;
  for (int i = threadIdx.x; i < V * kNormChunks; i += 32) s += part[2 * i + 1];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = (l2 > 0.f) ? l2 * s : 0.f;  // A3: l2 * sum(w^2), no 1/2
}

// grid = (V, chunks)
__global__ void __launch_bounds__(256) adam_kernel(const VarDev* __restrict__ vars, float* __restrict__ params, const float* __restrict__ grads,
                                                   float* __restrict__ m, float* __restrict__ vv, const float* __restrict__ norms, float l2,
                                                   float clipnorm, float alpha) {
  pdl_wait();
  const VarDev v = vars[blockIdx.x];
  const int n = v.rows * v.cols;
  const float k = (v.l2 && l2 > 0.f) ? 2.0f * l2 : 0.f;
  float nsq = 0.f;
#pragma unroll
  for (int c = 0; c < kNormChunks; ++c) nsq += norms[(blockIdx.x * kNormChunks + c) * 2];
  const float scale = (clipnorm > 0.f) ? clipnorm / fmaxf(sqrtf(nsq), clipnorm) : 1.0f;  // tf.clip_by_norm (A4)
  auto update = [&](float w, float g, float& mi, float& vi) {
    g = (g + k * w) * scale;
    mi = kAdamB1 * mi + (1.0f - kAdamB1) * g;
    vi = kAdamB2 * vi + (1.0f - kAdamB2) * g * g;
    return w - alpha * mi / (sqrtf(vi) + kAdamEps);  // A5
  };
  if (((v.cols | v.ld) & 3) == 0 && (v.off & 3) == 0) {  // 16-byte path (same element-wise arithmetic)
    const int c4 = v.cols >> 2, n4 = v.rows * c4;
    for (int j = blockIdx.y * 256 + threadIdx.x; j < n4; j += gridDim.y * 256) {
      const int r = j / c4, c = (j - r * c4) << 2;
      const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
      float4 w = *reinterpret_cast<const float4*>(params + idx);
      const float4 g = *reinterpret_cast<const float4*>(grads + idx);
      float4 mi = *reinterpret_cast<const float4*>(m + idx), vi = *reinterpret_cast<const float4*>(vv + idx);
      w.x = update(w.x, g.x, mi.x, vi.x); w.y = update(w.y, g.y, mi.y, vi.y);
      w.z = update(w.z, g.z, mi.z, vi.z); w.w = update(w.w, g.w, mi.w, vi.w);
      *reinterpret_cast<float4*>(m + idx) = mi;
      *reinterpret_cast<float4*>(vv + idx) = vi;
      *reinterpret_cast<float4*>(params + idx) = w;
    }
  } else {
    for (int i = blockIdx.y * 256 + threadIdx.x; i < n; i += gridDim.y * 256) {
      const int r = i / v.cols, c = i - r * v.cols;
      const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
      float mi = m[idx], vi = vv[idx];
      params[idx] = update(params[idx], grads[idx], mi, vi);
      m[idx] = mi;
      vv[idx] = vi;
    }
  }
}

int launch_regularization_loss(const VarDev* vars, int V, const float* params, float* norms, float l2, float* out, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(var_norms_kernel, dim3(V, kNormChunks), 256, 0, st, vars, params, nullptr, l2, V, norms));
  MFP_CUDA_OK(cudaGetLastError());
  MFP_CUDA_OK(launch_pdl(l2_loss_kernel, 1, 32, 0, st, norms, V, l2, out));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_optimizer(const VarDev* vars, int V, float* params, const float* grads, float* m, float* v, float* norms, int t, float lr, float clipnorm,
                     float l2, float* l2_loss_out, cudaStream_t st) {
  MFP_CUDA_OK(launch_pdl(var_norms_kernel, dim3(V, kNormChunks), 256, 0, st, vars, params, grads, l2, V, norms));
  MFP_CUDA_OK(cudaGetLastError());
  if (l2_loss_out) {
    MFP_CUDA_OK(launch_pdl(l2_loss_kernel, 1, 32, 0, st, norms, V, l2, l2_loss_out));
    MFP_CUDA_OK(cudaGetLastError());
  }
  const double alpha = (double)lr * sqrt(1.0 - pow(0.999, (double)t)) / (1.0 - pow(0.9, (double)t));
  MFP_CUDA_OK(launch_pdl(adam_kernel, dim3(V, 16), 256, 0, st, vars, params, grads, m, v, norms, l2, clipnorm, (float)alpha));
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
