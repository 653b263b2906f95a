// Adam(learning_rate, clipnorm=1.0) + Keras L2 regularisers (reference: train.py:71-77, architecture/utils.py:8-22;
// semantics SURVEY.md Appendix A3-A5): g <- g + 2*l2*w on regularised variables, per-VARIABLE clip_by_norm, then
// TF-form Adam (epsilon outside the bias-corrected sqrt).  Variables are strided 2-D views of the flat buffers.
#include "kernels.cuh"

namespace mfp {

// one CTA per variable: ||g + 2 l2 w||_2 and sum w^2 (fixed-order reduction -> deterministic)
__global__ void __launch_bounds__(256) var_norms_kernel(const VarDev* __restrict__ vars, const float* __restrict__ params,
                                                        const float* __restrict__ grads, float l2, int V, float* __restrict__ norms) {
  __shared__ float red[2][256];
  const VarDev v = vars[blockIdx.x];
  const int n = v.rows * v.cols;
  const float k = (v.l2 && l2 > 0.f) ? 2.0f * l2 : 0.f;
  float sg = 0.f, sw = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int r = i / v.cols, c = i - r * v.cols;
    const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
    const float w = params[idx];
    const float g = grads[idx] + k * w;
    sg += g * g;
    sw += w * w;
  }
  red[0][threadIdx.x] = sg;
  red[1][threadIdx.x] = sw;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      red[0][threadIdx.x] += red[0][threadIdx.x + o];
      red[1][threadIdx.x] += red[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    norms[blockIdx.x] = sqrtf(red[0][0]);
    norms[V + blockIdx.x] = (v.l2 && l2 > 0.f) ? red[1][0] : 0.f;
  }
}

__global__ void l2_loss_kernel(const float* __restrict__ norms, int V, float l2, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < V; ++i) s += norms[V + i];
    *out = (l2 > 0.f) ? l2 * s : 0.f;  // A3: l2 * sum(w^2), no 1/2
  }
}

// grid = (V, chunks)
__global__ void __launch_bounds__(256) adam_kernel(const VarDev* __restrict__ vars, float* __restrict__ params, const float* __restrict__ grads,
                                                   float* __restrict__ m, float* __restrict__ vv, const float* __restrict__ norms, float l2,
                                                   float clipnorm, float alpha) {
  const VarDev v = vars[blockIdx.x];
  const int n = v.rows * v.cols;
  const float k = (v.l2 && l2 > 0.f) ? 2.0f * l2 : 0.f;
  const float scale = (clipnorm > 0.f) ? clipnorm / fmaxf(norms[blockIdx.x], clipnorm) : 1.0f;  // tf.clip_by_norm (A4)
  for (int i = blockIdx.y * 256 + threadIdx.x; i < n; i += gridDim.y * 256) {
    const int r = i / v.cols, c = i - r * v.cols;
    const size_t idx = (size_t)v.off + (size_t)r * v.ld + c;
    const float w = params[idx];
    const float g = (grads[idx] + k * w) * scale;
    const float mi = kAdamB1 * m[idx] + (1.0f - kAdamB1) * g;
    const float vi = kAdamB2 * vv[idx] + (1.0f - kAdamB2) * g * g;
    m[idx] = mi;
    vv[idx] = vi;
    params[idx] = w - alpha * mi / (sqrtf(vi) + kAdamEps);  // A5
  }
}

// l2 * sum w^2 only (Keras test_step adds the regularisation losses to the reported loss as well)
__global__ void __launch_bounds__(256) var_w2_kernel(const VarDev* __restrict__ vars, const float* __restrict__ params, float l2, int V,
                                                     float* __restrict__ norms) {
  __shared__ float red[256];
  const VarDev v = vars[blockIdx.x];
  const int n = v.rows * v.cols;
  float sw = 0.f;
  if (v.l2 && l2 > 0.f)
    for (int i = threadIdx.x; i < n; i += 256) {
      const int r = i / v.cols, c = i - r * v.cols;
      const float w = params[(size_t)v.off + (size_t)r * v.ld + c];
      sw += w * w;
    }
  red[threadIdx.x] = sw;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) norms[V + blockIdx.x] = red[0];
}

int launch_regularization_loss(const VarDev* vars, int V, const float* params, float* norms, float l2, float* out, cudaStream_t st) {
  var_w2_kernel<<<V, 256, 0, st>>>(vars, params, l2, V, norms);
  MFP_CUDA_OK(cudaGetLastError());
  l2_loss_kernel<<<1, 32, 0, st>>>(norms, V, l2, out);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_optimizer(const VarDev* vars, int V, float* params, const float* grads, float* m, float* v, float* norms, int t, float lr, float clipnorm,
                     float l2, float* l2_loss_out, cudaStream_t st) {
  var_norms_kernel<<<V, 256, 0, st>>>(vars, params, grads, l2, V, norms);
  MFP_CUDA_OK(cudaGetLastError());
  if (l2_loss_out) {
    l2_loss_kernel<<<1, 32, 0, st>>>(norms, V, l2, l2_loss_out);
    MFP_CUDA_OK(cudaGetLastError());
  }
  const double alpha = (double)lr * sqrt(1.0 - pow(0.999, (double)t)) / (1.0 - pow(0.9, (double)t));
  adam_kernel<<<dim3(V, 16), 256, 0, st>>>(vars, params, grads, m, v, norms, l2, clipnorm, (float)alpha);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
