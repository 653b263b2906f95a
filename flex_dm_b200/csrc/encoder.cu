// Per-element multi-attribute embedding + additive fusion (reference: architecture/encoder.py:147-199) and its
// backward.  The numerical fields' Dense (512 -> D) runs on the tensor cores (gemm.cu); this file produces
// everything else of h0 in one pass: categorical gather-sum over sub-targets, <MASK>/<UNUSED> special rows
// and the Dense bias of unflagged rows.
#include "kernels.cuh"

namespace mfp {

constexpr int kTokPerCta = 8;

// thread d owns column d of every row it touches: coalesced 1 KB table-row reads, no cross-thread reduction
__global__ void __launch_bounds__(kD) embed_fwd_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs mod,
                                                       const unsigned char* __restrict__ flags, const float* __restrict__ params, int T,
                                                       float* __restrict__ h0) {
  const int d = threadIdx.x;
  const int t0 = blockIdx.x * kTokPerCta;
#pragma unroll 1
  for (int t = t0; t < min(T, t0 + kTokPerCta); ++t) {
    float acc = 0.0f;
    for (int f = 0; f < sc.F; ++f) {
      const FieldDev& fd = sc.f[f];
      if (fd.kind == 0) {
        const int* idx = reinterpret_cast<const int*>(mod.cols[f]) + (size_t)t * fd.C;
        for (int c = 0; c < fd.C; ++c) {
          int i = __ldg(idx + c);
          i = min(max(i, 0), fd.input_dim + 1);
          acc += __ldg(params + fd.table_off + (size_t)i * kD + d);  // encoder.py:157-160
        }
      } else {
        const int flag = flags[(size_t)fd.num_slot * T + t];
        acc += flag ? __ldg(params + fd.table_off + (size_t)(flag - 1) * kD + d)  // encoder.py:167-175
                    : __ldg(params + fd.bias_off + d);                            // Dense bias; x.W is added by the GEMM
      }
    }
    h0[(size_t)t * kD + d] = acc;
  }
}

// Backward of the above.  grid = (chunks, F).  Each CTA accumulates its token chunk into a shared-memory copy
// of the field's table (thread d owns column d -> no conflicts, no atomics), then flushes once with atomics.
// Numerical fields use a 3-row table {<MASK> row, <UNUSED> row, bias} and also emit dh0 masked by flag == 0,
// which is the B operand of the Dense kernel's wgrad GEMM.
__global__ void __launch_bounds__(kD) embed_bwd_kernel(const __grid_constant__ Schema sc, const __grid_constant__ BatchPtrs mod,
                                                       const unsigned char* __restrict__ flags, const float* __restrict__ dh0, int T,
                                                       int tok_per_chunk, float* __restrict__ grads, float* __restrict__ dh0_masked) {
  extern __shared__ float tab[];
  const int d = threadIdx.x;
  const int f = blockIdx.y;
  const FieldDev& fd = sc.f[f];
  const int rows = (fd.kind == 0) ? fd.input_dim + 2 : 3;
  for (int r = 0; r < rows; ++r) tab[r * kD + d] = 0.0f;
  const int t0 = blockIdx.x * tok_per_chunk;
  const int t1 = min(T, t0 + tok_per_chunk);
  if (fd.kind == 0) {
    const int* idx = reinterpret_cast<const int*>(mod.cols[f]);
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
      const float g = dh0[(size_t)t * kD + d];
      for (int c = 0; c < fd.C; ++c) {
        int i = __ldg(idx + (size_t)t * fd.C + c);
        i = min(max(i, 0), fd.input_dim + 1);
        tab[i * kD + d] += g;
      }
    }
    for (int r = 0; r < rows; ++r) {
      const float v = tab[r * kD + d];
      if (v != 0.0f) atomicAdd(grads + fd.table_off + (size_t)r * kD + d, v);
    }
  } else {
    float* masked = dh0_masked + (size_t)fd.num_slot * T * kD;
#pragma unroll 1
    for (int t = t0; t < t1; ++t) {
      const float g = dh0[(size_t)t * kD + d];
      const int flag = flags[(size_t)fd.num_slot * T + t];
      tab[(flag ? flag - 1 : 2) * kD + d] += g;
      masked[(size_t)t * kD + d] = flag ? 0.0f : g;
    }
    atomicAdd(grads + fd.table_off + d, tab[d]);
    atomicAdd(grads + fd.table_off + kD + d, tab[kD + d]);
    atomicAdd(grads + fd.bias_off + d, tab[2 * kD + d]);
  }
}

int launch_embed_fwd(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, const float* params, int T, float* h0, cudaStream_t st) {
  embed_fwd_kernel<<<(T + kTokPerCta - 1) / kTokPerCta, kD, 0, st>>>(sc, mod, flags, params, T, h0);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

int launch_embed_bwd(const Schema& sc, const BatchPtrs& mod, const unsigned char* flags, const float* dh0, int T, float* grads, float* dh0_masked,
                     cudaStream_t st) {
  int max_rows = 3;
  for (int f = 0; f < sc.F; ++f)
    if (sc.f[f].kind == 0) max_rows = max(max_rows, sc.f[f].input_dim + 2);
  const size_t smem = (size_t)max_rows * kD * sizeof(float);
  if (smem > 200 * 1024) { set_error("embed_bwd: vocabulary of %d rows does not fit the shared-memory table", max_rows); return MFP_ERR_UNSUPPORTED; }
  static size_t configured = 0;
  if (smem > configured) {
    MFP_CUDA_OK(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int chunks = min(64, (T + 63) / 64);
  const int tok_per_chunk = (T + chunks - 1) / chunks;
  embed_bwd_kernel<<<dim3(chunks, sc.F), kD, smem, st>>>(sc, mod, flags, dh0, T, tok_per_chunk, grads, dh0_masked);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
