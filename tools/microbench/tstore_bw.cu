// Microbenchmark: TMA *tensor* store bandwidth (cp.async.bulk.tensor.2d.global.shared) for different box shapes into a
// row-major fp32 matrix [rows][cols]:  usage: tstore_bw [cols] [box_cols] [box_rows] [nbuf] [writers]
// Each CTA runs `writers` independent issuing threads (like the 4 epilogue warps), each with nbuf staging buffers.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) tstore_kernel(const __grid_constant__ CUtensorMap tm, int rows, int cols, int box_cols, int box_rows, int nbuf,
                                                        int writers, int iters, int cb) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int chunk = box_cols * box_rows * 4;
  for (int i = threadIdx.x; i < writers * nbuf * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < writers) {
    const int tiles_c = cols / box_cols, tiles_r = rows / box_rows;
    const long long ntiles = (long long)tiles_c * tiles_r;
    long long pos = ((long long)(blockIdx.x * writers + w) * 977) % ntiles;
    for (int it = 0; it < iters; ++it) {
      const int b = it % nbuf;
      if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      else if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
      const int c0 = (int)(pos % tiles_c) * box_cols, r0 = (int)(pos / tiles_c) * box_rows;
      if (cb == 0)
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm)),
                     "r"(smem_u32(smem + (size_t)(w * nbuf + b) * chunk)), "r"(c0), "r"(r0)
                     : "memory");
      else
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&tm)),
                     "r"(smem_u32(smem + (size_t)(w * nbuf + b) * chunk)), "r"(0), "r"(c0 / 32), "r"(r0)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      pos = (pos + (long long)gridDim.x * writers) % ntiles;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  const int cols = argc > 1 ? atoi(argv[1]) : 768;
  const int box_cols = argc > 2 ? atoi(argv[2]) : 32;
  const int box_rows = argc > 3 ? atoi(argv[3]) : 32;
  const int nbuf = argc > 4 ? atoi(argv[4]) : 2;
  const int writers = argc > 5 ? atoi(argv[5]) : 4;
  const int cb = argc > 6 ? atoi(argv[6]) : 0;  // > 0: 3-D map (32, cols/32, rows), box (32, cb, box_rows), SWIZZLE_128B; box_cols must be 32*cb
  const int rows = 262144 * 2;
  float* buf; cudaMalloc(&buf, (size_t)rows * cols * 4); cudaMemset(buf, 0, (size_t)rows * cols * 4);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  CUtensorMap tm;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
  const CUtensorMapSwizzle sw = box_cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cb > 0) {
    const cuuint64_t d3[3] = {32, (cuuint64_t)cols / 32, (cuuint64_t)rows}, s3[2] = {128, (cuuint64_t)cols * 4};
    const cuuint32_t b3[3] = {32, (cuuint32_t)cb, (cuuint32_t)box_rows}, e3[3] = {1, 1, 1};
    r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, buf, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int chunk = box_cols * box_rows * 4;
  const size_t smem = (size_t)writers * nbuf * chunk;
  cudaFuncSetAttribute(tstore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = (int)((64u << 20) / chunk / writers) > 8000 ? 8000 : (int)((64u << 20) / chunk / writers);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    tstore_kernel<<<sms, 128, smem>>>(tm, rows, cols, box_cols, box_rows, nbuf, writers, iters, cb);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double total = (double)sms * writers * iters * chunk;
    if (rep == 2)
      printf("tensor store%s: cols %d box %dx%d (%d KB) nbuf %d writers %d: %.1f GB/s aggregate, %.1f GB/s per SM (%.3f ms) %s\n", cb ? " (3-D, swizzled 128 B blocks)" : "", cols, box_rows, box_cols,
             chunk / 1024, nbuf, writers, total / ms / 1e6, total / ms / 1e6 / sms, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
