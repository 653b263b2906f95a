// Microbenchmark, two questions about 128-byte-row traffic per SM:
//  (1) stg: how fast can 4 warps write [32 x 32] fp32 chunks of a row-major matrix with coalesced st.global.v4 (each
//      instruction covers 4 rows x 128 B) -- the alternative to TMA tensor stores of 32x32 boxes;
//  (2) tload: how fast do TMA *tensor* loads of K-major operand boxes (32 floats x R rows, 128-byte rows, SWIZZLE_128B) fill
//      shared memory, with `stages` boxes in flight.
// usage: rowrate stg [cols]   |   rowrate tload [cols] [box_rows] [stages]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) stg_kernel(float* out, int rows, int cols, int iters) {
  __shared__ __align__(128) float4 stage[4][256];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 256; i += 32) stage[w][i] = make_float4(1.f, 2.f, 3.f, 4.f);
  __syncwarp();
  const int tiles_c = cols / 32, tiles_r = rows / 32;
  const long long ntiles = (long long)tiles_c * tiles_r;
  long long pos = ((long long)(blockIdx.x * 4 + w) * 977) % ntiles;
  for (int it = 0; it < iters; ++it) {
    const int c0 = (int)(pos % tiles_c) * 32, r0 = (int)(pos / tiles_c) * 32;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + (lane >> 3), u = lane & 7;
      const float4 v = stage[w][r * 8 + (u ^ (r & 7))];
      *reinterpret_cast<float4*>(out + (size_t)(r0 + r) * cols + c0 + u * 4) = v;
    }
    pos = (pos + (long long)gridDim.x * 4) % ntiles;
  }
}

__global__ void __launch_bounds__(128, 1) tload_kernel(const __grid_constant__ CUtensorMap tm, int rows, int cols, int box_rows, int stages, int iters,
                                                       unsigned long long* sink, int poll, int issuers) {
  extern __shared__ __align__(1024) uint8_t smem_all[];
  __shared__ __align__(8) unsigned long long bars_all[32];
  const int wid = threadIdx.x >> 5;
  uint8_t* smem = smem_all + (size_t)wid * stages * (32 * box_rows * 4);
  unsigned long long* bars = bars_all + wid * 8;
  const int chunk = 32 * box_rows * 4;
  if ((threadIdx.x & 31) == 0 && wid < issuers) {
    for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && wid < issuers) {
    const int tiles_c = cols / 32, tiles_r = rows / box_rows;
    const long long ntiles = (long long)tiles_c * tiles_r;
    long long pos = ((long long)(blockIdx.x * issuers + wid) * 977) % ntiles;
    uint32_t phase = 0;
    for (int it = 0; it <= iters; ++it) {
      for (int s = 0; s < stages; ++s) {
        const uint32_t bar = smem_u32(&bars[s]);
        if (it > 0) {
          uint32_t done = 0;
          if (poll)  // non-blocking test in a spin loop instead of the (potentially suspending) try_wait
            while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
          else
            while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase) : "memory");
        }
        if (it < iters) {
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(chunk) : "memory");
          const int c0 = (int)(pos % tiles_c) * 32, r0 = (int)(pos / tiles_c) * box_rows;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                           smem_u32(smem + (size_t)s * chunk)),
                       "l"(reinterpret_cast<uint64_t>(&tm)), "r"(c0), "r"(r0), "r"(bar)
                       : "memory");
          pos = (pos + (long long)gridDim.x * issuers) % ntiles;
        }
      }
      if (it > 0) phase ^= 1u;
    }
    sink[blockIdx.x] = smem[0];
  }
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "stg";
  const int cols = argc > 2 ? atoi(argv[2]) : 768;
  const int rows = 262144 * 2;
  float* buf; cudaMalloc(&buf, (size_t)rows * cols * 4); cudaMemset(buf, 0, (size_t)rows * cols * 4);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms = 0;
  if (!strcmp(mode, "stg")) {
    const int iters = 4000;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a); stg_kernel<<<sms, 128>>>(buf, rows, cols, iters); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
    }
    const double total = (double)sms * 4 * iters * 4096;
    printf("stg.v4 coalesced 32x32 chunks, cols %d, 4 warps: %.1f GB/s aggregate, %.1f GB/s per SM (%.3f ms) %s\n", cols, total / ms / 1e6, total / ms / 1e6 / sms, ms,
           cudaGetErrorString(cudaGetLastError()));
  } else {
    const int box_rows = argc > 3 ? atoi(argv[3]) : 128, stages = argc > 4 ? atoi(argv[4]) : 4, poll = argc > 5 ? atoi(argv[5]) : 0, issuers = argc > 6 ? atoi(argv[6]) : 1;
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    CUtensorMap tm;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = ((EncodeTiledFn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    unsigned long long* sink; cudaMalloc(&sink, 148 * 8);
    const int chunk = 32 * box_rows * 4;
    cudaFuncSetAttribute(tload_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, issuers * stages * chunk);
    const int iters = 600;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(a); tload_kernel<<<sms, 128, issuers * stages * chunk>>>(tm, rows, cols, box_rows, stages, iters, sink, poll, issuers); cudaEventRecord(b); cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
    }
    const double total = (double)sms * iters * stages * chunk * issuers;
    printf("tensor load%s 32 x %d boxes (%d KB) x %d stages, cols %d: %.1f GB/s aggregate, %.1f GB/s per SM (%.3f ms) %s\n", issuers > 1 ? " [several issuing warps]" : (poll ? " [test_wait polling]" : ""), box_rows, chunk / 1024, stages, cols,
           total / ms / 1e6, total / ms / 1e6 / sms, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
