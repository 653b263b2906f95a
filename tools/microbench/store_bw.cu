// Microbenchmark: write bandwidth of bulk async stores (shared -> global), the datapath of the GEMM / attention epilogues.
// One CTA per SM, NBUF staging buffers of CHUNK bytes, cp.async.bulk.wait_group.read<NBUF-1> before a buffer is reused.
//   usage: store_bw [buffer_MB] [chunk_KB] [row_bytes: 0 = linear chunk, else rows of that many bytes at a 4x stride]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int NBUF>
__global__ void __launch_bounds__(128, 1) store_kernel(uint8_t* buf, size_t buf_bytes, int chunk, int row_bytes, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];
  for (int i = threadIdx.x; i < NBUF * chunk / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nchunks = buf_bytes / chunk;
    size_t pos = (size_t)blockIdx.x * 977 % nchunks;
    for (int it = 0; it < iters; ++it) {
      const int b = it % NBUF;
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
      if (row_bytes == 0) {
        bulk_store(buf + pos * chunk, smem_u32(smem + (size_t)b * chunk), chunk);
      } else {  // a [rows][row_bytes] box into a matrix whose pitch is 4 x row_bytes (like a 32-column chunk of a wide output)
        const int rows = chunk / row_bytes;
        uint8_t* base = buf + (pos / 4) * (size_t)chunk * 4 + (pos % 4) * row_bytes;
        for (int r = 0; r < rows; ++r) bulk_store(base + (size_t)r * row_bytes * 4, smem_u32(smem + (size_t)b * chunk + r * row_bytes), row_bytes);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      pos = (pos + gridDim.x) % nchunks;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 2048;
  const int chunk = (argc > 2 ? atoi(argv[2]) : 4) * 1024;
  const int row_bytes = argc > 3 ? atoi(argv[3]) : 0;
  const size_t bytes = mb << 20;
  uint8_t* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
  constexpr int NBUF = 2;
  cudaFuncSetAttribute(store_kernel<NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * chunk);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = (int)((size_t)(1u << 29) / chunk / 4);  // 128 MB per SM... bounded below
  const int it = iters > 4000 ? 4000 : iters;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    store_kernel<NBUF><<<sms, 128, NBUF * chunk>>>(buf, bytes, chunk, row_bytes, it);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double total = (double)sms * it * chunk;
    if (rep == 2)
      printf("store: buffer %zu MB chunk %d KB rows %d B x2 buffers: %.1f GB/s aggregate, %.1f GB/s per SM (%.3f ms) %s\n", mb, chunk / 1024, row_bytes,
             total / ms / 1e6, total / ms / 1e6 / sms, ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
