// Microbenchmark: how fast can all SMs stream L2-resident data into shared memory with bulk async copies (the TMA
// datapath the GEMM operand ring uses)?  One CTA per SM, STAGES x CHUNK bytes in flight, cycling over a buffer that
// fits in L2 (or not: pass a larger size to get the HBM number).   usage: l2_fill [buffer_MB] [chunk_KB] [stages]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(128, 1) fill_kernel(const uint8_t* buf, size_t buf_bytes, int chunk, int stages, int iters, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[8];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const size_t nchunks = buf_bytes / chunk;
    size_t pos = (size_t)blockIdx.x * 977 % nchunks;
    uint32_t phase = 0;
    for (int s = 0; s < stages; ++s) {
      mbar_expect_tx(smem_u32(&bars[s]), chunk);
      bulk_load(smem_u32(smem + (size_t)s * chunk), buf + pos * chunk, chunk, smem_u32(&bars[s]));
      pos = (pos + gridDim.x) % nchunks;
    }
    for (int it = 0; it < iters; ++it) {
      for (int s = 0; s < stages; ++s) {
        mbar_wait(smem_u32(&bars[s]), phase);
        mbar_expect_tx(smem_u32(&bars[s]), chunk);
        bulk_load(smem_u32(smem + (size_t)s * chunk), buf + pos * chunk, chunk, smem_u32(&bars[s]));
        pos = (pos + gridDim.x) % nchunks;
      }
      phase ^= 1u;
    }
    for (int s = 0; s < stages; ++s) mbar_wait(smem_u32(&bars[s]), phase);
    sink[blockIdx.x] = smem[threadIdx.x];
  }
}

// Cluster-of-2 variant: each CTA fetches half of every chunk and multicasts it into both CTAs' shared memory, so every SM
// receives `chunk` bytes per stage while only chunk/2 per SM leave L2.
__device__ __forceinline__ void bulk_load_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
fill_mc_kernel(const uint8_t* buf, size_t buf_bytes, int chunk, int stages, int iters, unsigned long long* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[8];
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    const size_t nchunks = buf_bytes / chunk;
    const int half = chunk / 2;
    size_t pos = (size_t)(blockIdx.x / 2) * 977 % nchunks;
    uint32_t phase = 0;
    for (int it = 0; it <= iters; ++it) {
      for (int s = 0; s < stages; ++s) {
        if (it > 0) mbar_wait(smem_u32(&bars[s]), phase);
        mbar_expect_tx(smem_u32(&bars[s]), chunk);
        bulk_load_mc(smem_u32(smem + (size_t)s * chunk + rank * half), buf + pos * chunk + rank * half, half, smem_u32(&bars[s]), (uint16_t)3);
        pos = (pos + gridDim.x / 2) % nchunks;
      }
      if (it > 0) phase ^= 1u;
    }
    for (int s = 0; s < stages; ++s) mbar_wait(smem_u32(&bars[s]), phase);
    sink[blockIdx.x] = smem[threadIdx.x];
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 32;
  const int chunk = (argc > 2 ? atoi(argv[2]) : 16) * 1024;
  const int stages = argc > 3 ? atoi(argv[3]) : 4;
  const size_t bytes = mb << 20;
  uint8_t* buf; unsigned long long* sink;
  cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes); cudaMalloc(&sink, 148 * 8);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * chunk);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 400;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    fill_kernel<<<sms, 128, stages * chunk>>>(buf, bytes, chunk, stages, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double total = (double)sms * (iters + 1) * stages * chunk;
    printf("buffer %zu MB chunk %d KB stages %d: %.1f GB/s aggregate, %.1f GB/s per SM (%.3f ms) %s\n", mb, chunk / 1024, stages, total / ms / 1e6,
           total / ms / 1e6 / sms, ms, cudaGetErrorString(cudaGetLastError()));
  }
  cudaFuncSetAttribute(fill_mc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, stages * chunk);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    fill_mc_kernel<<<sms & ~1, 128, stages * chunk>>>(buf, bytes, chunk, stages, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double total = (double)(sms & ~1) * (iters + 1) * stages * chunk;
    printf("multicast x2: buffer %zu MB chunk %d KB stages %d: %.1f GB/s received aggregate, %.1f GB/s per SM (%.3f ms) %s\n", mb, chunk / 1024, stages,
           total / ms / 1e6, total / ms / 1e6 / (sms & ~1), ms, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
