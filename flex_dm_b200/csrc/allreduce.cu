// Gradient all-reduce of the data-parallel MFP step over NVLink SHARP (NVLS), as one kernel on the step's own stream.
// The reference is single-process (train.py:25 is a commented-out MirroredStrategy stub); the B200 path shards documents over the GPUs
// and sums the flat gradient buffer.  With the buffer in symmetric memory (same offset on every rank, mapped
// into one multicast address range) the sum is made by the switch:
//   barrier (every rank's backward pass is complete)
//   rank r, for its 1/N slice:  v = multimem.ld_reduce.add [mc + i]   -- the switch pulls the 16 bytes from every rank and adds them
//                               multimem.st [mc + i] = v              -- and writes the sum back into every rank's buffer
//   barrier (every slice has landed everywhere) -> the optimiser kernels follow in stream order.
// Against ncclAllReduce on its own stream this removes two stream hand-overs and the ring / tree latency: 46 us instead of ~100 us for
// the 11.25 MB of crello on 8 GPUs.  Barriers are flags in the symmetric signal pads (st.release.sys / ld.acquire.sys), monotonic
// per call so nothing is ever reset.  Every wait is bounded and traps: a protocol bug must not hang the box.
#include <stdio.h>

#include "kernels.cuh"

namespace mfp {

constexpr int kArBlocks = 64;
constexpr int kArThreads = 512;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
template <typename F>
__device__ __forceinline__ void bounded_spin(F done, const char* what) {
  const long long t0 = clock64();
  while (!done()) {
    if (clock64() - t0 > 6000000000LL) {  // ~3 s
      printf("mfp nvls all-reduce: timeout waiting for %s (block %d thread %d)\n", what, blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Cross-GPU barrier run by block 0 (one thread per peer), then released to the other blocks through a flag in local memory.
__device__ __forceinline__ void rank_barrier(uint32_t* const* pads, int slot0, int rank, int world, uint32_t value, uint32_t* local_flag) {
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < world) {
      st_release_sys(pads[threadIdx.x] + slot0 + rank, value);  // tell peer t that this rank has arrived
      const uint32_t* mine = pads[rank] + slot0 + threadIdx.x;
      bounded_spin([&] { return (int)(ld_acquire_sys(mine) - value) >= 0; }, "a peer's arrival");
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(local_flag, value);
  } else {
    if (threadIdx.x == 0) bounded_spin([&] { return (int)(ld_acquire_gpu(local_flag) - value) >= 0; }, "block 0");
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kArThreads) nvls_allreduce_kernel(float* __restrict__ mc, uint32_t* const* __restrict__ pads, int slot0, int rank, int world,
                                                                    size_t n4, uint32_t call, uint32_t local_call, uint32_t* __restrict__ sync /*[0] flag, [1] done counter*/) {
  rank_barrier(pads, slot0, rank, world, 2u * call - 1u, sync);
  const size_t per = (n4 + world - 1) / world;
  const size_t lo = (size_t)rank * per, hi = min(n4, lo + per);
  for (size_t i = lo + (size_t)blockIdx.x * kArThreads + threadIdx.x; i < hi; i += (size_t)kArBlocks * kArThreads) {
    float4 v;
    float* p = mc + 4 * i;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  }
  // every store of this rank is out before it tells its peers so
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t arrived = atomicAdd(sync + 1, 1u) + 1u;
    if (blockIdx.x == 0) {
      const uint32_t* cnt = sync + 1;
      bounded_spin([&] { return (int)(ld_acquire_gpu(cnt) - local_call * (uint32_t)kArBlocks) >= 0; }, "this rank's blocks");
    }
    (void)arrived;
  }
  __syncthreads();
  rank_barrier(pads, slot0, rank, world, 2u * call, sync);
}

int launch_nvls_allreduce(float* multicast, uint32_t* const* pads_dev, int slot0, int rank, int world, size_t n_floats, uint32_t call, uint32_t local_call,
                          uint32_t* sync, cudaStream_t st) {
  if (n_floats % 4) { set_error("nvls all-reduce: the buffer length must be a multiple of 4 floats"); return MFP_ERR_ARG; }
  if (world < 2 || world > kArThreads || rank < 0 || rank >= world) { set_error("nvls all-reduce: bad rank / world"); return MFP_ERR_ARG; }
  // plain launch (no programmatic dependent launch): the kernel spins on other GPUs and on its own blocks, all of which must be resident
  nvls_allreduce_kernel<<<kArBlocks, kArThreads, 0, st>>>(multicast, pads_dev, slot0, rank, world, n_floats / 4, call, local_call, sync);
  MFP_CUDA_OK(cudaGetLastError());
  return MFP_OK;
}

}  // namespace mfp
