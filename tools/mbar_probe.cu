// GPU probe (not part of the product library): hand-off latency between two warps of a CTA through an mbarrier, for the
// different ways of waiting (mbarrier.try_wait = potentially-suspending, mbarrier.test_wait = pure polling), and through
// tcgen05.commit.  epic_tc.cu's per-layer chain crosses several such hops; this measures what one costs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/mbar_probe tools/mbar_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  if (MODE == 0) { while (!mbar_try_wait(bar, parity)) {} }
  else if (MODE == 1) { while (!mbar_test_wait(bar, parity)) {} }
  else { while (!mbar_try_wait_hint(bar, parity, 1u)) {} }
}

// warp 0 lane 0 <-> warp `other` lane 0 ping-pong; n_arrive threads arrive on the "ping" barrier (like 128 epilogue threads)
template <int MODE>
__global__ void __launch_bounds__(256, 1) pingpong(int iters, int other, int n_arrive, long long* out) {
  __shared__ uint64_t ping, pong;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&ping, n_arrive); mbar_init(&pong, 1); fence_barrier_init(); }
  __syncthreads();
  long long t0 = clock64();
  if (tid < n_arrive) {                      // "epilogue" side: all arrive, all wait for the answer
    for (int i = 0; i < iters; ++i) {
      mbar_arrive(&ping);
      wait<MODE>(&pong, i & 1);
    }
  } else if (warp == other && lane == 0) {   // "MMA warp" side: one thread waits and answers
    for (int i = 0; i < iters; ++i) {
      wait<MODE>(&ping, i & 1);
      mbar_arrive(&pong);
    }
  }
  long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
}

// the answer comes from tcgen05.commit (no MMA pending: the commit completes immediately) instead of a plain arrive
template <int MODE>
__global__ void __launch_bounds__(256, 1) commitpong(int iters, long long* out) {
  __shared__ uint64_t ping, pong;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&ping, 128); mbar_init(&pong, 1); fence_barrier_init(); }
  if (warp == 7) tmem_alloc(&tmem_base, 32);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  long long t0 = clock64();
  if (tid < 128) {
    for (int i = 0; i < iters; ++i) {
      tc_fence_before();
      mbar_arrive(&ping);
      wait<MODE>(&pong, i & 1);
      tc_fence_after();
    }
  } else if (warp == 5) {
    for (int i = 0; i < iters; ++i) {
      wait<MODE>(&ping, i & 1);
      tc_fence_after();
      if (elect_one()) mma_commit(&pong);
      __syncwarp();
    }
  }
  long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  tc_fence_before(); __syncthreads();
  if (warp == 7) tmem_dealloc(tmem_base, 32);
}

__global__ void __launch_bounds__(256, 1) barpong(int iters, long long* out) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) asm volatile("bar.sync 1, 256;" ::: "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main() {
  long long* d; CK(cudaMalloc(&d, 64));
  const int iters = 20000;
  long long h;
  const char* names[3] = {"try_wait (default)", "test_wait (polling)", "try_wait, 1 ns hint"};
  for (int n_arrive : {1, 32, 128}) {
    for (int mode = 0; mode < 3; ++mode) {
      if (mode == 0) pingpong<0><<<1, 256>>>(iters, 5, n_arrive, d);
      if (mode == 1) pingpong<1><<<1, 256>>>(iters, 5, n_arrive, d);
      if (mode == 2) pingpong<2><<<1, 256>>>(iters, 5, n_arrive, d);
      CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
      printf("ping-pong, %3d arriving threads, %-20s: %7.1f cycles per round trip (2 hops)\n", n_arrive, names[mode], (double)h / iters);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    if (mode == 0) commitpong<0><<<1, 256>>>(iters, d); else commitpong<1><<<1, 256>>>(iters, d);
    CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
    printf("128 threads arrive -> warp waits -> tcgen05.commit -> 128 wait, %-20s: %7.1f cycles per round trip\n", names[mode], (double)h / iters);
  }
  barpong<<<1, 256>>>(iters, d);
  CK(cudaDeviceSynchronize()); CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
  printf("bar.sync over 256 threads: %7.1f cycles each\n", (double)h / iters);
  return 0;
}
