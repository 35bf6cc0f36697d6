// TMEM load throughput by tcgen05.ld shape (timing only; data layout ignored).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
#define OUT32 "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define REGS32 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
__device__ __forceinline__ void ld_16x256b_x8(uint32_t a, uint32_t (&v)[32]) { asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 " REGS32 : OUT32 : "r"(a) : "memory"); }
__device__ __forceinline__ void ld_16x128b_x16(uint32_t a, uint32_t (&v)[32]) { asm volatile("tcgen05.ld.sync.aligned.16x128b.x16.b32 " REGS32 : OUT32 : "r"(a) : "memory"); }
__device__ __forceinline__ void ld_16x64b_x32(uint32_t a, uint32_t (&v)[32]) { asm volatile("tcgen05.ld.sync.aligned.16x64b.x32.b32 " REGS32 : OUT32 : "r"(a) : "memory"); }

__global__ void __launch_bounds__(256, 1) bench(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = tid + i;
  for (int c = 0; c < 8; ++c) tmem_st32(base + c * 32, v);
  tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int c = 0; c < 4; ++c) {
      const uint32_t a = base + (c & 1) * 64 + ((uint32_t)((c >> 1) * 16) << 16);
      if (mode == 0) tmem_ld32(base + c * 32, v);
      else if (mode == 1) ld_16x256b_x8(a, v);
      else if (mode == 2) ld_16x128b_x16(a, v);
      else ld_16x64b_x32(a, v);
      tmem_wait_ld();
      acc += __uint_as_float(v[0]) + __uint_as_float(v[31]) + __uint_as_float(v[13]);
    }
  }
  long long t1 = clock64();
  if ((tid & 31) == 0) out[warp] = t1 - t0;
  sink[tid] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
int main() {
  long long* d; float* sink; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&sink, 1024));
  const char* names[4] = {"32x32b.x32", "16x256b.x8", "16x128b.x16", "16x64b.x32"};
  for (int threads = 32; threads <= 256; threads *= 2)
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 2000;
      bench<<<1, threads>>>(mode, iters, d, sink);
      CK(cudaDeviceSynchronize());
      long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
      printf("%d warps  %-12s %8.1f cycles per 16 KB per warp  -> %.1f B/cycle/SM\n", threads / 32, names[mode], (double)h[0] / iters,
             16384.0 * (threads / 32) / ((double)h[0] / iters));
    }
  return 0;
}
