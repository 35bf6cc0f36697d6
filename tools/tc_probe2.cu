// Micro-benchmarks of TMEM load/store throughput as seen by epilogue warps (sizing the EPiC epilogues).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void __launch_bounds__(256, 1) bench(int mode, int iters, long long* out, float* sink) {
  __shared__ uint32_t tmem_base;
  __shared__ __align__(16) float bias[128];
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid < 128) bias[tid] = tid * 0.001f;
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base;
  const int wg = warp >> 2;
  const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + wg * 256;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = tid + i;
  for (int c = 0; c < 8; ++c) tmem_st32(base + c * 32, v);
  tmem_wait_st();
  __syncthreads();
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {            // 4 x (ld32 + wait)
      for (int c = 0; c < 4; ++c) { tmem_ld32(base + c * 32, v); tmem_wait_ld(); acc += __uint_as_float(v[it & 31]); }
    } else if (mode == 1) {     // 4 x ld32 then one wait (register heavy) -> emulate with 2+2
      uint32_t w[32];
      for (int c = 0; c < 4; c += 2) { tmem_ld32(base + c * 32, v); tmem_ld32(base + c * 32 + 32, w); tmem_wait_ld(); acc += __uint_as_float(v[it & 31]) + __uint_as_float(w[it & 31]); }
    } else if (mode == 2) {     // 4 x (st32) + wait
      for (int c = 0; c < 4; ++c) tmem_st32(base + c * 32, v);
      tmem_wait_st();
    } else if (mode == 3) {     // epi_h-like: ld, 32 x (add, lrelu), st back
      for (int c = 0; c < 4; ++c) {
        tmem_ld32(base + c * 32, v); tmem_wait_ld();
        const float4* bj = reinterpret_cast<const float4*>(&bias[c * 32]);
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = bj[i4];
          float a0 = __uint_as_float(v[i4*4+0]) + b.x, a1 = __uint_as_float(v[i4*4+1]) + b.y, a2 = __uint_as_float(v[i4*4+2]) + b.z, a3 = __uint_as_float(v[i4*4+3]) + b.w;
          a0 = fmaxf(a0, 0.01f * a0); a1 = fmaxf(a1, 0.01f * a1); a2 = fmaxf(a2, 0.01f * a2); a3 = fmaxf(a3, 0.01f * a3);
          v[i4*4+0] = __float_as_uint(a0); v[i4*4+1] = __float_as_uint(a1); v[i4*4+2] = __float_as_uint(a2); v[i4*4+3] = __float_as_uint(a3);
        }
        tmem_st32(base + c * 32, v);
      }
      tmem_wait_st();
    } else if (mode == 4) {     // 16-column granularity loads (x16)
      uint32_t w[16];
      for (int c = 0; c < 8; ++c) { tmem_ld16(base + c * 16, w); tmem_wait_ld(); acc += __uint_as_float(w[it & 15]); }
    }
  }
  long long t1 = clock64();
  if ((tid & 31) == 0) out[warp] = t1 - t0;
  sink[tid] = acc + __uint_as_float(v[3]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; float* sink; CK(cudaMalloc(&d, 64)); CK(cudaMalloc(&sink, 1024));
  const char* names[5] = {"4x(ld32+wait)", "2x(2xld32+wait)", "4xst32+wait", "epi_h-like ld+math+st", "8x(ld16+wait)"};
  for (int threads = 128; threads <= 256; threads += 128)
    for (int mode = 0; mode < 5; ++mode) {
      const int iters = 2000;
      bench<<<1, threads>>>(mode, iters, d, sink);
      CK(cudaDeviceSynchronize());
      long long h[8]; CK(cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost));
      printf("%d warps  %-26s %8.1f cycles per 128x128 fp32 tile-pass per warp (64 KB per warpgroup)\n", threads / 32, names[mode], (double)h[0] / iters);
    }
  return 0;
}
