// GPU probe (not part of the product library): register <-> (lane, column) map of tcgen05.ld.16x256b.x4, checked against
// a TMEM image written with tcgen05.st.32x32b (thread = lane, register = column).  The training GEMM epilogue relies on it:
// a quad of threads holds 8 consecutive fp32 columns of a row = one full 32-byte sector per global store.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tmem_ld_probe tools/tmem_ld_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(uint32_t* out) {
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) tmem_alloc(&tmem_base, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = (uint32_t)((warp * 32 + lane) * 1000 + c);       // lane * 1000 + column
  tmem_st32(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int lh = 0; lh < 2; ++lh) {
    uint32_t r[16];
    ld_16x256b_x4(tm + ((uint32_t)(warp * 32 + lh * 16) << 16), r);
    tmem_wait_ld();
    for (int k = 0; k < 16; ++k) out[((warp * 2 + lh) * 32 + lane) * 16 + k] = r[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 64);
}

int main() {
  uint32_t* d; CK(cudaMalloc(&d, 4 * 2 * 32 * 16 * 4));
  probe<<<1, 128>>>(d);
  CK(cudaDeviceSynchronize());
  static uint32_t h[4 * 2 * 32 * 16];
  CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int lh = 0; lh < 2; ++lh) for (int T = 0; T < 32; ++T) for (int k = 0; k < 16; ++k) {
    const int j = k >> 2, e = k & 1, hi = (k >> 1) & 1;
    const int row = w * 32 + lh * 16 + T / 4 + 8 * hi, col = 8 * j + 2 * (T % 4) + e;
    const uint32_t want = (uint32_t)(row * 1000 + col), got = h[((w * 2 + lh) * 32 + T) * 16 + k];
    if (want != got && bad++ < 12) printf("warp %d lh %d thread %d reg %d: got lane %u col %u, expected lane %d col %d\n", w, lh, T, k, got / 1000, got % 1000, row, col);
  }
  printf("16x256b.x4 map: reg 4j+2h+e of thread T = (lane T/4 + 8h, column 8j + 2(T%%4) + e): %s (%d mismatches)\n", bad ? "WRONG" : "confirmed", bad);
  return bad != 0;
}
