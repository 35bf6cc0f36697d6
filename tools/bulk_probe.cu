// GPU probe (not part of the product library): latency / throughput of 32 KB cp.async.bulk (global -> shared) copies of
// L2-resident weight images when all 148 SMs stream the SAME images at the same time -- the weight ring of epic_tc.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bulk_probe tools/bulk_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Smem {
  alignas(1024) uint8_t w[6][32768];
  uint64_t full[6];
};

// depth copies in flight per CTA; every CTA walks the same n_img images (same = 1) or its own (same = 0)
// mode 2: n_warps warps issue concurrently (lane 0 of each), each with its own `depth` slots of the 6
__global__ void __launch_bounds__(128, 1) bench2(const uint8_t* img, int n_img, int iters, int depth, int n_warps, uint32_t bytes, long long* out) {
  extern __shared__ uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  if (threadIdx.x == 0) { for (int i = 0; i < 6; ++i) mbar_init(&s.full[i], 1); fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && warp < n_warps) {
    long long t0 = clock64();
    int issued = 0, done = 0;
    while (done < iters) {
      while (issued < iters && issued - done < depth) {
        const int sl = warp * depth + issued % depth;
        mbar_arrive_expect_tx(&s.full[sl], bytes);
        bulk_copy_g2s(s.w[sl], img + (size_t)((issued + warp * 7) % n_img) * 32768, bytes, &s.full[sl]);
        ++issued;
      }
      const int sl = warp * depth + done % depth;
      mbar_wait(&s.full[sl], (done / depth) & 1);
      ++done;
    }
    long long t1 = clock64();
    out[blockIdx.x * 4 + warp] = t1 - t0;
  }
}

__global__ void __launch_bounds__(128, 1) bench(const uint8_t* img, int n_img, int iters, int depth, int same, uint32_t bytes, long long* out) {
  extern __shared__ uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  if (threadIdx.x == 0) { for (int i = 0; i < 6; ++i) mbar_init(&s.full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t* base = img + (same ? 0 : (size_t)blockIdx.x * n_img * 32768);
    long long t0 = clock64();
    int issued = 0, done = 0;
    long long lat = 0;
    long long t_issue[6];
    while (done < iters) {
      while (issued < iters && issued - done < depth) {
        const int sl = issued % depth;
        mbar_arrive_expect_tx(&s.full[sl], bytes);
        bulk_copy_g2s(s.w[sl], base + (size_t)(issued % n_img) * 32768, bytes, &s.full[sl]);
        t_issue[sl] = clock64();
        ++issued;
      }
      const int sl = done % depth;
      mbar_wait(&s.full[sl], (done / depth) & 1);
      lat += clock64() - t_issue[sl];
      ++done;
    }
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = lat;
  }
}

int main() {
  const int n_img = 27, nb = 148;
  uint8_t* img; CK(cudaMalloc(&img, (size_t)nb * n_img * 32768)); CK(cudaMemset(img, 1, (size_t)nb * n_img * 32768));
  long long* d; CK(cudaMalloc(&d, nb * 16));
  const int smem = sizeof(Smem) + 1024;
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 2000;
  CK(cudaFuncSetAttribute(bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long* d2; CK(cudaMalloc(&d2, nb * 32));
  for (int n_warps : {1, 2, 3})
    for (uint32_t bytes : {32768u, 16384u, 8192u, 2048u})
      for (int depth : {1, 2}) {
        for (int rep = 0; rep < 2; ++rep) { bench2<<<148, 128, smem>>>(img, n_img, iters, depth, n_warps, bytes, d2); CK(cudaDeviceSynchronize()); }
        long long h[592]; CK(cudaMemcpy(h, d2, 148 * 32, cudaMemcpyDeviceToHost));
        double tot = 0;
        for (int i = 0; i < 148; ++i) tot += h[4 * i];
        const double per = tot / 148 / iters;
        printf("148 CTAs, %d issuing warps x %d in flight, %5u B: %7.0f cycles per copy per warp -> %6.1f B/clk/SM\n", n_warps, depth, bytes, per,
               n_warps * bytes / per);
      }
  return 0;
}
