// GPU probe (not part of the product library): throughput of warp-level mma.sync (HMMA.16816.F32.BF16) and ldmatrix on
// sm_100a, alone and while another warp keeps the tcgen05 pipe busy -- the numbers behind moving the per-jet chain of
// epic_tc.cu (pooling, fc_global1/2, bias re-injection) from tcgen05 N=16 MMAs + CUDA-core GEMVs onto mma.sync.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/hmma_probe tools/hmma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"
using namespace pfm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

struct Smem {
  alignas(1024) uint8_t A[32768];
  alignas(1024) uint8_t B[32768];
  uint64_t bar;
  uint32_t tmem_base;
};

// mode bit 0: HMMA warps run; bit 1: tcgen05 warp runs; bit 2: HMMA loop includes 2 ldmatrix.x4 per 2 HMMA (realistic mix);
// bit 3: the HMMAs form ONE dependent chain per warp (latency) instead of 8 independent accumulators
__global__ void __launch_bounds__(384, 1) bench(int mode, int n_hmma_warps, int iters, int mma_iters, long long* out, float* sink) {
  extern __shared__ uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&s.bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  for (int i = tid; i < 65536 / 4; i += 384) reinterpret_cast<uint32_t*>(s.A)[i] = 0x3c003c00u + i;   // finite bf16 junk
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = s.tmem_base;
  float accsum = 0.f;
  long long t0 = 0, t1 = 0;
  if (warp >= 4 && warp < 4 + n_hmma_warps && (mode & 1)) {
    float d[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    uint32_t a[4] = {0x3c003c00u + lane, 0x3c003c00u, 0x3c013c00u, 0x3c003c02u}, b[4] = {0x3c003c00u, 0x3c003c01u + lane, 0, 0};
    const uint32_t arow = smem_u32(s.A) + (uint32_t)((warp - 4) * 4096 + (lane & 15) * 128 + (lane >> 4) * 16);
    t0 = clock64();
    if (mode & 8) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hmma(d[0], a, b[0], b[1]);
      }
    } else if (mode & 4) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint32_t fa[4], fb[4];
          ldsm4(arow + (uint32_t)(((i ^ (lane & 7)) & 7) << 4), fa);
          ldsm4t(arow + 2048u + (uint32_t)(((i ^ (lane & 7)) & 7) << 4), fb);
          hmma(d[2 * i], fa, fb[0], fb[1]);
          hmma(d[2 * i + 1], fa, fb[2], fb[3]);
        }
      }
    } else {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) hmma(d[i], a, b[0], b[1]);
      }
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) accsum += d[i][j];
  } else if (warp == 1 && (mode & 2)) {
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t da = desc_kmajor(smem_u32(s.A)), db = desc_kmajor(smem_u32(s.B));
    t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ks = (uint32_t)(k >> 2) * 1024u + (uint32_t)(k & 3) * 2u;
          if (mode & 16) mma_ts(tm + (it & 1) * 128, tm + 256 + k * 8, db + ks, idesc, 1u);      // A from TMEM: no shared-memory A fetch
          else mma_ss(tm + (it & 1) * 128, da + ks, db + ks, idesc, 1u);
        }
      }
      mma_commit(&s.bar);
    }
    __syncwarp();
    mbar_wait(&s.bar, 0);
    t1 = clock64();
  }
  if (lane == 0) out[warp] = t1 - t0;
  sink[tid] = accsum;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; float* sink; CK(cudaMalloc(&d, 128)); CK(cudaMalloc(&sink, 4096));
  const int smem = sizeof(Smem) + 1024;
  CK(cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int iters = 4000, mma_iters = 1000;
  struct Case { const char* name; int mode, nw; } cases[] = {
    {"HMMA independent x8, 4 warps (1/SMSP)", 1, 4}, {"HMMA independent x8, 8 warps (2/SMSP)", 1, 8},
    {"HMMA dependent chain, 4 warps", 1 | 8, 4}, {"HMMA dependent chain, 8 warps", 1 | 8, 8},
    {"ldmatrix+HMMA mix, 4 warps", 1 | 4, 4}, {"ldmatrix+HMMA mix, 8 warps", 1 | 4, 8},
    {"tcgen05 SS 128x128x16 alone", 2, 0}, {"tcgen05 TS 128x128x16 alone", 2 | 16, 0},
    {"tcgen05 SS + HMMA x8 on 8 warps", 3, 8}, {"tcgen05 TS + HMMA x8 on 8 warps", 3 | 16, 8},
    {"tcgen05 SS + ldmatrix/HMMA mix on 8 warps", 3 | 4, 8}, {"tcgen05 TS + ldmatrix/HMMA mix on 8 warps", 3 | 4 | 16, 8},
  };
  for (const Case& c : cases) {
    CK(cudaMemset(d, 0, 128));
    bench<<<1, 384, smem>>>(c.mode, c.nw, iters, mma_iters, d, sink);
    CK(cudaDeviceSynchronize());
    long long h[12]; CK(cudaMemcpy(h, d, 96, cudaMemcpyDeviceToHost));
    printf("%-46s", c.name);
    if (c.mode & 1) printf("  HMMA: %.2f cycles each per warp (warp 4), %.2f (warp %d)", (double)h[4] / (iters * 8.0), (double)h[4 + c.nw - 1] / (iters * 8.0), 4 + c.nw - 1);
    if (c.mode & 2) printf("  tcgen05: %.1f cycles per 128x128x16 MMA", (double)h[1] / (mma_iters * 8.0));
    printf("\n");
  }
  return 0;
}
