// GPU probe for the tcgen05 building blocks used by epic_tc.cu (not part of the product library):
//   1. SS MMA, K-major SW128 A and B, accumulating onto a tile preloaded into TMEM with tcgen05.st
//   2. TS MMA, A = bf16 pairs written to TMEM with tcgen05.st, B from shared memory
//   3. MN-major A (the h tile read "transposed") x K-major B with N = 16   (masked pooling as an MMA)
//   4. cp.async.bulk + mbarrier complete_tx
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../particle_fm_b200/csrc/tc_ptx.cuh"

using namespace pfm::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Smem {
  alignas(1024) uint8_t A[32768];     // 128 rows x 128 k  bf16, K-major SW128 (2 column blocks of 16 KB)
  alignas(1024) uint8_t B[32768];     // 128 n    x 128 k
  alignas(1024) uint8_t P[4096];      // 16 jets  x 128 rows(k), K-major SW128 (2 blocks of 2 KB)
  uint64_t bar_copy, bar_mma;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* gA, const uint8_t* gB, const uint8_t* gP, const float* C0,
                                                const float* Araw, float* D1, float* D2, float* D3) {
  extern __shared__ uint8_t raw[];
  Smem& s = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&s.bar_copy, 1); mbar_init(&s.bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&s.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  if (tid == 0) {                                   // 4. bulk copies
    mbar_arrive_expect_tx(&s.bar_copy, 32768 + 32768 + 4096);
    bulk_copy_g2s(s.A, gA, 32768, &s.bar_copy);
    bulk_copy_g2s(s.B, gB, 32768, &s.bar_copy);
    bulk_copy_g2s(s.P, gP, 4096, &s.bar_copy);
  }
  // preload C0 row `tid` into TMEM cols [0,128); A row `tid` as bf16 pairs into cols [256, 320)
  const uint32_t lane_addr = tm + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(C0[tid * 128 + c * 32 + i]);
    tmem_st32(lane_addr + c * 32, v);
  }
  for (int c = 0; c < 2; ++c) {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = pack_bf16x2(Araw[tid * 128 + c * 64 + 2 * i], Araw[tid * 128 + c * 64 + 2 * i + 1]);
    tmem_st32(lane_addr + 256 + c * 32, v);
  }
  tmem_wait_st();
  tc_fence_before();
  mbar_wait(&s.bar_copy, 0);
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pool = make_idesc_bf16(128, 16, 1, 0);
    const uint32_t a0 = smem_u32(s.A), b0 = smem_u32(s.B), p0 = smem_u32(s.P);
    for (int k = 0; k < 8; ++k) {                   // 1. D1 = C0 + A . B^T      (cols 0..127)
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      mma_ss(tm + 0, desc_kmajor(a0 + off), desc_kmajor(b0 + off), idesc, 1);
    }
    for (int k = 0; k < 8; ++k) {                   // 2. D2 = A(tmem) . B^T     (cols 128..255)
      const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
      mma_ts(tm + 128, tm + 256 + k * 8, desc_kmajor(b0 + off), idesc, k > 0);
    }
    for (int k = 0; k < 8; ++k) {                   // 3. D3[c][j] = sum_r A[r][c] P[j][r]   (cols 320..335)
      const uint64_t da = desc_mnmajor(a0 + k * 2048, 16384, 1024);
      const uint64_t db = desc_kmajor(p0 + (k >> 2) * 2048 + (k & 3) * 32);
      mma_ss(tm + 320, da, db, idesc_pool, k > 0);
    }
    mma_commit(&s.bar_mma);
  }
  mbar_wait(&s.bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < 4; ++c) {
    uint32_t v[32];
    tmem_ld32(lane_addr + c * 32, v);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D1[tid * 128 + c * 32 + i] = __uint_as_float(v[i]);
    tmem_ld32(lane_addr + 128 + c * 32, v);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) D2[tid * 128 + c * 32 + i] = __uint_as_float(v[i]);
  }
  {
    uint32_t v[16];
    tmem_ld16(lane_addr + 320, v);
    tmem_wait_ld();
    for (int i = 0; i < 16; ++i) D3[tid * 16 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf16r(float x) {
  uint32_t u; memcpy(&u, &x, 4);
  uint32_t r = u + 0x7FFF + ((u >> 16) & 1);
  r &= 0xFFFF0000u; float y; memcpy(&y, &r, 4); return y;
}
static uint16_t bf16bits(float x) { float y = bf16r(x); uint32_t u; memcpy(&u, &y, 4); return (uint16_t)(u >> 16); }

int main() {
  const int M = 128, N = 128, K = 128, J = 16;
  std::vector<float> A(M * K), B(N * K), C0(M * N), P(J * M, 0.f);
  srand(1);
  for (auto& v : A) v = bf16r((rand() / (float)RAND_MAX) * 2 - 1);
  for (auto& v : B) v = bf16r((rand() / (float)RAND_MAX) * 2 - 1);
  for (auto& v : C0) v = (rand() / (float)RAND_MAX) * 2 - 1;
  for (int r = 0; r < M; ++r) if (r < 117) P[(r * 5 / 37) * M + r] = 1.f;      // rows -> jets 0..15, tail rows unassigned
  std::vector<uint8_t> imA(32768), imB(32768), imP(4096, 0);
  for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) {
    *(uint16_t*)&imA[sw128_offset(r, c, 16384)] = bf16bits(A[r * K + c]);
    *(uint16_t*)&imB[sw128_offset(r, c, 16384)] = bf16bits(B[r * K + c]);
  }
  for (int j = 0; j < J; ++j) for (int r = 0; r < 128; ++r) *(uint16_t*)&imP[sw128_offset(j, r, 2048)] = bf16bits(P[j * M + r]);
  uint8_t *dA, *dB, *dP; float *dC0, *dAr, *dD1, *dD2, *dD3;
  CK(cudaMalloc(&dA, 32768)); CK(cudaMalloc(&dB, 32768)); CK(cudaMalloc(&dP, 4096));
  CK(cudaMalloc(&dC0, M * N * 4)); CK(cudaMalloc(&dAr, M * K * 4)); CK(cudaMalloc(&dD1, M * N * 4)); CK(cudaMalloc(&dD2, M * N * 4));
  CK(cudaMalloc(&dD3, M * J * 4));
  CK(cudaMemcpy(dA, imA.data(), 32768, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, imB.data(), 32768, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dP, imP.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dC0, C0.data(), M * N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAr, A.data(), M * K * 4, cudaMemcpyHostToDevice));
  size_t smem = sizeof(Smem) + 1024;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe<<<1, 128, smem>>>(dA, dB, dP, dC0, dAr, dD1, dD2, dD3);
  CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
  std::vector<float> D1(M * N), D2(M * N), D3(M * J);
  CK(cudaMemcpy(D1.data(), dD1, M * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D2.data(), dD2, M * N * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D3.data(), dD3, M * J * 4, cudaMemcpyDeviceToHost));
  double e1 = 0, e2 = 0, e3 = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double acc = 0; for (int k = 0; k < K; ++k) acc += (double)A[m * K + k] * B[n * K + k];
    e1 = fmax(e1, fabs(D1[m * N + n] - (acc + C0[m * N + n]))); e2 = fmax(e2, fabs(D2[m * N + n] - acc));
  }
  for (int c = 0; c < 128; ++c) for (int j = 0; j < J; ++j) {
    double acc = 0; for (int r = 0; r < M; ++r) acc += (double)A[r * K + c] * P[j * M + r];
    e3 = fmax(e3, fabs(D3[c * J + j] - acc));
  }
  printf("probe: SS+residual max err %.3e | TS (A in TMEM) max err %.3e | MN-major pooling max err %.3e\n", e1, e2, e3);
  printf("sample D1[0][0..3] = %f %f %f %f ; D3[0][0..3] = %f %f %f %f\n", D1[0], D1[1], D1[2], D1[3], D3[0], D3[1], D3[2], D3[3]);
  bool ok = e1 < 1e-3 && e2 < 1e-3 && e3 < 1e-3;
  printf(ok ? "PROBE OK\n" : "PROBE FAILED\n");
  return ok ? 0 : 2;
}
