// Device helpers shared by the fp32 CUDA-core kernels (inference: epic_simt.cu, training: epic_train.cu).
#pragma once
#include <cuda_runtime.h>

namespace pfm {

static constexpr int kThreads = 256;
static constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float lrelu(float v, float s) { return v > 0.f ? v : v * s; }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---------------------------------------------------------------------------------------------
// Row-block GEMM on CUDA cores.  Every warp owns RB consecutive rows of the current chunk and all
// `out` columns (lane l holds columns l, l+32, ...: TC per lane).  The k-major weight block
// Wt[K, ldo] streams through a double-buffered shared-memory stage shared by the 8 warps.
//   acc[r][i] = sum_k A[row0 + r][k] * Wt[k][lane + 32 i]
// A rows are zero in their padding columns [K, round_up(K,4)), so the k loop runs on multiples of 4.
// ---------------------------------------------------------------------------------------------
template <int TC, int RB, int LDA_CT, int LDO_CT, int KC_CT>
__device__ __forceinline__ void gemm_rows_impl(const float* __restrict__ A, int lda_rt, const float* __restrict__ Wt, int K,
                                               int ldo_rt, float* wbuf, int wbuf_half, int KC_rt, float (&acc)[RB][TC]) {
  // LDA_CT / LDO_CT / KC_CT != 0: leading dimensions and chunk depth known at compile time (the H = 128 nets): every
  // shared-memory access of the inner loop gets an immediate offset instead of integer multiply-adds per load
  const int lda = LDA_CT ? LDA_CT : lda_rt, ldo = LDO_CT ? LDO_CT : ldo_rt, KC = KC_CT ? KC_CT : KC_rt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int r = 0; r < RB; ++r)
#pragma unroll
    for (int i = 0; i < TC; ++i) acc[r][i] = 0.f;
  const int Kp = (K + 3) & ~3;
  const int n_chunks = (Kp + KC - 1) / KC;
  const float* Arow = A + (size_t)(warp * RB) * lda;
  // prologue: chunk 0
  {
    int kc = Kp < KC ? Kp : KC;
    int n16 = kc * ldo / 4;
    for (int i = tid; i < n16; i += kThreads) cp_async16(wbuf + i * 4, Wt + i * 4);
    cp_async_commit();
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int k0 = c * KC;
    if (c + 1 < n_chunks) {
      int k1 = k0 + KC;
      int kc = (Kp - k1) < KC ? (Kp - k1) : KC;
      int n16 = kc * ldo / 4;
      float* dst = wbuf + ((c + 1) & 1) * wbuf_half;
      const float* src = Wt + (size_t)k1 * ldo;
      for (int i = tid; i < n16; i += kThreads) cp_async16(dst + i * 4, src + i * 4);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = wbuf + (c & 1) * wbuf_half;
    const int kc = KC_CT ? KC_CT : ((Kp - k0) < KC ? (Kp - k0) : KC);      // compile-time path: K is a multiple of KC_CT
#pragma unroll
    for (int kk = 0; kk < kc; kk += 4) {
      float4 a[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = *reinterpret_cast<const float4*>(Arow + (size_t)r * lda + k0 + kk);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float w[TC];
#pragma unroll
        for (int i = 0; i < TC; ++i) w[i] = wb[(kk + q) * ldo + lane + 32 * i];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float av = q == 0 ? a[r].x : (q == 1 ? a[r].y : (q == 2 ? a[r].z : a[r].w));
#pragma unroll
          for (int i = 0; i < TC; ++i) acc[r][i] = fmaf(av, w[i], acc[r][i]);
        }
      }
    }
    __syncthreads();
  }
}



template <int TC, int RB>
__device__ __forceinline__ void gemm_rows(const float* __restrict__ A, int lda, const float* __restrict__ Wt, int K,
                                          int ldo, float* wbuf, int wbuf_half, int KC, float (&acc)[RB][TC]) {
  if constexpr (TC == 4) {
    if (lda == 132 && ldo == 128 && KC == 16 && (K & 15) == 0) {          // uniform over the block
      gemm_rows_impl<TC, RB, 132, 128, 16>(A, lda, Wt, K, ldo, wbuf, wbuf_half, KC, acc);
      return;
    }
  }
  if constexpr (TC == 5) {                                                 // H = 150 (LHCO): leading dimensions only
    if (lda == 156 && ldo == 152) { gemm_rows_impl<TC, RB, 156, 152, 0>(A, lda, Wt, K, ldo, wbuf, wbuf_half, KC, acc); return; }
  }
  if constexpr (TC == 10) {                                                // H = 300 (JetClass cond)
    if (lda == 304 && ldo == 300) { gemm_rows_impl<TC, RB, 304, 300, 0>(A, lda, Wt, K, ldo, wbuf, wbuf_half, KC, acc); return; }
  }
  gemm_rows_impl<TC, RB, 0, 0, 0>(A, lda, Wt, K, ldo, wbuf, wbuf_half, KC, acc);
}

}  // namespace pfm
