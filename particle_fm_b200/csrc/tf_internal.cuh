// Declarations shared by the droid-transformer translation units (tf_simt.cu, tf_tc.cu); not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pfm {

// Y[rows, N] = [R +] act(LN(X[rows, K]) . W^T + bias [+ jb[rowjet[row]]])
struct LinArgs {
  const float* X; int ldx; int K;
  const float* ln_g; const float* ln_b;
  const float* Wt; int ldo; int N;          // fp32 path: k-major Wt[k*ldo + o]; row 0 = first input column used
  const float* bias;
  const float* jb; int jb_stride; const int* rowjet;
  const float* R; int ldr;
  float* Y; int ldy;
  int act; float slope, eps;
  int rows;
  // bf16 tensor-core path: pre-swizzled weight image, 16 KB blocks [128 n x 64 k] in (n tile, k block) order
  const uint8_t* img; int img_kblocks;      // k blocks per n tile in the image
  int kb0;                                  // first k block used (k0 / 64)
};

// bf16 tcgen05 version of the fused linear (tf_tc.cu).  Requires K % 64 == 0, K <= 512, N % 128 == 0.
bool tf_tc_linear_supported(const LinArgs& a);
int tf_tc_linear(const LinArgs& a, int max_smem, cudaStream_t st);
// image[(nt*kblocks + kb)*16384 + swizzle(n, k)] = bf16(Wt[k*ldo + n])   (zero beyond `in`)
int tf_tc_pack(const float* Wt, int in, int out, int ldo, uint8_t* img, int kblocks, cudaStream_t st);

// masked self attention on tcgen05 (tf_attn_tc.cu): head dim 16, <= 256 particles per jet
int tf_attn_tc(const float* QKV, int ld, int D, int heads, const int* n_real, const int* rowoff, float* A, int lda, float scale,
               int B, int N, int sm_count, cudaStream_t st);

}  // namespace pfm
