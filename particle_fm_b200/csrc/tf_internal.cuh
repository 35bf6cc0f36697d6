// Declarations shared by the droid-transformer translation units (tf_simt.cu, tf_tc.cu); not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "../../include/pfm_b200.h"

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

// ---- handle -----------------------------------------------------------------------------------
struct TfLinear {
  int in = 0, out = 0, ldo = 0;
  float* Wt = nullptr;            // k-major Wt[k*ldo + o]
  float* b = nullptr;
  uint8_t* img = nullptr; int kblocks = 0;
  // training: row-major copy Wrow[o*ldw + k] (dX = dY . W) and where the gradients go in the flat buffer;
  // a linear fed from two state_dict entries (k_linear | v_linear) has two parts split at output row `split`
  float* Wrow = nullptr; int ldw = 0;
  uint8_t* img_bwd = nullptr; int kblocks_bwd = 0;   // bf16 image of Wrow viewed k-major (k = output row): dX on tcgen05
  int n_parts = 0, split = 0;
  size_t gw_off[2] = {0, 0}, gb_off[2] = {0, 0};
};
struct TfLN { int d = 0; float* g = nullptr; float* b = nullptr; size_t gg_off = 0, gb_off = 0; };
struct TfDense { TfLinear l1; TfLN ln; TfLinear l2; };
struct TfLayer { TfLinear qkv_or_q, kv, out; TfLN mha_ln; TfDense dense; TfLN n0, n1, n2; };
struct ParamSlot { int rows, cols; int kind; void* target; int col_off; };   // kind 0: linear W, 1: vector, 2: tokens
struct TfTape;

}  // namespace pfm

struct pfm_tf {
  pfm_tf_cfg cfg;
  int device, sm_count, max_smem;
  bool weights_set;
  pfm::TfDense ctxt, node, outp;
  std::vector<pfm::TfLayer> layers;       // full: L layers; cross: from_0..from_{L-1}, to_0..to_{L-1}
  pfm::TfLN final_norm;
  float* tok0;                       // [ntok, D]
  size_t tok0_goff;
  std::vector<pfm::ParamSlot> slots;      // canonical parameter order
  std::vector<float*> owned;
  // plan + workspaces
  int capB, capBN; size_t cap_rows;
  int *n_real, *rowoff, *n_total, *rowjet, *tokjet; uint16_t* ridx;
  float *xs, *x0, *v, *h, *H1, *QKV, *A, *tok, *tokA, *tokQ, *tokKV, *tokH1, *ctxin, *c1, *ctx, *jb; size_t jb_floats;
  int last_launches;
  int precision;                     // PFM_PREC_FP32 / PFM_PREC_BF16 (tcgen05 linears where the shape allows)
  std::vector<pfm::TfLinear*> all_linears;
  size_t grad_floats;                // size of the flat gradient (slot order)
  pfm::TfTape* tape;                 // training state (tf_train.cu)
};

namespace pfm {
float* tf_alloc(pfm_tf* h, size_t floats);
int tf_launch_linear(pfm_tf* h, const LinArgs& a, int ldo_class, bool allow_tc, cudaStream_t st);
__global__ void plan_count_kernel(const float* __restrict__ mask, int B, int N, int* __restrict__ n_real, uint16_t* __restrict__ ridx);
void tf_tape_destroy(pfm_tf* h);
}  // namespace pfm
