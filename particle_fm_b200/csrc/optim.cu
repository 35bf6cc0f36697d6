// Fused optimizer step over flat buffers: global-norm gradient clipping + AdamW (+ optional EMA of the weights) in two
// launches whatever the number of parameter tensors.
// Replaces, for one training step of the reference (configs/trainer: gradient_clip_val 0.5; configs/model/flow_matching.yaml:3-7
// torch.optim.AdamW; callbacks/ema.py:73-81):
//   torch.nn.utils.clip_grad_norm_  (total_norm = ||g||_2 over all parameters; g *= min(1, max_norm / (total_norm + 1e-6)))
//   torch.optim.AdamW.step          (decoupled weight decay, bias-corrected first / second moments)
//   EMA.apply_ema                   (ema -= (ema - w) * (1 - decay))
// which PyTorch runs as ~20 multi-tensor launches plus ~1.2 ms of host time per step for the 87 parameter tensors of the
// default net.  Bound: HBM, 4 reads + 3 writes of n floats (+ 2 for the EMA): 0.6 M parameters -> a few microseconds.
#include "pfm_internal.cuh"

namespace pfm {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
#pragma unroll
  for (int sft = 16; sft > 0; sft >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sft);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t);
  }
}

struct AdamP { float lr, b1, b2, eps, wd, max_norm, ema_keep; };

__global__ void __launch_bounds__(256) clip_adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                         float* __restrict__ v, float* __restrict__ ema, long long n,
                                                         const float* __restrict__ sumsq, float* __restrict__ norm_out,
                                                         const int* __restrict__ step_dev, int step_host, AdamP a) {
  // the step count lives on the device when the caller replays this launch from a CUDA graph (bumped by step_bump_kernel)
  const float stepf = (float)(step_dev ? *step_dev : step_host);
  const float bc1 = 1.f - powf(a.b1, stepf), bc2_sqrt = sqrtf(1.f - powf(a.b2, stepf));
  float coef = 1.f;
  if (a.max_norm > 0.f) {
    const float total = sqrtf(*sumsq);
    coef = fminf(a.max_norm / (total + 1e-6f), 1.f);       // torch.nn.utils.clip_grad_norm_
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = total;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float w = p[i] * (1.f - a.lr * a.wd);                   // decoupled weight decay (torch.optim.AdamW)
    const float mi = a.b1 * m[i] + (1.f - a.b1) * gi;
    const float vi = a.b2 * v[i] + (1.f - a.b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
    w -= (a.lr / bc1) * (mi / denom);
    p[i] = w;
    if (ema) { const float e = ema[i]; ema[i] = e - (e - w) * a.ema_keep; }     // ema.py:77-81
  }
}

__global__ void step_bump_kernel(int* step) { *step += 1; }

}  // namespace pfm

using namespace pfm;

extern "C" int pfm_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, float max_norm, int step, float* ema,
                              float ema_decay, float* workspace, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !workspace || n <= 0) { set_error("pfm_clip_adamw: bad argument"); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1184) blocks = 1184;                         // 8 CTAs per SM on 148 SMs, grid-stride beyond
  if (max_norm > 0.f) {
    PFM_CUDA_CHECK(cudaMemsetAsync(workspace, 0, sizeof(float), st));
    sumsq_kernel<<<blocks, 256, 0, st>>>(grads, n, workspace);
  }
  AdamP a;
  a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.max_norm = max_norm;
  a.ema_keep = 1.f - ema_decay;
  int* step_dev = nullptr;
  if (step < 1) {               // step <= 0: the count is kept in workspace[2] (an int) and incremented here -- graph-replay safe
    step_dev = reinterpret_cast<int*>(workspace + 2);
    step_bump_kernel<<<1, 1, 0, st>>>(step_dev);
  }
  clip_adamw_kernel<<<blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, ema, n, workspace, workspace + 1, step_dev, step, a);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}
