// PC-Droid-style set transformers, fp32 CUDA cores.
//
//   FullTransformerEncoder      droid_transformer.py:440-548   (configs/model/fm_droid_transformer.yaml)
//   FullCrossAttentionEncoder   droid_transformer.py:622-711   (configs/model/fm_droid_crossattention.yaml)
//
// Padding is skipped: the real particles of all jets are packed into [rows, D] activations in global memory
// (rows = sum of multiplicities); only keys are masked in the reference and every other op is per token, so real
// tokens never see padded ones and their results are exact.  (The reference does not mask its OUTPUT; padded slots
// of the result are returned as 0 here -- generate_data multiplies by the mask anyway.)
//
// One evaluation is a short program of three kernel kinds:
//   tf_linear_kernel   Y = [R +] act(LN(X) . W^T + bias [+ per-jet bias]) for a block of 64 rows: the optional
//                      LayerNorm is applied to the rows on load (every LayerNorm of these networks feeds a linear),
//                      the weights stream through a double-buffered cp.async stage in 32*TC-column tiles, the
//                      context / time columns of the concat-linears are hoisted into per-jet bias tables.
//   tf_attn_*_kernel   softmax(q.k / sqrt(dh)) v per (jet, head) over the jet's REAL tokens only.
//   small kernels      packing, context input, integrator update, unpacking.
// Strict (fp32) path; the tensor-core version of the linear kernel is the next step for these networks.
#include <cstdlib>
#include <cstring>

#include "pfm_internal.cuh"
#include "simt_common.cuh"
#include "tf_internal.cuh"

namespace pfm {

static inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// generic fused linear
// ---------------------------------------------------------------------------------------------
static constexpr int LIN_ROWS = 64;      // rows per CTA = 8 warps x 8 rows
static constexpr int LIN_KC = 16;

template <int TC>
__global__ void __launch_bounds__(kThreads) tf_linear_kernel(const LinArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int RB = 8, NT = 32 * TC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.x * LIN_ROWS;
  if (row0 >= a.rows) return;
  const int Kp = (a.K + 3) & ~3;
  const int lda = Kp + 4;
  float* Xs = smem;                          // [64][lda]
  float* wbuf = smem + LIN_ROWS * lda;       // 2 x [KC][NT]
  // ---- load the row block (zero padded), LayerNorm in place
  for (int i = tid; i < LIN_ROWS * lda; i += kThreads) {
    const int r = i / lda, c = i - r * lda;
    const int row = row0 + r;
    Xs[i] = (row < a.rows && c < a.K) ? a.X[(size_t)row * a.ldx + c] : 0.f;
  }
  __syncthreads();
  if (a.ln_g) {
    for (int r = warp * RB; r < warp * RB + RB; ++r) {
      float* xr = Xs + r * lda;
      float s = 0.f;
      for (int c = lane; c < a.K; c += 32) s += xr[c];
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      const float mean = s / (float)a.K;
      float v = 0.f;
      for (int c = lane; c < a.K; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
      const float rstd = rsqrtf(v / (float)a.K + a.eps);
      for (int c = lane; c < a.K; c += 32) xr[c] = (xr[c] - mean) * rstd * a.ln_g[c] + a.ln_b[c];
    }
    __syncthreads();
  }
  const float* Arow = Xs + (size_t)(warp * RB) * lda;
  const int n_chunks = (Kp + LIN_KC - 1) / LIN_KC;
  for (int n0 = 0; n0 < a.N; n0 += NT) {
    float acc[RB][TC];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int i = 0; i < TC; ++i) acc[r][i] = 0.f;
    // prologue: chunk 0 of this column tile
    auto stage = [&](int c, float* dst) {
      const int k0 = c * LIN_KC;
      const int kc = (Kp - k0) < LIN_KC ? (Kp - k0) : LIN_KC;
      const int n16 = kc * (NT / 4);
      for (int i = tid; i < n16; i += kThreads) {
        const int kk = i / (NT / 4), q = i - kk * (NT / 4);
        cp_async16(dst + kk * NT + q * 4, a.Wt + (size_t)(k0 + kk) * a.ldo + n0 + q * 4);
      }
      cp_async_commit();
    };
    stage(0, wbuf);
    for (int c = 0; c < n_chunks; ++c) {
      if (c + 1 < n_chunks) { stage(c + 1, wbuf + ((c + 1) & 1) * LIN_KC * NT); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncthreads();
      const float* wb = wbuf + (c & 1) * LIN_KC * NT;
      const int k0 = c * LIN_KC;
      const int kc = (Kp - k0) < LIN_KC ? (Kp - k0) : LIN_KC;
      for (int kk = 0; kk < kc; kk += 4) {
        float4 av[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) av[r] = *reinterpret_cast<const float4*>(Arow + (size_t)r * lda + k0 + kk);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float w[TC];
#pragma unroll
          for (int i = 0; i < TC; ++i) w[i] = wb[(kk + q) * NT + lane + 32 * i];
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float x = q == 0 ? av[r].x : (q == 1 ? av[r].y : (q == 2 ? av[r].z : av[r].w));
#pragma unroll
            for (int i = 0; i < TC; ++i) acc[r][i] = fmaf(x, w[i], acc[r][i]);
          }
        }
      }
      __syncthreads();
    }
    // ---- epilogue of this column tile
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int row = row0 + warp * RB + r;
      if (row >= a.rows) continue;
      const float* jbrow = a.jb ? a.jb + (size_t)(a.rowjet ? a.rowjet[row] : row) * a.jb_stride : nullptr;
#pragma unroll
      for (int i = 0; i < TC; ++i) {
        const int o = n0 + lane + 32 * i;
        if (o >= a.N) continue;
        float v = acc[r][i];
        if (a.bias) v += a.bias[o];
        if (jbrow) v += jbrow[o];
        if (a.act) v = v > 0.f ? v : v * a.slope;
        if (a.R) v += a.R[(size_t)row * a.ldr + o];
        a.Y[(size_t)row * a.ldy + o] = v;
      }
    }
  }
}

// Narrow output (N <= 8, e.g. the 512 -> 3 output projection): one warp per row, LayerNorm + N dot products by warp
// shuffles -- memory bound instead of wasting a 64-column tile on 3 outputs.
__global__ void tf_linear_smalln_kernel(const LinArgs a) {
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                      // [K][8]
  float* gs = sm + (size_t)a.K * 8;    // [K] LayerNorm gain, [K] shift
  float* bs = gs + a.K;
  for (int i = threadIdx.x; i < a.K * 8; i += blockDim.x) {
    const int c = i >> 3, o = i & 7;
    Ws[i] = o < a.N ? a.Wt[(size_t)c * a.ldo + o] : 0.f;
  }
  for (int i = threadIdx.x; i < a.K; i += blockDim.x) { gs[i] = a.ln_g ? a.ln_g[i] : 1.f; bs[i] = a.ln_g ? a.ln_b[i] : 0.f; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < a.rows; row += gridDim.x * wpb) {
    const float* x = a.X + (size_t)row * a.ldx;
    float xv[16];                      // K <= 512: 16 values per lane, columns lane + 32 i (coalesced)
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { const int c = lane + 32 * i; xv[i] = c < a.K ? __ldg(x + c) : 0.f; s += xv[i]; }
    float mean = 0.f, rstd = 1.f;
    if (a.ln_g) {
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
      mean = s / (float)a.K;
      float v = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) if (lane + 32 * i < a.K) { const float d = xv[i] - mean; v = fmaf(d, d, v); }
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
      rstd = rsqrtf(v / (float)a.K + a.eps);
    }
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < a.K) {
        const float xn = (xv[i] - mean) * rstd * gs[c] + bs[c];
        const float4 w0 = *reinterpret_cast<const float4*>(Ws + c * 8), w1 = *reinterpret_cast<const float4*>(Ws + c * 8 + 4);
        acc[0] = fmaf(xn, w0.x, acc[0]); acc[1] = fmaf(xn, w0.y, acc[1]); acc[2] = fmaf(xn, w0.z, acc[2]); acc[3] = fmaf(xn, w0.w, acc[3]);
        acc[4] = fmaf(xn, w1.x, acc[4]); acc[5] = fmaf(xn, w1.y, acc[5]); acc[6] = fmaf(xn, w1.z, acc[6]); acc[7] = fmaf(xn, w1.w, acc[7]);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o)
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], sh);
    if (lane < a.N) {
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o)
        if (o == lane) v = acc[o];
      if (a.bias) v += a.bias[lane];
      if (a.jb) v += a.jb[(size_t)(a.rowjet ? a.rowjet[row] : row) * a.jb_stride + lane];
      if (a.act) v = v > 0.f ? v : v * a.slope;
      if (a.R) v += a.R[(size_t)row * a.ldr + lane];
      a.Y[(size_t)row * a.ldy + lane] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// attention kernels (head dim dh <= 16)
// ---------------------------------------------------------------------------------------------
static constexpr int DH_MAX = 16;

// self attention: grid (B, heads); block = queries (real tokens of the jet, looped); K/V of the head in smem.
// DH = head dim (compile time: 16-byte shared loads); scores in log2 units, online softmax with one rescale per 8 keys.
template <int DH>
__global__ void tf_attn_self_kernel(const float* __restrict__ QKV, int ld, int D, const int* __restrict__ n_real,
                                    const int* __restrict__ rowoff, float* __restrict__ A, int lda, float scale) {
  extern __shared__ __align__(16) float sm[];
  const int jet = blockIdx.x, head = blockIdx.y;
  const int n = n_real[jet], r0 = rowoff[jet];
  if (n == 0) return;
  float* Ks = sm;                     // [n][DH]
  float* Vs = sm + (size_t)n * DH;    // [n][DH]
  for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
    const int t = i / (DH / 4), d4 = i - t * (DH / 4);
    const float* src = QKV + (size_t)(r0 + t) * ld + head * DH + d4 * 4;
    *reinterpret_cast<float4*>(Ks + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(src + D);
    *reinterpret_cast<float4*>(Vs + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(src + 2 * D);
  }
  __syncthreads();
  const float sl2 = scale * 1.4426950408889634f;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    float q[DH], o[DH];
    const float* qs = QKV + (size_t)(r0 + t) * ld + head * DH;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      const float4 v = *reinterpret_cast<const float4*>(qs + d4 * 4);
      q[d4 * 4 + 0] = v.x * sl2; q[d4 * 4 + 1] = v.y * sl2; q[d4 * 4 + 2] = v.z * sl2; q[d4 * 4 + 3] = v.w * sl2;
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < n; k0 += 8) {
      float sc[8];
      float bm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j < n ? k0 + j : n - 1;
        const float4* kp = reinterpret_cast<const float4*>(Ks + k * DH);
        float sv = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 kv = kp[d4];
          sv = fmaf(q[d4 * 4 + 0], kv.x, sv); sv = fmaf(q[d4 * 4 + 1], kv.y, sv);
          sv = fmaf(q[d4 * 4 + 2], kv.z, sv); sv = fmaf(q[d4 * 4 + 3], kv.w, sv);
        }
        sc[j] = k0 + j < n ? sv : -INFINITY;
        bm = fmaxf(bm, sc[j]);
      }
      const float mn = fmaxf(m, bm);
      const float corr = exp2f(m - mn);
      l *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = k0 + j < n ? k0 + j : n - 1;
        const float pj = exp2f(sc[j] - mn);
        l += pj;
        const float4* vp = reinterpret_cast<const float4*>(Vs + k * DH);
#pragma unroll
        for (int d4 = 0; d4 < DH / 4; ++d4) {
          const float4 vv = vp[d4];
          o[d4 * 4 + 0] = fmaf(pj, vv.x, o[d4 * 4 + 0]); o[d4 * 4 + 1] = fmaf(pj, vv.y, o[d4 * 4 + 1]);
          o[d4 * 4 + 2] = fmaf(pj, vv.z, o[d4 * 4 + 2]); o[d4 * 4 + 3] = fmaf(pj, vv.w, o[d4 * 4 + 3]);
        }
      }
      m = mn;
    }
    const float inv = 1.f / l;
    float* dst = A + (size_t)(r0 + t) * lda + head * DH;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4)
      *reinterpret_cast<float4*>(dst + d4 * 4) = make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
  }
}

// tokens <- sequence: grid B; thread = (head, token query); keys = the jet's real particles (KV [rows, 2D]: k | v)
__global__ void tf_attn_from_kernel(const float* __restrict__ Qt, int ldq, const float* __restrict__ KV, int ldkv, int D, int dh,
                                    int heads, int ntok, const int* __restrict__ n_real, const int* __restrict__ rowoff,
                                    float* __restrict__ At, int lda, float scale) {
  const int jet = blockIdx.x;
  const int n = n_real[jet], r0 = rowoff[jet];
  for (int i = threadIdx.x; i < heads * ntok; i += blockDim.x) {
    const int head = i / ntok, tq = i - head * ntok;
    const int qrow = jet * ntok + tq;
    float q[DH_MAX], o[DH_MAX];
#pragma unroll
    for (int d = 0; d < DH_MAX; ++d) { q[d] = d < dh ? Qt[(size_t)qrow * ldq + head * dh + d] * (scale * 1.4426950408889634f) : 0.f; o[d] = 0.f; }
    float m = -INFINITY, l = 0.f;
    for (int k = 0; k < n; ++k) {
      const float* kp = KV + (size_t)(r0 + k) * ldkv + head * dh;
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH_MAX; ++d)
        if (d < dh) s = fmaf(q[d], kp[d], s);
      const float mn = fmaxf(m, s);
      const float corr = exp2f(m - mn), p = exp2f(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int d = 0; d < DH_MAX; ++d)
        if (d < dh) o[d] = fmaf(p, kp[D + d], o[d] * corr);
      m = mn;
    }
    const float inv = 1.f / l;        // n == 0 -> 0/0 = NaN, like softmax over an empty key set in the reference
#pragma unroll
    for (int d = 0; d < DH_MAX; ++d)
      if (d < dh) At[(size_t)qrow * lda + head * dh + d] = o[d] * inv;
  }
}

// sequence <- tokens: thread = (row, head); keys = the ntok tokens of the row's jet (KVt [B*ntok, 2D]), no mask
__global__ void tf_attn_to_kernel(const float* __restrict__ Qs, int ldq, const float* __restrict__ KVt, int ldkv, int D, int dh,
                                  int heads, int ntok, const int* __restrict__ rowjet, int rows, float* __restrict__ As,
                                  int lda, float scale) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * heads) return;
  const int row = idx / heads, head = idx - row * heads;
  const int jet = rowjet[row];
  float q[DH_MAX], o[DH_MAX];
#pragma unroll
  for (int d = 0; d < DH_MAX; ++d) { q[d] = d < dh ? Qs[(size_t)row * ldq + head * dh + d] * (scale * 1.4426950408889634f) : 0.f; o[d] = 0.f; }
  float m = -INFINITY, l = 0.f;
  for (int k = 0; k < ntok; ++k) {
    const float* kp = KVt + (size_t)(jet * ntok + k) * ldkv + head * dh;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < DH_MAX; ++d)
      if (d < dh) s = fmaf(q[d], kp[d], s);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    l = l * corr + p;
#pragma unroll
    for (int d = 0; d < DH_MAX; ++d)
      if (d < dh) o[d] = fmaf(p, kp[D + d], o[d] * corr);
    m = mn;
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int d = 0; d < DH_MAX; ++d)
    if (d < dh) As[(size_t)row * lda + head * dh + d] = o[d] * inv;
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
__global__ void tf_rowoff_kernel(const int* __restrict__ n_real, int B, int* __restrict__ rowoff, int* __restrict__ n_total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int t = 0;
    for (int j = 0; j < B; ++j) { rowoff[j] = t; t += n_real[j]; }
    *n_total = t;
  }
}

// xs[row][:] = x[jet][particle][:], rowjet[row] = jet      (one warp per jet)
__global__ void tf_pack_kernel(const float* __restrict__ x, const int* __restrict__ n_real, const uint16_t* __restrict__ ridx,
                               const int* __restrict__ rowoff, int B, int N, int F, float* __restrict__ xs, int* __restrict__ rowjet) {
  const int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (jet >= B) return;
  const int n = n_real[jet], r0 = rowoff[jet];
  for (int r = lane; r < n; r += 32) rowjet[r0 + r] = jet;
  for (int i = lane; i < n * F; i += 32) {
    const int r = i / F, f = i - r * F;
    xs[(size_t)(r0 + r) * F + f] = x[((size_t)jet * N + ridx[(size_t)jet * N + r]) * F + f];
  }
}

__global__ void tf_unpack_kernel(const float* __restrict__ xs, const int* __restrict__ n_real, const uint16_t* __restrict__ ridx,
                                 const int* __restrict__ rowoff, int B, int N, int F, float* __restrict__ out) {
  const int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (jet >= B) return;
  const int n = n_real[jet], r0 = rowoff[jet];
  float* dst = out + (size_t)jet * N * F;
  for (int i = lane; i < N * F; i += 32) dst[i] = 0.f;
  __syncwarp();
  for (int i = lane; i < n * F; i += 32) {
    const int r = i / F, f = i - r * F;
    dst[(size_t)ridx[(size_t)jet * N + r] * F + f] = xs[(size_t)(r0 + r) * F + f];
  }
}

// context input rows: [t_code(row or shared) | cond(jet)]
__global__ void tf_ctxin_kernel(const float* __restrict__ t_code, int t_stride, int T, const float* __restrict__ cond, int C,
                                int rows, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = T + C;
  if (idx >= rows * W) return;
  const int r = idx / W, c = idx - r * W;
  out[idx] = c < T ? t_code[(size_t)r * t_stride + c] : cond[(size_t)r * C + (c - T)];
}

__global__ void tf_tokens_kernel(const float* __restrict__ tok0, int ntokD, int B, float* __restrict__ tok, int* __restrict__ tokjet,
                                 int ntok) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < B * ntokD) tok[idx] = tok0[idx % ntokD];
  if (idx < B * ntok) tokjet[idx] = idx / ntok;
}

// fixed-step integrator on the packed state (torchdyn order, oracle/ode_oracle.py): k = -v
//   mode 0 (Euler / midpoint second stage): x0 += dt*k; xc = x0      mode 1 (midpoint first stage): xc = x0 + 0.5*dt*k
__global__ void tf_update_kernel(float* __restrict__ x0, float* __restrict__ xc, const float* __restrict__ v, const float* __restrict__ dt,
                                 int step, int mode, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = dt[step], k = -v[i];
  if (mode == 1) {
    xc[i] = __fadd_rn(x0[i], __fmul_rn(__fmul_rn(0.5f, d), k));
  } else {
    const float xn = __fadd_rn(x0[i], __fmul_rn(d, k));
    x0[i] = xn;
    xc[i] = xn;
  }
}

}  // namespace pfm

// =============================================================================================
// handle (struct pfm_tf: tf_internal.cuh)
// =============================================================================================
namespace pfm {

float* tf_alloc(pfm_tf* h, size_t floats) {
  float* p = nullptr;
  if (cudaMalloc(&p, sizeof(float) * (floats ? floats : 1)) != cudaSuccess) return nullptr;
  cudaMemset(p, 0, sizeof(float) * (floats ? floats : 1));
  h->owned.push_back(p);
  return p;
}

static bool tf_make_linear(pfm_tf* h, TfLinear* L, int out, int in) {
  L->in = in; L->out = out; L->ldo = round_up_i(out, 64);
  L->Wt = tf_alloc(h, (size_t)(in + 4) * L->ldo);
  L->b = tf_alloc(h, L->ldo);
  L->ldw = round_up_i(in, 256);          // a multiple of the widest column tile of tf_linear_kernel (backward: dX = dY . Wrow)
  L->Wrow = tf_alloc(h, (size_t)(out + 4) * L->ldw + 64);
  if (!L->Wrow) return false;
  if (out % 128 == 0 && in >= 64) {      // eligible for the tensor-core path: bf16 image of 16 KB [128 n x 64 k] blocks
    L->kblocks = (in + 63) / 64;
    L->img = reinterpret_cast<uint8_t*>(tf_alloc(h, (size_t)(out / 128) * L->kblocks * 16384 / sizeof(float)));
    if (!L->img) return false;
  }
  if (out % 64 == 0 && in >= 128) {      // backward dX = dY . W on the tensor cores: K = out, N = input columns
    L->kblocks_bwd = out / 64;
    L->img_bwd = reinterpret_cast<uint8_t*>(tf_alloc(h, (size_t)(L->ldw / 128) * L->kblocks_bwd * 16384 / sizeof(float)));
    if (!L->img_bwd) return false;
  }
  h->all_linears.push_back(L);
  return L->Wt && L->b;
}
static bool tf_make_ln(pfm_tf* h, TfLN* n, int d) {
  n->d = d; n->g = tf_alloc(h, d); n->b = tf_alloc(h, d);
  return n->g && n->b;
}
static bool tf_make_dense(pfm_tf* h, TfDense* d, int inpt, int ctxt, int hddn, int outp) {
  return tf_make_linear(h, &d->l1, hddn, inpt + ctxt) && tf_make_ln(h, &d->ln, hddn) && tf_make_linear(h, &d->l2, outp, hddn);
}
static void slot_linear(pfm_tf* h, TfLinear* L, int col_off = 0, int out = -1) {
  const int o = out < 0 ? L->out : out;
  const int part = L->n_parts++;
  if (part == 1) L->split = col_off;
  L->gw_off[part] = h->grad_floats; h->grad_floats += (size_t)o * L->in;
  L->gb_off[part] = h->grad_floats; h->grad_floats += (size_t)o;
  h->slots.push_back({o, L->in, 0, L, col_off});
  h->slots.push_back({o, 1, 1, L->b + col_off, 0});
}
static void slot_ln(pfm_tf* h, TfLN* n) {
  n->gg_off = h->grad_floats; h->grad_floats += (size_t)n->d;
  n->gb_off = h->grad_floats; h->grad_floats += (size_t)n->d;
  h->slots.push_back({n->d, 1, 1, n->g, 0});
  h->slots.push_back({n->d, 1, 1, n->b, 0});
}
static void slot_dense(pfm_tf* h, TfDense* d) { slot_linear(h, &d->l1); slot_ln(h, &d->ln); slot_linear(h, &d->l2); }

// Wt[k*ldo + col_off + o] = W[o*in + k]
__global__ void tf_transpose_kernel(const float* __restrict__ W, float* __restrict__ Wt, int out, int in, int ldo, int col_off) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out * in) return;
  const int k = idx / out, o = idx - k * out;
  Wt[(size_t)k * ldo + col_off + o] = W[(size_t)o * in + k];
}

template <int TC>
static int launch_linear_tc(const pfm_tf* h, const LinArgs& a, cudaStream_t st) {
  const int Kp = (a.K + 3) & ~3;
  const size_t smem = sizeof(float) * ((size_t)LIN_ROWS * (Kp + 4) + 2 * LIN_KC * 32 * TC);
  if ((int)smem > h->max_smem) { set_error("transformer linear: K=%d needs %zu B of shared memory", a.K, smem); return PFM_ERR_UNSUPPORTED; }
  auto kern = tf_linear_kernel<TC>;
  PFM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(a.rows + LIN_ROWS - 1) / LIN_ROWS, kThreads, smem, st>>>(a);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// Y[rows, N] = [R +] act(LN(X[rows, K]) . W[k0 : k0+K]^T + bias [+ jb[rowjet]])
int tf_launch_linear(pfm_tf* h, const LinArgs& a, int ldo_class, bool allow_tc, cudaStream_t st) {
  h->last_launches++;
  // tensor cores for the per-token linears; the per-jet context / bias tables (a handful of rows) stay fp32
  if (allow_tc && h->precision == PFM_PREC_BF16 && a.rows >= 256 && tf_tc_linear_supported(a)) return tf_tc_linear(a, h->max_smem, st);
  if (a.N <= 8 && a.rows >= 64 && a.K <= 512) {
    const int blocks = (a.rows + 63) / 64 < 4 * h->sm_count ? (a.rows + 63) / 64 : 4 * h->sm_count;
    tf_linear_smalln_kernel<<<blocks, 256, sizeof(float) * (size_t)a.K * 10, st>>>(a);
    PFM_CUDA_CHECK(cudaGetLastError());
    return PFM_OK;
  }
  if (ldo_class % 256 == 0) return launch_linear_tc<8>(h, a, st);
  if (ldo_class % 128 == 0) return launch_linear_tc<4>(h, a, st);
  return launch_linear_tc<2>(h, a, st);
}

int run_linear(pfm_tf* h, cudaStream_t st, const float* X, int ldx, int K, const TfLN* ln, const TfLinear& L, int k0,
               bool use_bias, const float* jb, int jb_stride, const int* rowjet, const float* R, int ldr, float* Y,
               int ldy, int act, int rows) {
  if (rows <= 0) return PFM_OK;
  LinArgs a;
  a.X = X; a.ldx = ldx; a.K = K;
  a.ln_g = ln ? ln->g : nullptr; a.ln_b = ln ? ln->b : nullptr;
  a.Wt = L.Wt + (size_t)k0 * L.ldo; a.ldo = L.ldo; a.N = L.out;
  a.bias = use_bias ? L.b : nullptr;
  a.jb = jb; a.jb_stride = jb_stride; a.rowjet = rowjet;
  a.R = R; a.ldr = ldr; a.Y = Y; a.ldy = ldy;
  a.act = act; a.slope = h->cfg.neg_slope; a.eps = h->cfg.ln_eps;
  a.rows = rows;
  a.img = L.img; a.img_kblocks = L.kblocks; a.kb0 = k0 / 64;
  return tf_launch_linear(h, a, L.ldo, (k0 % 64) == 0, st);
}

static int tf_ensure(pfm_tf* h, int B, int N, int ctx_rows = 0) {
  const pfm_tf_cfg& c = h->cfg;
  {   // context rows: one per jet, or one per evaluation of an integration when the jets share the time code
    const size_t need = (size_t)(B > ctx_rows ? B : ctx_rows);
    if (need > h->cap_rows) {
      for (float** p : {&h->ctxin, &h->c1, &h->ctx, &h->jb}) { if (*p) cudaFree(*p); *p = nullptr; }
      const int n_tables = 2 + (int)h->layers.size();
      const int hmax = c.embd_hddn > c.dense_hddn ? c.embd_hddn : c.dense_hddn;
      h->jb_floats = (size_t)hmax;
      PFM_CUDA_CHECK(cudaMalloc(&h->ctxin, sizeof(float) * need * (c.t_dim + c.cond_dim)));
      PFM_CUDA_CHECK(cudaMalloc(&h->c1, sizeof(float) * need * c.embd_hddn));
      PFM_CUDA_CHECK(cudaMalloc(&h->ctx, sizeof(float) * need * c.ctxt_out));
      PFM_CUDA_CHECK(cudaMalloc(&h->jb, sizeof(float) * need * hmax * n_tables));
      h->cap_rows = need;
    }
  }
  if (B > h->capB) {
    for (void* p : {(void*)h->n_real, (void*)h->rowoff, (void*)h->tokjet}) if (p) cudaFree(p);
    PFM_CUDA_CHECK(cudaMalloc(&h->n_real, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&h->rowoff, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&h->tokjet, sizeof(int) * B * c.num_tokens));
    if (!h->n_total) PFM_CUDA_CHECK(cudaMalloc(&h->n_total, sizeof(int)));
    h->capB = B;
    for (float** p : {&h->tok, &h->tokA, &h->tokQ, &h->tokKV, &h->tokH1}) { if (*p) cudaFree(*p); *p = nullptr; }
    const size_t T4 = (size_t)B * c.num_tokens;
    const int D = c.model_dim;
    PFM_CUDA_CHECK(cudaMalloc(&h->tok, sizeof(float) * T4 * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->tokA, sizeof(float) * T4 * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->tokQ, sizeof(float) * T4 * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->tokKV, sizeof(float) * T4 * 2 * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->tokH1, sizeof(float) * T4 * c.dense_hddn));
  }
  if ((long long)B * N > h->capBN) {
    for (void* p : {(void*)h->ridx, (void*)h->rowjet}) if (p) cudaFree(p);
    for (float** p : {&h->xs, &h->x0, &h->v, &h->h, &h->H1, &h->QKV, &h->A}) { if (*p) cudaFree(*p); *p = nullptr; }
    const size_t rows = (size_t)B * N;
    const int D = c.model_dim;
    const int hmax = c.embd_hddn > c.dense_hddn ? c.embd_hddn : c.dense_hddn;
    PFM_CUDA_CHECK(cudaMalloc(&h->ridx, sizeof(uint16_t) * rows));
    PFM_CUDA_CHECK(cudaMalloc(&h->rowjet, sizeof(int) * rows));
    PFM_CUDA_CHECK(cudaMalloc(&h->xs, sizeof(float) * rows * c.feats));
    PFM_CUDA_CHECK(cudaMalloc(&h->x0, sizeof(float) * rows * c.feats));
    PFM_CUDA_CHECK(cudaMalloc(&h->v, sizeof(float) * rows * c.feats));
    PFM_CUDA_CHECK(cudaMalloc(&h->h, sizeof(float) * rows * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->H1, sizeof(float) * rows * hmax));
    PFM_CUDA_CHECK(cudaMalloc(&h->QKV, sizeof(float) * rows * 3 * D));
    PFM_CUDA_CHECK(cudaMalloc(&h->A, sizeof(float) * rows * D));
    h->capBN = B * N;
  }
  return PFM_OK;
}

// Context vector and hoisted bias tables for `Bc` context rows (jets, or -- when every jet shares the time and there is
// no conditioning -- the evaluations of a whole integration at once): table i, row r at  h->jb + (i*cap + r)*hmax.
static int tf_context(pfm_tf* h, cudaStream_t st, const float* t_code, int t_stride, const float* cond, int Bc, int cap) {
  const pfm_tf_cfg& c = h->cfg;
  const int D = c.model_dim, T = c.t_dim, C = c.cond_dim, F = c.feats, CO = c.ctxt_out;
  const int hmax = (int)h->jb_floats;
  const int t_in = c.add_time_to_input ? T : 0;
  int rc;
  tf_ctxin_kernel<<<(Bc * (T + C) + 255) / 256, 256, 0, st>>>(t_code, t_stride, T, cond, C, Bc, h->ctxin);
  h->last_launches++;
  if ((rc = run_linear(h, st, h->ctxin, T + C, T + C, nullptr, h->ctxt.l1, 0, true, nullptr, 0, nullptr, nullptr, 0, h->c1,
                       c.embd_hddn, 1, Bc)) != PFM_OK) return rc;
  if ((rc = run_linear(h, st, h->c1, c.embd_hddn, c.embd_hddn, &h->ctxt.ln, h->ctxt.l2, 0, true, nullptr, 0, nullptr, nullptr, 0,
                       h->ctx, CO, 0, Bc)) != PFM_OK) return rc;
  auto table = [&](int i) { return h->jb + (size_t)i * cap * hmax; };
  auto jet_bias = [&](const TfDense& d, int inpt, float* dst) -> int {     // b + W[:, inpt:inpt+CO] . ctx
    return run_linear(h, st, h->ctx, CO, CO, nullptr, d.l1, inpt, true, nullptr, 0, nullptr, nullptr, 0, dst, hmax, 0, Bc);
  };
  if ((rc = jet_bias(h->node, t_in + F, table(0))) != PFM_OK) return rc;
  if (t_in > 0)      // time columns of the per-particle input, hoisted: += W[:, 0:T] . t_code
    if ((rc = run_linear(h, st, h->ctxin, T + C, T, nullptr, h->node.l1, 0, false, nullptr, 0, nullptr, table(0), hmax, table(0),
                         hmax, 0, Bc)) != PFM_OK) return rc;
  if ((rc = jet_bias(h->outp, D, table(1))) != PFM_OK) return rc;
  for (size_t l = 0; l < h->layers.size(); ++l)
    if ((rc = jet_bias(h->layers[l].dense, D, table(2 + (int)l))) != PFM_OK) return rc;
  return PFM_OK;
}

// One evaluation of the network on the packed state h->xs -> h->v.   t_code: one row (t_stride 0) or one per jet.
// pre_ev >= 0: the tables were filled by tf_context for all evaluations of the integration (row pre_ev, capacity pre_cap).
static int tf_eval(pfm_tf* h, cudaStream_t st, const float* t_code, int t_rows, const float* cond, int B, int N, int rows,
                   int pre_ev = -1, int pre_cap = 0) {
  const pfm_tf_cfg& c = h->cfg;
  const int D = c.model_dim, T = c.t_dim, C = c.cond_dim, F = c.feats, CO = c.ctxt_out;
  const bool per_jet = (t_rows == B && B > 1) || C > 0;
  const int Bc = per_jet ? B : 1;
  const int t_stride = (t_rows == B && B > 1) ? T : 0;
  const int hmax = (int)h->jb_floats;
  const int jbs = per_jet ? hmax : 0;                 // row stride of the per-jet bias tables
  int rc;
  const int cap = pre_ev >= 0 ? pre_cap : B;
  if (pre_ev < 0 && (rc = tf_context(h, st, t_code, t_stride, cond, Bc, cap)) != PFM_OK) return rc;
  auto table = [&](int i) { return h->jb + ((size_t)i * cap + (pre_ev >= 0 ? pre_ev : 0)) * hmax; };
  const int* rj = per_jet ? h->rowjet : nullptr;       // shared tables: stride 0, any row index works
  const int* tj = per_jet ? h->tokjet : nullptr;
  auto dense_tail = [&](const TfDense& d, const TfLN* pre, const float* X, int K, float* jbt, const int* jets, float* Hbuf,
                        const float* R, float* Y, int ldy, int nrows) -> int {
    int r = run_linear(h, st, X, K, K, pre, d.l1, d.l1.in - CO - K >= 0 ? (d.l1.in - CO - K) : 0, false, jbt, jbs, jets, nullptr, 0,
                       Hbuf, d.l1.out, 1, nrows);
    if (r != PFM_OK) return r;
    return run_linear(h, st, Hbuf, d.l1.out, d.l1.out, &d.ln, d.l2, 0, true, nullptr, 0, nullptr, R, ldy, Y, ldy, 0, nrows);
  };
  // ---- node embedding: per-particle columns are the F features (time columns hoisted)
  if ((rc = dense_tail(h->node, nullptr, h->xs, F, table(0), rj, h->H1, nullptr, h->h, D, rows)) != PFM_OK) return rc;
  const int dh = D / c.num_heads;
  const float scale = 1.f / sqrtf((float)dh);
  if (c.kind == 0) {
    for (int l = 0; l < c.num_layers; ++l) {
      TfLayer& Ly = h->layers[l];
      if ((rc = run_linear(h, st, h->h, D, D, &Ly.n1, Ly.qkv_or_q, 0, true, nullptr, 0, nullptr, nullptr, 0, h->QKV, 3 * D, 0,
                           rows)) != PFM_OK) return rc;
      if (h->precision == PFM_PREC_BF16 && dh == 16 && N <= 256 && !getenv("PFM_TF_SIMT_ATTN")) {
        if ((rc = tf_attn_tc(h->QKV, 3 * D, D, c.num_heads, h->n_real, h->rowoff, h->A, D, scale, B, N, h->sm_count, st)) != PFM_OK) return rc;
      } else if (dh == 16)
        tf_attn_self_kernel<16><<<dim3(B, c.num_heads), 128, sizeof(float) * 2 * (size_t)N * dh, st>>>(h->QKV, 3 * D, D, h->n_real,
                                                                                                  h->rowoff, h->A, D, scale);
      else if (dh == 8)
        tf_attn_self_kernel<8><<<dim3(B, c.num_heads), 128, sizeof(float) * 2 * (size_t)N * dh, st>>>(h->QKV, 3 * D, D, h->n_real,
                                                                                                 h->rowoff, h->A, D, scale);
      else if (dh == 4)
        tf_attn_self_kernel<4><<<dim3(B, c.num_heads), 128, sizeof(float) * 2 * (size_t)N * dh, st>>>(h->QKV, 3 * D, D, h->n_real,
                                                                                                 h->rowoff, h->A, D, scale);
      else { set_error("self attention: head dim %d not supported (4, 8, 16)", dh); return PFM_ERR_UNSUPPORTED; }
      h->last_launches++;
      if ((rc = run_linear(h, st, h->A, D, D, &Ly.mha_ln, Ly.out, 0, true, nullptr, 0, nullptr, h->h, D, h->h, D, 0, rows)) != PFM_OK)
        return rc;
      if ((rc = dense_tail(Ly.dense, &Ly.n2, h->h, D, table(2 + l), rj, h->H1, h->h, h->h, D, rows)) != PFM_OK) return rc;
    }
  } else {
    const int nt = c.num_tokens, TR = B * nt;
    tf_tokens_kernel<<<(TR * D + 255) / 256, 256, 0, st>>>(h->tok0, nt * D, B, h->tok, h->tokjet, nt);
    h->last_launches++;
    for (int l = 0; l < c.num_layers; ++l) {
      TfLayer& Fr = h->layers[l];
      TfLayer& To = h->layers[c.num_layers + l];
      // tokens <- sequence
      if ((rc = run_linear(h, st, h->tok, D, D, &Fr.n1, Fr.qkv_or_q, 0, true, nullptr, 0, nullptr, nullptr, 0, h->tokQ, D, 0, TR)) != PFM_OK) return rc;
      if ((rc = run_linear(h, st, h->h, D, D, &Fr.n0, Fr.kv, 0, true, nullptr, 0, nullptr, nullptr, 0, h->QKV, 2 * D, 0, rows)) != PFM_OK) return rc;
      tf_attn_from_kernel<<<B, 64, 0, st>>>(h->tokQ, D, h->QKV, 2 * D, D, dh, c.num_heads, nt, h->n_real, h->rowoff, h->tokA, D, scale);
      h->last_launches++;
      if ((rc = run_linear(h, st, h->tokA, D, D, &Fr.mha_ln, Fr.out, 0, true, nullptr, 0, nullptr, h->tok, D, h->tok, D, 0, TR)) != PFM_OK) return rc;
      if ((rc = dense_tail(Fr.dense, &Fr.n2, h->tok, D, table(2 + l), tj, h->tokH1, h->tok, h->tok, D, TR)) != PFM_OK) return rc;
      // sequence <- tokens
      if ((rc = run_linear(h, st, h->h, D, D, &To.n1, To.qkv_or_q, 0, true, nullptr, 0, nullptr, nullptr, 0, h->A, D, 0, rows)) != PFM_OK) return rc;
      if ((rc = run_linear(h, st, h->tok, D, D, &To.n0, To.kv, 0, true, nullptr, 0, nullptr, nullptr, 0, h->tokKV, 2 * D, 0, TR)) != PFM_OK) return rc;
      tf_attn_to_kernel<<<(rows * c.num_heads + 255) / 256, 256, 0, st>>>(h->A, D, h->tokKV, 2 * D, D, dh, c.num_heads, nt, h->rowjet, rows,
                                                                        h->QKV, D, scale);
      h->last_launches++;
      if ((rc = run_linear(h, st, h->QKV, D, D, &To.mha_ln, To.out, 0, true, nullptr, 0, nullptr, h->h, D, h->h, D, 0, rows)) != PFM_OK) return rc;
      if ((rc = dense_tail(To.dense, &To.n2, h->h, D, table(2 + c.num_layers + l), rj, h->H1, h->h, h->h, D, rows)) != PFM_OK) return rc;
    }
  }
  // ---- output embedding (the full transformer applies its final LayerNorm first)
  const TfLN* fin = c.kind == 0 ? &h->final_norm : nullptr;
  if ((rc = run_linear(h, st, h->h, D, D, fin, h->outp.l1, 0, false, table(1), jbs, rj, nullptr, 0, h->H1, c.embd_hddn, 1, rows)) != PFM_OK)
    return rc;
  if ((rc = run_linear(h, st, h->H1, c.embd_hddn, c.embd_hddn, &h->outp.ln, h->outp.l2, 0, true, nullptr, 0, nullptr, nullptr, 0, h->v, F,
                       0, rows)) != PFM_OK) return rc;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

static int tf_plan(pfm_tf* h, cudaStream_t st, const float* x, const float* mask, int B, int N, int* rows_out, int ctx_rows = 0) {
  int rc = tf_ensure(h, B, N, ctx_rows);
  if (rc != PFM_OK) return rc;
  plan_count_kernel<<<(B + 7) / 8, 256, 0, st>>>(mask, B, N, h->n_real, h->ridx);
  tf_rowoff_kernel<<<1, 32, 0, st>>>(h->n_real, B, h->rowoff, h->n_total);
  tf_pack_kernel<<<(B + 7) / 8, 256, 0, st>>>(x, h->n_real, h->ridx, h->rowoff, B, N, h->cfg.feats, h->xs, h->rowjet);
  h->last_launches += 3;
  int rows = 0;
  PFM_CUDA_CHECK(cudaMemcpyAsync(&rows, h->n_total, sizeof(int), cudaMemcpyDeviceToHost, st));
  PFM_CUDA_CHECK(cudaStreamSynchronize(st));       // the grids of the evaluation are sized by the number of real particles
  *rows_out = rows;
  return PFM_OK;
}

}  // namespace pfm

using namespace pfm;

extern "C" {

int pfm_tf_create(const pfm_tf_cfg* cfg, int device, pfm_tf** out) {
  if (!cfg || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  *out = nullptr;
  const pfm_tf_cfg& c = *cfg;
  if (c.kind != 0 && c.kind != 1) { set_error("unknown transformer kind %d", c.kind); return PFM_ERR_INVALID; }
  if (c.model_dim <= 0 || c.num_heads <= 0 || c.model_dim % c.num_heads || c.model_dim / c.num_heads > DH_MAX) {
    set_error("model_dim %d / num_heads %d: head dim must divide and be <= %d", c.model_dim, c.num_heads, DH_MAX);
    return PFM_ERR_UNSUPPORTED;
  }
  if (c.feats <= 0 || c.t_dim <= 0 || c.num_layers <= 0 || c.ctxt_out <= 0) { set_error("invalid transformer dims"); return PFM_ERR_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); libpfm_b200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return PFM_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaSetDevice(device));
  pfm_tf* h = new pfm_tf();
  h->cfg = c; h->device = device; h->weights_set = false; h->tok0 = nullptr;
  h->grad_floats = 0; h->tok0_goff = 0; h->tape = nullptr;
  h->capB = 0; h->capBN = 0; h->cap_rows = 0;
  h->n_real = h->rowoff = h->n_total = h->rowjet = h->tokjet = nullptr; h->ridx = nullptr;
  h->xs = h->x0 = h->v = h->h = h->H1 = h->QKV = h->A = h->tok = h->tokA = h->tokQ = h->tokKV = h->tokH1 = h->ctxin = h->c1 = h->ctx = h->jb = nullptr;
  h->jb_floats = 0; h->last_launches = 0; h->precision = PFM_PREC_FP32;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&h->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  const int D = c.model_dim, CO = c.ctxt_out;
  const int inpt = c.feats + (c.add_time_to_input ? c.t_dim : 0);
  bool ok = tf_make_dense(h, &h->ctxt, c.t_dim + c.cond_dim, 0, c.embd_hddn, CO);
  slot_dense(h, &h->ctxt);
  const int nl = c.kind == 0 ? c.num_layers : 2 * c.num_layers;
  h->layers.resize(nl);
  if (c.kind == 0) {
    for (int l = 0; l < nl && ok; ++l) {
      TfLayer& L = h->layers[l];
      ok = ok && tf_make_linear(h, &L.qkv_or_q, 3 * D, D) && tf_make_ln(h, &L.mha_ln, D) && tf_make_linear(h, &L.out, D, D) &&
           tf_make_dense(h, &L.dense, D, CO, c.dense_hddn, D) && tf_make_ln(h, &L.n1, D) && tf_make_ln(h, &L.n2, D);
      slot_linear(h, &L.qkv_or_q); slot_ln(h, &L.mha_ln); slot_linear(h, &L.out); slot_dense(h, &L.dense);
      slot_ln(h, &L.n1); slot_ln(h, &L.n2);
    }
    ok = ok && tf_make_ln(h, &h->final_norm, D);
    slot_ln(h, &h->final_norm);
  } else {
    h->tok0 = tf_alloc(h, (size_t)c.num_tokens * D);
    ok = ok && h->tok0;
    h->tok0_goff = h->grad_floats; h->grad_floats += (size_t)c.num_tokens * D;
    h->slots.push_back({c.num_tokens, D, 2, h->tok0, 0});
    for (int l = 0; l < nl && ok; ++l) {
      TfLayer& L = h->layers[l];
      ok = ok && tf_make_linear(h, &L.qkv_or_q, D, D) && tf_make_linear(h, &L.kv, 2 * D, D) && tf_make_ln(h, &L.mha_ln, D) &&
           tf_make_linear(h, &L.out, D, D) && tf_make_dense(h, &L.dense, D, CO, c.dense_hddn, D) && tf_make_ln(h, &L.n0, D) &&
           tf_make_ln(h, &L.n1, D) && tf_make_ln(h, &L.n2, D);
      slot_linear(h, &L.qkv_or_q);                   // q_linear
      slot_linear(h, &L.kv, 0, D);                   // k_linear -> columns [0, D)
      slot_linear(h, &L.kv, D, D);                   // v_linear -> columns [D, 2D)
      slot_ln(h, &L.mha_ln); slot_linear(h, &L.out); slot_dense(h, &L.dense);
      slot_ln(h, &L.n0); slot_ln(h, &L.n1); slot_ln(h, &L.n2);
    }
  }
  ok = ok && tf_make_dense(h, &h->node, inpt, CO, c.embd_hddn, D) && tf_make_dense(h, &h->outp, D, CO, c.embd_hddn, c.feats);
  slot_dense(h, &h->node);
  slot_dense(h, &h->outp);
  if (!ok) { set_error("cudaMalloc failed for the transformer weights"); pfm_tf_destroy(h); return PFM_ERR_CUDA; }
  *out = h;
  return PFM_OK;
}

void pfm_tf_destroy(pfm_tf* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  tf_tape_destroy(h);
  for (float* p : h->owned) cudaFree(p);
  for (void* p : {(void*)h->n_real, (void*)h->rowoff, (void*)h->n_total, (void*)h->rowjet, (void*)h->tokjet, (void*)h->ridx,
                  (void*)h->xs, (void*)h->x0, (void*)h->v, (void*)h->h, (void*)h->H1, (void*)h->QKV, (void*)h->A, (void*)h->tok,
                  (void*)h->tokA, (void*)h->tokQ, (void*)h->tokKV, (void*)h->tokH1, (void*)h->ctxin, (void*)h->c1, (void*)h->ctx,
                  (void*)h->jb})
    if (p) cudaFree(p);
  delete h;
}

int pfm_tf_num_params(const pfm_tf* h) { return h ? (int)h->slots.size() : PFM_ERR_INVALID; }

int pfm_tf_param_shape(const pfm_tf* h, int i, int32_t* rows, int32_t* cols) {
  if (!h || i < 0 || i >= (int)h->slots.size()) { set_error("parameter index out of range"); return PFM_ERR_INVALID; }
  if (rows) *rows = h->slots[i].rows;
  if (cols) *cols = h->slots[i].cols;
  return PFM_OK;
}

int pfm_tf_set_weights(pfm_tf* h, const float* const* params, int n, void* stream) {
  if (!h || !params) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (n != (int)h->slots.size()) { set_error("expected %d parameter tensors, got %d", (int)h->slots.size(), n); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  for (int i = 0; i < n; ++i) {
    const ParamSlot& s = h->slots[i];
    if (!params[i]) { set_error("null pointer for parameter %d", i); return PFM_ERR_INVALID; }
    if (s.kind == 0) {
      TfLinear* L = reinterpret_cast<TfLinear*>(s.target);
      const int total = s.rows * s.cols;
      tf_transpose_kernel<<<(total + 255) / 256, 256, 0, st>>>(params[i], L->Wt, s.rows, s.cols, L->ldo, s.col_off);
      PFM_CUDA_CHECK(cudaMemcpy2DAsync(L->Wrow + (size_t)s.col_off * L->ldw, sizeof(float) * L->ldw, params[i], sizeof(float) * s.cols,
                                       sizeof(float) * s.cols, s.rows, cudaMemcpyDeviceToDevice, st));
    } else {
      PFM_CUDA_CHECK(cudaMemcpyAsync(s.target, params[i], sizeof(float) * (size_t)s.rows * s.cols, cudaMemcpyDeviceToDevice, st));
    }
  }
  for (TfLinear* L : h->all_linears) {
    if (L->img) { int rc = tf_tc_pack(L->Wt, L->in, L->out, L->ldo, L->img, L->kblocks, st); if (rc != PFM_OK) return rc; }
    if (L->img_bwd) { int rc = tf_tc_pack(L->Wrow, L->out, L->ldw, L->ldw, L->img_bwd, L->kblocks_bwd, st); if (rc != PFM_OK) return rc; }
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  h->weights_set = true;
  return PFM_OK;
}

int pfm_tf_set_precision(pfm_tf* h, int precision) {
  if (!h) { set_error("null handle"); return PFM_ERR_INVALID; }
  if (precision != PFM_PREC_FP32 && precision != PFM_PREC_BF16) { set_error("unknown precision %d", precision); return PFM_ERR_INVALID; }
  h->precision = precision;
  return PFM_OK;
}

static int tf_check_call(pfm_tf* h, const float* t_code, const float* cond, int B, int N) {
  if (!h->weights_set) { set_error("weights not set (call pfm_tf_set_weights first)"); return PFM_ERR_STATE; }
  if (B <= 0 || N <= 0 || N > 65535) { set_error("bad batch shape B=%d N=%d", B, N); return PFM_ERR_INVALID; }
  if (!t_code) { set_error("time code is NULL"); return PFM_ERR_INVALID; }
  if (h->cfg.cond_dim > 0 && !cond) { set_error("cond is NULL but the net is conditioned"); return PFM_ERR_INVALID; }
  if ((size_t)N * (h->cfg.model_dim / h->cfg.num_heads) * 2 * sizeof(float) > 200 * 1024) {
    set_error("a jet of %d particles does not fit the attention kernel's shared memory", N); return PFM_ERR_UNSUPPORTED;
  }
  return PFM_OK;
}

int pfm_tf_forward(pfm_tf* h, const float* t_code, int t_rows, const float* x, const float* mask, const float* cond, float* out,
                   int B, int N, void* stream) {
  if (!h || !x || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (t_rows != 1 && t_rows != B) { set_error("t_rows must be 1 or B"); return PFM_ERR_INVALID; }
  int rc = tf_check_call(h, t_code, cond, B, N);
  if (rc != PFM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  h->last_launches = 0;
  int rows = 0;
  if ((rc = tf_plan(h, st, x, mask, B, N, &rows)) != PFM_OK) return rc;
  {
    const size_t smem = sizeof(float) * 2 * (size_t)N * (h->cfg.model_dim / h->cfg.num_heads);
    const int lim = (int)(smem > 48 * 1024 ? smem : 48 * 1024);
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  }
  if ((rc = tf_eval(h, st, t_code, t_rows, cond, B, N, rows)) != PFM_OK) return rc;
  tf_unpack_kernel<<<(B + 7) / 8, 256, 0, st>>>(h->v, h->n_real, h->ridx, h->rowoff, B, N, h->cfg.feats, out);
  h->last_launches++;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_tf_sample(pfm_tf* h, float* x_inout, const float* mask, const float* cond, const float* t_codes, const float* dt,
                  int solver, int n_steps, int B, int N, void* stream) {
  if (!h || !x_inout || !dt) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (solver != PFM_SOLVER_EULER && solver != PFM_SOLVER_MIDPOINT) { set_error("unknown solver %d", solver); return PFM_ERR_INVALID; }
  if (n_steps <= 0) { set_error("n_steps must be positive"); return PFM_ERR_INVALID; }
  int rc = tf_check_call(h, t_codes, cond, B, N);
  if (rc != PFM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  h->last_launches = 0;
  int rows = 0;
  const int n_evals = n_steps * (solver == PFM_SOLVER_MIDPOINT ? 2 : 1);
  const bool shared_ctx = h->cfg.cond_dim == 0;      // every jet shares the time code: one context row per EVALUATION
  if ((rc = tf_plan(h, st, x_inout, mask, B, N, &rows, shared_ctx ? n_evals : 0)) != PFM_OK) return rc;
  const int ctx_cap = (int)h->cap_rows;
  if (shared_ctx && (rc = tf_context(h, st, t_codes, h->cfg.t_dim, nullptr, n_evals, ctx_cap)) != PFM_OK) return rc;
  {
    const size_t smem = sizeof(float) * 2 * (size_t)N * (h->cfg.model_dim / h->cfg.num_heads);
    const int lim = (int)(smem > 48 * 1024 ? smem : 48 * 1024);
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_self_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim));
  }
  const int F = h->cfg.feats, T = h->cfg.t_dim;
  const int nel = rows * F;
  if (nel > 0) PFM_CUDA_CHECK(cudaMemcpyAsync(h->x0, h->xs, sizeof(float) * nel, cudaMemcpyDeviceToDevice, st));
  const bool mid = solver == PFM_SOLVER_MIDPOINT;
  int ev = 0;
  for (int s = 0; s < n_steps; ++s) {
    for (int stage = 0; stage < (mid ? 2 : 1); ++stage, ++ev) {
      if ((rc = tf_eval(h, st, t_codes + (size_t)ev * T, 1, cond, B, N, rows, shared_ctx ? ev : -1, ctx_cap)) != PFM_OK) return rc;
      if (nel > 0) tf_update_kernel<<<(nel + 255) / 256, 256, 0, st>>>(h->x0, h->xs, h->v, dt, s, (mid && stage == 0) ? 1 : 0, nel);
      h->last_launches++;
    }
  }
  tf_unpack_kernel<<<(B + 7) / 8, 256, 0, st>>>(h->x0, h->n_real, h->ridx, h->rowoff, B, N, F, x_inout);
  h->last_launches++;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_tf_last_launches(const pfm_tf* h) { return h ? h->last_launches : PFM_ERR_INVALID; }

}  // extern "C"
