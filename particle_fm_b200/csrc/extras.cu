// The "next" rows of the scope table that sit either side of the vector-field kernels:
//   pfm_postprocess     generate_data's post-processing (inverse normalisation, log-pt, masking) fused with the
//                       device -> host hand-over: the kernel writes straight into the caller's (pinned host) buffer
//                       particle_fm/utils/data_generation.py:105-123
//   pfm_ot_assign /     CFM-OT mini-batch coupling: exact per-jet assignment between noise and data particles and the
//   pfm_ot_gather       resampling of the matched pairs          particle_fm/models/components/losses.py:165-189
//   pfm_mlp_*           jet-feature flow (small_cond_MLP_model behind CNF) and its fixed-step integrator, one launch
//                       particle_fm/models/flow_matching_no_sets.py:41-93, components/mlp.py:24-68
#include <cfloat>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "pfm_internal.cuh"
#include "simt_common.cuh"

namespace pfm {

// =============================================================================================
// generate_data post-processing
// =============================================================================================
struct PostParams {
  float scale[PFM_POST_MAX_FEATS], shift[PFM_POST_MAX_FEATS];
  int affine, log_col, F, N, first_only_col;
};

// One thread per particle slot: the F features of a slot are contiguous, 128-bit stores when F == 4, else scalar
// stores that the warp coalesces (a warp writes 32 * F consecutive floats).  `out` may be pinned host memory (UVA):
// the stores then travel over PCIe as posted writes and no separate D2H copy exists.
__global__ void postprocess_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ out,
                                   long long slots, PostParams pp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= slots) return;
  const float m = mask ? mask[i] : 1.f;
  const float* src = x + i * pp.F;
  float* dst = out + i * pp.F;
  for (int f = 0; f < pp.F; ++f) {
    float v = src[f];
    if (pp.affine) {
      // tensor[..., f] * (std/sigma) + mean: two roundings like the eager reference (no fused multiply-add)
      // pt_standardization quirk of the reference: inverse_normalize_tensor is handed the 2-D slice batch[..., 2], so its
      // ``tensor[..., 0]`` only touches particle 0 of every jet (data_generation.py:111-113, utils.py:198-199)
      if (f != pp.first_only_col || (i % pp.N) == 0) v = __fadd_rn(__fmul_rn(v, pp.scale[f]), pp.shift[f]);
      if (f == pp.log_col) v = 1.0f - expf(v);                  // data_generation.py:116-117
    }
    if (mask) v = __fmul_rn(v, m);                               // :118-119 (padded slots: x * 0)
    dst[f] = v;
  }
}

// =============================================================================================
// CFM-OT coupling: exact linear assignment, one warp per jet
// =============================================================================================
// Shortest-augmenting-path Hungarian algorithm (Kuhn-Munkres with potentials, O(n^3)) on the squared-distance cost
// between the n noise points (rows) and the n data points (columns) of one jet.  Uniform marginals make POT's
// ot.emd(a, b, M) (losses.py:180) an assignment problem: its vertex solutions are permutation matrices / n.
// Columns are distributed over the lanes (column j -> lane j % 32, slot j / 32); the per-column state (potential v,
// minv, way, owner row p) lives in registers, the row potentials u and both point sets in shared memory.  Costs are
// recomputed from the coordinates (F multiply-adds) instead of being stored: no n x n matrix exists anywhere.
// Arithmetic in double: with float32 coordinates the products and sums of three squares are exact enough that ties
// only arise for exactly coincident points (the zero padding), where every optimal solution maps to the same loss.
template <int Q>
__global__ void __launch_bounds__(128) ot_assign_kernel(const float* __restrict__ x0, const float* __restrict__ x1, int B, int N, int F,
                                                        int* __restrict__ sigma, double* __restrict__ cost_out) {
  extern __shared__ double ot_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + warp;
  if (k >= B) return;
  // per-warp carve-up: u[N+1] doubles | rowcol[N+1] ints (column assigned to row) | a[N*F] | b[N*F] floats
  const size_t per_warp = (size_t)(N + 1) * sizeof(double) + (size_t)(N + 2) / 2 * 2 * sizeof(int) + (size_t)2 * N * F * sizeof(float);
  unsigned char* base = reinterpret_cast<unsigned char*>(ot_smem) + (size_t)warp * ((per_warp + 15) / 16 * 16);
  double* u = reinterpret_cast<double*>(base);
  int* rowcol = reinterpret_cast<int*>(u + (N + 1));
  float* a = reinterpret_cast<float*>(rowcol + (N + 2) / 2 * 2);
  float* b = a + (size_t)N * F;
  for (int i = lane; i < N * F; i += 32) { a[i] = x0[(size_t)k * N * F + i]; b[i] = x1[(size_t)k * N * F + i]; }
  for (int i = lane; i <= N; i += 32) { u[i] = 0.0; rowcol[i] = -1; }
  __syncwarp();
  // column state of this lane: columns j = lane + 32 q (1-based index j + 1 in the classic formulation)
  double v[Q], minv[Q];
  int way[Q], p[Q];          // p: row owning the column (-1 = free); way: previous column on the alternating path (-1 = root)
  bool used[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) { v[q] = 0.0; p[q] = -1; }
  for (int i = 0; i < N; ++i) {
    // grow an alternating tree from row i until a free column is reached
#pragma unroll
    for (int q = 0; q < Q; ++q) { minv[q] = DBL_MAX; way[q] = -1; used[q] = false; }
    int i0 = i;            // row whose edges are relaxed next
    int j0 = -1;           // column just added to the tree (-1 = virtual root column holding row i)
    int jfree;
    for (;;) {
      const double ui0 = u[i0];
      double best = DBL_MAX;
      int bestj = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = lane + 32 * q;
        if (j < N && !used[q]) {
          double c = 0.0;
          for (int f = 0; f < F; ++f) { const double d = (double)a[i0 * F + f] - (double)b[j * F + f]; c = fma(d, d, c); }
          const double cur = c - ui0 - v[q];
          if (cur < minv[q]) { minv[q] = cur; way[q] = j0; }
          if (minv[q] < best) { best = minv[q]; bestj = j; }     // strict '<': lowest column index wins ties inside the lane
        }
      }
      // warp argmin (value, then lowest column index)
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oj = __shfl_xor_sync(0xffffffffu, bestj, s);
        if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
      }
      const double delta = best;
      const int j1 = bestj;
      // potentials: tree rows += delta, tree columns -= delta, the others' slack shrinks
      if (lane == 0) u[i] += delta;                              // the root row (virtual column)
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        const int j = lane + 32 * q;
        if (j < N) {
          if (used[q]) { u[p[q]] += delta; v[q] -= delta; }      // distinct rows: no write conflicts
          else minv[q] -= delta;
        }
      }
      __syncwarp();
      // add column j1 to the tree
      const int owner_lane = j1 & 31, owner_q = j1 >> 5;
      int pj1 = -1;
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (q == owner_q) { if (lane == owner_lane) used[q] = true; pj1 = p[q]; }
      pj1 = __shfl_sync(0xffffffffu, pj1, owner_lane);
      j0 = j1;
      if (pj1 < 0) { jfree = j1; break; }
      i0 = pj1;
    }
    // augment along the path: column j takes the row of way[j], back to the root
    int j = jfree;
    while (j >= 0) {
      const int ol = j & 31, oq = j >> 5;
      int wj = -1;
#pragma unroll
      for (int q = 0; q < Q; ++q) if (q == oq) wj = way[q];
      wj = __shfl_sync(0xffffffffu, wj, ol);
      int newrow;
      if (wj < 0) newrow = i;
      else {
        const int pl = wj & 31, pq = wj >> 5;
        int pw = -1;
#pragma unroll
        for (int q = 0; q < Q; ++q) if (q == pq) pw = p[q];
        newrow = __shfl_sync(0xffffffffu, pw, pl);
      }
#pragma unroll
      for (int q = 0; q < Q; ++q) if (q == oq && lane == ol) p[q] = newrow;
      j = wj;
    }
    __syncwarp();
  }
  // sigma[row] = column, total cost
  double tot = 0.0;
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = lane + 32 * q;
    if (j < N) {
      const int r = p[q];
      sigma[(size_t)k * N + r] = j;
      double c = 0.0;
      for (int f = 0; f < F; ++f) { const double d = (double)a[r * F + f] - (double)b[j * F + f]; c = fma(d, d, c); }
      tot += c;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
  if (cost_out && lane == 0) cost_out[k] = tot;
}

// x0p[k,m] = x0[k, i[k,m]], x1p[k,m] = x1[k, sigma[k, i[k,m]]], mask_ot[k,m] = mask[k, sigma[k, i[k,m]]]
// (losses.py:183-189: the pairs (i, j) drawn from the plan; x0p / x1p must not alias x0 / x1)
__global__ void ot_gather_kernel(const float* __restrict__ x0, const float* __restrict__ x1, const float* __restrict__ mask,
                                 const int* __restrict__ sigma, const int* __restrict__ pick, int N, int F, long long slots,
                                 float* __restrict__ x0p, float* __restrict__ x1p, float* __restrict__ mask_ot) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= slots) return;
  const long long k = s / N;
  const int i = pick[s];
  const int j = sigma[k * N + i];
  for (int f = 0; f < F; ++f) {
    x0p[s * F + f] = x0[(k * N + i) * F + f];
    x1p[s * F + f] = x1[(k * N + j) * F + f];
  }
  if (mask_ot) mask_ot[s] = mask ? mask[k * N + j] : 1.f;
}

// =============================================================================================
// jet-feature flow: conditional MLP vector field + fixed-step integrator, state resident per CTA
// =============================================================================================
static constexpr int MLP_ROWS = 32;        // rows (jets) per CTA tile = 8 warps x 4 rows
static constexpr int MLP_RB = 4;
static constexpr int MLP_KC = 16;
static constexpr int MLP_MAXW = 256;       // widest hidden layer
static constexpr int MLP_MAX_LIN = 32;

struct MlpLin {
  int K, out, ldo, TC;       // input width (incl. the concatenated time / cond columns), outputs, padded leading dimension, columns per lane
  int concat, act;           // input is [t | previous | cond];  activation after this linear
  const float* Wt;           // k-major [Kp, ldo], zero padded
  const float* b;            // [ldo]
};

struct MlpParams {
  int F, T, C, n_lin, act_kind, lda;
  MlpLin lin[MLP_MAX_LIN];
  const float* x_in; float* x_out; const float* cond; int B;
  const float* t_codes; int t_rows_per_eval;      // [n_evals, T] (one code per evaluation) or, forward only, [B, T]
  int n_evals, solver; const float* dt;
  int* counter;
};

__device__ __forceinline__ float mlp_act(float v, int kind) {
  switch (kind) {
    case PFM_ACT_ELU: return v > 0.f ? v : expm1f(v);
    case PFM_ACT_TANH: return tanhf(v);
    case PFM_ACT_RELU: return fmaxf(v, 0.f);
    case PFM_ACT_LEAKY_RELU: return v > 0.f ? v : 0.01f * v;
    case PFM_ACT_SILU: return v / (1.f + expf(-v));
    default: return v;
  }
}

template <int TC>
__device__ __forceinline__ void mlp_linear(const MlpParams& p, const MlpLin& L, const float* in, float* outb, int out_col0,
                                           float* wbuf, int rows) {
  float acc[MLP_RB][TC];
  gemm_rows<TC, MLP_RB>(in, p.lda, L.Wt, L.K, L.ldo, wbuf, MLP_KC * MLP_MAXW, MLP_KC, acc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < MLP_RB; ++r) {
    const int row = warp * MLP_RB + r;
#pragma unroll
    for (int i = 0; i < TC; ++i) {
      const int o = lane + 32 * i;
      if (o < L.out && row < rows) {
        float v = acc[r][i] + L.b[o];
        if (L.act) v = mlp_act(v, p.act_kind);
        outb[(size_t)row * p.lda + out_col0 + o] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads) mlp_flow_kernel(const MlpParams p) {
  extern __shared__ float mlp_smem[];
  float* bufA = mlp_smem;                                  // [32][lda]
  float* bufB = bufA + MLP_ROWS * p.lda;
  float* wbuf = bufB + MLP_ROWS * p.lda;                   // 2 x [KC][256]
  float* x0 = wbuf + 2 * MLP_KC * MLP_MAXW;                // [32][F]  state at the start of the step
  float* xc = x0 + MLP_ROWS * p.F;                         // [32][F]  network input of this evaluation
  float* cnd = xc + MLP_ROWS * p.F;                        // [32][C]
  __shared__ int s_tile;
  const int tid = threadIdx.x;
  const int n_tiles = (p.B + MLP_ROWS - 1) / MLP_ROWS;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_tile = atomicAdd(p.counter, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= n_tiles) break;
    const int r0 = tile * MLP_ROWS;
    const int rows = p.B - r0 < MLP_ROWS ? p.B - r0 : MLP_ROWS;
    for (int i = tid; i < MLP_ROWS * p.F; i += kThreads) {
      const int r = i / p.F;
      const float v = r < rows ? p.x_in[(size_t)r0 * p.F + i] : 0.f;
      x0[i] = v; xc[i] = v;
    }
    for (int i = tid; i < MLP_ROWS * p.C; i += kThreads) cnd[i] = (i / p.C) < rows ? p.cond[(size_t)r0 * p.C + i] : 0.f;
    __syncthreads();
    for (int ev = 0; ev < p.n_evals; ++ev) {
      float* in = bufA;
      float* out = bufB;
      const float* prev = xc;        // what the next concat wraps: [32][prev_w] with leading dimension prev_ld
      int prev_w = p.F, prev_ld = p.F;
      for (int l = 0; l < p.n_lin; ++l) {
        const MlpLin& L = p.lin[l];
        if (L.concat) {
          // in[row] = [t code | prev | cond | 0 padding]     (mlp.py:58-66 torch.cat([t, x, cond]))
          // prev lives either in xc (first block) or already at column T of `in` (written there by the previous linear)
          const int Kp = (L.K + 3) & ~3;
          for (int i = tid; i < MLP_ROWS * Kp; i += kThreads) {
            const int r = i / Kp, c = i - r * Kp;
            float v;
            if (c < p.T) {
              const int trow = p.t_rows_per_eval ? ev : (r0 + (r < rows ? r : 0));
              v = p.t_codes[(size_t)trow * p.T + c];
            } else if (c < p.T + prev_w) {
              if (prev != xc) continue;                     // already in place
              v = prev[r * prev_ld + (c - p.T)];
            } else if (c < p.T + prev_w + p.C) {
              v = cnd[r * p.C + (c - p.T - prev_w)];
            } else {
              v = 0.f;
            }
            in[(size_t)r * p.lda + c] = v;
          }
          __syncthreads();
        }
        const bool next_concat = (l + 1 < p.n_lin) && p.lin[l + 1].concat;
        const int col0 = next_concat ? p.T : 0;             // the next block's concat finds its x columns in place
        switch (L.TC) {
          case 1: mlp_linear<1>(p, L, in, out, col0, wbuf, MLP_ROWS); break;
          case 2: mlp_linear<2>(p, L, in, out, col0, wbuf, MLP_ROWS); break;
          case 4: mlp_linear<4>(p, L, in, out, col0, wbuf, MLP_ROWS); break;
          default: mlp_linear<8>(p, L, in, out, col0, wbuf, MLP_ROWS); break;
        }
        if (!next_concat) {
          // zero the K padding of the next linear's input (its K = L.out is not always a multiple of 4)
          const int o4 = (L.out + 3) & ~3;
          for (int i = tid; i < MLP_ROWS * (o4 - L.out); i += kThreads) {
            const int r = i / (o4 - L.out), c = L.out + i % (o4 - L.out);
            out[(size_t)r * p.lda + c] = 0.f;
          }
        }
        __syncthreads();
        prev = out; prev_w = L.out; prev_ld = p.lda;
        float* t = in; in = out; out = t;
      }
      // `in` now holds v = net(t, x, cond) in columns [0, F)
      if (p.solver < 0) {
        for (int i = tid; i < rows * p.F; i += kThreads) {
          const int r = i / p.F, f = i - r * p.F;
          p.x_out[(size_t)(r0 + r) * p.F + f] = in[(size_t)r * p.lda + f];
        }
      } else {
        // torchdyn fixed step in reversed time (oracle/ode_oracle.py): k = -v
        const bool mid = p.solver == PFM_SOLVER_MIDPOINT;
        const float dt = p.dt[mid ? (ev >> 1) : ev];
        const bool first_stage = mid && ((ev & 1) == 0);
        const float hdt = __fmul_rn(0.5f, dt);
        for (int i = tid; i < MLP_ROWS * p.F; i += kThreads) {
          const int r = i / p.F, f = i - r * p.F;
          const float k = -in[(size_t)r * p.lda + f];
          if (first_stage) {
            xc[i] = __fadd_rn(x0[i], __fmul_rn(hdt, k));
          } else {
            const float xn = __fadd_rn(x0[i], __fmul_rn(dt, k));
            x0[i] = xn; xc[i] = xn;
          }
        }
      }
      __syncthreads();
    }
    if (p.solver >= 0)
      for (int i = tid; i < rows * p.F; i += kThreads) p.x_out[(size_t)r0 * p.F + i] = x0[i];
  }
}

}  // namespace pfm

using namespace pfm;

struct pfm_mlp {
  pfm_mlp_cfg cfg;
  int device, sm_count;
  std::vector<int> K, out, concat, act, ldo, Kp;
  std::vector<size_t> w_off, b_off;
  float* store; size_t store_floats;
  int* counter;
  bool weights_set;
};

extern "C" {

int pfm_postprocess(const float* x, const float* mask, float* out, long long B, int N, int F, const float* scale,
                    const float* shift, int log_col, int first_only_col, void* stream) {
  if (!x || !out || B < 0 || N <= 0 || F <= 0) { set_error("pfm_postprocess: bad argument"); return PFM_ERR_INVALID; }
  if (F > PFM_POST_MAX_FEATS) { set_error("pfm_postprocess: F=%d exceeds %d", F, PFM_POST_MAX_FEATS); return PFM_ERR_INVALID; }
  if ((scale == nullptr) != (shift == nullptr)) { set_error("pfm_postprocess: scale and shift go together"); return PFM_ERR_INVALID; }
  if (B == 0) return PFM_OK;
  PostParams pp;
  memset(&pp, 0, sizeof(pp));
  pp.F = F; pp.N = N; pp.affine = scale ? 1 : 0; pp.log_col = scale ? log_col : -1; pp.first_only_col = first_only_col;
  for (int f = 0; f < F && scale; ++f) { pp.scale[f] = scale[f]; pp.shift[f] = shift[f]; }
  const long long slots = B * N;
  postprocess_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, mask, out, slots, pp);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_ot_assign(const float* x0, const float* x1, int B, int N, int F, int32_t* sigma, double* cost, void* stream) {
  if (!x0 || !x1 || !sigma || B <= 0 || N <= 0 || F <= 0) { set_error("pfm_ot_assign: bad argument"); return PFM_ERR_INVALID; }
  const int Q = (N + 31) / 32;
  if (Q > 10) { set_error("pfm_ot_assign: N=%d exceeds 320 particles", N); return PFM_ERR_UNSUPPORTED; }
  const size_t per_warp = ((size_t)(N + 1) * sizeof(double) + (size_t)(N + 2) / 2 * 2 * sizeof(int) + (size_t)2 * N * F * sizeof(float) + 15) / 16 * 16;
  int warps = 4;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem > 220 * 1024) { set_error("pfm_ot_assign: N=%d x F=%d does not fit shared memory", N, F); return PFM_ERR_UNSUPPORTED; }
  const int grid = (B + warps - 1) / warps;
  cudaStream_t st = (cudaStream_t)stream;
#define PFM_OT_CASE(QQ)                                                                                              \
  case QQ:                                                                                                           \
    if (smem > 48 * 1024) PFM_CUDA_CHECK(cudaFuncSetAttribute(ot_assign_kernel<QQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    ot_assign_kernel<QQ><<<grid, warps * 32, smem, st>>>(x0, x1, B, N, F, sigma, cost);                             \
    break;
  switch (Q) {
    PFM_OT_CASE(1) PFM_OT_CASE(2) PFM_OT_CASE(3) PFM_OT_CASE(4) PFM_OT_CASE(5)
    PFM_OT_CASE(6) PFM_OT_CASE(7) PFM_OT_CASE(8) PFM_OT_CASE(9) PFM_OT_CASE(10)
  }
#undef PFM_OT_CASE
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_ot_gather(const float* x0, const float* x1, const float* mask, const int32_t* sigma, const int32_t* pick, int B, int N,
                  int F, float* x0p, float* x1p, float* mask_ot, void* stream) {
  if (!x0 || !x1 || !sigma || !pick || !x0p || !x1p || B <= 0 || N <= 0 || F <= 0) { set_error("pfm_ot_gather: bad argument"); return PFM_ERR_INVALID; }
  if (x0p == x0 || x1p == x1) { set_error("pfm_ot_gather: outputs must not alias the inputs"); return PFM_ERR_INVALID; }
  const long long slots = (long long)B * N;
  ot_gather_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x0, x1, mask, sigma, pick, N, F, slots, x0p,
                                                                                      x1p, mask_ot);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// ---------------------------------------------------------------------------------------------
int pfm_mlp_create(const pfm_mlp_cfg* cfg, const int32_t* out_widths, const int32_t* concat, const int32_t* act, int device,
                   pfm_mlp** out) {
  if (!cfg || !out_widths || !concat || !act || !out) { set_error("pfm_mlp_create: null argument"); return PFM_ERR_INVALID; }
  *out = nullptr;
  if (cfg->n_linears < 1 || cfg->n_linears > MLP_MAX_LIN) { set_error("pfm_mlp_create: 1..%d linears", MLP_MAX_LIN); return PFM_ERR_INVALID; }
  if (cfg->features < 1 || cfg->features > 32 || cfg->t_dim < 0 || cfg->cond_dim < 0) { set_error("pfm_mlp_create: bad widths"); return PFM_ERR_INVALID; }
  if (cfg->act < 0 || cfg->act > PFM_ACT_SILU) { set_error("pfm_mlp_create: unknown activation %d", cfg->act); return PFM_ERR_INVALID; }
  if (!concat[0]) { set_error("pfm_mlp_create: the first linear takes [t | x | cond]"); return PFM_ERR_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s): libpfm_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return PFM_ERR_CUDA;
  }
  PFM_CUDA_CHECK(cudaSetDevice(device));
  pfm_mlp* h = new (std::nothrow) pfm_mlp();
  if (!h) { set_error("out of host memory"); return PFM_ERR_INVALID; }
  h->cfg = *cfg; h->device = device; h->weights_set = false; h->store = nullptr; h->counter = nullptr;
  cudaDeviceProp prop;
  PFM_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  int prev = cfg->features;
  size_t off = 0;
  for (int l = 0; l < cfg->n_linears; ++l) {
    const int K = concat[l] ? cfg->t_dim + prev + cfg->cond_dim : prev;
    const int o = out_widths[l];
    if (o < 1 || o > MLP_MAXW || K > MLP_MAXW + cfg->t_dim + cfg->cond_dim) { delete h; set_error("pfm_mlp_create: layer width out of range"); return PFM_ERR_UNSUPPORTED; }
    int TC = (o + 31) / 32;
    TC = TC <= 1 ? 1 : (TC <= 2 ? 2 : (TC <= 4 ? 4 : 8));
    const int Kp = (K + 3) & ~3;
    h->K.push_back(K); h->out.push_back(o); h->concat.push_back(concat[l]); h->act.push_back(act[l]);
    h->ldo.push_back(32 * TC); h->Kp.push_back(Kp);
    h->w_off.push_back(off); off += (size_t)(Kp + MLP_KC) * 32 * TC;      // slack rows: the last chunk is copied whole
    h->b_off.push_back(off); off += 32 * TC;
    prev = o;
  }
  if (prev != cfg->features) { delete h; set_error("pfm_mlp_create: the last linear must map back to `features`"); return PFM_ERR_INVALID; }
  h->store_floats = off;
  PFM_CUDA_CHECK(cudaMalloc(&h->store, off * sizeof(float)));
  PFM_CUDA_CHECK(cudaMemset(h->store, 0, off * sizeof(float)));
  PFM_CUDA_CHECK(cudaMalloc(&h->counter, sizeof(int)));
  *out = h;
  return PFM_OK;
}

void pfm_mlp_destroy(pfm_mlp* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->store) cudaFree(h->store);
  if (h->counter) cudaFree(h->counter);
  delete h;
}

int pfm_mlp_linear_shape(const pfm_mlp* h, int i, int32_t* out_features, int32_t* in_features) {
  if (!h || i < 0 || i >= h->cfg.n_linears) { set_error("pfm_mlp_linear_shape: bad index"); return PFM_ERR_INVALID; }
  *out_features = h->out[i]; *in_features = h->K[i];
  return PFM_OK;
}

namespace pfm {
// W [out, K] row-major -> k-major zero-padded copy
__global__ void mlp_pack_kernel(const float* __restrict__ W, const float* __restrict__ b, int out, int K, int ldo, float* __restrict__ Wt,
                                float* __restrict__ bp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * ldo; i += gridDim.x * blockDim.x) {
    const int k = i / ldo, o = i - k * ldo;
    Wt[i] = o < out ? W[(size_t)o * K + k] : 0.f;
  }
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < ldo; o += gridDim.x * blockDim.x) bp[o] = o < out ? b[o] : 0.f;
}
}  // namespace pfm

int pfm_mlp_set_weights(pfm_mlp* h, const float* const* weights, const float* const* biases, int n, void* stream) {
  if (!h || !weights || !biases || n != h->cfg.n_linears) { set_error("pfm_mlp_set_weights: expected %d linears", h ? h->cfg.n_linears : -1); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  for (int l = 0; l < n; ++l) {
    if (!weights[l] || !biases[l]) { set_error("pfm_mlp_set_weights: null tensor %d", l); return PFM_ERR_INVALID; }
    mlp_pack_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(weights[l], biases[l], h->out[l], h->K[l], h->ldo[l], h->store + h->w_off[l],
                                                         h->store + h->b_off[l]);
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  h->weights_set = true;
  return PFM_OK;
}

static int mlp_run(pfm_mlp* h, const float* x_in, float* x_out, const float* cond, const float* t_codes, int t_rows_per_eval,
                   const float* dt, int solver, int n_evals, int B, cudaStream_t st) {
  if (!h->weights_set) { set_error("weights not set (call pfm_mlp_set_weights first)"); return PFM_ERR_STATE; }
  if (B <= 0) { set_error("B must be positive"); return PFM_ERR_INVALID; }
  if (h->cfg.cond_dim > 0 && !cond) { set_error("cond is NULL but the net is conditioned"); return PFM_ERR_INVALID; }
  if (h->cfg.t_dim > 0 && !t_codes) { set_error("time code is NULL"); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  MlpParams p;
  memset(&p, 0, sizeof(p));
  p.F = h->cfg.features; p.T = h->cfg.t_dim; p.C = h->cfg.cond_dim; p.n_lin = h->cfg.n_linears; p.act_kind = h->cfg.act;
  int kmax = 0;
  for (int l = 0; l < p.n_lin; ++l) {
    MlpLin& L = p.lin[l];
    L.K = h->K[l]; L.out = h->out[l]; L.ldo = h->ldo[l]; L.TC = h->ldo[l] / 32; L.concat = h->concat[l]; L.act = h->act[l];
    L.Wt = h->store + h->w_off[l]; L.b = h->store + h->b_off[l];
    if (h->Kp[l] > kmax) kmax = h->Kp[l];
    if (L.out + p.T + p.C + 4 > kmax) kmax = L.out + p.T + p.C + 4;
  }
  p.lda = ((kmax + 3) & ~3) + 4;
  p.x_in = x_in; p.x_out = x_out; p.cond = cond; p.B = B;
  p.t_codes = t_codes; p.t_rows_per_eval = t_rows_per_eval;
  p.n_evals = n_evals; p.solver = solver; p.dt = dt; p.counter = h->counter;
  const size_t smem = sizeof(float) * ((size_t)2 * MLP_ROWS * p.lda + 2 * MLP_KC * MLP_MAXW + (size_t)2 * MLP_ROWS * p.F + (size_t)MLP_ROWS * (p.C > 0 ? p.C : 1));
  PFM_CUDA_CHECK(cudaFuncSetAttribute(mlp_flow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PFM_CUDA_CHECK(cudaMemsetAsync(h->counter, 0, sizeof(int), st));
  const int tiles = (B + MLP_ROWS - 1) / MLP_ROWS;
  const int grid = tiles < 2 * h->sm_count ? tiles : 2 * h->sm_count;
  mlp_flow_kernel<<<grid, kThreads, smem, st>>>(p);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_mlp_forward(pfm_mlp* h, const float* t_code, int t_rows, const float* x, const float* cond, float* out, int B, void* stream) {
  if (!h || !x || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (t_rows != 1 && t_rows != B) { set_error("t_rows must be 1 or B (got %d, B=%d)", t_rows, B); return PFM_ERR_INVALID; }
  return mlp_run(h, x, out, cond, t_code, t_rows == 1 ? 1 : 0, nullptr, -1, 1, B, (cudaStream_t)stream);
}

int pfm_mlp_sample(pfm_mlp* h, float* x_inout, const float* cond, const float* t_codes, const float* dt, int solver, int n_steps,
                   int B, void* stream) {
  if (!h || !x_inout || !dt) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (solver != PFM_SOLVER_EULER && solver != PFM_SOLVER_MIDPOINT) { set_error("unknown solver %d", solver); return PFM_ERR_INVALID; }
  if (n_steps <= 0) { set_error("n_steps must be positive"); return PFM_ERR_INVALID; }
  const int n_evals = n_steps * (solver == PFM_SOLVER_MIDPOINT ? 2 : 1);
  return mlp_run(h, x_inout, x_inout, cond, t_codes, 1, dt, solver, n_evals, B, (cudaStream_t)stream);
}

}  // extern "C"
