// Training backward of the EPiC vector field, fp32 CUDA cores.
//
// The reference gets this from torch autograd over ~250 ops per evaluation; here it is
// three kinds of kernels over the activations the training forward (epic_simt.cu, TRAIN) saved:
//
//   epic_bwd_kernel   one persistent CTA per jet group (same plan as the forward).  Keeps the gradient
//                     w.r.t. the hidden features dh [rows, H] in shared memory and walks the network in
//                     reverse: head, EPiC layers L-1..0 (fc_local2, fc_local1, pooled-gradient broadcast,
//                     fc_global2, fc_global1), stem.  Per-particle  dX = dY . W  products reuse the forward's
//                     row-block GEMM with the row-major weight copies.  It writes
//                       dact[stage][row][:]  = gradient at the pre-activation of every per-particle linear
//                       dbeff[jet][:]        = gradient of every linear's per-jet effective bias
//                                              (b + W_t.time + W_c.cond (+ W_glob.g)); for the per-jet
//                                              (global) linears this IS the pre-activation gradient
//   xty_kernel        every weight gradient is a product  dW[o][c] = sum_r Y[r][o] * X[r][c]  over particle rows
//                     (main blocks of the per-particle linears) or over jets (time / cond / global-vector
//                     blocks, the per-jet linears, biases).  One launch walks a job table.
//   small helpers     gradient seed from an incoming dL/dv, scatter of dL/dx, loss finalisation.
//
// Gradients are produced w.r.t. the FOLDED weights W = g*v/||v|| in the flat layout
// [W_0 | b_0 | W_1 | b_1 | ...]; the host maps them onto weight_g / weight_v (torch._weight_norm backward), see
// SURVEY A.6.
#include <cstdlib>

#include "pfm_internal.cuh"
#include "simt_common.cuh"

namespace pfm {

int simt_caps_for_train(const pfm_epic* h, int N, int* R_cap, int* J_cap, int* TC, int* RB, int* KC);

static constexpr int kJMax = 16;     // upper bound of J_cap (simt_shape)
static constexpr int kGradChunks = 3;

struct BwdParams {
  int F, Kx, xin_off, H, Hp, LDH, Z, Zp, L, n_lin, LDP;
  int R_cap, J_cap, KC;
  float sum_scale, slope;
  const Lin* lin;
  const int* n_real; const int2* groups; const int* n_groups; int* counter; const int* rowoff;
  const float* act; float* dact; size_t stage_stride; int Hp_act;
  const float* jact; int junit, jstride;
  const float* dpre3;
  float* dbeff; int bstride;
  float* dxs;                 // [rows, Kx] or nullptr
  float* dh_spill;            // spill mode: dh lives in a per-CTA slab of global memory
  int o_dh, o_tA, o_tB, o_wbuf, o_db1, o_db2, o_dG, o_pg1, o_pg2, o_din, o_int, total_floats;
  int wbuf_floats;
};

__device__ __forceinline__ float dlrelu(float post, float slope) { return post > 0.f ? 1.f : slope; }

// d[r][i]: incoming gradient of the warp's RB rows (row = c0 + warp*RB + r, column lane + 32 i).  Multiplies by
// leaky_relu'(saved post-activation of `stage`), stores the pre-activation gradient to the chunk buffer `tdst`
// (A operand of the next product) and to dact[stage], and accumulates the per-jet column sums into db[jet][:].
// Saved post-activations of the warp's RB rows of `stage`, loaded ahead of the GEMM whose result they will mask (the
// loads are L2 / HBM latency: issued early, they overlap the GEMM instead of stalling the epilogue row by row).
template <int TC, int RB>
__device__ __forceinline__ void prefetch_act(const BwdParams& p, float (&a)[RB][TC], int c0, int R, int row_g0, int stage) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* a_st = p.act + (size_t)stage * p.stage_stride;
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const int row = c0 + warp * RB + r;
#pragma unroll
    for (int i = 0; i < TC; ++i) {
      const int o = lane + 32 * i;
      a[r][i] = (row < R && o < p.H) ? __ldg(a_st + (size_t)(row_g0 + row) * p.Hp_act + o) : 1.f;
    }
  }
}

template <int TC, int RB>
__device__ __forceinline__ void mask_store(const BwdParams& p, float (&d)[RB][TC], const float (&act)[RB][TC], int c0, int R,
                                           int row_g0, int stage, float* tdst, float* db, const short* rjet) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* d_st = p.dact + (size_t)stage * p.stage_stride;
  int cur = -1;
  float s[TC];
#pragma unroll
  for (int i = 0; i < TC; ++i) s[i] = 0.f;
#pragma unroll
  for (int r = 0; r < RB; ++r) {
    const int row = c0 + warp * RB + r;
    if (row < R) {
      const int j = rjet[row];
      if (j != cur) {
        if (cur >= 0) {
#pragma unroll
          for (int i = 0; i < TC; ++i) {
            const int o = lane + 32 * i;
            if (o < p.H) atomicAdd(&db[cur * p.Hp + o], s[i]);
            s[i] = 0.f;
          }
        }
        cur = j;
      }
#pragma unroll
      for (int i = 0; i < TC; ++i) {
        const int o = lane + 32 * i;
        if (o < p.H) {
          const size_t gi = (size_t)(row_g0 + row) * p.Hp_act + o;
          const float v = d[r][i] * dlrelu(act[r][i], p.slope);
          d[r][i] = v;
          tdst[(warp * RB + r) * p.LDH + o] = v;
          d_st[gi] = v;
          s[i] += v;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < TC; ++i) {
        const int o = lane + 32 * i;
        if (o < p.H) tdst[(warp * RB + r) * p.LDH + o] = 0.f;
        d[r][i] = 0.f;
      }
    }
  }
  if (cur >= 0) {
#pragma unroll
    for (int i = 0; i < TC; ++i) {
      const int o = lane + 32 * i;
      if (o < p.H) atomicAdd(&db[cur * p.Hp + o], s[i]);
    }
  }
}

// Backward of one unit's per-jet (global) MLP.  unit 0 = stem (fc_g1, fc_g2, no residual, pool order (sum, mean));
// unit l+1 = EPiC layer l (fc_global1, fc_global2 + residual, pool order (mean, sum, global)).
//   in : dG[j][z]  gradient w.r.t. the unit's OUTPUT global vector (carry from above; for a layer the caller has
//                  already added W_glob^T . db1)
//   out: pg2 / pg1 pre-activation gradients of the two linears (-> dbeff), din = gradient w.r.t. the pooled input,
//        dh += broadcast of the pooled gradient, dG = gradient w.r.t. the unit's INPUT global vector
__device__ __forceinline__ void global_backward(const BwdParams& p, const Lin& Ga, const Lin& Gb, int unit, int j0, int nj,
                                                const int* jrow0, const short* rjet, int R, float* dh, float* dG, float* pg1,
                                                float* pg2, float* din) {
  const int tid = threadIdx.x;
  const int H = p.H, Z = p.Z;
  // pg2 = dG * lrelu'(g)
  for (int i = tid; i < nj * Z; i += kThreads) {
    const int j = i / Z, z = i - j * Z;
    const float g = p.jact[(size_t)(j0 + j) * p.jstride + (size_t)unit * p.junit + p.LDP + p.Hp_act + z];
    pg2[j * p.Zp + z] = dG[j * p.Zp + z] * dlrelu(g, p.slope);
  }
  __syncthreads();
  // dg1[j][k] = sum_z W_b[z][m_off + k] pg2[j][z];  pg1 = dg1 * lrelu'(g1)
  for (int i = tid; i < nj * H; i += kThreads) {
    const int j = i / H, k = i - j * H;
    float a = 0.f;
    for (int z = 0; z < Z; ++z) a = fmaf(__ldg(Gb.Wr + (size_t)z * Gb.ldr + k), pg2[j * p.Zp + z], a);
    const float g1 = p.jact[(size_t)(j0 + j) * p.jstride + (size_t)unit * p.junit + p.LDP + k];
    pg1[j * p.Hp + k] = a * dlrelu(g1, p.slope);
  }
  __syncthreads();
  // din[j][k] = sum_o W_a[o][m_off + k] pg1[j][o]
  // every thread owns column k = tid (and tid + 256, 512 ... while < m_len); 4 weight loads are in flight per step
  for (int k = tid; k < Ga.m_len; k += kThreads) {
    float acc[kJMax];
#pragma unroll
    for (int j = 0; j < kJMax; ++j) acc[j] = 0.f;
    const float* wk = Ga.Wr + k;
    int o = 0;
    for (; o + 4 <= H; o += 4) {
      float w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = __ldg(wk + (size_t)(o + q) * Ga.ldr);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < kJMax; ++j)
          if (j < nj) acc[j] = fmaf(w[q], pg1[j * p.Hp + o + q], acc[j]);
    }
    for (; o < H; ++o) {
      const float w = __ldg(wk + (size_t)o * Ga.ldr);
#pragma unroll
      for (int j = 0; j < kJMax; ++j)
        if (j < nj) acc[j] = fmaf(w, pg1[j * p.Hp + o], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kJMax; ++j)
      if (j < nj) din[j * p.LDP + k] = acc[j];
  }
  __syncthreads();
  // pooled gradient back onto the particles:  S -> (mean = S/n, sum = S*s)
  const int o_mean = unit == 0 ? H : 0, o_sum = unit == 0 ? 0 : H;
  // (combined once per jet, in place over the mean section, then one add per particle element: warp = row, lane = column)
  for (int i = tid; i < nj * H; i += kThreads) {
    const int j = i / H, c = i - j * H;
    const float n = (float)(jrow0[j + 1] - jrow0[j]);
    din[j * p.LDP + o_mean + c] = din[j * p.LDP + o_mean + c] / n + p.sum_scale * din[j * p.LDP + o_sum + c];
  }
  __syncthreads();
  for (int row = tid >> 5; row < R; row += kWarps) {
    const float* comb = din + (int)rjet[row] * p.LDP + o_mean;
    float* d = dh + (size_t)row * p.LDH;
    for (int c = tid & 31; c < H; c += 32) d[c] += comb[c];
  }
  // gradient w.r.t. the incoming global vector: through fc_global1's global columns and the residual
  if (unit > 0) {
    for (int i = tid; i < nj * Z; i += kThreads) {
      const int j = i / Z, z = i - j * Z;
      dG[j * p.Zp + z] = din[j * p.LDP + 2 * H + z] + pg2[j * p.Zp + z];
    }
  }
  // effective-bias gradients of the two per-jet linears
  for (int i = tid; i < nj * H; i += kThreads) {
    const int j = i / H, o = i - j * H;
    p.dbeff[(size_t)(j0 + j) * p.bstride + Ga.bias_off + o] = pg1[j * p.Hp + o];
  }
  for (int i = tid; i < nj * Z; i += kThreads) {
    const int j = i / Z, z = i - j * Z;
    p.dbeff[(size_t)(j0 + j) * p.bstride + Gb.bias_off + z] = pg2[j * p.Zp + z];
  }
  __syncthreads();
}

// SPILL (compile time): dh lives in global memory; otherwise every access to it is a shared-memory instruction
template <int TC, int RB, bool SPILL>
__global__ void __launch_bounds__(kThreads, 1) epic_bwd_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float smem[];
  float* dh = SPILL ? p.dh_spill + (size_t)blockIdx.x * p.R_cap * p.LDH : smem + p.o_dh;       // [R_cap, LDH]  gradient w.r.t. the hidden features entering the current unit
  float* tA = smem + p.o_tA;       // [8*RB, LDH]   chunk of pre-activation gradients (A operand)
  float* tB = smem + p.o_tB;       // [8*RB, LDH]
  float* wbuf = smem + p.o_wbuf;
  float* db1 = smem + p.o_db1;     // [J_cap, Hp]   per-jet sums of the fc_local1 / fc_l1 pre-activation gradients
  float* db2 = smem + p.o_db2;     // [J_cap, Hp]   ... fc_local2 / fc_l2
  float* dG = smem + p.o_dG;       // [J_cap, Zp]
  float* pg1 = smem + p.o_pg1;     // [J_cap, Hp]
  float* pg2 = smem + p.o_pg2;     // [J_cap, Zp]
  float* din = smem + p.o_din;     // [J_cap, LDP]
  int* ints = reinterpret_cast<int*>(smem + p.o_int);
  int* jrow0 = ints;
  int* s_group = ints + p.J_cap + 1;
  short* rjet = reinterpret_cast<short*>(ints + p.J_cap + 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, Z = p.Z, F = p.F, LDH = p.LDH;
  const int CR = kWarps * RB;
  const int n_groups = *p.n_groups;
  const Lin* lin = p.lin;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_group[0] = atomicAdd(p.counter, 1);
    __syncthreads();
    const int gidx = s_group[0];
    if (gidx >= n_groups) break;
    const int2 grp = p.groups[gidx];
    const int j0 = grp.x, nj = grp.y;
    if (tid == 0) {
      int r = 0;
      for (int j = 0; j < nj; ++j) { jrow0[j] = r; r += p.n_real[j0 + j]; }
      jrow0[nj] = r;
    }
    __syncthreads();
    const int R = jrow0[nj];
    const int row_g0 = p.rowoff[j0];
    for (int j = 0; j < nj; ++j) {
      const int r0 = jrow0[j], n = jrow0[j + 1] - r0;
      for (int i = tid; i < n; i += kThreads) rjet[r0 + i] = (short)j;
    }
    for (int i = tid; i < p.R_cap * LDH; i += kThreads) dh[i] = 0.f;
    for (int i = tid; i < CR * LDH; i += kThreads) { tA[i] = 0.f; tB[i] = 0.f; }
    for (int i = tid; i < p.J_cap * p.Zp; i += kThreads) dG[i] = 0.f;
    __syncthreads();

    // ---------------- head: v = lrelu(fc_l3(h_L));  dpre3 is given ----------------
    {
      const Lin L3 = lin[p.n_lin - 1];
      const float* w3 = L3.Wt + (size_t)L3.m_off * L3.ldo;
      for (int i = tid; i < R * H; i += kThreads) {
        const int row = i / H, k = i - row * H;
        const float* d3 = p.dpre3 + (size_t)(row_g0 + row) * F;
        float a = 0.f;
        for (int f = 0; f < F; ++f) a = fmaf(__ldg(w3 + (size_t)k * L3.ldo + f), d3[f], a);
        dh[(size_t)row * LDH + k] = a;
      }
      for (int i = tid; i < nj * F; i += kThreads) {
        const int j = i / F, f = i - j * F;
        float a = 0.f;
        for (int r = jrow0[j]; r < jrow0[j + 1]; ++r) a += p.dpre3[(size_t)(row_g0 + r) * F + f];
        p.dbeff[(size_t)(j0 + j) * p.bstride + L3.bias_off + f] = a;
      }
      __syncthreads();
    }

    // ---------------- EPiC layers in reverse ----------------
    for (int l = p.L - 1; l >= 0; --l) {
      const Lin Ga = lin[LIN_LAYER0 + 4 * l + 0], Gb = lin[LIN_LAYER0 + 4 * l + 1];
      const Lin La = lin[LIN_LAYER0 + 4 * l + 2], Lb = lin[LIN_LAYER0 + 4 * l + 3];
      const int st_u = 2 + 2 * l, st_h = 3 + 2 * l;
      for (int i = tid; i < p.J_cap * p.Hp; i += kThreads) { db1[i] = 0.f; db2[i] = 0.f; }
      __syncthreads();
      for (int c0 = 0; c0 < R; c0 += CR) {
        float d[RB][TC];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int row = c0 + warp * RB + r;
#pragma unroll
          for (int i = 0; i < TC; ++i) {
            const int o = lane + 32 * i;
            d[r][i] = (row < R && o < H) ? dh[(size_t)row * LDH + o] : 0.f;
          }
        }
        float a_h[RB][TC], a_u[RB][TC];
        prefetch_act<TC, RB>(p, a_h, c0, R, row_g0, st_h);
        prefetch_act<TC, RB>(p, a_u, c0, R, row_g0, st_u);
        mask_store<TC, RB>(p, d, a_h, c0, R, row_g0, st_h, tA, db2, rjet);     // d pre(fc_local2)
        __syncwarp();
        gemm_rows<TC, RB>(tA, LDH, Lb.Wr, Lb.out, Lb.ldr, wbuf, p.wbuf_floats, p.KC, d);     // du = dpre2 . W2
        mask_store<TC, RB>(p, d, a_u, c0, R, row_g0, st_u, tB, db1, rjet);     // d pre(fc_local1)
        __syncwarp();
        gemm_rows<TC, RB>(tB, LDH, La.Wr, La.out, La.ldr, wbuf, p.wbuf_floats, p.KC, d);     // dpre1 . W1(main)
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int row = c0 + warp * RB + r;
          if (row < R) {
#pragma unroll
            for (int i = 0; i < TC; ++i) {
              const int o = lane + 32 * i;
              if (o < H) dh[(size_t)row * LDH + o] = tA[(warp * RB + r) * LDH + o] + d[r][i];     // residual + through fc_local1
            }
          }
        }
        __syncwarp();
      }
      __syncthreads();
      // gradient w.r.t. the layer's new global vector: carry + W_glob^T . db1
      for (int i = tid; i < nj * Z; i += kThreads) {
        const int j = i / Z, z = i - j * Z;
        const float* w = La.Wt + (size_t)(La.g_off + z) * La.ldo;
        float a = dG[j * p.Zp + z];
        for (int o = 0; o < H; o += 8) {          // 8 loads in flight (H is a multiple of 4; the tail is guarded)
          float wv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) wv[q] = o + q < H ? __ldg(w + o + q) : 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (o + q < H) a = fmaf(wv[q], db1[j * p.Hp + o + q], a);
        }
        dG[j * p.Zp + z] = a;
      }
      for (int i = tid; i < nj * H; i += kThreads) {
        const int j = i / H, o = i - j * H;
        p.dbeff[(size_t)(j0 + j) * p.bstride + La.bias_off + o] = db1[j * p.Hp + o];
        p.dbeff[(size_t)(j0 + j) * p.bstride + Lb.bias_off + o] = db2[j * p.Hp + o];
      }
      __syncthreads();
      global_backward(p, Ga, Gb, l + 1, j0, nj, jrow0, rjet, R, dh, dG, pg1, pg2, din);
    }

    // ---------------- stem ----------------
    {
      const Lin L1 = lin[LIN_L1], L2 = lin[LIN_L2], G1 = lin[LIN_G1], G2 = lin[LIN_G2];
      global_backward(p, G1, G2, 0, j0, nj, jrow0, rjet, R, dh, dG, pg1, pg2, din);
      for (int i = tid; i < p.J_cap * p.Hp; i += kThreads) { db1[i] = 0.f; db2[i] = 0.f; }
      __syncthreads();
      for (int c0 = 0; c0 < R; c0 += CR) {
        float d[RB][TC];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const int row = c0 + warp * RB + r;
#pragma unroll
          for (int i = 0; i < TC; ++i) {
            const int o = lane + 32 * i;
            d[r][i] = (row < R && o < H) ? dh[(size_t)row * LDH + o] : 0.f;
          }
        }
        float a_h[RB][TC], a_u[RB][TC];
        prefetch_act<TC, RB>(p, a_h, c0, R, row_g0, 1);
        prefetch_act<TC, RB>(p, a_u, c0, R, row_g0, 0);
        mask_store<TC, RB>(p, d, a_h, c0, R, row_g0, 1, tA, db2, rjet);        // d pre(fc_l2)
        __syncwarp();
        gemm_rows<TC, RB>(tA, LDH, L2.Wr, L2.out, L2.ldr, wbuf, p.wbuf_floats, p.KC, d);
#pragma unroll
        for (int r = 0; r < RB; ++r)
#pragma unroll
          for (int i = 0; i < TC; ++i) {
            const int o = lane + 32 * i;
            if (o < H) d[r][i] += tA[(warp * RB + r) * LDH + o];               // residual h1
          }
        mask_store<TC, RB>(p, d, a_u, c0, R, row_g0, 0, tB, db1, rjet);        // d pre(fc_l1)
        __syncwarp();
        if (p.dxs) {     // gradient w.r.t. the per-particle input columns
          for (int r = 0; r < RB; ++r) {
            const int row = c0 + warp * RB + r;
            if (row >= R) break;
            for (int c = 0; c < p.Kx; ++c) {
              const float* w = L1.Wt + (size_t)(L1.m_off + p.xin_off + c) * L1.ldo;
              float a = 0.f;
              for (int o = lane; o < H; o += 32) a = fmaf(__ldg(w + o), tB[(warp * RB + r) * LDH + o], a);
#pragma unroll
              for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
              if (lane == 0) p.dxs[(size_t)(row_g0 + row) * p.Kx + c] = a;
            }
          }
        }
        __syncwarp();
      }
      __syncthreads();
      for (int i = tid; i < nj * H; i += kThreads) {
        const int j = i / H, o = i - j * H;
        p.dbeff[(size_t)(j0 + j) * p.bstride + L1.bias_off + o] = db1[j * p.Hp + o];
        p.dbeff[(size_t)(j0 + j) * p.bstride + L2.bias_off + o] = db2[j * p.Hp + o];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// weight gradients:  dW[o*ldw + col0 + c] += sum_r Y[r*ldy + o] * X[r*ldx + c]      (o < out, c < K, r < rows)
// grid = (row chunks, tiles of all jobs); one block = one 128 x 128 output tile over one chunk of rows.
// ---------------------------------------------------------------------------------------------
static constexpr int XT = 128;        // output tile edge
static constexpr int XR = 16;         // rows staged per iteration
static constexpr int X_CHUNK = 512;   // rows per block (2048 was measured 20 % slower: too few blocks)

__device__ __forceinline__ void xty_tile(const XtyJob& J, int rows, int t_local) {
  __shared__ __align__(16) float sY[XR][XT];
  __shared__ __align__(16) float sX[XR][XT];
  const int r_begin = blockIdx.x * X_CHUNK;
  if (r_begin >= rows) return;
  const int r_end = min(rows, r_begin + X_CHUNK);
  const int o0 = (t_local / J.tiles_k) * XT, k0 = (t_local % J.tiles_k) * XT;
  const int tid = threadIdx.x;
  const int to = tid >> 4, tk = tid & 15;      // 16 x 16 threads, 8 x 8 outputs each
  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += XR) {
    for (int i = tid; i < XR * XT; i += 256) {
      const int rr = i / XT, c = i - rr * XT;
      const int r = r0 + rr;
      float y = 0.f, x = 0.f;
      if (r < r_end) {
        if (o0 + c < J.out) y = J.Y[(size_t)r * J.ldy + o0 + c];
        if (k0 + c < J.K) x = J.X[(size_t)r * J.ldx + k0 + c];
      }
      sY[rr][c] = y;
      sX[rr][c] = x;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < XR; ++rr) {
      float y[8], x[8];
      *reinterpret_cast<float4*>(&y[0]) = *reinterpret_cast<const float4*>(&sY[rr][to * 8]);
      *reinterpret_cast<float4*>(&y[4]) = *reinterpret_cast<const float4*>(&sY[rr][to * 8 + 4]);
      *reinterpret_cast<float4*>(&x[0]) = *reinterpret_cast<const float4*>(&sX[rr][tk * 8]);
      *reinterpret_cast<float4*>(&x[4]) = *reinterpret_cast<const float4*>(&sX[rr][tk * 8 + 4]);
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(y[a], x[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int o = o0 + to * 8 + a;
    if (o >= J.out) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int k = k0 + tk * 8 + b;
      if (k < J.K) atomicAdd(J.dW + (size_t)o * J.ldw + J.col0 + k, acc[a][b]);
    }
  }
}

__global__ void __launch_bounds__(256) xty_kernel(const XtyJob* __restrict__ jobs, int n_jobs, const int* __restrict__ n_total) {
  // locate the job of this tile (a few dozen jobs: every thread scans, it is cheap)
  int jb = 0;
  const int tile = blockIdx.y;
  for (int i = 1; i < n_jobs; ++i)
    if (jobs[i].tile0 <= tile) jb = i;
  const XtyJob J = jobs[jb];
  xty_tile(J, J.rows >= 0 ? J.rows : *n_total, tile - J.tile0);
}

// one job passed by value (the droid training path launches one product per linear)
__global__ void __launch_bounds__(256) xty_one_kernel(const XtyJob J) { xty_tile(J, J.rows, blockIdx.y); }

// PFM_XTY_SIMT=1 selects the fp32 CUDA-core weight-gradient kernels (debugging / A-B measurements)
bool xty_use_simt() {
  static const bool v = [] { const char* e = getenv("PFM_XTY_SIMT"); return e && e[0] == '1'; }();
  return v;
}

int xty_launch_one(const float* Y, int ldy, const float* X, int ldx, float* dW, int ldw, int out, int K, int col0, int rows,
                   cudaStream_t st) {
  if (rows <= 0 || out <= 0 || K <= 0) return PFM_OK;
  XtyJob J;
  J.Y = Y; J.X = X; J.dW = dW; J.ldy = ldy; J.ldx = ldx; J.ldw = ldw; J.out = out; J.K = K; J.col0 = col0; J.rows = rows; J.tile0 = 0;
  J.tiles_o = (out + XT - 1) / XT; J.tiles_k = (K + XT - 1) / XT;
  dim3 grid((unsigned)((rows + X_CHUNK - 1) / X_CHUNK), (unsigned)(J.tiles_o * J.tiles_k));
  xty_one_kernel<<<grid, 256, 0, st>>>(J);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct BwdShape { int TC, RB; bool spill; BwdParams p; size_t smem; };

static int bwd_plan(const pfm_epic* h, int N, int TC, int RB, int KC, int J_cap, int R_cap_force, BwdShape* s) {
  const pfm_epic_cfg& c = h->cfg;
  const int H = c.hid, Z = c.latent;
  const int Hp = (H + 3) & ~3, Zp = (Z + 3) & ~3;
  const int LDH = Hp + 4;
  const int LDP = ((2 * H + Z + 3) & ~3);
  const int CR = kWarps * RB;
  const int budget = h->max_smem_optin / 4;
  const int wbuf = ((KC * Hp + 32 * TC + 32) + 3) & ~3;
  const int per_jet = 3 * Hp + 2 * Zp + LDP;
  const int fixed = 2 * CR * LDH + 2 * wbuf + J_cap * per_jet + (J_cap + 8) + 64;
  int R_cap = (budget - fixed) / (LDH + 1);
  if (R_cap > 1024) R_cap = 1024;
  bool spill = false;
  if (R_cap < N || (R_cap_force > 0 && R_cap_force > R_cap)) {     // spill mode: dh in global memory
    spill = true;
    R_cap = (budget - fixed) / 1;
    const int want = N > 512 ? N : 512;
    if (R_cap > want) R_cap = want;
  }
  if (R_cap_force > 0) {
    if (R_cap_force > R_cap) { set_error("internal: backward plan smaller than the forced capacity"); return PFM_ERR_INVALID; }
    R_cap = R_cap_force;
  }
  if (R_cap < N) {
    set_error("training: a jet of %d particles does not fit the backward kernel's shared-memory budget (%d rows at hid=%d)",
              N, R_cap, H);
    return PFM_ERR_UNSUPPORTED;
  }
  s->spill = spill;
  s->TC = TC; s->RB = RB;
  BwdParams& p = s->p;
  memset(&p, 0, sizeof(p));
  p.F = c.feats; p.H = H; p.Hp = Hp; p.LDH = LDH; p.Z = Z; p.Zp = Zp; p.L = c.layers; p.n_lin = h->n_lin; p.LDP = LDP;
  p.R_cap = R_cap; p.J_cap = J_cap; p.KC = KC;
  p.sum_scale = c.sum_scale; p.slope = c.neg_slope;
  p.wbuf_floats = wbuf;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  p.o_dh = spill ? 0 : take(R_cap * LDH);
  p.o_tA = take(CR * LDH);
  p.o_tB = take(CR * LDH);
  p.o_wbuf = take(2 * wbuf);
  p.o_db1 = take(J_cap * Hp);
  p.o_db2 = take(J_cap * Hp);
  p.o_dG = take(J_cap * Zp);
  p.o_pg1 = take(J_cap * Hp);
  p.o_pg2 = take(J_cap * Zp);
  p.o_din = take(J_cap * LDP);
  p.o_int = take(J_cap + 8 + (R_cap + 1) / 2);
  p.total_floats = o;
  s->smem = (size_t)o * 4;
  if ((int)s->smem > h->max_smem_optin) {
    set_error("training: backward shared-memory plan %zu B exceeds the device limit %d B", s->smem, h->max_smem_optin);
    return PFM_ERR_UNSUPPORTED;
  }
  return PFM_OK;
}

// One plan for the forward and the backward kernel (the smaller of the two capacities), and the strides of the
// saved-activation arrays for a batch of B jets x N particles.
int train_layout(const pfm_epic* h, int B, int N, TrainLayout* lay) {
  const pfm_epic_cfg& c = h->cfg;
  int R_f, J_f, TC, RB, KC;
  int rc = simt_caps_for_train(h, N, &R_f, &J_f, &TC, &RB, &KC);
  if (rc != PFM_OK) return rc;
  if (J_f > kJMax) J_f = kJMax;
  BwdShape bs;
  rc = bwd_plan(h, N, TC, RB, KC, J_f, 0, &bs);
  if (rc != PFM_OK) return rc;
  lay->R_cap = R_f < bs.p.R_cap ? R_f : bs.p.R_cap;
  lay->J_cap = J_f;
  lay->Hp = (c.hid + 3) & ~3;
  lay->Zp = (c.latent + 3) & ~3;
  lay->LDP = (2 * c.hid + c.latent + 3) & ~3;
  lay->junit = lay->LDP + lay->Hp + lay->Zp;
  lay->jstride = (c.layers + 1) * lay->junit;
  lay->stage_stride = (size_t)B * N * lay->Hp;
  return PFM_OK;
}

template <int TC, int RB, bool SPILL>
static int launch_bwd_s(const BwdShape& s, int grid, cudaStream_t st) {
  auto kern = epic_bwd_kernel<TC, RB, SPILL>;
  PFM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
  kern<<<grid, kThreads, s.smem, st>>>(s.p);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}
template <int TC, int RB>
static int launch_bwd(const BwdShape& s, int grid, cudaStream_t st) {
  return s.p.dh_spill ? launch_bwd_s<TC, RB, true>(s, grid, st) : launch_bwd_s<TC, RB, false>(s, grid, st);
}

__global__ void reset_counter_kernel(int* counter) { *counter = 0; }

// per-particle / per-jet pre-activation gradients on the fp32 CUDA-core kernel
static int train_backward_simt(pfm_epic* h, const TrainBwdArgs& a, cudaStream_t st) {
  int R_f, J_f, TC, RB, KC;
  int rc = simt_caps_for_train(h, a.N, &R_f, &J_f, &TC, &RB, &KC);
  if (rc != PFM_OK) return rc;
  BwdShape s;
  rc = bwd_plan(h, a.N, TC, RB, KC, a.lay.J_cap, a.lay.R_cap, &s);
  if (rc != PFM_OK) return rc;
  BwdParams& p = s.p;
  p.Kx = a.Kx; p.xin_off = a.xin_off;
  p.lin = h->lin_dev;
  p.n_real = h->plan.n_real; p.groups = h->plan.groups; p.n_groups = h->plan.n_groups; p.counter = h->plan.counter;
  p.rowoff = h->plan.rowoff;
  p.act = h->act; p.dact = h->dact; p.stage_stride = a.lay.stage_stride; p.Hp_act = a.lay.Hp;
  p.jact = h->jact; p.junit = a.lay.junit; p.jstride = a.lay.jstride;
  p.dpre3 = h->dpre3;
  p.dbeff = h->dbeff; p.bstride = h->bstride;
  p.dxs = a.want_dx ? h->dxs : nullptr;
  p.dh_spill = nullptr;
  if (s.spill) {
    const size_t need = (size_t)(h->sm_count < a.B ? h->sm_count : a.B) * p.R_cap * p.LDH;
    if (need > h->dh_spill_cap) {
      if (h->dh_spill) cudaFree(h->dh_spill);
      h->dh_spill = nullptr; h->dh_spill_cap = 0;
      PFM_CUDA_CHECK(cudaMalloc(&h->dh_spill, sizeof(float) * need));
      h->dh_spill_cap = need;
    }
    p.dh_spill = h->dh_spill;
  }
  reset_counter_kernel<<<1, 1, 0, st>>>(h->plan.counter);
  const int grid = h->sm_count < a.B ? h->sm_count : a.B;
  if (TC == 4) rc = launch_bwd<4, 8>(s, grid, st);
  else if (TC == 5) rc = launch_bwd<5, 8>(s, grid, st);
  else rc = launch_bwd<10, 8>(s, grid, st);
  if (rc != PFM_OK) return rc;
  h->last_launches += 2;
  return PFM_OK;
}

int train_backward(pfm_epic* h, const TrainBwdArgs& a, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  int rc;
  if (tt_enabled(h)) {
    rc = tt_train_backward(h, a, st);                    // tensor-core path: epic_train_tc.cu
  } else {
    if (h->train_tc) {                                   // the forward ran on the tensor-core path: its plan has no CTA groups
      rc = train_plan_groups(h, a.B, a.lay, st);
      if (rc != PFM_OK) return rc;
      h->train_tc = false;
    }
    rc = train_backward_simt(h, a, st);
  }
  if (rc != PFM_OK) return rc;
  if (!a.grad_flat) return PFM_OK;

  // ---- weight-gradient job table ----
  // The jobs are cut into kGradChunks launches by linear index; an event after each lets a data-parallel host start the
  // all-reduce of that slice of the flat gradient while the next chunk's weight gradients are still being computed.
  std::vector<XtyJob> jobs;
  int tile = 0;
  const int n_chunks = h->n_lin < kGradChunks ? h->n_lin : kGradChunks;
  std::vector<int> chunk_job0(n_chunks + 1, 0), chunk_tiles(n_chunks, 0);
  h->grad_chunk_off.assign(n_chunks + 1, 0);
  size_t maxrows_jet = (size_t)a.B, maxrows_part = (size_t)a.B * a.N;
  auto add = [&](const float* Y, int ldy, int out, const float* X, int ldx, int K, int rows, float* dW, int ldw, int col0) {
    if (K <= 0 || out <= 0) return;
    XtyJob j;
    j.Y = Y; j.X = X; j.dW = dW; j.ldy = ldy; j.ldx = ldx; j.ldw = ldw; j.out = out; j.K = K; j.col0 = col0; j.rows = rows;
    j.tile0 = tile; j.tiles_o = (out + XT - 1) / XT; j.tiles_k = (K + XT - 1) / XT;
    tile += j.tiles_o * j.tiles_k;
    jobs.push_back(j);
  };
  const TrainLayout& lay = a.lay;
  const int Hp = lay.Hp;
  size_t off = 0;
  int cur_chunk = 0;
  for (int i = 0; i < h->n_lin; ++i) {
    const int chunk = (int)((long long)i * n_chunks / h->n_lin);
    if (chunk != cur_chunk) {                       // close the previous chunk
      chunk_tiles[cur_chunk] = tile; chunk_job0[chunk] = (int)jobs.size(); h->grad_chunk_off[chunk] = (long long)off;
      cur_chunk = chunk; tile = 0;
    }
    const Lin& L = h->lin_host[i];
    float* gW = a.grad_flat + off;
    float* gb = gW + (size_t)L.out * L.in;
    off += (size_t)L.out * L.in + L.out;
    const float* dbe = h->dbeff + L.bias_off;                 // [B][bstride] slice of this linear
    // bias, time, cond columns: sums over jets
    add(dbe, h->bstride, L.out, h->ones, 0, 1, a.B, gb, 1, 0);
    if (L.t_len > 0) add(dbe, h->bstride, L.out, a.t_code, a.t_ld, L.t_len, a.B, gW, L.in, L.t_off);
    if (L.c_len > 0) add(dbe, h->bstride, L.out, a.cond, a.cond_dim, L.c_len, a.B, gW, L.in, L.c_off);
    const bool is_l1 = (i == LIN_L1), is_l3 = (i == h->n_lin - 1);
    int r = (i >= LIN_LAYER0 && !is_l3) ? ((i - LIN_LAYER0) & 3) : -1;
    int l = (i >= LIN_LAYER0 && !is_l3) ? ((i - LIN_LAYER0) >> 2) : -1;
    if (is_l1) {
      if (a.t_in > 0) add(dbe, h->bstride, L.out, a.t_code_in, a.t_in, a.t_in, a.B, gW, L.in, L.m_off);     // hoisted input-time columns
      add(h->dact, Hp, L.out, h->yact, a.Kx, a.Kx, -1, gW, L.in, L.m_off + a.xin_off);
    } else if (i == LIN_L2) {
      add(h->dact + lay.stage_stride, Hp, L.out, h->act, Hp, c.hid, -1, gW, L.in, L.m_off);
    } else if (i == LIN_G1) {
      add(dbe, h->bstride, L.out, h->jact, lay.jstride, L.m_len, a.B, gW, L.in, L.m_off);
    } else if (i == LIN_G2) {
      add(dbe, h->bstride, L.out, h->jact + lay.LDP, lay.jstride, L.m_len, a.B, gW, L.in, L.m_off);
    } else if (is_l3) {
      add(h->dpre3, c.feats, L.out, h->act + (size_t)(1 + 2 * c.layers) * lay.stage_stride, Hp, c.hid, -1, gW, L.in, L.m_off);
    } else if (r == 0) {        // fc_global1: input = saved pool row of unit l+1
      add(dbe, h->bstride, L.out, h->jact + (size_t)(l + 1) * lay.junit, lay.jstride, L.m_len, a.B, gW, L.in, L.m_off);
    } else if (r == 1) {        // fc_global2: input = g1 of unit l+1
      add(dbe, h->bstride, L.out, h->jact + (size_t)(l + 1) * lay.junit + lay.LDP, lay.jstride, L.m_len, a.B, gW, L.in, L.m_off);
    } else if (r == 2) {        // fc_local1: particle block from h_l, global block from the unit's new global vector
      add(h->dact + (size_t)(2 + 2 * l) * lay.stage_stride, Hp, L.out, h->act + (size_t)(1 + 2 * l) * lay.stage_stride, Hp,
          c.hid, -1, gW, L.in, L.m_off);
      add(dbe, h->bstride, L.out, h->jact + (size_t)(l + 1) * lay.junit + lay.LDP + Hp, lay.jstride, L.g_len, a.B, gW, L.in, L.g_off);
    } else {                    // fc_local2: input = u_l
      add(h->dact + (size_t)(3 + 2 * l) * lay.stage_stride, Hp, L.out, h->act + (size_t)(2 + 2 * l) * lay.stage_stride, Hp,
          c.hid, -1, gW, L.in, L.m_off);
    }
  }
  chunk_tiles[cur_chunk] = tile; chunk_job0[n_chunks] = (int)jobs.size(); h->grad_chunk_off[n_chunks] = (long long)off;
  const size_t bytes = sizeof(XtyJob) * jobs.size();
  if (h->jobs_cap < bytes) {
    if (h->jobs_dev) cudaFree(h->jobs_dev);
    h->jobs_dev = nullptr; h->jobs_cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(&h->jobs_dev, bytes));
    h->jobs_cap = bytes;
  }
  rc = upload_table(h, 2, h->jobs_dev, jobs.data(), bytes, st);      // pinned staging, skipped when unchanged (CUDA-graph safe)
  if (rc != PFM_OK) return rc;
  const size_t maxrows = maxrows_part > maxrows_jet ? maxrows_part : maxrows_jet;
  while ((int)h->grad_ev.size() < n_chunks) {
    cudaEvent_t e;
    PFM_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->grad_ev.push_back(e);
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int nj = chunk_job0[c + 1] - chunk_job0[c];
    if (nj > 0 && chunk_tiles[c] > 0) {
      const XtyJob* jd = reinterpret_cast<const XtyJob*>(h->jobs_dev) + chunk_job0[c];
      if (!xty_use_simt()) {                         // tensor cores, 3-term bf16 split (fp32-accurate): xty_tc.cu
        int rc = xty_tc_launch(jd, nj, chunk_tiles[c], h->plan.n_total, (int)maxrows, h->sm_count, st);
        if (rc != PFM_OK) return rc;
      } else {
        dim3 grid2((unsigned)((maxrows + X_CHUNK - 1) / X_CHUNK), (unsigned)chunk_tiles[c]);
        xty_kernel<<<grid2, 256, 0, st>>>(jd, nj, h->plan.n_total);
      }
      h->last_launches += 1;
    }
    PFM_CUDA_CHECK(cudaEventRecord(h->grad_ev[c], st));
  }
  h->grad_chunks = n_chunks;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // namespace pfm
