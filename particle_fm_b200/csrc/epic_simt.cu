// fp32 CUDA-core ("strict") path of the EPiC vector field + fixed-step integrator.
//
// One persistent CTA owns a GROUP of consecutive jets whose REAL particles (padding is skipped --
// exact) fit its shared-memory row budget, and keeps for the whole integration
//   - the ODE state of those particles            xs / x0s   [rows, feats]
//   - the per-particle hidden features             hs         [rows, hid]    fp32
//   - the per-jet pooled / global vectors, biases  (small)
// so that a full sample() is ONE launch: jets are independent, no grid-wide sync is needed.
// Per evaluation it performs exactly the arithmetic of EPiC_encoder.forward (epic.py:304-391) in the
// hoisted form: every concat-linear  W.[time | main | global | cond] + b  is evaluated as
//   W_main . main  +  (b + W_time . time)  [bias table, per evaluation]  +  W_cond . cond  [per jet]
//   (+ W_glob . g  per jet, for fc_local1),
// which is the same sum in a different association order (fp32 rounding only).
#include "pfm_internal.cuh"
#include "simt_common.cuh"

namespace pfm {

struct SimtParams {
  int F, Kx, Kxp, x_ld, xin_off, H, Hp, LDH, Z, L, n_lin;
  int R_cap, J_cap, KC, LDX;
  float sum_scale, slope;
  const Lin* lin;
  const float* tbias; const float* cbias; int bstride; int tbias_per_jet;
  const int* n_real; const uint16_t* ridx; const int2* groups; const int* n_groups; int* counter;
  const int* jetmap;           // position -> jet (inference plan; nullptr = identity, training)
  const float* x_in; float* x_out; int B, N;
  int n_evals, solver, n_steps; const float* dt;
  int step_kind; const float* coef; const float* noise; long long noise_step_stride; int jet0;   // diffusion step program (RunArgs)
  // training forward (TRAIN instantiation): interpolation inputs, saved activations, loss
  const float* x1; const float* tjet; const float* noise0; const float* noise1; int loss_kind; float sigma;
  float* act; size_t act_stage_stride; int Hp_act;          // act[stage][row][Hp_act]
  float* yact;                                              // yact[row][Kx]   network input of every real particle
  float* jact; int junit; int jstride; int LDP_act;         // jact[jet][unit][pool input (LDP) | g1 (Hp) | g]
  float* dpre3; float* loss_acc; const int* rowoff; const int* n_total;
  // spill mode (jets too large for shared memory, e.g. LHCO 279 x H150): the hidden features live in a per-CTA slab
  // of global memory (L2-resident) instead of shared memory; everything else is unchanged
  float* hs_spill;
  // shared-memory carve-up (float offsets)
  int o_xs, o_x0, o_hs, o_tmp, o_wbuf, o_pool, o_g, o_g1, o_bl1, o_bl2, o_v, o_int, total_floats;
  int wbuf_floats, LDB, LDP;
};

// per-jet effective bias of one linear: time table row + cond table row (+ W_glob . g)
__device__ __forceinline__ float bias_of(const SimtParams& p, const Lin& L, int eval, int jet_global, int o) {
  const int trow = p.tbias_per_jet ? jet_global : eval;
  float b = p.tbias[(size_t)trow * p.bstride + L.bias_off + o];
  if (p.cbias) b += p.cbias[(size_t)jet_global * p.bstride + L.bias_off + o];
  return b;
}

// a + sum_k w[k*ldo] * in[k]  (one output column of a k-major weight block, weights from L2): 8 loads in flight per step
// -- these per-jet products are latency-bound, not bandwidth-bound; same summation order as a plain loop
__device__ __forceinline__ float gemv_col(const float* __restrict__ w, int ldo, const float* __restrict__ in, int K, float a) {
  int k = 0;
  for (; k + 8 <= K; k += 8) {
    float wv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) wv[q] = __ldg(w + (size_t)(k + q) * ldo);
#pragma unroll
    for (int q = 0; q < 8; ++q) a = fmaf(wv[q], in[k + q], a);
  }
  for (; k < K; ++k) a = fmaf(__ldg(w + (size_t)k * ldo), in[k], a);
  return a;
}

template <int TC, int RB, bool TRAIN, bool SPILL>
__global__ void __launch_bounds__(kThreads, 1) epic_simt_kernel(const SimtParams p) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem + p.o_xs;      // [R_cap, LDX]   current network input (state or midpoint state)
  float* x0 = smem + p.o_x0;      // [R_cap, F]     state at the start of the step
  float* hs = SPILL ? p.hs_spill + (size_t)blockIdx.x * p.R_cap * p.LDH : smem + p.o_hs;      // [R_cap, LDH]; SPILL is a compile-time flag so that the common case compiles to shared-memory instructions
  float* tmp = smem + p.o_tmp;    // [8*RB, LDH]
  float* wbuf = smem + p.o_wbuf;  // 2 stages
  float* pool = smem + p.o_pool;  // [J_cap, LDP]   (LDP >= 2H + Z)
  float* gv = smem + p.o_g;       // [J_cap, Z]
  float* g1 = smem + p.o_g1;      // [J_cap, Hp]
  float* bl1 = smem + p.o_bl1;    // [J_cap, LDB]
  float* bl2 = smem + p.o_bl2;    // [J_cap, LDB]
  float* vbuf = smem + p.o_v;     // [R_cap, F]     network output
  int* ints = reinterpret_cast<int*>(smem + p.o_int);
  int* jrow0 = ints;                       // [J_cap + 1] first row of every jet of the group
  int* s_group = ints + p.J_cap + 1;       // [2]
  int* jid = ints + p.J_cap + 4;           // [J_cap] batch index of the group's jets
  short* rjet = reinterpret_cast<short*>(ints + 2 * p.J_cap + 4);   // [R_cap] jet (inside the group) of a row

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, Z = p.Z, F = p.F, LDH = p.LDH, LDX = p.LDX, LDB = p.LDB, LDP = p.LDP;
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
      for (int j = 0; j < nj; ++j) {
        jid[j] = p.jetmap ? p.jetmap[j0 + j] : j0 + j;
        jrow0[j] = r;
        r += p.n_real[jid[j]];
      }
      jrow0[nj] = r;
    }
    __syncthreads();
    const int R = jrow0[nj];
    const int row_g0 = TRAIN ? p.rowoff[j0] : 0;     // first packed row of this group in the saved-activation arrays
    // zero everything that later relies on zero padding, then load the group's particles
    for (int i = tid; i < p.R_cap * LDX; i += kThreads) xs[i] = 0.f;
    for (int i = tid; i < p.R_cap * LDH; i += kThreads) hs[i] = 0.f;
    for (int i = tid; i < CR * LDH; i += kThreads) tmp[i] = 0.f;
    __syncthreads();
    for (int j = 0; j < nj; ++j) {
      const int r0 = jrow0[j], n = jrow0[j + 1] - r0;
      for (int i = tid; i < n; i += kThreads) rjet[r0 + i] = (short)j;
      for (int i = tid; i < n * p.Kx; i += kThreads) {
        const int r = i / p.Kx, c = i - r * p.Kx;
        const int part = p.ridx[(size_t)jid[j] * p.N + r];
        const size_t gi = ((size_t)jid[j] * p.N + part) * p.x_ld + c;
        float v;
        if (TRAIN && p.loss_kind >= 0) {
          // flow-matching interpolation y(x1, t, noise) and target u_t (losses.py:56-62, :115-119, :320-326);
          // the target is parked in x0 until the loss is evaluated
          const float x = p.x1[gi], t = p.tjet[jid[j]], z = p.noise0[gi];
          float u;
          if (p.loss_kind == PFM_LOSS_FM_OT) {
            v = (1.f - t) * x + (p.sigma + (1.f - p.sigma) * t) * z;
            u = (1.f - p.sigma) * z - x;
          } else if (p.loss_kind == PFM_LOSS_CFM) {
            v = ((1.f - t) * x + t * z) + p.sigma * p.noise1[gi];
            u = z - x;
          } else {
            v = x + t * z;
            u = z;
          }
          x0[(r0 + r) * F + c] = u;
        } else {
          v = p.x_in[gi];
          if (p.solver >= 0) x0[(r0 + r) * F + c] = v;
        }
        xs[(r0 + r) * LDX + c] = v;
        if (TRAIN) p.yact[(size_t)(row_g0 + r0 + r) * p.Kx + c] = v;
      }
    }
    __syncthreads();

    for (int ev = 0; ev < p.n_evals; ++ev) {
      // ---------------- stem: fc_l1, fc_l2 (epic.py:360-366) ----------------
      {
        const Lin L1 = lin[LIN_L1], L2 = lin[LIN_L2];
        for (int i = tid; i < nj * H; i += kThreads) {
          const int j = i / H, o = i - j * H;
          bl1[j * LDB + o] = bias_of(p, L1, ev, jid[j], o);
          bl2[j * LDB + o] = bias_of(p, L2, ev, jid[j], o);
        }
        __syncthreads();
        for (int c0 = 0; c0 < R; c0 += CR) {
          float acc[RB][TC];
          gemm_rows<TC, RB>(xs + (size_t)c0 * LDX, LDX, L1.Wt + (size_t)(L1.m_off + p.xin_off) * L1.ldo, p.Kx, L1.ldo,
                            wbuf, p.wbuf_floats, p.KC, acc);
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int row = c0 + warp * RB + r;
            if (row < R) {
              const float* bj = bl1 + rjet[row] * LDB;
#pragma unroll
              for (int i = 0; i < TC; ++i) {
                const int o = lane + 32 * i;
                if (o < H) {
                  const float hv = lrelu(acc[r][i] + bj[o], p.slope);
                  tmp[(warp * RB + r) * LDH + o] = hv;
                  if (TRAIN) p.act[(size_t)(row_g0 + row) * p.Hp_act + o] = hv;                       // stage 0: h1
                }
              }
            }
          }
          __syncwarp();
          gemm_rows<TC, RB>(tmp, LDH, L2.Wt + (size_t)L2.m_off * L2.ldo, H, L2.ldo, wbuf, p.wbuf_floats, p.KC, acc);
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int row = c0 + warp * RB + r;
            if (row < R) {
              const float* bj = bl2 + rjet[row] * LDB;
#pragma unroll
              for (int i = 0; i < TC; ++i) {
                const int o = lane + 32 * i;
                if (o < H) {
                  const float hv = lrelu(acc[r][i] + bj[o] + tmp[(warp * RB + r) * LDH + o], p.slope);
                  hs[(size_t)row * LDH + o] = hv;
                  if (TRAIN) p.act[p.act_stage_stride + (size_t)(row_g0 + row) * p.Hp_act + o] = hv;  // stage 1: h0
                }
              }
            }
          }
          __syncwarp();
        }
        __syncthreads();
      }
      // ---------------- stem pooling + fc_g1, fc_g2 (epic.py:369-380) ----------------
      {
        for (int i = tid; i < nj * H; i += kThreads) {
          const int j = i / H, o = i - j * H;
          const int r0 = jrow0[j], r1 = jrow0[j + 1];
          float s = 0.f;
          for (int r = r0; r < r1; ++r) s += hs[(size_t)r * LDH + o];
          pool[j * LDP + o] = s * p.sum_scale;                    // (sum, mean) order in the stem, :373
          pool[j * LDP + H + o] = s / (float)(r1 - r0);
          if (TRAIN) {                                                                                 // unit 0: pool input
            p.jact[(size_t)jid[j] * p.jstride + o] = pool[j * LDP + o];
            p.jact[(size_t)jid[j] * p.jstride + H + o] = pool[j * LDP + H + o];
          }
        }
        __syncthreads();
        const Lin G1 = lin[LIN_G1], G2 = lin[LIN_G2];
        for (int i = tid; i < nj * H; i += kThreads) {
          const int j = i / H, o = i - j * H;
          float a = bias_of(p, G1, ev, jid[j], o);
          const float* w = G1.Wt + (size_t)G1.m_off * G1.ldo + o;
          const float* in = pool + j * LDP;
          a = gemv_col(w, G1.ldo, in, 2 * H, a);
          g1[j * p.Hp + o] = lrelu(a, p.slope);
          if (TRAIN) p.jact[(size_t)jid[j] * p.jstride + p.LDP_act + o] = g1[j * p.Hp + o];        // unit 0: g1
        }
        __syncthreads();
        for (int i = tid; i < nj * Z; i += kThreads) {
          const int j = i / Z, o = i - j * Z;
          float a = bias_of(p, G2, ev, jid[j], o);
          const float* w = G2.Wt + (size_t)G2.m_off * G2.ldo + o;
          const float* in = g1 + j * p.Hp;
          a = gemv_col(w, G2.ldo, in, H, a);
          gv[j * Z + o] = lrelu(a, p.slope);                       // no residual in the stem, :378-380
          if (TRAIN) p.jact[(size_t)jid[j] * p.jstride + p.LDP_act + p.Hp_act + o] = gv[j * Z + o]; // unit 0: g
        }
        __syncthreads();
      }
      // ---------------- EPiC layers (epic.py:85-203) ----------------
      for (int l = 0; l < p.L; ++l) {
        const Lin Ga = lin[LIN_LAYER0 + 4 * l + 0], Gb = lin[LIN_LAYER0 + 4 * l + 1];
        const Lin La = lin[LIN_LAYER0 + 4 * l + 2], Lb = lin[LIN_LAYER0 + 4 * l + 3];
        for (int i = tid; i < nj * H; i += kThreads) {
          const int j = i / H, o = i - j * H;
          const int r0 = jrow0[j], r1 = jrow0[j + 1];
          float s = 0.f;
          for (int r = r0; r < r1; ++r) s += hs[(size_t)r * LDH + o];
          pool[j * LDP + o] = s / (float)(r1 - r0);                // (mean, sum, global) order, :164-171
          pool[j * LDP + H + o] = s * p.sum_scale;
          if (TRAIN) {                                                                                // unit l+1: pool input
            float* ja = p.jact + (size_t)jid[j] * p.jstride + (size_t)(l + 1) * p.junit;
            ja[o] = pool[j * LDP + o];
            ja[H + o] = pool[j * LDP + H + o];
          }
        }
        for (int i = tid; i < nj * Z; i += kThreads) {
          const int j = i / Z, o = i - j * Z;
          pool[j * LDP + 2 * H + o] = gv[j * Z + o];
          if (TRAIN) p.jact[(size_t)jid[j] * p.jstride + (size_t)(l + 1) * p.junit + 2 * H + o] = gv[j * Z + o];
        }
        __syncthreads();
        for (int i = tid; i < nj * H; i += kThreads) {            // fc_global1, :180-182
          const int j = i / H, o = i - j * H;
          float a = bias_of(p, Ga, ev, jid[j], o);
          const float* w = Ga.Wt + (size_t)Ga.m_off * Ga.ldo + o;
          const float* in = pool + j * LDP;
          a = gemv_col(w, Ga.ldo, in, 2 * H + Z, a);
          g1[j * p.Hp + o] = lrelu(a, p.slope);
          if (TRAIN) p.jact[(size_t)jid[j] * p.jstride + (size_t)(l + 1) * p.junit + p.LDP_act + o] = g1[j * p.Hp + o];
        }
        __syncthreads();
        for (int i = tid; i < nj * Z; i += kThreads) {            // fc_global2 + residual, :184-186
          const int j = i / Z, o = i - j * Z;
          float a = bias_of(p, Gb, ev, jid[j], o);
          const float* w = Gb.Wt + (size_t)Gb.m_off * Gb.ldo + o;
          const float* in = g1 + j * p.Hp;
          a = gemv_col(w, Gb.ldo, in, H, a);
          // the new global vector is written to the pool row first (gv is still being read as the residual
          // by other threads only through gv[j*Z+o] of the SAME (j,o) -> safe in place)
          gv[j * Z + o] = lrelu(a + gv[j * Z + o], p.slope);
          if (TRAIN) p.jact[(size_t)jid[j] * p.jstride + (size_t)(l + 1) * p.junit + p.LDP_act + p.Hp_act + o] = gv[j * Z + o];
        }
        __syncthreads();
        for (int i = tid; i < nj * H; i += kThreads) {            // per-jet bias of fc_local1 / fc_local2
          const int j = i / H, o = i - j * H;
          float a = bias_of(p, La, ev, jid[j], o);
          const float* w = La.Wt + (size_t)La.g_off * La.ldo + o;
          const float* in = gv + j * Z;
          a = gemv_col(w, La.ldo, in, Z, a);
          bl1[j * LDB + o] = a;
          bl2[j * LDB + o] = bias_of(p, Lb, ev, jid[j], o);
        }
        __syncthreads();
        for (int c0 = 0; c0 < R; c0 += CR) {                      // fc_local1, fc_local2 + residual, :189-200
          float acc[RB][TC];
          gemm_rows<TC, RB>(hs + (size_t)c0 * LDH, LDH, La.Wt + (size_t)La.m_off * La.ldo, H, La.ldo, wbuf,
                            p.wbuf_floats, p.KC, acc);
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int row = c0 + warp * RB + r;
            if (row < R) {
              const float* bj = bl1 + rjet[row] * LDB;
#pragma unroll
              for (int i = 0; i < TC; ++i) {
                const int o = lane + 32 * i;
                if (o < H) {
                  const float uv = lrelu(acc[r][i] + bj[o], p.slope);
                  tmp[(warp * RB + r) * LDH + o] = uv;
                  if (TRAIN) p.act[(size_t)(2 + 2 * l) * p.act_stage_stride + (size_t)(row_g0 + row) * p.Hp_act + o] = uv;
                }
              }
            }
          }
          __syncwarp();
          gemm_rows<TC, RB>(tmp, LDH, Lb.Wt + (size_t)Lb.m_off * Lb.ldo, H, Lb.ldo, wbuf, p.wbuf_floats, p.KC, acc);
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const int row = c0 + warp * RB + r;
            if (row < R) {
              const float* bj = bl2 + rjet[row] * LDB;
#pragma unroll
              for (int i = 0; i < TC; ++i) {
                const int o = lane + 32 * i;
                if (o < H) {
                  float* hp = hs + (size_t)row * LDH + o;
                  *hp = lrelu(acc[r][i] + bj[o] + *hp, p.slope);
                  if (TRAIN) p.act[(size_t)(3 + 2 * l) * p.act_stage_stride + (size_t)(row_g0 + row) * p.Hp_act + o] = *hp;
                }
              }
            }
          }
          __syncwarp();
        }
        __syncthreads();
      }
      // ---------------- head fc_l3 + leaky_relu (epic.py:387-389) ----------------
      {
        const Lin L3 = lin[p.n_lin - 1];
        const float* w3 = L3.Wt + (size_t)L3.m_off * L3.ldo;
        for (int i = tid; i < R * F; i += kThreads) {
          const int row = i / F, f = i - row * F;
          float a = bias_of(p, L3, ev, jid[rjet[row]], f);
          const float* hr = hs + (size_t)row * LDH;
          a = gemv_col(w3 + f, L3.ldo, hr, H, a);
          vbuf[i] = lrelu(a, p.slope);
        }
        __syncthreads();
      }
      // ---------------- integrator (torchdyn fixed step, restated in oracle/ode_oracle.py) ----------------
      if (p.solver >= 0 && (TRAIN || p.step_kind == 0)) {
        const bool mid = p.solver == PFM_SOLVER_MIDPOINT;
        const int step = mid ? (ev >> 1) : ev;
        const float dt = p.dt[step];
        const bool first_stage = mid && ((ev & 1) == 0);
        const float hdt = __fmul_rn(0.5f, dt);
        for (int i = tid; i < R * F; i += kThreads) {
          const int row = i / F, f = i - row * F;
          const float k = -vbuf[i];
          if (first_stage) {
            xs[row * LDX + f] = __fadd_rn(x0[i], __fmul_rn(hdt, k));     // x + 0.5*dt*k1
          } else {
            const float xn = __fadd_rn(x0[i], __fmul_rn(dt, k));         // x + dt*f_(...)
            x0[i] = xn;
            xs[row * LDX + f] = xn;
          }
        }
        __syncthreads();
      } else if (!TRAIN && p.solver >= 0) {
        // ---------------- diffusion step programs (components/solver.py, flow_matching_module.py:62-69) ----------------
        const float4 cf = *reinterpret_cast<const float4*>(p.coef + (size_t)ev * 4);
        if (p.step_kind == PFM_STEP_PF_ODE) {
          // f(t, x) = -0.5 * beta * (x - v / noise_rate); reversed time: k = -f; x is the state the net was evaluated at
          const bool mid = p.solver == PFM_SOLVER_MIDPOINT;
          const float dt = p.dt[mid ? (ev >> 1) : ev];
          const bool first_stage = mid && ((ev & 1) == 0);
          const float hdt = __fmul_rn(0.5f, dt);
          const float mhb = __fmul_rn(-0.5f, cf.x);
          for (int i = tid; i < R * F; i += kThreads) {
            const int row = i / F, f = i - row * F;
            const float xin = xs[row * LDX + f];
            const float k = -__fmul_rn(mhb, __fadd_rn(xin, -__fdiv_rn(vbuf[i], cf.y)));
            if (first_stage) {
              xs[row * LDX + f] = __fadd_rn(x0[i], __fmul_rn(hdt, k));
            } else {
              const float xn = __fadd_rn(x0[i], __fmul_rn(dt, k));
              x0[i] = xn;
              xs[row * LDX + f] = xn;
            }
          }
        } else if (p.step_kind == PFM_STEP_DDIM) {
          // pred = (x - nr * v) / sr;  x <- next_sr * pred + next_nr * v;  the last step returns pred  (solver.py:16-19, :79-92)
          const bool last = ev == p.n_evals - 1;
          for (int i = tid; i < R * F; i += kThreads) {
            const int row = i / F, f = i - row * F;
            const float v = vbuf[i];
            const float pred = __fdiv_rn(__fadd_rn(x0[i], -__fmul_rn(cf.y, v)), cf.x);
            const float xn = last ? pred : __fadd_rn(__fmul_rn(cf.z, pred), __fmul_rn(cf.w, v));
            x0[i] = xn;
            xs[row * LDX + f] = xn;
          }
        } else {
          // Euler-Maruyama (solver.py:122-131): s = -v / nr; x += 0.5*beta*(x + 2 s)*delta_t; x += sqrt(beta*delta_t) * noise
          const float hb = __fmul_rn(0.5f, cf.x);
          const float* nz = p.noise + (size_t)ev * p.noise_step_stride;
          for (int i = tid; i < R * F; i += kThreads) {
            const int row = i / F, f = i - row * F;
            const int j = rjet[row];
            const int part = p.ridx[(size_t)jid[j] * p.N + (row - jrow0[j])];
            const float sc = __fdiv_rn(-vbuf[i], cf.y);
            float xn = x0[i];
            xn = __fadd_rn(xn, __fmul_rn(__fmul_rn(hb, __fadd_rn(xn, __fmul_rn(2.f, sc))), cf.z));
            xn = __fadd_rn(xn, __fmul_rn(cf.w, nz[((size_t)(p.jet0 + jid[j]) * p.N + part) * F + f]));
            x0[i] = xn;
            xs[row * LDX + f] = xn;
          }
        }
        __syncthreads();
      }
    }
    if (TRAIN && p.loss_kind < 0) {
      // plain forward with saved activations: park leaky_relu'(pre3) (sign(v) == sign(pre3)); the backward
      // entry point multiplies it with the incoming gradient
      for (int i = tid; i < R * F; i += kThreads) p.dpre3[(size_t)row_g0 * F + i] = vbuf[i] > 0.f ? 1.f : p.slope;
    }
    if (TRAIN && p.loss_kind >= 0) {
      // masked squared error sum((v - u)^2) (losses.py:75-76) and the gradient seed at the head pre-activation:
      // d loss / d pre3 = 2 (v - u) / sum(mask) * leaky_relu'(pre3)   (sign(v) == sign(pre3))
      const float inv_n = 1.f / (float)(*p.n_total);
      float part = 0.f;
      for (int i = tid; i < R * F; i += kThreads) {
        const float v = vbuf[i], d = v - x0[i];
        part += d * d;
        p.dpre3[(size_t)row_g0 * F + i] = 2.f * d * inv_n * (v > 0.f ? 1.f : p.slope);
      }
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sft);
      if (lane == 0) atomicAdd(p.loss_acc, part);
      continue;
    }
    // ---------------- write back: real particles from the resident buffers, padding = 0 ----------------
    {
      const float* src = p.solver >= 0 ? x0 : vbuf;
      for (int j = 0; j < nj; ++j) {
        const int n = jrow0[j + 1] - jrow0[j];
        const float fill = n == 0 ? __int_as_float(0x7fc00000) : 0.f;   // empty jet -> NaN like the reference
        float* dst = p.x_out + (size_t)jid[j] * p.N * F;
        for (int i = tid; i < p.N * F; i += kThreads) dst[i] = fill;
      }
      __syncthreads();
      for (int i = tid; i < R * F; i += kThreads) {
        const int row = i / F, f = i - row * F;
        const int j = rjet[row];
        const int part = p.ridx[(size_t)jid[j] * p.N + (row - jrow0[j])];
        p.x_out[((size_t)jid[j] * p.N + part) * F + f] = src[i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct SimtShape { int TC, RB, KC, R_cap, J_cap; bool spill; SimtParams p; size_t smem; };

static int simt_shape(const pfm_epic* h, int N, int Kx, SimtShape* s, int R_cap_force = 0, int J_cap_force = 0) {
  const pfm_epic_cfg& c = h->cfg;
  const int H = c.hid, Z = c.latent, F = c.feats;
  if (H > 320) { set_error("fp32 path supports hid <= 320 (got %d)", H); return PFM_ERR_UNSUPPORTED; }
  s->TC = H <= 128 ? 4 : (H <= 160 ? 5 : 10);
  s->RB = 8;
  const int Hp = (H + 3) & ~3;
  const int LDH = Hp + 4;
  const int Kxmax = c.input_dim > F ? c.input_dim : F;
  const int LDX = ((Kxmax + 3) & ~3) + 4;
  const int ldo_max = Hp;
  int J_cap = (H <= 160 && Z <= 32) ? 16 : 4;
  const int LDB = Hp, LDP = ((2 * H + Z + 3) & ~3);
  const int CR = kWarps * s->RB;
  const int budget = h->max_smem_optin / 4;   // floats
  int KC = 16;
  int R_cap = 0;
  for (int pass = 0; pass < 4; ++pass) {
    const int wbuf = KC * ldo_max + 32 * s->TC + 32;   // + slack for the unguarded column reads
    const int per_jet = LDP + Z + Hp + 2 * LDB;
    const int fixed = CR * LDH + 2 * wbuf + J_cap * per_jet + (2 * J_cap + 8) + 64;
    const int per_row = LDX + 2 * F + LDH + 1;          // xs, x0, vbuf, hs, rjet(short, rounded up)
    R_cap = (budget - fixed) / per_row;
    s->KC = KC;
    if (pass == 0) {            // size the per-jet arrays for ~3x the jets that fit at full multiplicity
      int want = R_cap > 0 ? (3 * R_cap + N - 1) / N : 1;
      int jc = want < 2 ? 2 : (want > J_cap ? J_cap : want);
      if (jc != J_cap) { J_cap = jc; continue; }
    }
    if (R_cap >= N || KC == 8) break;
    KC = 8;
  }
  bool spill = false;
  if (R_cap < N) {
    // spill mode: hs moves to global memory; the remaining per-row state is tiny
    spill = true;
    const int wbuf = s->KC * ldo_max + 32 * s->TC + 32;
    const int per_jet = LDP + Z + Hp + 2 * LDB;
    const int fixed = CR * LDH + 2 * wbuf + J_cap * per_jet + (J_cap + 8) + 64;
    const int per_row = LDX + 2 * F + 1;
    R_cap = (budget - fixed) / per_row;
    const int want = N > 512 ? N : 512;
    if (R_cap > want) R_cap = want;
    if (R_cap < N) {
      set_error("fp32 path: a jet of %d particles does not fit even in spill mode (%d rows at hid=%d)", N, R_cap, H);
      return PFM_ERR_UNSUPPORTED;
    }
  }
  s->spill = spill;
  if (R_cap > 1024) R_cap = 1024;
  if (R_cap_force > 0) {          // training: forward and backward kernels share one plan
    if (R_cap_force > R_cap || J_cap_force > J_cap) { set_error("internal: forced group capacity exceeds the forward plan"); return PFM_ERR_INVALID; }
    R_cap = R_cap_force; J_cap = J_cap_force;
  }
  s->R_cap = R_cap; s->J_cap = J_cap;
  SimtParams& p = s->p;
  memset(&p, 0, sizeof(p));
  p.F = F; p.H = H; p.Hp = Hp; p.LDH = LDH; p.Z = Z; p.L = c.layers; p.n_lin = h->n_lin;
  p.R_cap = R_cap; p.J_cap = J_cap; p.KC = s->KC; p.LDX = LDX; p.LDB = LDB; p.LDP = LDP;
  p.sum_scale = c.sum_scale; p.slope = c.neg_slope;
  const int wbuf = s->KC * ldo_max + 32 * s->TC + 32;
  p.wbuf_floats = (wbuf + 3) & ~3;
  int o = 0;
  auto take = [&](int n) { int r = o; o += (n + 3) & ~3; return r; };
  p.o_xs = take(R_cap * LDX);
  p.o_x0 = take(R_cap * F);
  p.o_hs = spill ? 0 : take(R_cap * LDH);
  p.o_tmp = take(CR * LDH);
  p.o_wbuf = take(2 * p.wbuf_floats);
  p.o_pool = take(J_cap * LDP);
  p.o_g = take(J_cap * Z);
  p.o_g1 = take(J_cap * Hp);
  p.o_bl1 = take(J_cap * LDB);
  p.o_bl2 = take(J_cap * LDB);
  p.o_v = take(R_cap * F);
  p.o_int = take(2 * J_cap + 8 + (R_cap + 1) / 2);
  p.total_floats = o;
  s->smem = (size_t)o * 4;
  if ((int)s->smem > h->max_smem_optin) {
    set_error("fp32 path: shared-memory plan %zu B exceeds the device limit %d B", s->smem, h->max_smem_optin);
    return PFM_ERR_UNSUPPORTED;
  }
  (void)Kx;
  return PFM_OK;
}

int simt_plan_caps(const pfm_epic* h, int N, int* R_cap, int* J_cap) {
  SimtShape s;
  int rc = simt_shape(h, N, 0, &s);
  if (rc != PFM_OK) return rc;
  *R_cap = s.R_cap; *J_cap = s.J_cap;
  return PFM_OK;
}

static int ensure_spill(pfm_epic* h, SimtShape& s, int grid) {
  s.p.hs_spill = nullptr;
  if (!s.spill) return PFM_OK;
  const size_t need = (size_t)grid * s.R_cap * s.p.LDH;
  if (need > h->hs_spill_cap) {
    if (h->hs_spill) cudaFree(h->hs_spill);
    h->hs_spill = nullptr; h->hs_spill_cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(&h->hs_spill, sizeof(float) * need));
    h->hs_spill_cap = need;
  }
  s.p.hs_spill = h->hs_spill;
  return PFM_OK;
}

template <int TC, int RB, bool TRAIN, bool SPILL>
static int launch_simt_s(const SimtShape& s, int grid, cudaStream_t st) {
  auto kern = epic_simt_kernel<TC, RB, TRAIN, SPILL>;
  PFM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.smem));
  kern<<<grid, kThreads, s.smem, st>>>(s.p);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}
template <int TC, int RB, bool TRAIN>
static int launch_simt(const pfm_epic* h, const SimtShape& s, int grid, cudaStream_t st) {
  (void)h;
  return s.p.hs_spill ? launch_simt_s<TC, RB, TRAIN, true>(s, grid, st) : launch_simt_s<TC, RB, TRAIN, false>(s, grid, st);
}

int simt_run(pfm_epic* h, const RunArgs& a, cudaStream_t st) {
  SimtShape s;
  int rc = simt_shape(h, a.N, a.Kx, &s);
  if (rc != PFM_OK) return rc;
  SimtParams& p = s.p;
  p.Kx = a.Kx; p.Kxp = (a.Kx + 3) & ~3; p.x_ld = a.Kx; p.xin_off = a.xin_off;
  p.lin = h->lin_dev;
  p.tbias = h->tbias; p.cbias = a.has_cbias ? h->cbias : nullptr; p.bstride = h->bstride;
  p.tbias_per_jet = a.tbias_per_jet;
  p.n_real = h->plan.n_real; p.ridx = h->plan.ridx; p.groups = h->plan.groups; p.n_groups = h->plan.n_groups;
  p.counter = h->plan.counter; p.jetmap = a.jetmap;
  p.x_in = a.x_in; p.x_out = a.x_out; p.B = a.B; p.N = a.N;
  p.n_evals = a.n_evals; p.solver = a.solver; p.n_steps = a.n_steps; p.dt = a.dt;
  p.step_kind = a.step_kind; p.coef = a.coef; p.noise = a.noise; p.noise_step_stride = a.noise_step_stride; p.jet0 = a.jet0;
  // persistent CTAs: one per SM, but never more than there can be groups (every group has >= 1 jet)
  int grid = h->sm_count < a.B ? h->sm_count : a.B;
  rc = ensure_spill(h, s, grid);
  if (rc != PFM_OK) return rc;
  if (s.TC == 4) return launch_simt<4, 8, false>(h, s, grid, st);
  if (s.TC == 5) return launch_simt<5, 8, false>(h, s, grid, st);
  return launch_simt<10, 8, false>(h, s, grid, st);
}

// training forward: same kernel with the interpolation prologue (loss_kind >= 0), the activation saves and the
// loss / gradient-seed epilogue
int simt_train_forward(pfm_epic* h, const TrainFwdArgs& a, cudaStream_t st) {
  SimtShape s;
  int rc = simt_shape(h, a.N, a.Kx, &s, a.lay.R_cap, a.lay.J_cap);
  if (rc != PFM_OK) return rc;
  SimtParams& p = s.p;
  p.Kx = a.Kx; p.Kxp = (p.Kx + 3) & ~3; p.x_ld = p.Kx; p.xin_off = a.xin_off;
  p.lin = h->lin_dev;
  p.tbias = h->tbias; p.cbias = a.has_cbias ? h->cbias : nullptr; p.bstride = h->bstride; p.tbias_per_jet = a.tbias_per_jet;
  p.n_real = h->plan.n_real; p.ridx = h->plan.ridx; p.groups = h->plan.groups; p.n_groups = h->plan.n_groups;
  p.counter = h->plan.counter;
  p.x_in = a.x_in; p.x_out = a.x_out; p.B = a.B; p.N = a.N;
  p.n_evals = 1; p.solver = -1; p.n_steps = 0; p.dt = nullptr;
  p.step_kind = 0; p.coef = nullptr; p.noise = nullptr; p.noise_step_stride = 0; p.jet0 = 0;
  p.x1 = a.x_in; p.tjet = a.t; p.noise0 = a.noise0; p.noise1 = a.noise1; p.loss_kind = a.loss_kind; p.sigma = a.sigma;
  p.act = h->act; p.act_stage_stride = a.lay.stage_stride; p.Hp_act = a.lay.Hp;
  p.yact = h->yact;
  p.jact = h->jact; p.junit = a.lay.junit; p.jstride = a.lay.jstride; p.LDP_act = a.lay.LDP;
  p.dpre3 = h->dpre3; p.loss_acc = h->loss_acc; p.rowoff = h->plan.rowoff; p.n_total = h->plan.n_total;
  int grid = h->sm_count < a.B ? h->sm_count : a.B;
  rc = ensure_spill(h, s, grid);
  if (rc != PFM_OK) return rc;
  if (s.TC == 4) return launch_simt<4, 8, true>(h, s, grid, st);
  if (s.TC == 5) return launch_simt<5, 8, true>(h, s, grid, st);
  return launch_simt<10, 8, true>(h, s, grid, st);
}

// group capacity of the forward kernel (the training plan takes the minimum with the backward kernel's)
int simt_caps_for_train(const pfm_epic* h, int N, int* R_cap, int* J_cap, int* TC, int* RB, int* KC) {
  SimtShape s;
  int rc = simt_shape(h, N, 0, &s);
  if (rc != PFM_OK) return rc;
  *R_cap = s.R_cap; *J_cap = s.J_cap; *TC = s.TC; *RB = s.RB; *KC = s.KC;
  return PFM_OK;
}

}  // namespace pfm
