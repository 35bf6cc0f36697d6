// bf16 tensor-core path (tcgen05 / TMEM / bulk async copies) of the EPiC vector field + integrator.
//
// One persistent CTA owns a group of jets whose real particles fill up to two 128-row tiles and runs
// the WHOLE integration for them (one launch per sample()).  Per-particle state never leaves the SM:
//   TMEM   cols [0,256)    fp32 hidden features h of tile A / tile B  = residual stream AND accumulator
//                          of fc_local2 / fc_l2 (the MMA accumulates straight onto the residual)
//          cols [256,512)  fp32 accumulators of fc_local1 per tile; the epilogue overwrites them in place
//                          with the bf16 activations u that feed fc_local2 as a TMEM A-operand.
//                          The pooling / global-MLP accumulators (48 columns) alias tile B's region.
//   SMEM   h tiles as bf16 (K-major SWIZZLE_128B = A operand of fc_local1, and read "transposed" as an
//          MN-major A operand by the pooling MMA), a 3-slot ring of 32 KB pre-swizzled weight images fed
//          by cp.async.bulk, the jet-indicator matrix P, per-jet biases and global vectors.
//   REGS   every epilogue thread owns one particle: its ODE state x lives in registers for all steps.
// Masked mean/sum pooling is an MMA:  S[c][jet] = sum_rows h[row][c] * P[jet][row]  (M=128, N=16),
// and the 256->128 part of fc_global1 is two more N=16 MMAs (W_mean . S, W_sum . S), so the serial
// per-jet path costs ~5% extra tensor time instead of a CUDA-core GEMV.
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one lane), warps 4-7 / 8-11 = epilogue
// warpgroups of tile A / tile B (TMEM lane quadrant = warp % 4).
#include <cstdlib>

#include "pfm_internal.cuh"
#include "tc_ptx.cuh"

namespace pfm {

using namespace tc;

static constexpr int TCH = 128;          // hidden width this path is specialised for
static constexpr int TC_ROWS = 256;      // 2 tiles of 128 particles
static constexpr int TC_J = 16;          // jets per group (N of the pooling MMA)
static constexpr int TC_ZMAX = 16;        // latent width (10 / 16 in the configs); the small-weight pack is laid out for 16
static constexpr int TC_KXMAX = 8;          // per-particle input columns / features (3 JetNet, 8 JetClass)
// per-unit small-weight pack, one bulk copy:  W_gg | W_glob (bf16, [z][o])  and  W_g2 (fp32, [z][k], rows padded to
// 132 floats so that 8 lanes reading 8 different rows with LDS.128 hit 8 different bank groups)
struct SpkPack {
  __nv_bfloat16 gg[16][128];     // W_gg[z][o]   = fc_global1[o][2H + z]        (layers)
  __nv_bfloat16 gl[16][128];     // W_glob[z][o] = fc_local1[o][H + z]          (layers)
  float g2[16][128 + 4];         // W_g2[z][k]   = fc_global2[z][k]  (fc_g2 for the stem)
};
static constexpr uint32_t TC_SPK = sizeof(SpkPack);
static_assert(sizeof(SpkPack) % 16 == 0, "bulk copies move multiples of 16 bytes");
static constexpr int TC_SBIAS = 384 + 16;               // one unit's slice of the time-bias table
static constexpr int TC_THREADS = 384;
static constexpr int TC_NSLOT = 3;
static constexpr uint32_t TC_MAT = 32768;   // one 128x128 bf16 weight image

template <int FP>
struct TcSmem {
  alignas(1024) uint8_t h[2][TC_MAT];          // bf16 h tiles
  alignas(1024) uint8_t w[TC_NSLOT][TC_MAT];   // weight ring
  alignas(1024) uint8_t P[8192];               // [16 jets x 256 rows] bf16, K-major SW128, 4 blocks of 2 KB
  union alignas(1024) {
    uint8_t St[8192];                          // [16 jets x 256 k] bf16, K-major SW128, 4 blocks of 2 KB: k < 128 mean, k >= 128 sum.
    float g1[TC_J][TCH];                       // B operand of the fc_global1 MMAs; dead once they complete -> reused for g1
  } sg;
  float bl1[TC_J][TCH];
  float bl2[TC_J][TCH];
  alignas(16) float gv[TC_J][TC_ZMAX];
  alignas(16) float w1s[TCH][FP];              // fc_l1 weights of the particle features, [column][feature]
  alignas(16) float w3s[TCH][FP];              // fc_l3 (k-major)
  alignas(16) SpkPack spk;                     // small weights of the current unit
  alignas(16) float sbias[TC_SBIAS];           // time-bias slice of the current unit (4 consecutive linears)
  float inv_n[TC_J];
  int boff[128];                               // bias-table offset of every linear (copied once: no global descriptor loads in the loop)
  int jrow0[TC_J + 1];
  int jid[TC_J];                               // batch index of the group's jets (the plan bin-packs jets out of order)
  int group;
  uint32_t tmem_base;
  uint64_t full[TC_NSLOT], empty[TC_NSLOT];
  uint64_t hready[2], accU_full[2], u_ready[2][2], accH_full[2];      // u_ready[tile][half of the 128 u columns]
  uint64_t pool_full, glob_go, glob_full, d_free, wg1_ready;
  uint64_t spk_full, spk_empty;
};

static_assert(sizeof(TcSmem<8>) + 1024 <= 232448, "TcSmem exceeds the 227 KB per-block shared-memory limit of sm_100");
template <int N> struct PrintSize;
struct TcParams {
  int F, Kx, x_ld, xin_off, Z, L, n_lin, n_items;
  float sum_scale, slope;
  const Lin* lin;
  const uint8_t* wimg;
  const uint8_t* spk;        // [L+1] small-weight packs
  int boff_stem, boff_layer0, boff_layer_stride, bias_chunk_floats;   // where a unit's 4 linears sit in a bias-table row
  const float* tbias; const float* cbias; int bstride; int tbias_per_jet;
  const int* n_real; const uint16_t* ridx; const int2* groups; const int* n_groups; int* counter; const int* jetmap;
  const float* x_in; float* x_out; int B, N;
  int n_evals, solver, n_steps; const float* dt;
  long long* prof;          // [3][20] debug phase timers (PFM_TC_PROF)
};

__device__ __forceinline__ float lrelu_tc(float v, float s) { return fmaxf(v, v * s); }   // 0 < s < 1

__device__ __forceinline__ float tc_bias_of(const TcParams& p, const Lin& L, int eval, int jet_global, int o) {
  const int trow = p.tbias_per_jet ? jet_global : eval;
  float b = p.tbias[(size_t)trow * p.bstride + L.bias_off + o];
  if (p.cbias) b += p.cbias[(size_t)jet_global * p.bstride + L.bias_off + o];
  return b;
}

// packed fp32x2 math (FADD2 / FMUL2 on sm_100): leaky_relu(v + b) for two neighbouring columns
__device__ __forceinline__ void bias_lrelu2(uint32_t& v0, uint32_t& v1, float b0, float b1, unsigned long long slope2) {
  unsigned long long v = ((unsigned long long)v1 << 32) | v0;
  const unsigned long long b = ((unsigned long long)__float_as_uint(b1) << 32) | __float_as_uint(b0);
  unsigned long long a, t;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(v), "l"(b));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(a), "l"(slope2));
  v0 = __float_as_uint(fmaxf(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)t)));
  v1 = __float_as_uint(fmaxf(__uint_as_float((uint32_t)(a >> 32)), __uint_as_float((uint32_t)(t >> 32))));
}

// bf16x2( leaky_relu( bf16(v + b) ) ) for two neighbouring columns: packed add in fp32, one cvt, packed mul + max in bf16
__device__ __forceinline__ uint32_t bias_lrelu_bf16x2(uint32_t v0, uint32_t v1, float b0, float b1, uint32_t slope_bf2) {
  const unsigned long long v = ((unsigned long long)v1 << 32) | v0;
  const unsigned long long b = ((unsigned long long)__float_as_uint(b1) << 32) | __float_as_uint(b0);
  unsigned long long a;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(a) : "l"(v), "l"(b));
  const uint32_t x = pack_bf16x2(__uint_as_float((uint32_t)a), __uint_as_float((uint32_t)(a >> 32)));
  uint32_t t, r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(t) : "r"(x), "r"(slope_bf2));
  asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(t));
  return r;
}

__device__ __forceinline__ void ebar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

// descriptor advance (units of 16 bytes) of K-step k inside a 128-wide K-major SW128 image: 64-column blocks are
// 16 KB apart, a K=16 step is 32 bytes inside the swizzled 128-byte row
__device__ __forceinline__ constexpr uint32_t kstep16(int k) { return (uint32_t)(k >> 2) * 1024u + (uint32_t)(k & 3) * 2u; }

// 8 MMAs: D[128 x 128] (+)= A[128 x 128] . B[128 x 128]^T, descriptors of the two K-major SW128 images precomputed
// (called by the whole, converged MMA warp; one elected lane issues -- operands stay in uniform registers)
__device__ __forceinline__ void issue_ss_128(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc_first) {
  if (elect_one()) {
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ss(d, a_desc + kstep16(k), b_desc + kstep16(k), idesc, (acc_first || k > 0) ? 1u : 0u);
  }
  __syncwarp();
}
__device__ __forceinline__ void commit_to(uint64_t* bar) {
  if (elect_one()) mma_commit(bar);
  __syncwarp();
}

// Phase profiler (debug, PFM_TC_PROF=1): block 0 accumulates clock64() deltas per phase for one thread of each role.
// It also records an event trace (slot, clock) of evaluation TRACE_EV of the block's first group.
#define PROF_T(slot) do { if (PROF && prof_on) { const long long _n = clock64(); prof[slot] += _n - prof_t; prof_t = _n; \
    if (trace_on && ev == TRACE_EV && trace_n < TRACE_MAX) { p.prof[60 + (prof_role * TRACE_MAX + trace_n) * 2] = slot; \
      p.prof[61 + (prof_role * TRACE_MAX + trace_n) * 2] = _n; ++trace_n; } } } while (0)
static constexpr int TRACE_EV = 6, TRACE_MAX = 160;

// SIMPLE: one time code for the whole batch and no conditioning (the sampling configuration of the headline
// workload): every bias comes from the staged slice of the time-bias table, which removes all data-dependent
// branches from the serial per-jet phases (a uniform branch costs ~30-40 cycles of fetch bubble, and those phases
// are latency-bound).
template <int FP, bool PROF, bool SIMPLE>
__global__ void __launch_bounds__(TC_THREADS, 1) epic_tc_kernel(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte aligned view of the dynamic shared memory, derived by pointer arithmetic on the __shared__ array so
  // that the compiler keeps the shared address space (LDS/STS instead of generic LD/ST)
  TcSmem<FP>& s = *reinterpret_cast<TcSmem<FP>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Lin* lin = p.lin;
  const int L = p.L, Z = p.Z, F = p.F;
  long long prof[20];
  long long prof_t = 0;
  const bool prof_on = PROF && blockIdx.x == 0 && (tid == 32 || tid == 128 || tid == 256);   // lane 0 of the MMA warp, of epilogue A, of epilogue B
  const int prof_role = tid == 32 ? 0 : (tid == 128 ? 1 : 2);
  int trace_n = 0;
  bool trace_on = PROF;
  if (PROF) {
#pragma unroll
    for (int i = 0; i < 20; ++i) prof[i] = 0;
    prof_t = clock64();
  }

  if (tid == 0) {
    for (int i = 0; i < TC_NSLOT; ++i) { mbar_init(&s.full[i], 1); mbar_init(&s.empty[i], 1); }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s.hready[t], 128); mbar_init(&s.accU_full[t], 1); mbar_init(&s.u_ready[t][0], 128); mbar_init(&s.u_ready[t][1], 128);
      mbar_init(&s.accH_full[t], 1);
    }
    mbar_init(&s.pool_full, 1); mbar_init(&s.glob_go, 256); mbar_init(&s.glob_full, 1); mbar_init(&s.d_free, 256);
    mbar_init(&s.wg1_ready, 256);
    mbar_init(&s.spk_full, 1); mbar_init(&s.spk_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s.tmem_base, 512);
  {   // small fp32 weights used by every particle thread: fc_l1 feature rows, fc_l3
    const Lin L1 = lin[LIN_L1], L3 = lin[p.n_lin - 1];
    for (int i = tid; i < p.Kx * TCH; i += TC_THREADS) {
      const int k = i / TCH, o = i - k * TCH;
      s.w1s[o][k] = L1.Wt[(size_t)(L1.m_off + p.xin_off + k) * L1.ldo + o];
    }
    for (int i = tid; i < TCH * FP; i += TC_THREADS)
      if ((i % FP) >= p.Kx) (&s.w1s[0][0])[i] = 0.f;
    for (int i = tid; i < TCH * FP; i += TC_THREADS) {
      const int c = i / FP, f = i - c * FP;
      s.w3s[c][f] = f < L3.ldo ? L3.Wt[(size_t)(L3.m_off + c) * L3.ldo + f] : 0.f;
    }
    for (int i = tid; i < p.n_lin && i < 128; i += TC_THREADS) s.boff[i] = lin[i].bias_off;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  const int n_groups = *p.n_groups;

  // running use counters of every barrier (parity = count & 1); each role only advances the ones it uses
  uint32_t ring_it = 0;                                     // producer / MMA: weight items consumed so far
  uint32_t c_hready[2] = {0, 0}, c_uready[2] = {0, 0}, c_globgo = 0, c_dfree = 0, c_wg1 = 0;   // MMA side
  uint32_t ring_e = 0;                                      // epilogue: mirror of the weight-ring position (tile A stages W_g1 into TMEM)
  uint32_t c_accH = 0, c_accU = 0, c_pool = 0, c_glob = 0, c_spk = 0;                         // epilogue side
  uint32_t spk_it = 0;                                      // producer: small-weight packs issued so far

  for (;;) {
    if (PROF && trace_n > 0) trace_on = false;     // trace the block's first group only
    __syncthreads();
    if (tid == 0) s.group = atomicAdd(p.counter, 1);
    __syncthreads();
    const int gidx = s.group;
    if (gidx >= n_groups) break;
    const int2 grp = p.groups[gidx];
    const int j0 = grp.x, nj = grp.y;

    if (warp == 0) {
      // ================================ weight producer (whole warp, elected lane issues) ================================
      {
        const bool stage_bias = !p.tbias_per_jet;
        const uint32_t bias_bytes = stage_bias ? (uint32_t)p.bias_chunk_floats * 4u : 0u;
        for (int ev = 0; ev < p.n_evals; ++ev) {
          for (int it = 0; it < p.n_items; ++it, ++ring_it) {
            // the small-weight pack + bias slice of unit u travel just before the unit's global-MLP images
            const int u = it == 1 ? 0 : ((it >= 3 && ((it - 3) & 3) == 0) ? 1 + ((it - 3) >> 2) : -1);
            if (u >= 0) {
              mbar_wait(&s.spk_empty, (spk_it & 1) ^ 1);
              ++spk_it;
              if (elect_one()) {
                mbar_arrive_expect_tx(&s.spk_full, TC_SPK + bias_bytes);
                bulk_copy_g2s(&s.spk, p.spk + (size_t)u * TC_SPK, TC_SPK, &s.spk_full);
                if (stage_bias)
                  bulk_copy_g2s(s.sbias, p.tbias + (size_t)ev * p.bstride + (u == 0 ? p.boff_stem : p.boff_layer0 + (u - 1) * p.boff_layer_stride),
                                bias_bytes, &s.spk_full);
              }
              __syncwarp();
            }
            const uint32_t slot = ring_it % TC_NSLOT, round = ring_it / TC_NSLOT;
            mbar_wait(&s.empty[slot], (round & 1) ^ 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(&s.full[slot], TC_MAT);
              bulk_copy_g2s(s.w[slot], p.wimg + (size_t)it * TC_MAT, TC_MAT, &s.full[slot]);
            }
            __syncwarp();
          }
        }
      }
    } else if (warp == 1) {
      // ================================ MMA issuer (whole warp, elected lane issues) ================================
      {
        const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
        const uint32_t idesc_pool = make_idesc_bf16(128, 16, 1, 0);
        const uint32_t idesc_glob = make_idesc_bf16(128, 16, 0, 0);
        const uint64_t hA = desc_kmajor(smem_u32(s.h[0])), hB = desc_kmajor(smem_u32(s.h[1]));
        const uint64_t hAt = desc_mnmajor(smem_u32(s.h[0]), 16384u, 1024u), hBt = desc_mnmajor(smem_u32(s.h[1]), 16384u, 1024u);
        const uint64_t pdesc = desc_kmajor(smem_u32(s.P)), sdesc = desc_kmajor(smem_u32(s.sg.St));
        const uint64_t wdesc0 = desc_kmajor(smem_u32(s.w[0]));
        const uint32_t accH0 = tm, accH1 = tm + 128, accU0 = tm + 256, accU1 = tm + 384;
        // pooling / global accumulators: 4 independent 16-column accumulators each (a dependent chain of N=16 MMAs on
        // ONE accumulator is latency-bound, ~70 cycles per MMA); the readers add the 4 partial results
        const uint32_t dpool = tm + 384, dglob = tm + 384 + 64;
        auto wait_full = [&](uint32_t it) { mbar_wait(&s.full[it % TC_NSLOT], (it / TC_NSLOT) & 1); };
        auto wslot = [&](uint32_t it) { return wdesc0 + (uint64_t)((it % TC_NSLOT) * (TC_MAT >> 4)); };
        for (int ev = 0; ev < p.n_evals; ++ev) {
          // ---- stem fc_l2: accH[t] (holds h1) += h1 . W_l2^T
          const uint32_t it_l2 = ring_it++;
          for (int t = 0; t < 2; ++t) {
            PROF_T(0);
            mbar_wait(&s.hready[t], c_hready[t]++ & 1);
            tc_fence_after();
            PROF_T(1);
            if (t == 0) wait_full(it_l2);
            PROF_T(2);
            issue_ss_128(t ? accH1 : accH0, t ? hB : hA, wslot(it_l2), idesc, true);
            commit_to(&s.accH_full[t]);
          }
          commit_to(&s.empty[it_l2 % TC_NSLOT]);
          for (int gi = 0; gi <= L; ++gi) {
            uint32_t it_w1 = 0;
            if (gi != 1) {   // a new version of h is complete: pool it   S[c][jet] = sum_rows h[row][c] P[jet][row]
#pragma unroll
              for (int t = 0; t < 2; ++t) {          // tile A's half is issued as soon as tile A is ready
                PROF_T(0);
                mbar_wait(&s.hready[t], c_hready[t]++ & 1);
                tc_fence_after();
                PROF_T(3);
                if (elect_one()) {
#pragma unroll
                  for (int k = 0; k < 8; ++k) {      // 16 rows of h per step: +2 KB in the MN-major view; P: 2 KB per 64 rows
                    const uint64_t da = (t ? hBt : hAt) + (uint64_t)(k * 128);
                    const uint64_t db = pdesc + (uint64_t)((t * 2 + (k >> 2)) * 128 + (k & 3) * 2);
                    mma_ss(dpool + (uint32_t)((k & 3) * 16), da, db, idesc_pool, (t | (k >> 2)) ? 1u : 0u);
                  }
                }
                __syncwarp();
              }
              commit_to(&s.pool_full);
            }
            // ring order: fc_global1 mean | sum | fc_local1 | fc_local2
            // ---- 256 -> 128 part of fc_g1 / fc_global1:  D[o][jet] = W_mean[o][:] . mean[jet][:] + W_sum[o][:] . sum[jet][:]
            // (issued before fc_local1: it is on the serial per-jet chain, fc_local1's result is not needed before the chain ends)
            const uint32_t it_gm = ring_it++, it_gs = ring_it++;
            if (gi >= 1) it_w1 = ring_it++;
            // The 128 x 256 weight block is the A operand: read from shared memory it costs 64 cycles per K = 16 step
            // whatever N is (operand fetch at 64 B/clk), and these 16 steps sit on the serial per-jet chain.  Tile A's
            // epilogue warps therefore copy the two images into TMEM (the fc_local1 accumulator of tile A is free until
            // this unit's fc_local1) while the pooling MMAs run, and the MMAs take A from TMEM: ~8 cycles per step.
            PROF_T(0);
            mbar_wait(&s.wg1_ready, c_wg1++ & 1);
            tc_fence_after();
            if (elect_one()) { mbar_arrive(&s.empty[it_gm % TC_NSLOT]); mbar_arrive(&s.empty[it_gs % TC_NSLOT]); }   // images consumed
            __syncwarp();
            PROF_T(6);
            mbar_wait(&s.glob_go, c_globgo++ & 1);
            tc_fence_after();
            PROF_T(5);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 16; ++kk)
                mma_ts(dglob + (uint32_t)((kk & 3) * 16), accU0 + (uint32_t)kk * 8u, sdesc + (uint64_t)((kk >> 2) * 128 + (kk & 3) * 2), idesc_glob,
                       (kk >> 2) ? 1u : 0u);
            }
            __syncwarp();
            commit_to(&s.glob_full);
            if (gi >= 1) {   // fc_local1 of tile A
              PROF_T(0);
              wait_full(it_w1);
              PROF_T(4);
              issue_ss_128(accU0, hA, wslot(it_w1), idesc, false);
              commit_to(&s.accU_full[0]);
            }
            PROF_T(0);
            mbar_wait(&s.d_free, c_dfree++ & 1);      // pooling / global accumulators (aliasing accU of tile B) consumed
            tc_fence_after();
            PROF_T(7);
            if (gi >= 1) {
              issue_ss_128(accU1, hB, wslot(it_w1), idesc, false);
              commit_to(&s.accU_full[1]);
              commit_to(&s.empty[it_w1 % TC_NSLOT]);
              const uint32_t it_w2 = ring_it++;
              // fc_local2: accH[t] += u[t] (bf16 in TMEM) . W2^T.  The fc_local1 epilogue hands over u in two halves of 64
              // columns, so the first four K steps of both tiles run while the second half is still being written.
#pragma unroll
              for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                  PROF_T(0);
                  mbar_wait(&s.u_ready[t][half], c_uready[t] & 1);
                  tc_fence_after();
                  PROF_T(8 + t);
                  if (t == 0 && half == 0) wait_full(it_w2);
                  PROF_T(10);
                  const uint64_t wd = wslot(it_w2);
                  if (elect_one()) {
#pragma unroll
                    for (int k = half * 4; k < half * 4 + 4; ++k)
                      mma_ts(t ? accH1 : accH0, (t ? accU1 : accU0) + (uint32_t)k * 8u, wd + kstep16(k), idesc, 1u);
                  }
                  __syncwarp();
                  if (half == 1) { commit_to(&s.accH_full[t]); ++c_uready[t]; }
                }
              }
              commit_to(&s.empty[it_w2 % TC_NSLOT]);
            }
          }
        }
      }
    } else if (warp >= 4) {
      // ================================ epilogue / particle threads ================================
      const int et = tid - 128;
      const int wg = et >> 7;                    // tile
      const int r = et & 127;                    // row in tile = TMEM lane
      const int row = wg * 128 + r;
      const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
      const uint32_t accH = lane_base + wg * 128, accU = lane_base + 256 + wg * 128;
      const uint32_t dpool = lane_base + 384, dglob = lane_base + 384 + 64;
      uint8_t* hrow = s.h[wg] + (r >> 3) * 1024 + (r & 7) * 128;       // this particle's 128-byte swizzled row (per 64-col block)

      if (et == 0) {
        int acc = 0;
        for (int j = 0; j < TC_J; ++j) s.jid[j] = p.jetmap ? p.jetmap[j0 + (j < nj ? j : 0)] : j0 + (j < nj ? j : 0);
        for (int j = 0; j < nj; ++j) {
          s.jrow0[j] = acc;
          const int n = p.n_real[s.jid[j]];
          s.inv_n[j] = 1.f / (float)n;           // n == 0 -> inf -> NaN confined to that jet, like the reference
          acc += n;
        }
        for (int j = nj; j <= TC_J; ++j) s.jrow0[j] = acc;
        for (int j = nj; j < TC_J; ++j) s.inv_n[j] = 0.f;
      }
      // P = 0, h = 0: rows without a particle are never written afterwards, so they stay finite (0) in every
      // operand the pooling MMA reads; their TMEM lanes hold finite junk that no real row ever sees
      for (int i = et; i < 8192 / 16; i += 256) reinterpret_cast<uint4*>(s.P)[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < 2 * (int)TC_MAT / 16; i += 256) reinterpret_cast<uint4*>(&s.h[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < TC_J * TC_ZMAX; i += 256) (&s.gv[0][0])[i] = 0.f;
      ebar();
      const int R = s.jrow0[nj];
      const bool valid = row < R;
      int myjet = 0;
      for (int j = 1; j < nj; ++j) myjet += (row >= s.jrow0[j]) ? 1 : 0;
      if (!valid) myjet = 0;
      if (valid) *reinterpret_cast<__nv_bfloat16*>(s.P + sw128_offset(myjet, row, 2048)) = __float2bfloat16(1.0f);
      const int jg = s.jid[myjet];
      // bias of a linear of the CURRENT unit: staged slice of the time table (+ per-jet cond table), or the
      // slow direct path when every jet has its own time (training-style forward)
      auto unit_bias = [&](int lin_idx, int voff, int jet_global, int o) -> float {
        if (SIMPLE) return s.sbias[voff + o];
        float b = p.tbias_per_jet ? p.tbias[(size_t)jet_global * p.bstride + s.boff[lin_idx] + o] : s.sbias[voff + o];
        if (p.cbias) b += p.cbias[(size_t)jet_global * p.bstride + s.boff[lin_idx] + o];
        return b;
      };
      float x0[FP], xc[FP], vout[FP];
#pragma unroll
      for (int f = 0; f < FP; ++f) { x0[f] = 0.f; xc[f] = 0.f; vout[f] = 0.f; }
      int part = 0;
      if (valid) {
        part = p.ridx[(size_t)jg * p.N + (row - s.jrow0[myjet])];
        const float* src = p.x_in + ((size_t)jg * p.N + part) * p.x_ld;
#pragma unroll
        for (int f = 0; f < FP; ++f)
          if (f < p.Kx) { xc[f] = src[f]; x0[f] = xc[f]; }
      }
      const uint32_t hrow_addr = smem_u32(hrow), rx16 = (uint32_t)(r & 7) << 4, vpred = valid ? 1u : 0u;
      const uint32_t bl1_addr = smem_u32(&s.bl1[myjet][0]), bl2_addr = smem_u32(&s.bl2[myjet][0]);
      const uint32_t slope_bf2 = pack_bf16x2(p.slope, p.slope);
      const unsigned long long slope2 = ((unsigned long long)__float_as_uint(p.slope) << 32) | __float_as_uint(p.slope);
      const int ZP = (Z + 3) & ~3;

      const uint32_t w3_addr = smem_u32(&s.w3s[0][0]), st_addr = smem_u32(s.sg.St);

      // store 32 fp32 columns [32c, 32c+32) of this particle's row as bf16 into the swizzled h tile
      auto store_h_bf16 = [&](const uint32_t (&v)[32], int c, uint32_t pred) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[q * 8 + 0]), __uint_as_float(v[q * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7]));
          const int c16 = c * 4 + q;
          sts128_if(hrow_addr + (uint32_t)((c16 >> 3) * 16384) + ((uint32_t)((c16 & 7) << 4) ^ rx16), pk.x, pk.y, pk.z, pk.w, pred);
        }
      };
      // one 32-column chunk of the residual update: h = lrelu(acc + bias) -> TMEM fp32 (in place) + shared bf16
      auto epi_h_chunk = [&](uint32_t (&v)[32], int c, uint32_t spred) {
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = lds128(bl2_addr + (uint32_t)(c * 128 + i4 * 16));
          bias_lrelu2(v[i4 * 4 + 0], v[i4 * 4 + 1], b.x, b.y, slope2);
          bias_lrelu2(v[i4 * 4 + 2], v[i4 * 4 + 3], b.z, b.w, slope2);
        }
        tmem_st32(accH + c * 32, v);
        store_h_bf16(v, c, spred);
      };
      // one 32-column chunk of the fc_local1 epilogue: u = lrelu(acc + bias) -> bf16 pairs in place in TMEM
      // (the activation is applied AFTER the rounding to bf16, on packed pairs: max(x, s*x) in bf16x2 halves the
      // ALU-pipe work of this epilogue; u is only ever consumed as bf16)
      auto epi1_chunk = [&](uint32_t (&v)[32], int c) {
        uint32_t u16[16];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = lds128(bl1_addr + (uint32_t)(c * 128 + i4 * 16));
          u16[i4 * 2 + 0] = bias_lrelu_bf16x2(v[i4 * 4 + 0], v[i4 * 4 + 1], b.x, b.y, slope_bf2);
          u16[i4 * 2 + 1] = bias_lrelu_bf16x2(v[i4 * 4 + 2], v[i4 * 4 + 3], b.z, b.w, slope_bf2);
        }
        tmem_st16(accU + c * 16, u16);        // columns [16c, 16c+16) were already read (16c+16 <= 32c+32)
      };

      // NOTE on code size: the layer loop below is deliberately kept compact (chunk loops rolled, one call site per
      // epilogue, head outside the loop).  With 12 warps at different program counters the instruction cache is a
      // first-order resource: the fully unrolled version of this loop body (~50 KB of SASS) ran ~4x slower per
      // instruction than this one in the serial per-jet phases.
      for (int ev = 0; ev < p.n_evals; ++ev) {
        // ---------------- unit 0 pack (stem biases + fc_g2) has landed; per-jet stem biases ----------------
        PROF_T(0);
        mbar_wait(&s.spk_full, c_spk++ & 1);
        PROF_T(1);
        for (int i = et; i < nj * TCH; i += 256) {
          const int j = i >> 7, o = i & 127;
          s.bl1[j][o] = unit_bias(LIN_L1, 0, s.jid[j], o);
          s.bl2[j][o] = unit_bias(LIN_L2, 128, s.jid[j], o);
        }
        ebar();
        // ---------------- fc_l1 on CUDA cores (K = a few features): h1 -> TMEM (fp32) + shared (bf16) ----------------
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
#pragma unroll
          for (int i4 = 0; i4 < 8; ++i4) {
            const float4 b = lds128(bl1_addr + (uint32_t)(c * 128 + i4 * 16));
            float a[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float* w = &s.w1s[c * 32 + i4 * 4 + q][0];
              if (FP == 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w);
                a[q] = fmaf(w4.x, xc[0], a[q]); a[q] = fmaf(w4.y, xc[1], a[q]); a[q] = fmaf(w4.z, xc[2], a[q]);
                a[q] = fmaf(w4.w, xc[3], a[q]);
              } else {
#pragma unroll
                for (int f = 0; f < FP; ++f) a[q] = fmaf(w[f], xc[f], a[q]);
              }
              v[i4 * 4 + q] = __float_as_uint(lrelu_tc(a[q], p.slope));
            }
          }
          tmem_st32(accH + c * 32, v);
          store_h_bf16(v, c, vpred);
        }
        tmem_wait_st();
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&s.hready[wg]);
        PROF_T(2);

        // [A] (off the critical path) pre[j] = bias_g1[o] + W_gg[o] . g_prev[j] of unit gi, for this thread's jets;
        // waits for the unit's small-weight pack first
        float pre[2][4];
        auto compute_pre = [&](int gi) {
          const int Ga = gi == 0 ? LIN_G1 : LIN_LAYER0 + 4 * (gi - 1) + 0;
          const int off_ga = gi == 0 ? 256 : 0;
          PROF_T(0);
          if (gi >= 1) mbar_wait(&s.spk_full, c_spk++ & 1);                        // unit gi's pack (unit 0: waited at eval start)
          PROF_T(5);
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int jb = (wg + 2 * bb) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) pre[bb][q] = 0.f;
            if (jb < nj) {
#pragma unroll
              for (int q = 0; q < 4; ++q) pre[bb][q] = (SIMPLE || jb + q < nj) ? unit_bias(Ga, off_ga, s.jid[jb + q], r) : 0.f;
              if (gi >= 1) {
#pragma unroll
                for (int z4 = 0; z4 < TC_ZMAX / 4; ++z4) {       // rows z >= Z of the pack are zero: no bound check
                  {
                    float w[4];
#pragma unroll
                    for (int zz = 0; zz < 4; ++zz) w[zz] = __bfloat162float(s.spk.gg[z4 * 4 + zz][r]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const float4 g = *reinterpret_cast<const float4*>(&s.gv[jb + q][z4 * 4]);
                      pre[bb][q] = fmaf(w[0], g.x, pre[bb][q]); pre[bb][q] = fmaf(w[1], g.y, pre[bb][q]);
                      pre[bb][q] = fmaf(w[2], g.z, pre[bb][q]); pre[bb][q] = fmaf(w[3], g.w, pre[bb][q]);
                    }
                  }
                }
              }
            }
          }
          PROF_T(8);
        };

        float sreg[2][4];                          // this thread's pooled sums S[c = r][its jets] (reused by layer 0)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb)
#pragma unroll
          for (int q = 0; q < 4; ++q) sreg[bb][q] = 0.f;

        compute_pre(0);
        ++ring_e;                                  // the stem's fc_l2 image
#pragma unroll 1
        for (int gi = 0;; ++gi) {
          // ======== residual update epilogue of the h version unit gi pools: the stem's fc_l2 (gi = 0), fc_local2 of layer
          // gi-2 (gi >= 2; layer 0 pools the same h as the stem, so nothing at gi = 1); gi = L+1: the last layer's.
          // The TMEM load of chunk c+1 is in flight while chunk c is processed.
          if (gi != 1) {
            const bool last = gi == L + 1;
            PROF_T(0);
            mbar_wait(&s.accH_full[wg], c_accH++ & 1);
            tc_fence_after();
            PROF_T(3);
            uint32_t va[32], vb[32];
            const uint32_t spred = last ? 0u : vpred;
            tmem_ld32(accH, va);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              tmem_wait_ld();
              tmem_ld32(accH + cc * 64 + 32, vb);
              epi_h_chunk(va, 2 * cc, spred);
              tmem_wait_ld();
              if (cc == 0) tmem_ld32(accH + 64, va);
              epi_h_chunk(vb, 2 * cc + 1, spred);
            }
            PROF_T(4);
            tmem_wait_st();
            if (last) break;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&s.hready[wg]);
            PROF_T(18);
          }
          // ======== tile A's warps stage this unit's W_g1 (mean | sum images of the ring) into TMEM as the A operand of the
          // fc_global1 MMAs: thread = output row o, 2 x 128 k values = 128 packed columns of tile A's fc_local1 accumulator
          {
            const uint32_t it_gm = ring_e++, it_gs = ring_e++;
            if (gi >= 1) ring_e += 2;              // fc_local1, fc_local2 images
            {                                      // tile A's warps copy the mean image, tile B's the sum image (same TMEM lanes)
              {
                const int m = wg;
                const uint32_t itw = m ? it_gs : it_gm;
                mbar_wait(&s.full[itw % TC_NSLOT], (itw / TC_NSLOT) & 1);
                const uint32_t img = smem_u32(s.w[itw % TC_NSLOT]) + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                  uint32_t v[32];
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float4 q = lds128(img + (uint32_t)(half * 16384) + ((uint32_t)(c << 4) ^ rx16));
                    v[c * 4 + 0] = __float_as_uint(q.x); v[c * 4 + 1] = __float_as_uint(q.y);
                    v[c * 4 + 2] = __float_as_uint(q.z); v[c * 4 + 3] = __float_as_uint(q.w);
                  }
                  tmem_st32(lane_base + 256u + (uint32_t)(m * 64 + half * 32), v);      // tile A's fc_local1 accumulator
                }
              }
              tmem_wait_st();
              tc_fence_before();
              mbar_arrive(&s.wg1_ready);
            }
          }
          // ======== global phase gi: 0 = stem (fc_g1, fc_g2), gi >= 1 = EPiC layer gi-1 (fc_global1/2) ========
          // Per-jet work is split over the 256 threads as (o = r) x (batches of 4 jets: wg, wg + 2).
          const int Gb = gi == 0 ? LIN_G2 : LIN_LAYER0 + 4 * (gi - 1) + 1;
          const int off_gb = gi == 0 ? 384 : 128;                                 // slice offset inside sbias
          // ---- [B] pooled sums -> pre-scaled bf16 B operand  St[j][0:128) = S/n (mean), St[j][128:256) = S*s (sum)
          if (gi != 1) {
            mbar_wait(&s.pool_full, c_pool++ & 1);
            tc_fence_after();
            PROF_T(6);
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
              const int jb = (wg + 2 * bb) * 4;
              if (jb < nj) {
                uint32_t v0[4], v1[4], v2[4], v3[4];
                tmem_ld4(dpool + jb, v0); tmem_ld4(dpool + 16 + jb, v1); tmem_ld4(dpool + 32 + jb, v2); tmem_ld4(dpool + 48 + jb, v3);
                tmem_wait_ld();
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  sreg[bb][q] = (__uint_as_float(v0[q]) + __uint_as_float(v1[q])) + (__uint_as_float(v2[q]) + __uint_as_float(v3[q]));
              }
            }
          }
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int jb = (wg + 2 * bb) * 4;
            if (jb < nj) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int j = jb + q;
                const uint32_t a = st_addr + sw128_offset(j, r, 2048);       // (j, k = r); (j, k = 128 + r) is 4 KB further
                const __nv_bfloat16 m = __float2bfloat16(sreg[bb][q] * s.inv_n[j]), sm = __float2bfloat16(sreg[bb][q] * p.sum_scale);
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(a), "h"(*reinterpret_cast<const unsigned short*>(&m)) : "memory");
                asm volatile("st.shared.b16 [%0], %1;" ::"r"(a + 4096u), "h"(*reinterpret_cast<const unsigned short*>(&sm)) : "memory");
              }
            }
          }
          PROF_T(7);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&s.glob_go);
          PROF_T(19);
          // ---- [C] g1[j][o] = lrelu(W_mean . mean + W_sum . sum (+ W_gg . g) + bias)     (epic.py:180-182, :375-377)
          mbar_wait(&s.glob_full, c_glob++ & 1);
          tc_fence_after();
          PROF_T(9);
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int jb = (wg + 2 * bb) * 4;
            if (jb < nj) {
              uint32_t v0[4], v1[4], v2[4], v3[4];
              tmem_ld4(dglob + jb, v0); tmem_ld4(dglob + 16 + jb, v1); tmem_ld4(dglob + 32 + jb, v2); tmem_ld4(dglob + 48 + jb, v3);
              tmem_wait_ld();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float d = (__uint_as_float(v0[q]) + __uint_as_float(v1[q])) + (__uint_as_float(v2[q]) + __uint_as_float(v3[q]));
                s.sg.g1[jb + q][r] = lrelu_tc(d + pre[bb][q], p.slope);
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&s.d_free);
          ebar();
          PROF_T(10);
          // ---- [D+E] one warp per jet, no block barrier in between:
          //   fc_g2 / fc_global2 (+ residual for the layers): lane = (k half, z), K = 2 x 64, one shuffle;
          //   then the jet's biases of fc_local1 (incl. W_glob . g, g broadcast by shuffles) and fc_local2
          {
            const int l = gi - 1;
            const int La = LIN_LAYER0 + 4 * l + 2, Lb = LIN_LAYER0 + 4 * l + 3;
            for (int j = (et >> 5); j < nj; j += 8) {
              const int lane_ = et & 31, z = lane_ & 15, kh = lane_ >> 4;
              const int zc = z < Z ? z : Z - 1;
              const uint32_t wa = smem_u32(&s.spk.g2[zc][kh * 64]), ga = smem_u32(&s.sg.g1[j][kh * 64]);
              float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const float4 w = lds128(wa + k * 16), g = lds128(ga + k * 16);
                a0 = fmaf(w.x, g.x, a0); a1 = fmaf(w.y, g.y, a1); a2 = fmaf(w.z, g.z, a2); a3 = fmaf(w.w, g.w, a3);
              }
              float acc = (a0 + a1) + (a2 + a3);
              acc += __shfl_xor_sync(0xffffffffu, acc, 16);
              acc += unit_bias(Gb, off_gb, s.jid[j], zc);
              if (gi >= 1) acc += s.gv[j][zc];
              const float gz = z < Z ? lrelu_tc(acc, p.slope) : 0.f;
              __syncwarp();                        // every lane has read the old g before it is overwritten
              if (kh == 0 && z < Z) s.gv[j][z] = gz;
              if (gi >= 1) {
                // lane = 4 consecutive outputs: one 8-byte load of W_glob per z, 16-byte bias loads / stores
                const int o4 = lane_ * 4;
                float b1v[4], b2v[4];
                if (SIMPLE) {
                  const float4 t1 = *reinterpret_cast<const float4*>(&s.sbias[128 + ZP + o4]);
                  const float4 t2 = *reinterpret_cast<const float4*>(&s.sbias[256 + ZP + o4]);
                  b1v[0] = t1.x; b1v[1] = t1.y; b1v[2] = t1.z; b1v[3] = t1.w;
                  b2v[0] = t2.x; b2v[1] = t2.y; b2v[2] = t2.z; b2v[3] = t2.w;
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    b1v[i] = unit_bias(La, 128 + ZP, s.jid[j], o4 + i);
                    b2v[i] = unit_bias(Lb, 256 + ZP, s.jid[j], o4 + i);
                  }
                }
#pragma unroll
                for (int zz = 0; zz < TC_ZMAX; ++zz) {           // rows zz >= Z of W_glob are zero and gz is 0 there
                  const float g = __shfl_sync(0xffffffffu, gz, zz);
                  const uint2 w = *reinterpret_cast<const uint2*>(&s.spk.gl[zz][o4]);
                  b1v[0] = fmaf(__uint_as_float(w.x << 16), g, b1v[0]);
                  b1v[1] = fmaf(__uint_as_float(w.x & 0xffff0000u), g, b1v[1]);
                  b1v[2] = fmaf(__uint_as_float(w.y << 16), g, b1v[2]);
                  b1v[3] = fmaf(__uint_as_float(w.y & 0xffff0000u), g, b1v[3]);
                }
                *reinterpret_cast<float4*>(&s.bl1[j][o4]) = make_float4(b1v[0], b1v[1], b1v[2], b1v[3]);
                *reinterpret_cast<float4*>(&s.bl2[j][o4]) = make_float4(b2v[0], b2v[1], b2v[2], b2v[3]);
              }
            }
          }
          mbar_arrive(&s.spk_empty);               // pack + bias slice of this unit are dead: the producer may refill
          ebar();
          PROF_T(13);
          if (gi >= 1) {
            // ======== fc_local1 epilogue: u = lrelu(acc + bias) -> bf16 pairs in place in TMEM ========
            mbar_wait(&s.accU_full[wg], c_accU++ & 1);
            tc_fence_after();
            PROF_T(14);
            uint32_t va[32], vb[32];
            tmem_ld32(accU, va);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              tmem_wait_ld();
              tmem_ld32(accU + cc * 64 + 32, vb);
              epi1_chunk(va, 2 * cc);
              tmem_wait_ld();
              if (cc == 0) tmem_ld32(accU + 64, va);
              epi1_chunk(vb, 2 * cc + 1);
              tmem_wait_st();                      // u columns [32 cc, 32 cc + 32) = K steps 4 cc .. 4 cc + 3 of fc_local2
              tc_fence_before();
              mbar_arrive(&s.u_ready[wg][cc]);
            }
            PROF_T(15);
          }
          if (gi < L) compute_pre(gi + 1);         // next unit's [A] while the tensor pipe runs fc_local2
        }
        // ---------------- head fc_l3 on CUDA cores: h_L (fp32) is re-read from TMEM ----------------
#pragma unroll
        for (int f = 0; f < FP; ++f) vout[f] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(accH + c * 32, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float a = __uint_as_float(v[i]);
            if (FP == 4) {
              const float4 w4 = lds128(w3_addr + (uint32_t)((c * 32 + i) * 16));
              vout[0] = fmaf(w4.x, a, vout[0]); vout[1] = fmaf(w4.y, a, vout[1]); vout[2] = fmaf(w4.z, a, vout[2]);
              vout[3] = fmaf(w4.w, a, vout[3]);
            } else {
#pragma unroll
              for (int f = 0; f < FP; ++f) vout[f] = fmaf(s.w3s[c * 32 + i][f], a, vout[f]);
            }
          }
        }
        // ---------------- head bias + activation, integrator step (thread-local) ----------------
        // (head bias and step size are loaded HERE, from L2, not at the start of the evaluation: five registers that were
        // live across the whole layer loop were what spilled)
        const float dt_ev = p.solver >= 0 ? p.dt[p.solver == PFM_SOLVER_MIDPOINT ? (ev >> 1) : ev] : 0.f;
#pragma unroll
        for (int f = 0; f < FP; ++f)
          if (f < F) {
            float b3 = p.tbias[(size_t)((!SIMPLE && p.tbias_per_jet) ? jg : ev) * p.bstride + s.boff[p.n_lin - 1] + f];
            if (!SIMPLE && p.cbias) b3 += p.cbias[(size_t)jg * p.bstride + s.boff[p.n_lin - 1] + f];
            vout[f] = valid ? lrelu_tc(vout[f] + b3, p.slope) : 0.f;
          }
        if (p.solver >= 0) {
          const bool mid = p.solver == PFM_SOLVER_MIDPOINT;
          const float dt = dt_ev;
          const bool first_stage = mid && ((ev & 1) == 0);
          const float hdt = __fmul_rn(0.5f, dt);
#pragma unroll
          for (int f = 0; f < FP; ++f) {
            const float k = -vout[f];
            if (first_stage) {
              xc[f] = __fadd_rn(x0[f], __fmul_rn(hdt, k));
            } else {
              x0[f] = __fadd_rn(x0[f], __fmul_rn(dt, k));
              xc[f] = x0[f];
            }
          }
        }
        PROF_T(16);
        ebar();      // bl1/bl2 of this evaluation are dead before the next one rewrites them
        PROF_T(17);
      }
      // ---------------- write back ----------------
      for (int j = 0; j < nj; ++j) {
        const int n = s.jrow0[j + 1] - s.jrow0[j];
        const float fill = n == 0 ? __int_as_float(0x7fc00000) : 0.f;
        float* dst = p.x_out + (size_t)s.jid[j] * p.N * F;
        for (int i = et; i < p.N * F; i += 256) dst[i] = fill;
      }
      ebar();
      if (valid) {
        float* dst = p.x_out + ((size_t)jg * p.N + part) * F;
#pragma unroll
        for (int f = 0; f < FP; ++f)
          if (f < F) dst[f] = p.solver >= 0 ? x0[f] : vout[f];
      }
    }
  }
  if (PROF && prof_on && p.prof) {
    const int role = prof_role;
    for (int i = 0; i < 20; ++i) p.prof[role * 20 + i] = prof[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

// ---------------------------------------------------------------------------------------------
// weight images: bf16, K-major SWIZZLE_128B, one 32 KB image per 128 x 128 block, in ring order
//   stem: fc_l2 | fc_g1[mean cols] | fc_g1[sum cols]      layer l: fc_local1 | fc_global1[mean] | fc_global1[sum] | fc_local2
// ---------------------------------------------------------------------------------------------
struct ImgSrc { const float* Wt; int ldo; int k0; };   // image[n][k] = Wt[(k0 + k) * ldo + n]

__global__ void pack_images_kernel(const ImgSrc* __restrict__ src, uint8_t* __restrict__ img) {
  const ImgSrc S = src[blockIdx.x];
  uint8_t* out = img + (size_t)blockIdx.x * TC_MAT;
  for (int i = threadIdx.x; i < 128 * 128; i += blockDim.x) {
    const int k = i >> 7, n = i & 127;        // consecutive threads -> consecutive n: coalesced reads of the k-major copy
    const float w = S.Wt[(size_t)(S.k0 + k) * S.ldo + n];
    *reinterpret_cast<__nv_bfloat16*>(out + sw128_offset(n, k, 16384)) = __float2bfloat16(w);
  }
}

// small-weight pack of unit u (SpkPack): W_gg[z][o] = fc_global1[o][2H + z], W_glob[z][o] = fc_local1[o][H + z] (bf16),
// W_g2[z][k] = fc_global2[z][k] (fc_g2 for the stem; fp32); zero padding to 16 rows
struct SpkSrc { const float* gg; int gg_ldo; const float* gl; int gl_ldo; const float* g2; int g2_ldo; int Z; };

__global__ void pack_spk_kernel(const SpkSrc* __restrict__ src, SpkPack* __restrict__ out) {
  const SpkSrc S = src[blockIdx.x];
  SpkPack& o = out[blockIdx.x];
  for (int i = threadIdx.x; i < 16 * TCH; i += blockDim.x) {
    const int z = i >> 7, c = i & 127;
    o.gg[z][c] = __float2bfloat16((S.gg && z < S.Z) ? S.gg[(size_t)z * S.gg_ldo + c] : 0.f);
    o.gl[z][c] = __float2bfloat16((S.gl && z < S.Z) ? S.gl[(size_t)z * S.gl_ldo + c] : 0.f);
    o.g2[z][c] = z < S.Z ? S.g2[(size_t)c * S.g2_ldo + z] : 0.f;
    if (c < 4) o.g2[z][TCH + c] = 0.f;
  }
}

int tc_supported(const pfm_epic* h, int N) {
  const pfm_epic_cfg& c = h->cfg;
  if (c.hid != TCH) { set_error("PFM_PREC_BF16 needs hid == 128 (got %d); use PFM_PREC_FP32", c.hid); return PFM_ERR_UNSUPPORTED; }
  if (c.latent > TC_ZMAX) { set_error("PFM_PREC_BF16 needs latent <= %d (got %d)", TC_ZMAX, c.latent); return PFM_ERR_UNSUPPORTED; }
  if (c.feats > TC_KXMAX) { set_error("PFM_PREC_BF16 needs feats <= %d (got %d)", TC_KXMAX, c.feats); return PFM_ERR_UNSUPPORTED; }
  if (c.layers < 1 || c.layers > 30) { set_error("PFM_PREC_BF16 needs 1..30 EPiC layers (got %d)", c.layers); return PFM_ERR_UNSUPPORTED; }
  if (N > TC_ROWS) { set_error("PFM_PREC_BF16: a jet of %d particles exceeds the %d-row group; use PFM_PREC_FP32", N, TC_ROWS); return PFM_ERR_UNSUPPORTED; }
  if ((int)(sizeof(TcSmem<8>) + 1024) > h->max_smem_optin) { set_error("PFM_PREC_BF16: not enough shared memory per block"); return PFM_ERR_UNSUPPORTED; }
  return PFM_OK;
}

int tc_plan_caps(const pfm_epic* h, int N, int* R_cap, int* J_cap) {
  (void)h; (void)N;
  *R_cap = TC_ROWS; *J_cap = TC_J;
  return PFM_OK;
}

int tc_pack_weights(pfm_epic* h, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  const int n_items = 3 + 4 * c.layers;
  std::vector<ImgSrc> src(n_items);
  auto mk = [&](int lin_idx, int k_off) {
    const Lin& L = h->lin_host[lin_idx];
    ImgSrc s; s.Wt = L.Wt; s.ldo = L.ldo; s.k0 = L.m_off + k_off; return s;
  };
  int it = 0;
  src[it++] = mk(LIN_L2, 0);
  src[it++] = mk(LIN_G1, TCH);      // stem concat order is (sum, mean): the mean block is second (epic.py:373)
  src[it++] = mk(LIN_G1, 0);
  for (int l = 0; l < c.layers; ++l) {
    src[it++] = mk(LIN_LAYER0 + 4 * l + 0, 0);       // fc_global1: (mean, sum, global) order (epic.py:164-171); first in the
    src[it++] = mk(LIN_LAYER0 + 4 * l + 0, TCH);     //  ring: tile A's warps stage them into TMEM as soon as they land
    src[it++] = mk(LIN_LAYER0 + 4 * l + 2, 0);       // fc_local1, particle columns
    src[it++] = mk(LIN_LAYER0 + 4 * l + 3, 0);       // fc_local2
  }
  std::vector<SpkSrc> spk(c.layers + 1);
  for (int u = 0; u <= c.layers; ++u) {
    SpkSrc q;
    memset(&q, 0, sizeof(q));
    q.Z = c.latent;
    if (u == 0) {
      const Lin& G2 = h->lin_host[LIN_G2];
      q.g2 = G2.Wt + (size_t)G2.m_off * G2.ldo; q.g2_ldo = G2.ldo;
    } else {
      const Lin& Ga = h->lin_host[LIN_LAYER0 + 4 * (u - 1) + 0];
      const Lin& Gb = h->lin_host[LIN_LAYER0 + 4 * (u - 1) + 1];
      const Lin& La = h->lin_host[LIN_LAYER0 + 4 * (u - 1) + 2];
      q.gg = Ga.Wt + (size_t)(Ga.m_off + 2 * TCH) * Ga.ldo; q.gg_ldo = Ga.ldo;
      q.gl = La.Wt + (size_t)La.g_off * La.ldo; q.gl_ldo = La.ldo;
      q.g2 = Gb.Wt + (size_t)Gb.m_off * Gb.ldo; q.g2_ldo = Gb.ldo;
    }
    spk[u] = q;
  }
  const size_t img_bytes = (size_t)n_items * TC_MAT;
  const size_t spk_bytes = (size_t)(c.layers + 1) * TC_SPK;
  const size_t bytes = img_bytes + spk_bytes;
  const size_t aux = sizeof(ImgSrc) * n_items + sizeof(SpkSrc) * spk.size();
  bool upload = false;
  if (h->tc_bytes < bytes + aux) {
    if (h->tc_store) cudaFree(h->tc_store);
    h->tc_store = nullptr; h->tc_bytes = 0;
    PFM_CUDA_CHECK(cudaMalloc(&h->tc_store, bytes + aux));
    h->tc_bytes = bytes + aux;
    upload = true;
  }
  uint8_t* base = reinterpret_cast<uint8_t*>(h->tc_store);
  ImgSrc* dsrc = reinterpret_cast<ImgSrc*>(base + bytes);
  SpkSrc* dspk = reinterpret_cast<SpkSrc*>(base + bytes + sizeof(ImgSrc) * n_items);
  if (upload) {      // the source tables only hold pointers into the handle's fp32 weight store: uploaded once per allocation,
                     // so that a repack (every sample() re-syncs the weights) is two launches and never blocks the host
    PFM_CUDA_CHECK(cudaMemcpyAsync(dsrc, src.data(), sizeof(ImgSrc) * n_items, cudaMemcpyHostToDevice, st));
    PFM_CUDA_CHECK(cudaMemcpyAsync(dspk, spk.data(), sizeof(SpkSrc) * spk.size(), cudaMemcpyHostToDevice, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));     // src / spk are host temporaries
  }
  pack_images_kernel<<<n_items, 256, 0, st>>>(dsrc, base);
  pack_spk_kernel<<<c.layers + 1, 256, 0, st>>>(dspk, reinterpret_cast<SpkPack*>(base + img_bytes));
  PFM_CUDA_CHECK(cudaGetLastError());
  h->tc_dirty = false;
  return PFM_OK;
}

template <int FP, bool PROF, bool SIMPLE>
static int launch_tc(const TcParams& p, int grid, cudaStream_t st) {
  auto kern = epic_tc_kernel<FP, PROF, SIMPLE>;
  const int smem = (int)sizeof(TcSmem<FP>) + 1024;
  PFM_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  kern<<<grid, TC_THREADS, smem, st>>>(p);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int tc_run(pfm_epic* h, const RunArgs& a, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  if (a.Kx > TC_KXMAX) {
    set_error("PFM_PREC_BF16: %d per-particle input columns exceed %d (add_time_to_input through pfm_epic_forward); "
              "use PFM_PREC_FP32 or the sampling entry point, which hoists the time columns", a.Kx, TC_KXMAX);
    return PFM_ERR_UNSUPPORTED;
  }
  if (!h->tc_store || h->tc_dirty) { int rc = tc_pack_weights(h, st); if (rc != PFM_OK) return rc; }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.F = c.feats; p.Kx = a.Kx; p.x_ld = a.Kx; p.xin_off = a.xin_off; p.Z = c.latent; p.L = c.layers; p.n_lin = h->n_lin;
  p.n_items = 3 + 4 * c.layers;
  p.sum_scale = c.sum_scale; p.slope = c.neg_slope;
  p.lin = h->lin_dev; p.wimg = reinterpret_cast<const uint8_t*>(h->tc_store);
  p.spk = p.wimg + (size_t)p.n_items * TC_MAT;
  {
    const int ZP = (c.latent + 3) & ~3;
    p.boff_stem = h->lin_host[LIN_L1].bias_off;
    p.boff_layer0 = h->lin_host[LIN_LAYER0].bias_off;
    p.boff_layer_stride = 3 * TCH + ZP;
    p.bias_chunk_floats = 3 * TCH + ZP;
  }
  p.tbias = h->tbias; p.cbias = a.has_cbias ? h->cbias : nullptr; p.bstride = h->bstride; p.tbias_per_jet = a.tbias_per_jet;
  p.n_real = h->plan.n_real; p.ridx = h->plan.ridx; p.groups = h->plan.groups; p.n_groups = h->plan.n_groups;
  p.counter = h->plan.counter; p.jetmap = a.jetmap;
  p.x_in = a.x_in; p.x_out = a.x_out; p.B = a.B; p.N = a.N;
  p.n_evals = a.n_evals; p.solver = a.solver; p.n_steps = a.n_steps; p.dt = a.dt;
  const int grid = h->sm_count < a.B ? h->sm_count : a.B;
  const int kmax = a.Kx > c.feats ? a.Kx : c.feats;
  const bool simple = !p.tbias_per_jet && !p.cbias;
  if (getenv("PFM_TC_PROF")) {        // debug: phase timers of block 0, printed after the kernel
    static long long* dprof = nullptr;
    const int n_ll = 60 + 3 * TRACE_MAX * 2;
    if (!dprof) PFM_CUDA_CHECK(cudaMalloc(&dprof, sizeof(long long) * n_ll));
    PFM_CUDA_CHECK(cudaMemsetAsync(dprof, 0, sizeof(long long) * n_ll, st));
    p.prof = dprof;
    int rc = kmax <= 4 ? (simple ? launch_tc<4, true, true>(p, grid, st) : launch_tc<4, true, false>(p, grid, st))
                       : launch_tc<8, true, false>(p, grid, st);
    if (rc != PFM_OK) return rc;
    static long long hp[60 + 3 * TRACE_MAX * 2];
    PFM_CUDA_CHECK(cudaMemcpyAsync(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));
    if (getenv("PFM_TC_TRACE")) {
      long long t0 = 0;
      for (int r = 0; r < 3; ++r) { const long long t = hp[61 + r * TRACE_MAX * 2]; if (t && (!t0 || t < t0)) t0 = t; }
      for (int r = 0; r < 3; ++r)
        for (int i = 0; i < TRACE_MAX; ++i) {
          const long long slot = hp[60 + (r * TRACE_MAX + i) * 2], t = hp[61 + (r * TRACE_MAX + i) * 2];
          if (t) fprintf(stderr, "[pfm tc trace] %lld %d %lld\n", t - t0, r, slot);
        }
    }
    const char* roles[3] = {"mma", "epiA", "epiB"};
    for (int r = 0; r < 3; ++r) {
      long long tot = 0;
      for (int i = 0; i < 20; ++i) tot += hp[r * 20 + i];
      fprintf(stderr, "[pfm tc prof] %-4s total %lld cyc:", roles[r], tot);
      for (int i = 0; i < 20; ++i) fprintf(stderr, " %d:%.1f%%", i, tot ? 100.0 * hp[r * 20 + i] / tot : 0.0);
      fprintf(stderr, "\n");
    }
    return PFM_OK;
  }
  if (kmax <= 4) return simple ? launch_tc<4, false, true>(p, grid, st) : launch_tc<4, false, false>(p, grid, st);
  return simple ? launch_tc<8, false, true>(p, grid, st) : launch_tc<8, false, false>(p, grid, st);
}

}  // namespace pfm
