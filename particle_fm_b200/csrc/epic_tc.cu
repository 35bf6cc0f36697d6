// bf16 tensor-core path (tcgen05 / TMEM / bulk async copies + warp-level mma.sync) of the EPiC vector field + integrator.
//
// One persistent CTA owns a group of jets whose real particles fill up to two 128-row tiles and runs
// the WHOLE integration for them (one launch per sample()).  Per-particle state never leaves the SM:
//   TMEM   cols [0,256)    fp32 hidden features h of tile A / tile B  = residual stream AND accumulator
//                          of fc_local2 / fc_l2 (the MMA accumulates straight onto the residual)
//          cols [256,512)  fp32 accumulators of fc_local1 per tile; the epilogue overwrites them in place
//                          with the bf16 activations u that feed fc_local2 as a TMEM A-operand.
//   SMEM   h tiles as bf16 (K-major SWIZZLE_128B = A operand of fc_local1), a 3-slot ring of 32 KB pre-swizzled
//          weight images fed by cp.async.bulk, per-jet pooled sums / global vectors / biases.
//   REGS   every epilogue thread owns one particle: its ODE state x lives in registers for all steps.
// The two per-particle GEMMs of a layer (128 x 128 x 128 per tile) are tcgen05.mma with TMEM accumulators.
// The per-jet chain between them -- masked mean/sum pooling, fc_global1 (2H+Z -> H), fc_global2 (H -> Z, residual),
// re-injection of the global vector as a per-jet bias of fc_local1 -- runs on the EPILOGUE warps with warp-level
// mma.sync (M = 16 jets is exactly one fragment row block), so that it needs no round trip through the tcgen05 pipe:
//   pooling      S[jet][c] = sum_rows P[jet][row] h[row][c]: every warp reduces its own 32 rows right after it stored them
//                (A = 0/1 indicator fragments built once per group, B = ldmatrix.trans of the swizzled h tile); the partial
//                sums go to a slot per (warp, jet) pair and are added in a fixed order (deterministic, no atomics)
//   fc_global1   [16 jets x 272] . [272 x 128]: warp w owns 16 outputs; B fragments come straight out of the ring images
//   fc_global2   [16 x 128] . [128 x 16], every warp redundantly (the result stays in registers: no exchange)
//   re-injection [16 x 16] . [16 x 128]: A fragment = the new global vector, converted in registers
// (Round 1 ran pooling and fc_global1 as N = 16 tcgen05 MMAs and the rest as CUDA-core GEMVs: four dependent hops through
// mbarriers / TMEM per layer, ~4 k of the ~9 k cycles of a layer.)
// Warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one lane), warps 2-5 / 6-9 = epilogue
// warpgroups of tile A / tile B (TMEM lane quadrant = warp % 4).
#include <cstdlib>

#include "pfm_internal.cuh"
#include "tc_ptx.cuh"

namespace pfm {

using namespace tc;

static constexpr int TCH = 128;          // hidden width this path is specialised for
static constexpr int TC_ROWS = 256;      // 2 tiles of 128 particles
static constexpr int TC_J = 16;          // jets per group (M of the per-jet mma.sync fragments)
static constexpr int TC_ZMAX = 16;        // latent width (10 / 16 in the configs); the small-weight pack is laid out for 16
static constexpr int TC_KXMAX = 8;          // per-particle input columns / features (3 JetNet, 8 JetClass)
// row strides (in bf16 elements) of the small row-major mma.sync operands, chosen so that the 8 rows of an ldmatrix
// 8x8 block start in different bank groups
static constexpr int TC_GG_LD = 24;      // [o][z] weights, [jet][z] global vectors: 48-byte rows
static constexpr int TC_G2_LD = 136;     // W_g2 [z][k], g1 [jet][k]: 272-byte rows
static constexpr int TC_ST_LD = 264;     // St [jet][mean 128 | sum 128]: 528-byte rows
// per-unit small-weight pack, one bulk copy (unit 0 = stem: only g2 = fc_g2 is used)
struct SpkPack {
  __nv_bfloat16 gg[TCH][TC_GG_LD];       // W_gg[o][z]   = fc_global1[o][2H + z]   B operand of K step 16 of fc_global1
  __nv_bfloat16 gl[TCH][TC_GG_LD];       // W_glob[o][z] = fc_local1[o][H + z]     B operand of the re-injection
  __nv_bfloat16 g2[TC_ZMAX][TC_G2_LD];   // W_g2[z][k]   = fc_global2[z][k]        B operand of fc_global2 (rows z >= Z zero)
};
static constexpr uint32_t TC_SPK = sizeof(SpkPack);
static_assert(sizeof(SpkPack) % 16 == 0, "bulk copies move multiples of 16 bytes");
static constexpr int TC_SBIAS = 384 + 16;               // one unit's slice of the time-bias table
static constexpr int TC_THREADS = 320;
static constexpr int TC_NSLOT = 3;
static constexpr uint32_t TC_MAT = 32768;   // one 128x128 bf16 weight image
static constexpr int TC_PAIRS = 24;         // (32-row block, jet) pairs of a group: jets are contiguous, so at most 16 + 8 - 1

template <int FP>
struct TcSmem {
  alignas(1024) uint8_t h[2][TC_MAT];          // bf16 h tiles
  alignas(1024) uint8_t w[TC_NSLOT][TC_MAT];   // weight ring
  alignas(16) float Spart[TC_PAIRS][TCH];      // pooled partial sums: slot (jet + 32-row block) <- that block's rows of the jet
  union alignas(16) {
    __nv_bfloat16 St[TC_J][TC_ST_LD];          // A operand of fc_global1: [jet][mean | sum], dead once fc_global1 is done ->
    float bl1[TC_J][TCH];                      //  reused for the per-jet bias of fc_local1 (b + W_t . t + W_glob . g)
  } sb;
  alignas(16) __nv_bfloat16 g1[TC_J][TC_G2_LD];   // fc_global1 output, A operand of fc_global2
  alignas(16) __nv_bfloat16 Sg[TC_J][TC_GG_LD];   // global vectors as bf16: K columns 256.. of fc_global1's A operand
  alignas(16) float bl2[TC_J][TCH];            // per-jet bias of fc_local2 / fc_l2
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
  uint64_t spk_full, spk_empty;
};

static_assert(sizeof(TcSmem<8>) + 1024 <= 232448, "TcSmem exceeds the 227 KB per-block shared-memory limit of sm_100");
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
#define PROF_T(slot) do { if (PROF && prof_on) { const long long _n = clock64(); prof[slot] += _n - prof_t; prof_t = _n; } } while (0)

// SIMPLE: one time code for the whole batch and no conditioning (the sampling configuration of the headline
// workload): every bias comes from the staged slice of the time-bias table, which removes all data-dependent
// branches from the serial per-jet phases.
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
  const bool prof_on = PROF && blockIdx.x == 0 && (tid == 32 || tid == 128 || tid == 256);   // lane 0 of the MMA warp, of a warp of epilogue A / B
  const int prof_role = tid == 32 ? 0 : (tid == 128 ? 1 : 2);
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
  uint32_t c_hready[2] = {0, 0}, c_uready[2] = {0, 0};      // MMA side
  uint32_t ring_e = 0;                                      // epilogue: mirror of the weight-ring position (it reads the fc_global1 images)
  uint32_t c_accH = 0, c_accU = 0, c_spk = 0;               // epilogue side
  uint32_t spk_it = 0;                                      // producer: small-weight packs issued so far

  for (;;) {
    __syncthreads();
    if (tid == 0) s.group = atomicAdd(p.counter, 1);
    __syncthreads();
    const int gidx = s.group;
    if (gidx >= n_groups) break;
    const int2 grp = p.groups[gidx];
    const int j0 = grp.x, nj = grp.y;

    if (warp == 0) {
      // ================================ weight producer (whole warp, elected lane issues) ================================
      // ring order per evaluation:  stem  fc_l2 | fc_g1 mean | fc_g1 sum     layer l  fc_local1 | fc_global1 mean | sum | fc_local2
      const bool stage_bias = !p.tbias_per_jet;
      const uint32_t bias_bytes = stage_bias ? (uint32_t)p.bias_chunk_floats * 4u : 0u;
      for (int ev = 0; ev < p.n_evals; ++ev) {
        for (int it = 0; it < p.n_items; ++it, ++ring_it) {
          // the small-weight pack + bias slice of unit u travel just before the unit's first image
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
    } else if (warp == 1) {
      // ================================ MMA issuer (whole warp, elected lane issues) ================================
      const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t hA = desc_kmajor(smem_u32(s.h[0])), hB = desc_kmajor(smem_u32(s.h[1]));
      const uint64_t wdesc0 = desc_kmajor(smem_u32(s.w[0]));
      const uint32_t accH0 = tm, accH1 = tm + 128, accU0 = tm + 256, accU1 = tm + 384;
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
        ring_it += 2;                                  // fc_g1 images: read (and released) by the epilogue warps
        for (int l = 0; l < L; ++l) {
          const uint32_t it_w1 = ring_it++;
          ring_it += 2;                                // fc_global1 images
          const uint32_t it_w2 = ring_it++;
          // ---- fc_local1: accU[t] = h_l . W1^T as soon as the tile's new h is in shared memory; its result is only needed
          // once the per-jet chain has produced the bias, so these MMAs hide under the chain
          for (int t = 0; t < 2; ++t) {
            PROF_T(0);
            mbar_wait(&s.hready[t], c_hready[t]++ & 1);
            tc_fence_after();
            PROF_T(3);
            if (t == 0) wait_full(it_w1);
            PROF_T(4);
            issue_ss_128(t ? accU1 : accU0, t ? hB : hA, wslot(it_w1), idesc, false);
            commit_to(&s.accU_full[t]);
          }
          commit_to(&s.empty[it_w1 % TC_NSLOT]);
          // ---- fc_local2: accH[t] += u[t] (bf16 in TMEM) . W2^T.  The fc_local1 epilogue hands over u in two halves of 64
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
    } else {
      // ================================ epilogue / particle threads (warps 2..9) ================================
      const int et = tid - 64;
      const int wg = et >> 7;                    // tile
      const int q = warp & 3;                    // TMEM lane quadrant this warp may access
      const int r = q * 32 + lane;               // row in tile = TMEM lane
      const int row = wg * 128 + r;
      const int w8 = wg * 4 + q;                 // 32-row block of the group owned by this warp
      const int g8 = lane >> 2, t4 = lane & 3;   // mma.sync fragment coordinates
      const uint32_t lane_base = tm + ((uint32_t)(q * 32) << 16);
      const uint32_t accH = lane_base + wg * 128, accU = lane_base + 256 + wg * 128;
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
      // h = 0: rows without a particle are never written afterwards, so they stay finite (0) in every operand the
      // pooling fragments read; their TMEM lanes hold finite junk that no real row ever sees.  The per-jet operand
      // arrays are zeroed once so that the fragment rows of jets the group does not have hold finite values too.
      for (int i = et; i < 2 * (int)TC_MAT / 16; i += 256) reinterpret_cast<uint4*>(&s.h[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.sb) / 16; i += 256) reinterpret_cast<uint4*>(&s.sb)[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.g1) / 16; i += 256) reinterpret_cast<uint4*>(&s.g1[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.Sg) / 16; i += 256) reinterpret_cast<uint4*>(&s.Sg[0][0])[i] = make_uint4(0, 0, 0, 0);
      ebar();
      const int R = s.jrow0[nj];
      const bool valid = row < R;
      int myjet = 0;
      for (int j = 1; j < nj; ++j) myjet += (row >= s.jrow0[j]) ? 1 : 0;
      if (!valid) myjet = 0;
      const int jg = s.jid[myjet];
      // pooling A fragments: 0/1 indicators P[jet][row] of this warp's 32 rows (two K = 16 steps), and which of this lane's
      // two fragment rows (jets g8, g8 + 8) have particles in the warp's rows at all
      uint32_t ind[2][4];
      {
        auto in_jet = [&](int j, int rr) -> uint32_t { return (rr >= s.jrow0[j] && rr < s.jrow0[j + 1]) ? 0x3F80u : 0u; };   // bf16 1.0
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int rb = w8 * 32 + kk * 16 + 2 * t4;
          ind[kk][0] = in_jet(g8, rb) | (in_jet(g8, rb + 1) << 16);
          ind[kk][1] = in_jet(g8 + 8, rb) | (in_jet(g8 + 8, rb + 1) << 16);
          ind[kk][2] = in_jet(g8, rb + 8) | (in_jet(g8, rb + 9) << 16);
          ind[kk][3] = in_jet(g8 + 8, rb + 8) | (in_jet(g8 + 8, rb + 9) << 16);
        }
      }
      const bool pres_lo = s.jrow0[g8] < w8 * 32 + 32 && s.jrow0[g8 + 1] > w8 * 32;
      const bool pres_hi = s.jrow0[g8 + 8] < w8 * 32 + 32 && s.jrow0[g8 + 9] > w8 * 32;
      const bool jet_lo = g8 < nj, jet_hi = g8 + 8 < nj;         // fragment rows that are jets of this group
      const int jg_lo = s.jid[g8], jg_hi = s.jid[g8 + 8];
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
      const uint32_t bl1_addr = smem_u32(&s.sb.bl1[myjet][0]), bl2_addr = smem_u32(&s.bl2[myjet][0]);
      const uint32_t slope_bf2 = pack_bf16x2(p.slope, p.slope);
      const unsigned long long slope2 = ((unsigned long long)__float_as_uint(p.slope) << 32) | __float_as_uint(p.slope);
      const int ZP = (Z + 3) & ~3;
      const uint32_t w3_addr = smem_u32(&s.w3s[0][0]);
      // per-lane ldmatrix row addresses (see tc_ptx.cuh for the fragment layouts)
      const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;            // A operands [jet][k]: matrices (rows 0-7 | 8-15) x (k lo | hi)
      const int akh = lane >> 4;
      const int brow = (lane & 7) + (lane >> 4) * 8;                  // B operands [n][k]: matrices (n 0-7, k lo | hi), (n 8-15, k lo | hi)
      const int bkh = (lane >> 3) & 1;
      const int o_b = w8 * 16 + brow;                                 // this warp's 16 outputs of fc_global1 / the re-injection
      const uint32_t st_lane = smem_u32(&s.sb.St[0][0]) + (uint32_t)(arow * TC_ST_LD * 2 + akh * 16);
      const uint32_t sg_lane = smem_u32(&s.Sg[0][0]) + (uint32_t)(arow * TC_GG_LD * 2 + akh * 16);
      const uint32_t g1_lane = smem_u32(&s.g1[0][0]) + (uint32_t)(arow * TC_G2_LD * 2 + akh * 16);
      const uint32_t wg2_lane = smem_u32(&s.spk.g2[0][0]) + (uint32_t)(brow * TC_G2_LD * 2 + bkh * 16);
      const uint32_t wgg_lane = smem_u32(&s.spk.gg[0][0]) + (uint32_t)(o_b * TC_GG_LD * 2 + bkh * 16);
      const uint32_t wgl_lane = smem_u32(&s.spk.gl[0][0]) + (uint32_t)(o_b * TC_GG_LD * 2 + bkh * 16);
      const uint32_t img_row = (uint32_t)((o_b >> 3) * 1024 + (o_b & 7) * 128);     // row o_b of a K-major SW128 image
      const uint32_t img_swz = (uint32_t)(o_b & 7);
      // pooling B operand: this warp's rows of the h tile through ldmatrix.trans; matrices (rows lo | hi) x (8-column unit lo | hi)
      const uint32_t pool_lane = smem_u32(s.h[wg]) + (uint32_t)((q * 4 + ((lane >> 3) & 1)) * 1024 + (lane & 7) * 128);
      const uint32_t pool_swz = (uint32_t)(lane & 7), pool_cu = (uint32_t)(lane >> 4);
      const uint32_t spart_lane = smem_u32(&s.Spart[0][0]) + (uint32_t)(((g8 + w8) * TCH + 2 * t4) * 4);

      // store 32 fp32 columns [32c, 32c+32) of this particle's row as bf16 into the swizzled h tile
      auto store_h_bf16 = [&](const uint32_t (&v)[32], int c, uint32_t pred) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[qq * 8 + 0]), __uint_as_float(v[qq * 8 + 1]));
          pk.y = pack_bf16x2(__uint_as_float(v[qq * 8 + 2]), __uint_as_float(v[qq * 8 + 3]));
          pk.z = pack_bf16x2(__uint_as_float(v[qq * 8 + 4]), __uint_as_float(v[qq * 8 + 5]));
          pk.w = pack_bf16x2(__uint_as_float(v[qq * 8 + 6]), __uint_as_float(v[qq * 8 + 7]));
          const int c16 = c * 4 + qq;
          sts128_if(hrow_addr + (uint32_t)((c16 >> 3) * 16384) + ((uint32_t)((c16 & 7) << 4) ^ rx16), pk.x, pk.y, pk.z, pk.w, pred);
        }
      };
      // masked pooling of columns [32c, 32c+32) over this warp's 32 rows (just stored): 8 mma.sync, partial sums -> Spart
      auto pool_chunk = [&](int c) {
        float pa[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { pa[i][0] = 0.f; pa[i][1] = 0.f; pa[i][2] = 0.f; pa[i][3] = 0.f; }
        __syncwarp();
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
          for (int np = 0; np < 2; ++np) {
            const uint32_t cu = (uint32_t)(c * 4 + np * 2) + pool_cu;            // 8-column unit (16 bytes) inside the row
            uint32_t fb[4];
            ldsm_x4_trans(pool_lane + (uint32_t)(kk * 2048) + (cu >> 3) * 16384u + (((cu & 7u) ^ pool_swz) << 4), fb);
            hmma_bf16(pa[np * 2], ind[kk][0], ind[kk][1], ind[kk][2], ind[kk][3], fb[0], fb[1]);
            hmma_bf16(pa[np * 2 + 1], ind[kk][0], ind[kk][1], ind[kk][2], ind[kk][3], fb[2], fb[3]);
          }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const uint32_t a = spart_lane + (uint32_t)((c * 32 + nt * 8) * 4);
          sts64_if(a, pa[nt][0], pa[nt][1], pres_lo);
          sts64_if(a + 8u * TCH * 4u, pa[nt][2], pa[nt][3], pres_hi);
        }
      };
      // one 32-column chunk of the residual update: h = lrelu(acc + bias) -> TMEM fp32 (in place) + shared bf16 (+ pooling)
      auto epi_h_chunk = [&](uint32_t (&v)[32], int c, bool last) {
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = lds128(bl2_addr + (uint32_t)(c * 128 + i4 * 16));
          bias_lrelu2(v[i4 * 4 + 0], v[i4 * 4 + 1], b.x, b.y, slope2);
          bias_lrelu2(v[i4 * 4 + 2], v[i4 * 4 + 3], b.z, b.w, slope2);
        }
        tmem_st32(accH + c * 32, v);
        if (!last) {
          store_h_bf16(v, c, vpred);
          pool_chunk(c);
        }
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

      float greg[2][4];                          // global vectors of the group in C-fragment layout: [z tile][jet g8: z 2t4, 2t4+1 | jet g8+8: same]
      for (int ev = 0; ev < p.n_evals; ++ev) {
        // ---------------- unit 0 pack (stem biases + fc_g2) has landed; per-jet stem biases ----------------
        PROF_T(0);
        mbar_wait(&s.spk_full, c_spk++ & 1);
        PROF_T(1);
        float b3[FP];                              // head bias and step size: loaded now, used at the end of the evaluation
#pragma unroll
        for (int f = 0; f < FP; ++f) {
          b3[f] = 0.f;
          if (f < F) {
            b3[f] = p.tbias[(size_t)((!SIMPLE && p.tbias_per_jet) ? jg : ev) * p.bstride + s.boff[p.n_lin - 1] + f];
            if (!SIMPLE && p.cbias) b3[f] += p.cbias[(size_t)jg * p.bstride + s.boff[p.n_lin - 1] + f];
          }
        }
        const float dt_ev = p.solver >= 0 ? p.dt[p.solver == PFM_SOLVER_MIDPOINT ? (ev >> 1) : ev] : 0.f;
        for (int i = et; i < nj * TCH; i += 256) {
          const int j = i >> 7, o = i & 127;
          s.sb.bl1[j][o] = unit_bias(LIN_L1, 0, s.jid[j], o);
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
            for (int qq = 0; qq < 4; ++qq) {
              const float* w = &s.w1s[c * 32 + i4 * 4 + qq][0];
              if (FP == 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(w);
                a[qq] = fmaf(w4.x, xc[0], a[qq]); a[qq] = fmaf(w4.y, xc[1], a[qq]); a[qq] = fmaf(w4.z, xc[2], a[qq]);
                a[qq] = fmaf(w4.w, xc[3], a[qq]);
              } else {
#pragma unroll
                for (int f = 0; f < FP; ++f) a[qq] = fmaf(w[f], xc[f], a[qq]);
              }
              v[i4 * 4 + qq] = __float_as_uint(lrelu_tc(a[qq], p.slope));
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
            tmem_ld32(accH, va);
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
              tmem_wait_ld();
              tmem_ld32(accH + cc * 64 + 32, vb);
              epi_h_chunk(va, 2 * cc, last);
              tmem_wait_ld();
              if (cc == 0) tmem_ld32(accH + 64, va);
              epi_h_chunk(vb, 2 * cc + 1, last);
            }
            PROF_T(4);
            tmem_wait_st();
            if (last) break;
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&s.hready[wg]);           // MMA warp: fc_local1 of this tile may start
            ebar();                               // every warp's partial sums are in Spart; bl1 of the previous unit is dead
            PROF_T(5);
            // ---- pooled sums -> pre-scaled bf16 A operand of fc_global1:  St[j][0:128) = S/n (mean), St[j][128:256) = S*s (sum)
            for (int i = et; i < nj * TCH; i += 256) {
              const int j = i >> 7, c = i & 127;
              const int r0 = s.jrow0[j], r1 = s.jrow0[j + 1];
              float S = 0.f;
              if (r1 > r0)
                for (int w = r0 >> 5; w <= ((r1 - 1) >> 5); ++w) S += s.Spart[j + w][c];
              s.sb.St[j][c] = __float2bfloat16(S * s.inv_n[j]);
              s.sb.St[j][TCH + c] = __float2bfloat16(S * p.sum_scale);
            }
            PROF_T(6);
          }
          // ======== global phase gi: 0 = stem (fc_g1, fc_g2), gi >= 1 = EPiC layer gi-1 (fc_global1/2) ========
          if (gi >= 1) ++ring_e;                   // fc_local1 image
          const uint32_t it_gm = ring_e++, it_gs = ring_e++;
          if (gi >= 1) ++ring_e;                   // fc_local2 image
          const int Ga = gi == 0 ? LIN_G1 : LIN_LAYER0 + 4 * (gi - 1) + 0, off_ga = gi == 0 ? 256 : 0;
          const int Gb = gi == 0 ? LIN_G2 : LIN_LAYER0 + 4 * (gi - 1) + 1, off_gb = gi == 0 ? 384 : 128;
          if (gi >= 1) mbar_wait(&s.spk_full, c_spk++ & 1);       // unit gi's pack + bias slice (unit 0: waited at eval start)
          ebar();                                  // St (and the previous unit's Sg) complete
          PROF_T(7);
          // ---- fc_g1 / fc_global1:  g1[j][o] = lrelu(W_mean . mean + W_sum . sum (+ W_gg . g) + bias), this warp: 16 outputs
          // (epic.py:180-182, :375-377)
          {
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
#pragma unroll
            for (int m = 0; m < 2; ++m) {          // mean image, then sum image
              const uint32_t itw = m ? it_gs : it_gm;
              mbar_wait(&s.full[itw % TC_NSLOT], (itw / TC_NSLOT) & 1);
              const uint32_t img = smem_u32(s.w[itw % TC_NSLOT]) + img_row;
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) {
                const uint32_t kc = (uint32_t)(ks * 2 + bkh);      // 16-byte unit (8 k values) inside the 128-k image row
                uint32_t fa[4], fb[4];
                ldsm_x4(st_lane + (uint32_t)((m * 8 + ks) * 32), fa);
                ldsm_x4(img + (kc >> 3) * 16384u + (((kc & 7u) ^ img_swz) << 4), fb);
                hmma_bf16(acc[0], fa[0], fa[1], fa[2], fa[3], fb[0], fb[1]);
                hmma_bf16(acc[1], fa[0], fa[1], fa[2], fa[3], fb[2], fb[3]);
              }
            }
            if (gi >= 1) {                         // the 16 global-vector columns
              uint32_t fa[4], fb[4];
              ldsm_x4(sg_lane, fa);
              ldsm_x4(wgg_lane, fb);
              hmma_bf16(acc[0], fa[0], fa[1], fa[2], fa[3], fb[0], fb[1]);
              hmma_bf16(acc[1], fa[0], fa[1], fa[2], fa[3], fb[2], fb[3]);
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const int o = w8 * 16 + nt * 8 + 2 * t4;
              const float b0 = unit_bias(Ga, off_ga, jg_lo, o), b1 = unit_bias(Ga, off_ga, jg_lo, o + 1);
              const float c0 = SIMPLE ? b0 : unit_bias(Ga, off_ga, jg_hi, o), c1 = SIMPLE ? b1 : unit_bias(Ga, off_ga, jg_hi, o + 1);
              sts32_if(smem_u32(&s.g1[g8][o]), pack_bf16x2(lrelu_tc(acc[nt][0] + b0, p.slope), lrelu_tc(acc[nt][1] + b1, p.slope)), jet_lo);
              sts32_if(smem_u32(&s.g1[g8 + 8][o]), pack_bf16x2(lrelu_tc(acc[nt][2] + c0, p.slope), lrelu_tc(acc[nt][3] + c1, p.slope)), jet_hi);
            }
          }
          ebar();                                  // g1 complete; St and the two ring images are dead
          PROF_T(8);
          if (et == 0) { mbar_arrive(&s.empty[it_gm % TC_NSLOT]); mbar_arrive(&s.empty[it_gs % TC_NSLOT]); }
          // ---- fc_g2 / fc_global2 (+ residual for the layers), every warp redundantly: the new global vectors stay in registers
          {
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              uint32_t fa[4], fb[4];
              ldsm_x4(g1_lane + (uint32_t)(ks * 32), fa);
              ldsm_x4(wg2_lane + (uint32_t)(ks * 32), fb);
              hmma_bf16(acc[0], fa[0], fa[1], fa[2], fa[3], fb[0], fb[1]);
              hmma_bf16(acc[1], fa[0], fa[1], fa[2], fa[3], fb[2], fb[3]);
            }
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int z = nt * 8 + 2 * t4 + e, zc = z < Z ? z : Z - 1;
                const float blo = unit_bias(Gb, off_gb, jg_lo, zc), bhi = SIMPLE ? blo : unit_bias(Gb, off_gb, jg_hi, zc);
                float vlo = acc[nt][e] + blo, vhi = acc[nt][2 + e] + bhi;
                if (gi >= 1) { vlo += greg[nt][e]; vhi += greg[nt][2 + e]; }
                greg[nt][e] = z < Z ? lrelu_tc(vlo, p.slope) : 0.f;
                greg[nt][2 + e] = z < Z ? lrelu_tc(vhi, p.slope) : 0.f;
              }
            }
          }
          const uint32_t ga0 = pack_bf16x2(greg[0][0], greg[0][1]), ga1 = pack_bf16x2(greg[0][2], greg[0][3]);
          const uint32_t ga2 = pack_bf16x2(greg[1][0], greg[1][1]), ga3 = pack_bf16x2(greg[1][2], greg[1][3]);
          if (w8 == 0) {                           // bf16 copy for the next unit's fc_global1 (read after that unit's barrier)
            sts32_if(smem_u32(&s.Sg[g8][2 * t4]), ga0, true);
            sts32_if(smem_u32(&s.Sg[g8 + 8][2 * t4]), ga1, true);
            sts32_if(smem_u32(&s.Sg[g8][8 + 2 * t4]), ga2, true);
            sts32_if(smem_u32(&s.Sg[g8 + 8][8 + 2 * t4]), ga3, true);
          }
          PROF_T(9);
          if (gi >= 1) {
            // ---- re-injection: per-jet biases of fc_local1 (b + W_t . t + W_glob . g) and fc_local2, this warp's 16 outputs
            const int La = LIN_LAYER0 + 4 * (gi - 1) + 2, Lb = LIN_LAYER0 + 4 * (gi - 1) + 3;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
            uint32_t fb[4];
            ldsm_x4(wgl_lane, fb);
            hmma_bf16(acc[0], ga0, ga1, ga2, ga3, fb[0], fb[1]);
            hmma_bf16(acc[1], ga0, ga1, ga2, ga3, fb[2], fb[3]);
#pragma unroll
            for (int nt = 0; nt < 2; ++nt) {
              const int o = w8 * 16 + nt * 8 + 2 * t4;
              const float a0 = unit_bias(La, 128 + ZP, jg_lo, o), a1 = unit_bias(La, 128 + ZP, jg_lo, o + 1);
              const float d0 = unit_bias(Lb, 256 + ZP, jg_lo, o), d1 = unit_bias(Lb, 256 + ZP, jg_lo, o + 1);
              sts64_if(smem_u32(&s.sb.bl1[g8][o]), acc[nt][0] + a0, acc[nt][1] + a1, jet_lo);
              sts64_if(smem_u32(&s.bl2[g8][o]), d0, d1, jet_lo);
              const float a2 = SIMPLE ? a0 : unit_bias(La, 128 + ZP, jg_hi, o), a3 = SIMPLE ? a1 : unit_bias(La, 128 + ZP, jg_hi, o + 1);
              const float d2 = SIMPLE ? d0 : unit_bias(Lb, 256 + ZP, jg_hi, o), d3 = SIMPLE ? d1 : unit_bias(Lb, 256 + ZP, jg_hi, o + 1);
              sts64_if(smem_u32(&s.sb.bl1[g8 + 8][o]), acc[nt][2] + a2, acc[nt][3] + a3, jet_hi);
              sts64_if(smem_u32(&s.bl2[g8 + 8][o]), d2, d3, jet_hi);
            }
          }
          mbar_arrive(&s.spk_empty);               // pack + bias slice of this unit are dead: the producer may refill
          PROF_T(10);
          if (gi >= 1) {
            ebar();                                // bl1 / bl2 complete
            PROF_T(11);
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
#pragma unroll
        for (int f = 0; f < FP; ++f)
          if (f < F) vout[f] = valid ? lrelu_tc(vout[f] + b3[f], p.slope) : 0.f;
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
// (fc_local1 first: its MMAs start as soon as the new h is stored; the fc_global1 images are read a little later by the
// epilogue warps through ldmatrix -- same K-major SWIZZLE_128B layout, [o][k] rows are exactly mma.sync B fragments)
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

// small-weight pack of unit u (SpkPack): W_gg[o][z] = fc_global1[o][2H + z], W_glob[o][z] = fc_local1[o][H + z],
// W_g2[z][k] = fc_global2[z][k] (fc_g2 for the stem); bf16, zero padding (z >= Z, pad columns)
struct SpkSrc { const float* gg; int gg_ldo; const float* gl; int gl_ldo; const float* g2; int g2_ldo; int Z; };

__global__ void pack_spk_kernel(const SpkSrc* __restrict__ src, SpkPack* __restrict__ out) {
  const SpkSrc S = src[blockIdx.x];
  SpkPack& o = out[blockIdx.x];
  for (int i = threadIdx.x; i < TCH * TC_GG_LD; i += blockDim.x) {
    const int c = i / TC_GG_LD, z = i - c * TC_GG_LD;           // source: k-major fp32 copies, [z][c]
    o.gg[c][z] = __float2bfloat16((S.gg && z < S.Z) ? S.gg[(size_t)z * S.gg_ldo + c] : 0.f);
    o.gl[c][z] = __float2bfloat16((S.gl && z < S.Z) ? S.gl[(size_t)z * S.gl_ldo + c] : 0.f);
  }
  for (int i = threadIdx.x; i < TC_ZMAX * TC_G2_LD; i += blockDim.x) {
    const int z = i / TC_G2_LD, c = i - z * TC_G2_LD;
    o.g2[z][c] = __float2bfloat16((z < S.Z && c < TCH) ? S.g2[(size_t)c * S.g2_ldo + z] : 0.f);
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
    src[it++] = mk(LIN_LAYER0 + 4 * l + 2, 0);       // fc_local1, particle columns
    src[it++] = mk(LIN_LAYER0 + 4 * l + 0, 0);       // fc_global1: (mean, sum, global) order (epic.py:164-171)
    src[it++] = mk(LIN_LAYER0 + 4 * l + 0, TCH);
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
    const int n_ll = 60;
    if (!dprof) PFM_CUDA_CHECK(cudaMalloc(&dprof, sizeof(long long) * n_ll));
    PFM_CUDA_CHECK(cudaMemsetAsync(dprof, 0, sizeof(long long) * n_ll, st));
    p.prof = dprof;
    int rc = kmax <= 4 ? (simple ? launch_tc<4, true, true>(p, grid, st) : launch_tc<4, true, false>(p, grid, st))
                       : launch_tc<8, true, false>(p, grid, st);
    if (rc != PFM_OK) return rc;
    static long long hp[60];
    PFM_CUDA_CHECK(cudaMemcpyAsync(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));
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
