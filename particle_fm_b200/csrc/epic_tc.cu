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
// all in the "transposed" orientation D^T[feature][jet] (M = 16 features, N = 8 jets per fragment: a typical group has 3-4
// jets, so one N tile; groups with 9..16 jets take a second one):
//   pooling      S^T[c][jet] = sum_rows h[row][c] P[jet][row]: warp w owns columns 16w..16w+15 and reduces over all rows of
//                the group (A = ldmatrix.trans of the swizzled h tiles, B = the 0/1 indicator matrix P) -> complete sums in
//                one warp, deterministic, scaled (mean, sum) and written as the bf16 operand St of fc_global1
//   fc_global1   [16 outputs of warp w x 272] . [272 x jets]: A fragments come straight out of the ring images
//   fc_global2   [16 z x 128] . [128 x jets], every warp redundantly (the result stays in registers: no exchange)
//   re-injection [16 outputs x 16 z] . [16 z x jets]: B fragment = the new global vector, transposed with movmatrix
// (Round 1 ran pooling and fc_global1 as N = 16 tcgen05 MMAs and the rest as CUDA-core GEMVs: four dependent hops through
// mbarriers / TMEM per layer, ~4 k of the ~9 k cycles of a layer.)
// Warp roles: warps 0 and 10 = weight producers (one thread sustains only one cp.async.bulk per ~690 cycles whatever its size
// -- tools/bulk_probe.cu -- so the ring is fed by two warps, even / odd items), warp 1 = MMA issuer (one lane),
// warps 2-5 / 6-9 = epilogue warpgroups of tile A / tile B (TMEM lane quadrant = warp % 4).
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
static constexpr int TC_THREADS = 352;
static constexpr int TC_NSLOT = 3;
static constexpr uint32_t TC_MAT = 32768;   // one 128x128 bf16 weight image

template <int FP>
struct TcSmem {
  alignas(1024) uint8_t h[2][TC_MAT];          // bf16 h tiles
  alignas(1024) uint8_t w[TC_NSLOT][TC_MAT];   // weight ring
  alignas(16) __nv_bfloat16 P[TC_J][TC_ST_LD]; // 0/1 indicator P[jet][row] of the group (B operand of the pooling)
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
  alignas(16) float sbias[2][TC_SBIAS];        // time-bias slice of the current / next unit (4 consecutive linears), cp.async by the epilogue warps
  float inv_n[TC_J];
  int boff[128];                               // bias-table offset of every linear (copied once: no global descriptor loads in the loop)
  int jrow0[TC_J + 1];
  int jid[TC_J];                               // batch index of the group's jets (the plan bin-packs jets out of order)
  int group;
  uint32_t tmem_base;
  uint64_t full[TC_NSLOT], empty[TC_NSLOT];
  uint64_t hready[2][4], accU_full[2], u_ready[2][2], accH_full[2];   // hready[tile][32-column chunk of h], u_ready[tile][half of the 128 u columns]
  uint64_t spk_full, spk_empty;
  uint32_t issued[2];                          // ring positions issued so far by the even / odd producer warp (+1)
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
      for (int c = 0; c < 4; ++c) mbar_init(&s.hready[t][c], 128);
      mbar_init(&s.accU_full[t], 1); mbar_init(&s.u_ready[t][0], 128); mbar_init(&s.u_ready[t][1], 128);
      mbar_init(&s.accH_full[t], 1);
    }
    mbar_init(&s.spk_full, 1); mbar_init(&s.spk_empty, 256);
    s.issued[0] = 0; s.issued[1] = 0;
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
  uint32_t c_hready[2] = {0, 0}, c_uready[2] = {0, 0};      // MMA side (one count per tile: its four chunk barriers advance together)
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

    if (warp == 0 || warp == 10) {
      // ================================ weight producers (whole warp, elected lane issues) ================================
      // ring order per evaluation:  stem  fc_l2 | fc_g1 mean | fc_g1 sum     layer l  fc_local1 | fc_global1 mean | sum | fc_local2
      // warp 0 issues the even ring positions, warp 10 the odd ones and the small-weight packs
      const uint32_t mine = warp == 0 ? 0u : 1u;
      for (int ev = 0; ev < p.n_evals; ++ev) {
        for (int it = 0; it < p.n_items; ++it, ++ring_it) {
          // the small-weight pack of unit u travels just before the unit's fc_global1 images (after the layer's fc_local1 image,
          // which is needed much earlier: fc_local1 runs under the residual epilogue, before the previous unit's chain has ended)
          const int u = it == 1 ? 0 : ((it >= 4 && ((it - 4) & 3) == 0) ? 1 + ((it - 4) >> 2) : -1);
          if (u >= 0 && mine == 1u) {
            mbar_wait(&s.spk_empty, (spk_it & 1) ^ 1);
            ++spk_it;
            if (elect_one()) {
              mbar_arrive_expect_tx(&s.spk_full, TC_SPK);
              bulk_copy_g2s(&s.spk, p.spk + (size_t)u * TC_SPK, TC_SPK, &s.spk_full);
            }
            __syncwarp();
          }
          if ((ring_it & 1u) != mine) continue;
          const uint32_t slot = ring_it % TC_NSLOT, round = ring_it / TC_NSLOT;
          // The parity test below is only valid while this warp is at most one phase ahead of the slot's barrier.  The slot's
          // previous user (ring position - 3) belongs to the OTHER producer: wait until it has been issued (its own wait
          // proved the phase before that complete).
          if (ring_it >= 3u) {
            uint32_t seen;
            do {
              asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(&s.issued[mine ^ 1u])) : "memory");
            } while (seen < ring_it - 2u);
          }
          mbar_wait(&s.empty[slot], (round & 1) ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&s.full[slot], TC_MAT);
            bulk_copy_g2s(s.w[slot], p.wimg + (size_t)it * TC_MAT, TC_MAT, &s.full[slot]);
            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(&s.issued[mine])), "r"(ring_it + 1u) : "memory");
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
        // ---- stem fc_l2: accH[t] (holds h1) += h1 . W_l2^T.  Every version of h is handed over in four 32-column chunks
        // (two K steps each); the stem arrives on all four at once (its TMEM stores must be complete before the first MMA
        // accumulates onto them), the layers chunk by chunk so that fc_local1 runs WHILE the residual epilogue still works.
        const uint32_t it_l2 = ring_it++;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            PROF_T(0);
            mbar_wait(&s.hready[t][c], c_hready[t] & 1);
            tc_fence_after();
            PROF_T(1);
            if (t == 0 && c == 0) wait_full(it_l2);
            PROF_T(2);
            const uint64_t ad = t ? hB : hA, wd = wslot(it_l2);
            if (elect_one()) {
              mma_ss(t ? accH1 : accH0, ad + kstep16(2 * c), wd + kstep16(2 * c), idesc, 1u);
              mma_ss(t ? accH1 : accH0, ad + kstep16(2 * c + 1), wd + kstep16(2 * c + 1), idesc, 1u);
            }
            __syncwarp();
            if (c == 3) { commit_to(&s.accH_full[t]); ++c_hready[t]; }
          }
        }
        commit_to(&s.empty[it_l2 % TC_NSLOT]);
        ring_it += 2;                                  // fc_g1 images: read (and released) by the epilogue warps
        for (int l = 0; l < L; ++l) {
          const uint32_t it_w1 = ring_it++;
          ring_it += 2;                                // fc_global1 images
          const uint32_t it_w2 = ring_it++;
          // ---- fc_local1: accU[t] = h_l . W1^T, two K steps per 32-column chunk of h as the residual epilogue stores them:
          // the GEMM runs under the (ALU-bound) epilogue, so the tensor pipe is free for the mma.sync of the per-jet chain
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              PROF_T(0);
              mbar_wait(&s.hready[t][c], c_hready[t] & 1);
              tc_fence_after();
              PROF_T(3);
              if (t == 0 && c == 0) wait_full(it_w1);
              PROF_T(4);
              const uint64_t ad = t ? hB : hA, wd = wslot(it_w1);
              if (elect_one()) {
                mma_ss(t ? accU1 : accU0, ad + kstep16(2 * c), wd + kstep16(2 * c), idesc, c ? 1u : 0u);
                mma_ss(t ? accU1 : accU0, ad + kstep16(2 * c + 1), wd + kstep16(2 * c + 1), idesc, 1u);
              }
              __syncwarp();
              if (c == 3) { commit_to(&s.accU_full[t]); ++c_hready[t]; }
            }
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
      const int w8 = wg * 4 + q;                 // this warp's slice of the per-jet chain: features 16 w8 .. 16 w8 + 15
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
      // arrays are zeroed once so that the fragment columns of jets the group does not have hold finite values too.
      for (int i = et; i < 2 * (int)TC_MAT / 16; i += 256) reinterpret_cast<uint4*>(&s.h[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.P) / 16; i += 256) reinterpret_cast<uint4*>(&s.P[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.sb) / 16; i += 256) reinterpret_cast<uint4*>(&s.sb)[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.g1) / 16; i += 256) reinterpret_cast<uint4*>(&s.g1[0][0])[i] = make_uint4(0, 0, 0, 0);
      for (int i = et; i < (int)sizeof(s.Sg) / 16; i += 256) reinterpret_cast<uint4*>(&s.Sg[0][0])[i] = make_uint4(0, 0, 0, 0);
      ebar();
      const int R = s.jrow0[nj];
      const bool valid = row < R;
      int myjet = 0;
      for (int j = 1; j < nj; ++j) myjet += (row >= s.jrow0[j]) ? 1 : 0;
      if (!valid) myjet = 0;
      if (valid) s.P[myjet][row] = __float2bfloat16(1.0f);
      const int jg = s.jid[myjet];
      const int NB = (nj + 7) >> 3;              // 8-jet fragment column blocks of the per-jet chain (1 or 2)
      const int pool_ks = (R + 15) >> 4;         // 16-row K steps of the pooling that hold particles
      // bias of a linear of the CURRENT unit: staged slice of the time table (+ per-jet cond table), or the
      // slow direct path when every jet has its own time (training-style forward)
      // time-bias slices: unit k of the running count lives in sbias[k & 1]; the epilogue threads fetch the next unit's slice
      // with cp.async while the current unit runs
      const bool stage_bias = !p.tbias_per_jet;
      const int bias_n16 = (p.bias_chunk_floats * 4 + 15) >> 4;
      uint32_t c_unit = 0;                       // units completed so far by this CTA (selects the sbias buffer)
      auto fetch_bias = [&](int ev_, int u_, uint32_t buf) {
        if (stage_bias && et < bias_n16) {
          const float* src = p.tbias + (size_t)ev_ * p.bstride + (u_ == 0 ? p.boff_stem : p.boff_layer0 + (u_ - 1) * p.boff_layer_stride) + et * 4;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&s.sbias[buf][et * 4])), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      const float* sbias_cur = &s.sbias[0][0];
      auto unit_bias = [&](int lin_idx, int voff, int jet_global, int o) -> float {
        if (SIMPLE) return sbias_cur[voff + o];
        float b = p.tbias_per_jet ? p.tbias[(size_t)jet_global * p.bstride + s.boff[lin_idx] + o] : sbias_cur[voff + o];
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
      // ---- per-lane ldmatrix row addresses (fragment layouts: tc_ptx.cuh).  The chain runs transposed: M = 16 features.
      //  A operands [m][k] row-major (weights): matrices (m lo, k lo) (m hi, k lo) (m lo, k hi) (m hi, k hi)
      //  B operands [n = jet][k] row-major, one 8-jet block, TWO K steps per x4: matrices (k0 lo) (k0 hi) (k1 lo) (k1 hi)
      const int l7 = lane & 7, mi = lane >> 3;
      const int a_m = l7 + (mi & 1) * 8, a_kh = mi >> 1;
      const int o_a = w8 * 16 + a_m;                                  // this lane's weight row in fc_global1 / the re-injection
      const uint32_t img_row = (uint32_t)((o_a >> 3) * 1024 + (o_a & 7) * 128);     // row o_a of a K-major SW128 image
      const uint32_t img_swz = (uint32_t)(o_a & 7);
      const uint32_t wgg_lane = smem_u32(&s.spk.gg[0][0]) + (uint32_t)(o_a * TC_GG_LD * 2 + a_kh * 16);
      const uint32_t wgl_lane = smem_u32(&s.spk.gl[0][0]) + (uint32_t)(o_a * TC_GG_LD * 2 + a_kh * 16);
      const uint32_t wg2_lane = smem_u32(&s.spk.g2[0][0]) + (uint32_t)(a_m * TC_G2_LD * 2 + a_kh * 16);
      const uint32_t st_lane = smem_u32(&s.sb.St[0][0]) + (uint32_t)(l7 * TC_ST_LD * 2 + mi * 16);
      const uint32_t p_lane = smem_u32(&s.P[0][0]) + (uint32_t)(l7 * TC_ST_LD * 2 + mi * 16);
      const uint32_t g1_lane = smem_u32(&s.g1[0][0]) + (uint32_t)(l7 * TC_G2_LD * 2 + mi * 16);
      const uint32_t sg_lane = smem_u32(&s.Sg[0][0]) + (uint32_t)(l7 * TC_GG_LD * 2 + (mi & 1) * 16);
      // pooling A operand = h^T through ldmatrix.trans: stored 8x8 blocks are (8 rows) x (8 columns); matrices
      // (c lo, rows lo) (c hi, rows lo) (c lo, rows hi) (c hi, rows hi) of the warp's 16 columns and a 16-row K step
      const uint32_t pool_cu = (uint32_t)(2 * w8 + (mi & 1));         // 8-column unit (16 bytes) inside the 128-byte row
      const uint32_t pool_lane = (pool_cu >> 3) * 16384u + (uint32_t)((mi >> 1) * 1024 + l7 * 128) + (((pool_cu & 7u) ^ (uint32_t)l7) << 4);
      const uint32_t h_base = smem_u32(&s.h[0][0]);

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
      // one 32-column chunk of the residual update: h = lrelu(acc + bias) -> TMEM fp32 (in place) + shared bf16, then the
      // chunk (two K steps of fc_local1) is handed to the MMA warp
      auto epi_h_chunk = [&](uint32_t (&v)[32], int c, bool last) {
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 b = lds128(bl2_addr + (uint32_t)(c * 128 + i4 * 16));
          bias_lrelu2(v[i4 * 4 + 0], v[i4 * 4 + 1], b.x, b.y, slope2);
          bias_lrelu2(v[i4 * 4 + 2], v[i4 * 4 + 3], b.z, b.w, slope2);
        }
        // hand chunk c-1 over now: its shared-memory stores were issued a whole chunk of ALU work ago, so the proxy fence
        // does not wait (a fence right behind the stores costs ~200 cycles per chunk)
        if (!last && c > 0) { fence_proxy_async(); mbar_arrive(&s.hready[wg][c - 1]); }
        tmem_st32(accH + c * 32, v);
        if (!last) {
          store_h_bf16(v, c, vpred);
          if (c == 3) { fence_proxy_async(); mbar_arrive(&s.hready[wg][3]); }
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

      // global vectors of the group, fp32, in the D^T fragment layout: greg[jet block][z g8: jets 2t4, 2t4+1 | z g8+8: same]
      float greg[2][4];
      fetch_bias(0, 0, 0);
      for (int ev = 0; ev < p.n_evals; ++ev) {
        // ---------------- unit 0 pack (stem biases + fc_g2) has landed; per-jet stem biases ----------------
        PROF_T(0);
        mbar_wait(&s.spk_full, c_spk++ & 1);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        ebar();                                    // unit 0's bias slice (every thread waited for its own pieces) is visible
        sbias_cur = &s.sbias[c_unit & 1][0];
        fetch_bias(ev, 1, (c_unit + 1) & 1);
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
        tmem_wait_st();                            // fc_l2 accumulates onto h1 in TMEM: all of it must be there first
        fence_proxy_async();
        tc_fence_before();
#pragma unroll
        for (int c = 0; c < 4; ++c) mbar_arrive(&s.hready[wg][c]);
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
            if (last) { tmem_wait_st(); break; }
            ebar();                               // the whole new h is in shared memory; bl1 of the previous unit is dead
            PROF_T(5);
            // ---- masked pooling, this warp's 16 columns over all rows of the group:  S^T[c][jet] = sum_rows h[row][c] P[jet][row]
            // -> pre-scaled bf16 B operand of fc_global1:  St[jet][0:128) = S/n (mean), St[jet][128:256) = S*s (sum)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              if (nb < NB) {
                // batches of 4 K steps (64 rows), 4 independent accumulators; the fragments of batch b+1 are requested before
                // the mma.sync of batch b are issued (the volatile asm keeps program order: no compiler pipelining)
                float pa[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { pa[i][0] = 0.f; pa[i][1] = 0.f; pa[i][2] = 0.f; pa[i][3] = 0.f; }
                uint32_t fA[2][4][4], fB[2][2][4];
                const int nbatch = (pool_ks + 3) >> 2;
                auto load_batch = [&](int bb, uint32_t (&A)[4][4], uint32_t (&Bf)[2][4]) {
                  const uint32_t hk = h_base + pool_lane + (uint32_t)((bb >> 1) * (int)TC_MAT + (bb & 1) * 8192);   // 64 rows = 8 KB inside a tile's 64-column block
#pragma unroll
                  for (int i = 0; i < 4; ++i) ldsm_x4_trans(hk + (uint32_t)(i * 2048), A[i]);
                  ldsm_x4(p_lane + (uint32_t)(nb * 8 * TC_ST_LD * 2 + bb * 128), Bf[0]);
                  ldsm_x4(p_lane + (uint32_t)(nb * 8 * TC_ST_LD * 2 + bb * 128 + 64), Bf[1]);
                };
                load_batch(0, fA[0], fB[0]);
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                  if (bb < nbatch) {
                    if (bb + 1 < nbatch) load_batch(bb + 1, fA[(bb + 1) & 1], fB[(bb + 1) & 1]);
                    uint32_t (&A)[4][4] = fA[bb & 1];
                    uint32_t (&Bf)[2][4] = fB[bb & 1];
                    hmma_bf16(pa[0], A[0][0], A[0][1], A[0][2], A[0][3], Bf[0][0], Bf[0][1]);
                    hmma_bf16(pa[1], A[1][0], A[1][1], A[1][2], A[1][3], Bf[0][2], Bf[0][3]);
                    hmma_bf16(pa[2], A[2][0], A[2][1], A[2][2], A[2][3], Bf[1][0], Bf[1][1]);
                    hmma_bf16(pa[3], A[3][0], A[3][1], A[3][2], A[3][3], Bf[1][2], Bf[1][3]);
                  }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = nb * 8 + 2 * t4 + e;
                  const float inv = s.inv_n[j];
                  const float slo = (pa[0][e] + pa[1][e]) + (pa[2][e] + pa[3][e]);                   // column 16 w8 + g8
                  const float shi = (pa[0][2 + e] + pa[1][2 + e]) + (pa[2][2 + e] + pa[3][2 + e]);   // column 16 w8 + g8 + 8
                  const uint32_t a = smem_u32(&s.sb.St[j][w8 * 16 + g8]);
                  sts16_if(a, __float2bfloat16(slo * inv), j < nj);
                  sts16_if(a + 16u, __float2bfloat16(shi * inv), j < nj);
                  sts16_if(a + 2u * TCH, __float2bfloat16(slo * p.sum_scale), j < nj);
                  sts16_if(a + 2u * TCH + 16u, __float2bfloat16(shi * p.sum_scale), j < nj);
                }
              }
            }
            PROF_T(6);
          }
          // ======== global phase gi: 0 = stem (fc_g1, fc_g2), gi >= 1 = EPiC layer gi-1 (fc_global1/2) ========
          if (gi >= 1) ++ring_e;                   // fc_local1 image
          const uint32_t it_gm = ring_e++, it_gs = ring_e++;
          if (gi >= 1) ++ring_e;                   // fc_local2 image
          const int Ga = gi == 0 ? LIN_G1 : LIN_LAYER0 + 4 * (gi - 1) + 0, off_ga = gi == 0 ? 256 : 0;
          const int Gb = gi == 0 ? LIN_G2 : LIN_LAYER0 + 4 * (gi - 1) + 1, off_gb = gi == 0 ? 384 : 128;
          if (gi >= 1) {                           // unit gi's pack + bias slice (unit 0: waited at eval start)
            mbar_wait(&s.spk_full, c_spk++ & 1);
            asm volatile("cp.async.wait_group 0;" ::: "memory");
          }
          ebar();                                  // St (and the previous unit's Sg), this unit's bias slice complete
          if (gi >= 1) {
            sbias_cur = &s.sbias[c_unit & 1][0];
            if (gi < L) fetch_bias(ev, gi + 1, (c_unit + 1) & 1);
            else if (ev + 1 < p.n_evals) fetch_bias(ev + 1, 0, (c_unit + 1) & 1);
          }
          PROF_T(7);
          // ---- fc_g1 / fc_global1, this warp's 16 outputs:  g1[j][o] = lrelu(W_mean . mean + W_sum . sum (+ W_gg . g) + bias)
          // (epic.py:180-182, :375-377)
          {
            mbar_wait(&s.full[it_gm % TC_NSLOT], (it_gm / TC_NSLOT) & 1);
            mbar_wait(&s.full[it_gs % TC_NSLOT], (it_gs / TC_NSLOT) & 1);
            const uint32_t img_m = smem_u32(s.w[it_gm % TC_NSLOT]) + img_row, img_s = smem_u32(s.w[it_gs % TC_NSLOT]) + img_row;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              if (nb < NB) {
                // 16 K steps over the two images in batches of 4, 4 independent accumulators, loads one batch ahead
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
                uint32_t fA[2][4][4], fB[2][2][4];
                auto load_batch = [&](int bb, uint32_t (&A)[4][4], uint32_t (&Bf)[2][4]) {   // K steps 4 bb .. 4 bb + 3 of the 16
                  const uint32_t img = (bb >> 1) ? img_s : img_m;
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const uint32_t kc = (uint32_t)(((bb & 1) * 4 + i) * 2 + a_kh);      // 16-byte unit (8 k values) inside the image row
                    ldsm_x4(img + (kc >> 3) * 16384u + (((kc & 7u) ^ img_swz) << 4), A[i]);
                  }
                  ldsm_x4(st_lane + (uint32_t)(nb * 8 * TC_ST_LD * 2 + bb * 128), Bf[0]);
                  ldsm_x4(st_lane + (uint32_t)(nb * 8 * TC_ST_LD * 2 + bb * 128 + 64), Bf[1]);
                };
                load_batch(0, fA[0], fB[0]);
                uint32_t fga[4], fgb[2];
                if (gi >= 1) {                     // the 16 global-vector columns (17th K step)
                  ldsm_x4(wgg_lane, fga);
                  ldsm_x2(sg_lane + (uint32_t)(nb * 8 * TC_GG_LD * 2), fgb);
                }
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                  if (bb < 3) load_batch(bb + 1, fA[(bb + 1) & 1], fB[(bb + 1) & 1]);
                  uint32_t (&A)[4][4] = fA[bb & 1];
                  uint32_t (&Bf)[2][4] = fB[bb & 1];
                  hmma_bf16(acc[0], A[0][0], A[0][1], A[0][2], A[0][3], Bf[0][0], Bf[0][1]);
                  hmma_bf16(acc[1], A[1][0], A[1][1], A[1][2], A[1][3], Bf[0][2], Bf[0][3]);
                  hmma_bf16(acc[2], A[2][0], A[2][1], A[2][2], A[2][3], Bf[1][0], Bf[1][1]);
                  hmma_bf16(acc[3], A[3][0], A[3][1], A[3][2], A[3][3], Bf[1][2], Bf[1][3]);
                }
                if (gi >= 1) hmma_bf16(acc[0], fga[0], fga[1], fga[2], fga[3], fgb[0], fgb[1]);
                const int o = w8 * 16 + g8;        // D^T rows of this lane: o, o + 8; columns: jets 2 t4, 2 t4 + 1 of the block
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = nb * 8 + 2 * t4 + e;
                  const int jgl = SIMPLE ? 0 : s.jid[j];
                  const float v0 = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]) + unit_bias(Ga, off_ga, jgl, o);
                  const float v1 = (acc[0][2 + e] + acc[1][2 + e]) + (acc[2][2 + e] + acc[3][2 + e]) + unit_bias(Ga, off_ga, jgl, o + 8);
                  const uint32_t a = smem_u32(&s.g1[j][o]);
                  sts16_if(a, __float2bfloat16(lrelu_tc(v0, p.slope)), j < nj);
                  sts16_if(a + 16u, __float2bfloat16(lrelu_tc(v1, p.slope)), j < nj);
                }
              }
            }
          }
          ebar();                                  // g1 complete; St and the two ring images are dead
          PROF_T(8);
          if (et == 0) { mbar_arrive(&s.empty[it_gm % TC_NSLOT]); mbar_arrive(&s.empty[it_gs % TC_NSLOT]); }
          // ---- fc_g2 / fc_global2 (+ residual for the layers), every warp redundantly: the new global vectors stay in
          // registers; then the re-injection for this warp's 16 outputs: per-jet biases of fc_local1 (b + W_t . t + W_glob . g)
          // and fc_local2
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            if (nb < NB) {
              float acc[4][4];
#pragma unroll
              for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
              uint32_t fA[8][4], fB[4][4];         // all 8 K steps requested up front
#pragma unroll
              for (int ks = 0; ks < 8; ++ks) ldsm_x4(wg2_lane + (uint32_t)(ks * 32), fA[ks]);
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) ldsm_x4(g1_lane + (uint32_t)(nb * 8 * TC_G2_LD * 2 + k2 * 64), fB[k2]);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                hmma_bf16(acc[ks & 3], fA[ks][0], fA[ks][1], fA[ks][2], fA[ks][3], fB[ks >> 1][(ks & 1) * 2], fB[ks >> 1][(ks & 1) * 2 + 1]);
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int j = nb * 8 + 2 * t4 + e;
                const int jgl = SIMPLE ? 0 : s.jid[j];
                const int z0 = g8 < Z ? g8 : Z - 1, z1 = g8 + 8 < Z ? g8 + 8 : Z - 1;
                float v0 = (acc[0][e] + acc[1][e]) + (acc[2][e] + acc[3][e]) + unit_bias(Gb, off_gb, jgl, z0);
                float v1 = (acc[0][2 + e] + acc[1][2 + e]) + (acc[2][2 + e] + acc[3][2 + e]) + unit_bias(Gb, off_gb, jgl, z1);
                if (gi >= 1) { v0 += greg[nb][e]; v1 += greg[nb][2 + e]; }
                greg[nb][e] = g8 < Z ? lrelu_tc(v0, p.slope) : 0.f;
                greg[nb][2 + e] = g8 + 8 < Z ? lrelu_tc(v1, p.slope) : 0.f;
              }
              // bf16 copy, transposed to [jet g8][z 2t4, 2t4+1 | 8 + 2t4, ...]: B fragment of the re-injection, and the
              // global-vector columns of the next unit's fc_global1 operand
              const uint32_t gt0 = movmatrix_trans(pack_bf16x2(greg[nb][0], greg[nb][1]));
              const uint32_t gt1 = movmatrix_trans(pack_bf16x2(greg[nb][2], greg[nb][3]));
              if (w8 == 0) {
                sts32_if(smem_u32(&s.Sg[nb * 8 + g8][2 * t4]), gt0, true);
                sts32_if(smem_u32(&s.Sg[nb * 8 + g8][8 + 2 * t4]), gt1, true);
              }
              if (gi >= 1) {
                const int La = LIN_LAYER0 + 4 * (gi - 1) + 2, Lb = LIN_LAYER0 + 4 * (gi - 1) + 3;
                float bb[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t fa[4];
                ldsm_x4(wgl_lane, fa);
                hmma_bf16(bb, fa[0], fa[1], fa[2], fa[3], gt0, gt1);
                const int o = w8 * 16 + g8;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                  const int j = nb * 8 + 2 * t4 + e;
                  const int jgl = SIMPLE ? 0 : s.jid[j];
                  if (j < nj) {
                    s.sb.bl1[j][o] = bb[e] + unit_bias(La, 128 + ZP, jgl, o);
                    s.sb.bl1[j][o + 8] = bb[2 + e] + unit_bias(La, 128 + ZP, jgl, o + 8);
                    s.bl2[j][o] = unit_bias(Lb, 256 + ZP, jgl, o);
                    s.bl2[j][o + 8] = unit_bias(Lb, 256 + ZP, jgl, o + 8);
                  }
                }
              }
            }
          }
          mbar_arrive(&s.spk_empty);               // the pack of this unit is dead: the producer may refill
          ++c_unit;
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
