// Training forward / backward of the EPiC vector field with the per-particle GEMMs on the tensor cores (hid == 128).
//
// In training every activation has to reach HBM anyway (the backward needs it), so unlike the sampler the network is NOT
// kept resident in one CTA: it runs as a short program of kernels over the PACKED real particles of the whole batch,
//   rowlin2_tc_kernel  Y[rows,128] = epi( X[rows,128] . W^T )                 every 128 x 128 per-particle linear and its
//                      transpose (dX = dY . W), tcgen05.mma with fp32 accumulation in TMEM.  fp32 accuracy is kept with the
//                      3-term split x = hi + lo (two bf16): X W^T ~ Xh Wh^T + Xh Wl^T + Xl Wh^T (the dropped Xl Wl^T term is
//                      2^-16 relative), so loss and gradients stay inside the fp32 gates (1e-5 / 1e-4) of the CUDA-core path.
//                      Epilogue: + per-jet effective bias | + residual | + per-jet broadcast vector, leaky_relu,
//                      * leaky_relu'(saved sign bits), sign bits of the result.
//   per-jet kernels    pooling + global MLP + effective biases (forward) and their backward, one CTA per jet, CUDA cores
//                      (1% of the FLOPs), plus the K = 3 stem / N = 3 head.
// Bound: HBM.  One 128-row tile moves 64 KB in and 64 KB out per operand array against 24 MMAs (1.5 k cycles).
// The kernels fill the SAME saved-activation / gradient arrays as the fp32 CUDA-core kernels (epic_simt.cu TRAIN,
// epic_train.cu), so the weight-gradient jobs (xty_tc.cu) and the weight-norm chain rule are shared.
// Reference: autograd over EPiC_encoder.forward / EPiC_layer.forward (epic.py:304-391, 85-203) and the flow-matching
// losses (losses.py:38-77, 101-136, 308-342).
#include <cstdlib>
#include <cstring>

#include "pfm_internal.cuh"
#include "tc_ptx.cuh"

namespace pfm {

using namespace tc;

static constexpr int TT_H = 128;
static constexpr uint32_t TT_IMG = 32768;          // one bf16 128 x 128 K-major SW128 image

__device__ __forceinline__ float tt_lrelu(float v, float s) { return v > 0.f ? v : v * s; }
// Accumulator column of output feature n.  tcgen05.ld.16x256b gives thread T of a warp the columns 8j + 2(T%4) + {0,1}
// (j = 0..3) of a 32-column group; the weight images are packed with their N rows permuted inside every group so that
// those eight registers are the features 4(T%4) + {0..3} and 16 + 4(T%4) + {0..3}: two float4 per thread, and a quad of
// threads covers 64 contiguous bytes of a global row per instruction (full 32-byte sectors, half the requests of a
// float2 layout).
__host__ __device__ __forceinline__ int tt_ncol(int n) {
  const int f = n & 31;
  return (n & ~31) | (8 * (2 * (f >> 4) + ((f >> 1) & 1)) + 2 * ((f >> 2) & 3) + (f & 1));
}
__device__ __forceinline__ float tt_dlrelu(float post, float s) { return post > 0.f ? 1.f : s; }

// ---------------------------------------------------------------------------------------------
// weight images: for GEMM g (0 = fc_l2, 1 + 2l = fc_local1 main block, 2 + 2l = fc_local2)
//   [g][0] forward  B[n = o][k]     = W[o][m_off + k]     (Y = X . W^T)
//   [g][1] backward B[n = k][K = o] = W[o][m_off + k]     (dX = dY . W)
// each as a hi image followed by a lo image (W = hi + lo in bf16)
// ---------------------------------------------------------------------------------------------
struct TtImgSrc { const float* Wt; int ldo; int k0; };   // W[o][k] = Wt[(k0 + k) * ldo + o]

__global__ void tt_pack_kernel(const TtImgSrc* __restrict__ src, uint8_t* __restrict__ img) {
  const int im = blockIdx.x >> 3, part = blockIdx.x & 7;    // 8 CTAs per image
  const TtImgSrc S = src[im >> 1];
  const int transposed = im & 1;
  uint8_t* hi = img + (size_t)im * 2 * TT_IMG;
  uint8_t* lo = hi + TT_IMG;
  for (int i = part * 2048 + threadIdx.x; i < (part + 1) * 2048; i += blockDim.x) {
    const int k = i >> 7, o = i & 127;                   // consecutive threads -> consecutive o: coalesced reads
    const float w = S.Wt[(size_t)(S.k0 + k) * S.ldo + o];
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    // the GEMM's N index (o forward, k transposed) is stored at accumulator column tt_ncol(.): see rowlin2_tc_kernel's epilogue
    const uint32_t off = transposed ? sw128_offset(tt_ncol(k), o, 16384) : sw128_offset(tt_ncol(o), k, 16384);
    *reinterpret_cast<__nv_bfloat16*>(hi + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo + off) = l;
  }
}

// ---------------------------------------------------------------------------------------------
// One GEMM pass:  Y[rows,128] = epi( X[rows,128] . W^T )   (rowlin2_tc_kernel below)
// A 128-row tile of X is one contiguous 64 KB block of global memory, fetched by bulk async copies (no registers), split into
// bf16 hi / lo K-major SW128 operands and multiplied with 24 tcgen05.mma (3-term split) into one of two TMEM accumulators.
// Epilogue (in this order): + bias[jet] | + R | + bc[jet] | leaky_relu | * leaky_relu'(sign bits E) ; optionally the sign
// bits of the result are written (they are all the backward needs of a saved activation: 16 bytes per row instead of 512).
// ---------------------------------------------------------------------------------------------
struct RowLinP {
  const float* X;          // [rows][128]
  const uint8_t* Wimg;     // hi | lo
  const float* bias;       // per-jet effective bias rows, already offset to the linear's slice (nullptr: none)
  int bias_ld;
  const float* R;          // [rows][128] residual (nullptr: none)
  const float* bc;         // [B][128] per-jet vector (nullptr: none)
  int act;                 // leaky_relu on the result
  const uint32_t* E;       // [rows][4] sign bits of a saved post-activation: Y *= leaky_relu'(.) (nullptr: none)
  float* Y;                // [rows][128]
  uint32_t* sgn_out;       // [rows][4] sign bits of Y (nullptr: none)
  const int* rowjet;       // [rows]
  const int* n_total;
  float slope;
  long long* prof;         // debug (PFM_TT_PROF): per-phase cycle counters of CTA 0
};

// ---------------------------------------------------------------------------------------------
// rowlin2_tc_kernel: warp-specialised so that its three streams (operand conversion, MMA, epilogue) overlap instead of taking
// turns on the same warps (the first version of this pass did, and spent 17.8 k cycles per 128-row tile of a residual pass
// where its HBM share allows 8.4 k):
//   warps 0-7  (256 threads) "converters": fp32 half tiles (64 rows, 32 KB) arrive in a 3-slot ring by cp.async.bulk and are
//              split into the bf16 hi / lo SWIZZLE_128B operand; thread 32 refills the slot it just drained, an elected lane
//              of warp 0 issues the 24 MMAs of the tile into one of two TMEM accumulators;
//   warps 8-15 (256 threads) epilogue: tcgen05.ld.16x256b hands a QUAD of lanes 8 columns of a row (layout checked by
//              tools/tmem_ld_probe.cu); the weight images are packed with permuted N rows (tt_ncol) so that those are
//              two groups of 4 consecutive features per lane: residual loads and result stores are float4 accesses, 64
//              contiguous bytes per quad, straight from / to global memory -- no shared-memory staging, no CTA-wide
//              barriers; the operands of the next chunk (residual, sign words, row -> jet) are in flight while a chunk
//              is processed.
// Hand-offs: raw_full[3] (bulk-copy bytes), mma_done[2] (tcgen05.commit; read by the epilogue AND by the converters, whose
// operand buffer is single), acc_free[2] (8 epilogue warps).
// ---------------------------------------------------------------------------------------------
struct RowLin2Smem {
  alignas(1024) uint8_t W[2][TT_IMG];
  alignas(1024) uint8_t A[2][TT_IMG];           // hi, lo
  alignas(1024) float raw[3][64 * TT_H];        // ring of fp32 half tiles as they lie in global memory
  uint64_t mbar_w, raw_full[3], mma_done[2], acc_free[2];
  uint32_t tmem;
};

static constexpr int RL2_CONV_WARPS = 8;                    // converter warps (the 8 epilogue warps follow them)
static constexpr int RL2_CONV = 32 * RL2_CONV_WARPS;
static constexpr int RL2_THREADS = RL2_CONV + 256;
static constexpr int RL2_AHEAD = 1;                         // epilogue operands are fetched this many chunks ahead (1 or 2)

__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

// operands of one epilogue chunk (16 rows x 32 features of a warp: this thread's rows A = qr, B = qr + 8; features
// 4 (T%4) + {0..3} and 16 + 4 (T%4) + {0..3} of the group)
struct Rl2Pre { float4 r[4]; uint32_t e[2]; int jet[2]; };

// which operands a pass has: compile-time, so that each of the five passes of a training step runs straight-line epilogue code
enum : int { RL2_BIAS = 1, RL2_RES = 2, RL2_BC = 4, RL2_ACT = 8, RL2_E = 16, RL2_SGN = 32 };

template <int FL>
__global__ void __launch_bounds__(RL2_THREADS, 1) rowlin2_tc_kernel(const RowLinP p) {
  constexpr bool HAS_BIAS = (FL & RL2_BIAS) != 0, HAS_RES = (FL & RL2_RES) != 0, HAS_BC = (FL & RL2_BC) != 0, HAS_ACT = (FL & RL2_ACT) != 0,
                 HAS_E = (FL & RL2_E) != 0, HAS_SGN = (FL & RL2_SGN) != 0, HAS_PJ = HAS_BIAS || HAS_BC;
  static_assert(!(HAS_BIAS && HAS_BC), "a pass has one per-jet vector");
  extern __shared__ uint8_t smem_raw[];
  RowLin2Smem& s = *reinterpret_cast<RowLin2Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // Launched with programmatic stream serialization: everything up to griddepcontrol.wait (TMEM allocation, barrier
  // initialisation -- nothing that reads global memory) overlaps the tail of the kernel before this one in the stream.
  if (warp == 0) tmem_alloc(&s.tmem, 256);
  if (tid == 0) {
    mbar_init(&s.mbar_w, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&s.raw_full[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&s.mma_done[i], 1); mbar_init(&s.acc_free[i], 8); }
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem;
  pdl_trigger();
  pdl_wait();
  const int rows = *p.n_total;
  // Tiles of 128 rows, round-robin over the CTAs.  The tiles of the last, incomplete round (R < grid of them: with 660
  // tiles on 148 SMs every pass would end with 68 CTAs working and 80 idle) are handed out as 64-row half tiles when that
  // gives every CTA at most one: a half tile still takes a full M = 128 MMA, but a tile's time is conversion, epilogue and
  // HBM traffic, which halve.
  const int n_tiles = (rows + 127) >> 7;
  const int G = (int)gridDim.x, cta = (int)blockIdx.x;
  const int full_rounds = n_tiles / G, R = n_tiles - full_rounds * G;
  const bool split = R > 0 && 2 * R <= G;
  const int my_tiles = full_rounds + (split ? (cta < 2 * R ? 1 : 0) : (cta < R ? 1 : 0));
  if (my_tiles == 0) {                                    // uniform: nothing to do for this CTA
    if (warp == 0) tmem_dealloc(tm, 256);
    return;
  }
  const int n_half = 2 * my_tiles;
  auto tile_rows = [&](int t_local, int& row0, int& rend) {          // rows [row0, rend) of this CTA's t_local-th tile
    if (split && t_local >= full_rounds) { row0 = full_rounds * G * 128 + cta * 64; rend = row0 + 64; }
    else { row0 = (cta + t_local * G) * 128; rend = row0 + 128; }
    if (rend > rows) rend = rows;
  };

  if (warp < RL2_CONV_WARPS) {
    // ------------------------------------------------------------------ converters / MMA issue / loads
    auto load_half = [&](int hc) {                         // one thread: the valid rows of half tile hc, one bulk copy
      int row0, rend;
      tile_rows(hc >> 1, row0, rend);
      const int r0 = row0 + (hc & 1) * 64;
      int n = rend - r0;
      n = n < 0 ? 0 : (n > 64 ? 64 : n);
      const uint32_t bytes = (uint32_t)n * TT_H * 4u;
      uint64_t* bar = &s.raw_full[hc % 3];
      mbar_arrive_expect_tx(bar, bytes);
      if (bytes) bulk_copy_g2s(s.raw[hc % 3], p.X + (size_t)r0 * TT_H, bytes, bar);
    };
    if (tid == 0) {
      mbar_arrive_expect_tx(&s.mbar_w, 2 * TT_IMG);
      bulk_copy_g2s(s.W[0], p.Wimg, TT_IMG, &s.mbar_w);
      bulk_copy_g2s(s.W[1], p.Wimg + TT_IMG, TT_IMG, &s.mbar_w);
    }
    if (tid == 32) for (int hc = 0; hc < 3 && hc < n_half; ++hc) load_half(hc);
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    long long pt = 0, pc[5] = {0, 0, 0, 0, 0};
    const bool prof = p.prof && blockIdx.x == 0 && tid == 0;
#define RL2_PROF(k) do { if (prof) { const long long n_ = clock64(); pc[k] += n_ - pt; pt = n_; } } while (0)
    if (prof) pt = clock64();
    for (int t = 0; t < my_tiles; ++t) {
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int hc = 2 * t + half;
        mbar_wait(&s.raw_full[hc % 3], (uint32_t)((hc / 3) & 1));
        RL2_PROF(0);
        if (half == 0 && t >= 1) mbar_wait(&s.mma_done[(t - 1) & 1], (uint32_t)(((t - 1) >> 1) & 1));    // operand buffer free
        RL2_PROF(1);
        const float* src = s.raw[hc % 3];
        // raw fp32 [64][128] -> bf16 hi / lo, K-major SWIZZLE_128B (thread = 4 consecutive columns of a row)
#pragma unroll 8
        for (int i = 0; i < 2048 / RL2_CONV; ++i) {
          const int idx = tid + RL2_CONV * i, rl = idx >> 5, c = (idx & 31) * 4;
          const float4 v = *reinterpret_cast<const float4*>(&src[rl * TT_H + c]);
          const uint32_t off = sw128_offset(half * 64 + rl, c, 16384);
          uint2 h2, l2;
          split_bf16x4(v, h2, l2);
          *reinterpret_cast<uint2*>(s.A[0] + off) = h2;
          *reinterpret_cast<uint2*>(s.A[1] + off) = l2;
        }
        fence_proxy_async();
        RL2_PROF(2);
        named_bar_sync(1, RL2_CONV);
        if (tid == 32 && hc + 3 < n_half) load_half(hc + 3);     // the slot is drained: refill it
        RL2_PROF(3);
      }
      if (warp == 0) {
        if (t >= 2) mbar_wait(&s.acc_free[t & 1], (uint32_t)(((t >> 1) - 1) & 1));      // the epilogue of tile t - 2 has read it
        tc_fence_after();
        if (elect_one()) {
          if (t == 0) mbar_wait(&s.mbar_w, 0);
          const uint64_t ah = desc_kmajor(smem_u32(s.A[0])), al = desc_kmajor(smem_u32(s.A[1]));
          const uint64_t wh = desc_kmajor(smem_u32(s.W[0])), wl = desc_kmajor(smem_u32(s.W[1]));
          const uint32_t acc = tm + (uint32_t)((t & 1) * 128);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t d = (uint64_t)((k >> 2) * 1024 + (k & 3) * 2);
            mma_ss(acc, ah + d, wh + d, idesc, k ? 1u : 0u);
            mma_ss(acc, ah + d, wl + d, idesc, 1u);
            mma_ss(acc, al + d, wh + d, idesc, 1u);
          }
          mma_commit(&s.mma_done[t & 1]);
        }
        __syncwarp();
      }
      RL2_PROF(4);
    }
    if (prof) for (int k = 0; k < 5; ++k) p.prof[k] = pc[k];
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3, hf = (warp - RL2_CONV_WARPS) >> 2;     // TMEM lane quadrant (= warp % 4), column half
    const int qr = lane >> 2, qf = (lane & 3) * 4;          // row inside the 8-row block, first feature inside the group
    const float* pj = HAS_BIAS ? p.bias : p.bc;            // the per-jet vector of this pass (never both)
    const int pj_ld = HAS_BIAS ? p.bias_ld : TT_H;
    Rl2Pre pre[2 * RL2_AHEAD];
    // chunk c of tile t: lane half lh = c >> 1 (16 rows), feature group g = 2 hf + (c & 1) (32 features)
    auto prefetch = [&](int t_local, int c, Rl2Pre& o) {
      int row0, rend;
      tile_rows(t_local, row0, rend);
      const int g = hf * 2 + (c & 1);
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int r = row0 + q * 32 + (c >> 1) * 16 + qr + 8 * h2;
        const bool ok = t_local < my_tiles && r < rend;
        o.jet[h2] = 0; o.e[h2] = 0xffffffffu;
        if (ok) {
          if (HAS_PJ) o.jet[h2] = __ldg(p.rowjet + r);
          if (HAS_E) o.e[h2] = __ldcg(p.E + (size_t)r * 4 + g);
          if (HAS_RES) {
            const float4* src = reinterpret_cast<const float4*>(p.R + (size_t)r * TT_H + g * 32 + qf);
            o.r[2 * h2] = __ldcg(src);
            o.r[2 * h2 + 1] = __ldcg(src + 4);
          }
        }
      }
    };
    auto process = [&](int t_local, int c, const Rl2Pre& o) {
      int row0, rend;
      tile_rows(t_local, row0, rend);
      const int g = hf * 2 + (c & 1);
      uint32_t v[16];
      tmem_ld_16x256b_x4(tm + ((uint32_t)(q * 32 + (c >> 1) * 16) << 16) + (uint32_t)((t_local & 1) * 128 + g * 32), v);
      float4 b[4];
      if (HAS_PJ) {
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const float4* src = reinterpret_cast<const float4*>(pj + (size_t)o.jet[h2] * pj_ld + g * 32 + qf);
          b[2 * h2] = __ldg(src);
          b[2 * h2 + 1] = __ldg(src + 4);
        }
      }
      tmem_wait_ld();
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int r = row0 + q * 32 + (c >> 1) * 16 + qr + 8 * h2;
        const bool ok = r < rend;
        uint32_t sb = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {                      // features g*32 + 16 k + qf + {0..3}: registers 8k + 2 h2 + {0, 1, 4, 5}
          float4 a = make_float4(__uint_as_float(v[8 * k + 2 * h2]), __uint_as_float(v[8 * k + 2 * h2 + 1]),
                                 __uint_as_float(v[8 * k + 2 * h2 + 4]), __uint_as_float(v[8 * k + 2 * h2 + 5]));
          const int bit = 16 * k + qf;
          if (HAS_BIAS) { const float4 w = b[2 * h2 + k]; a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w; }
          if (HAS_RES) { const float4 w = o.r[2 * h2 + k]; a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w; }
          if (HAS_BC) { const float4 w = b[2 * h2 + k]; a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w; }
          if (HAS_ACT) { a.x = tt_lrelu(a.x, p.slope); a.y = tt_lrelu(a.y, p.slope); a.z = tt_lrelu(a.z, p.slope); a.w = tt_lrelu(a.w, p.slope); }
          if (HAS_E) {
            const uint32_t e = o.e[h2] >> bit;
            a.x *= (e & 1u) ? 1.f : p.slope; a.y *= (e & 2u) ? 1.f : p.slope; a.z *= (e & 4u) ? 1.f : p.slope; a.w *= (e & 8u) ? 1.f : p.slope;
          }
          if (ok) __stcg(reinterpret_cast<float4*>(p.Y + (size_t)r * TT_H + g * 32 + bit), a);
          if (HAS_SGN) sb |= ((a.x > 0.f ? 1u : 0u) | (a.y > 0.f ? 2u : 0u) | (a.z > 0.f ? 4u : 0u) | (a.w > 0.f ? 8u : 0u)) << bit;
        }
        if (HAS_SGN) {                                     // the quad holds the 32 features of one sign word
          sb |= __shfl_xor_sync(0xffffffffu, sb, 1);
          sb |= __shfl_xor_sync(0xffffffffu, sb, 2);
          if (ok && (lane & 3) == 0) p.sgn_out[(size_t)r * 4 + g] = sb;
        }
      }
    };
    long long pt = 0, pc[2] = {0, 0};
    const bool prof = p.prof && blockIdx.x == 0 && tid == RL2_CONV;
    if (prof) pt = clock64();
    prefetch(0, 0, pre[0]);
    if (RL2_AHEAD == 2) prefetch(0, 1, pre[1]);
    for (int t = 0; t < my_tiles; ++t) {
      mbar_wait(&s.mma_done[t & 1], (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      RL2_PROF(0);
      if (RL2_AHEAD == 2) {
        prefetch(t, 2, pre[2]);
        process(t, 0, pre[0]);
        prefetch(t, 3, pre[3]);
        process(t, 1, pre[1]);
        prefetch(t + 1, 0, pre[0]);
        process(t, 2, pre[2]);
        prefetch(t + 1, 1, pre[1]);
        process(t, 3, pre[3]);
      } else {
        prefetch(t, 1, pre[1]);
        process(t, 0, pre[0]);
        prefetch(t, 2, pre[0]);
        process(t, 1, pre[1]);
        prefetch(t, 3, pre[1]);
        process(t, 2, pre[0]);
        prefetch(t + 1, 0, pre[0]);
        process(t, 3, pre[1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_free[t & 1]);
      RL2_PROF(1);
    }
    if (prof) { p.prof[5] = pc[0]; p.prof[6] = pc[1]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------
// rowoff = exclusive prefix sum of n_real, n_total.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024) tt_rows_kernel(const int* __restrict__ n_real, int B, int* __restrict__ rowoff,
                                                       int* __restrict__ n_total) {
  __shared__ int part[1024];
  const int tid = threadIdx.x;
  const int per = (B + 1023) / 1024;
  int s = 0;
  for (int i = 0; i < per; ++i) { const int j = tid * per + i; if (j < B) s += n_real[j]; }
  part[tid] = s;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {                     // Hillis-Steele inclusive scan
    const int v = tid >= d ? part[tid - d] : 0;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  int off = part[tid] - s;
  for (int i = 0; i < per; ++i) {
    const int j = tid * per + i;
    if (j < B) { rowoff[j] = off; off += n_real[j]; }
  }
  if (tid == 1023) *n_total = part[1023];
}

// rowjet[rowoff[j] + r] = j
__global__ void tt_rowjet_kernel(const int* __restrict__ n_real, const int* __restrict__ rowoff, int B, int* __restrict__ rowjet) {
  const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= B) return;
  const int n = n_real[j], r0 = rowoff[j];
  for (int r = threadIdx.x & 31; r < n; r += 32) rowjet[r0 + r] = j;
}

// beff[j][:] = tbias[row(j)][:] + cbias[j][:]      (effective bias of every linear before the global-vector term)
__global__ void tt_beff_kernel(const float* __restrict__ tbias, int per_jet, const float* __restrict__ cbias, int bstride,
                               float* __restrict__ beff) {
  const int j = blockIdx.x;
  const float* tb = tbias + (size_t)(per_jet ? j : 0) * bstride;
  for (int i = threadIdx.x; i < bstride; i += blockDim.x)
    beff[(size_t)j * bstride + i] = tb[i] + (cbias ? cbias[(size_t)j * bstride + i] : 0.f);
}

// sign bits of saved post-activations (only needed when the forward was produced by the CUDA-core kernels)
__global__ void tt_sign_kernel(const float* __restrict__ act, size_t stage_stride, uint32_t* __restrict__ sgn, size_t sgn_stride,
                               const int* __restrict__ n_total) {
  const int rows = *n_total;
  const int stage = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * 4; i += gridDim.x * blockDim.x) {
    const float* a = act + (size_t)stage * stage_stride + (size_t)(i >> 2) * TT_H + (i & 3) * 32;
    uint32_t b = 0;
    for (int k = 0; k < 32; ++k) b |= (a[k] > 0.f ? 1u : 0u) << k;
    sgn[(size_t)stage * sgn_stride + i] = b;
  }
}

struct TtCommon {
  const Lin* lin; int n_lin, L, H, Z, F, Kx, xin_off, N, B, bstride;
  float sum_scale, slope;
  const int* n_real; const uint16_t* ridx; const int* rowoff; const int* n_total; const int* rowjet;
  float* act; float* dact; size_t stage_stride;
  uint32_t* sgn; size_t sgn_stride;
  float* yact; float* jact; int junit, jstride, LDP, Hp, Zp;
  float* dpre3; float* dbeff; float* beff; float* bc; float* dGc; float* dxs;
  float* loss_acc;
  // loss inputs
  const float* x_in; float* x_out; const float* tjet; const float* noise0; const float* noise1; int loss_kind; float sigma;
};

// stem: y (flow-matching interpolation or the given input) -> yact;  h1 = lrelu(fc_l1(y)) -> act[0] (+ sign bits).
// One warp per row, lane = 4 consecutive columns; a warp works on TWO rows at a time so that their chains of dependent
// index loads (row -> jet -> particle slot -> inputs) overlap.
__global__ void __launch_bounds__(256, 4) tt_stem_kernel(const TtCommon p) {
  const int lane = threadIdx.x & 31;
  const int rows = *p.n_total;
  const Lin L1 = p.lin[LIN_L1];
  const int wg = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
  for (int r0 = wg * 2; r0 < rows; r0 += nw * 2) {
    const bool two = r0 + 1 < rows;
    const int rr[2] = {r0, two ? r0 + 1 : r0};
    int j[2], part[2];
    float tj[2];
    float4 a[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) j[i] = p.rowjet[rr[i]];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      part[i] = p.ridx[(size_t)j[i] * p.N + (rr[i] - p.rowoff[j[i]])];
      tj[i] = p.loss_kind >= 0 ? p.tjet[j[i]] : 0.f;
      a[i] = *reinterpret_cast<const float4*>(p.beff + (size_t)j[i] * p.bstride + L1.bias_off + lane * 4);
    }
    for (int c = 0; c < p.Kx; ++c) {
      float y[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const size_t gi = ((size_t)j[i] * p.N + part[i]) * p.Kx + c;
        if (p.loss_kind >= 0) {                                  // losses.py:56, :115-116, :320
          const float x = p.x_in[gi], t = tj[i], z = p.noise0[gi];
          if (p.loss_kind == PFM_LOSS_FM_OT) y[i] = (1.f - t) * x + (p.sigma + (1.f - p.sigma) * t) * z;
          else if (p.loss_kind == PFM_LOSS_CFM) y[i] = ((1.f - t) * x + t * z) + p.sigma * p.noise1[gi];
          else y[i] = x + t * z;
        } else {
          y[i] = p.x_in[gi];
        }
      }
      const float4 w = __ldg(reinterpret_cast<const float4*>(L1.Wt + (size_t)(L1.m_off + p.xin_off + c) * L1.ldo + lane * 4));
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (lane == 0 && (i == 0 || two)) p.yact[(size_t)rr[i] * p.Kx + c] = y[i];
        a[i].x = fmaf(w.x, y[i], a[i].x); a[i].y = fmaf(w.y, y[i], a[i].y); a[i].z = fmaf(w.z, y[i], a[i].z); a[i].w = fmaf(w.w, y[i], a[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 o = a[i];
      o.x = tt_lrelu(o.x, p.slope); o.y = tt_lrelu(o.y, p.slope); o.z = tt_lrelu(o.z, p.slope); o.w = tt_lrelu(o.w, p.slope);
      uint32_t b = ((o.x > 0.f ? 1u : 0u) | (o.y > 0.f ? 2u : 0u) | (o.z > 0.f ? 4u : 0u) | (o.w > 0.f ? 8u : 0u)) << ((lane & 7) * 4);
      b |= __shfl_xor_sync(0xffffffffu, b, 1); b |= __shfl_xor_sync(0xffffffffu, b, 2); b |= __shfl_xor_sync(0xffffffffu, b, 4);
      if (i == 0 || two) {
        *reinterpret_cast<float4*>(p.act + (size_t)rr[i] * TT_H + lane * 4) = o;
        if ((lane & 7) == 0) p.sgn[(size_t)rr[i] * 4 + (lane >> 3)] = b;
      }
    }
  }
}

// Column sums over rows [r0, r0 + n) of a [rows][128] array by a group of 32 * G threads (t = 0 .. 32 G - 1): thread t sums
// the four columns 4 (t % 32) .. + 3 of rows t / 32, t / 32 + G, ... with 8 independent 16-byte loads in flight, the G partial
// rows are combined through `red` ([G][128] floats of shared memory).  The caller synchronises the CTA afterwards and reads
// the sums with tt_colsum_get.  (A jet has <= 150 rows: with G = 4 / 8 a thread issues all its loads in 1-3 batches, where one
// thread per column needed up to 10 dependent batches of HBM latency.)
template <int G>
__device__ __forceinline__ void tt_colsum_put(const float* __restrict__ a, int r0, int n, int t, float* __restrict__ red) {
  const int c4 = t & 31, rg = t >> 5;
  float4 s[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4* base = reinterpret_cast<const float4*>(a + (size_t)r0 * TT_H) + c4;
  int r = rg;
  for (; r + 7 * G < n; r += 8 * G) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldcg(base + (size_t)(r + q * G) * (TT_H / 4));
      s[q].x += v.x; s[q].y += v.y; s[q].z += v.z; s[q].w += v.w;
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {                            // tail: still independent loads
    if (r + q * G < n) {
      const float4 v = __ldcg(base + (size_t)(r + q * G) * (TT_H / 4));
      s[q].x += v.x; s[q].y += v.y; s[q].z += v.z; s[q].w += v.w;
    }
  }
  float4 o;
  o.x = ((s[0].x + s[1].x) + (s[2].x + s[3].x)) + ((s[4].x + s[5].x) + (s[6].x + s[7].x));
  o.y = ((s[0].y + s[1].y) + (s[2].y + s[3].y)) + ((s[4].y + s[5].y) + (s[6].y + s[7].y));
  o.z = ((s[0].z + s[1].z) + (s[2].z + s[3].z)) + ((s[4].z + s[5].z) + (s[6].z + s[7].z));
  o.w = ((s[0].w + s[1].w) + (s[2].w + s[3].w)) + ((s[4].w + s[5].w) + (s[6].w + s[7].w));
  *reinterpret_cast<float4*>(red + rg * TT_H + c4 * 4) = o;
}
template <int G>
__device__ __forceinline__ float tt_colsum_get(const float* __restrict__ red, int col) {
  float s = red[col];
#pragma unroll
  for (int g = 1; g < G; ++g) s += red[g * TT_H + col];
  return s;
}

// Forward of one per-jet unit (unit 0 = stem fc_g1 / fc_g2, unit l+1 = EPiC layer l): pooling of the unit's input h,
// global MLP, effective biases of the layer's two local linears (epic.py:369-380, :160-196).
// One CTA of 256 threads per jet, four CTAs per SM: a chain of short phases, each about one global round trip long.
__global__ void __launch_bounds__(256, 4) tt_jet_fwd_kernel(const TtCommon p, int unit, long long* prof) {
  long long pt = prof ? clock64() : 0, pc[6] = {0, 0, 0, 0, 0, 0};
#define JF_PROF(k) do { if (prof) { const long long n_ = clock64(); pc[k] += n_ - pt; pt = n_; } } while (0)
  __shared__ float pool[2 * TT_H + 32];
  __shared__ float half2[TT_H];
  __shared__ float g1s[TT_H];
  __shared__ float gs[32];
  __shared__ __align__(16) float red[8 * TT_H];
  const int j = blockIdx.x, tid = threadIdx.x, col = tid & 127, hp = tid >> 7;
  const int H = p.H, Z = p.Z;
  const int n = p.n_real[j], r0 = p.rowoff[j];
  const int l = unit - 1;
  const Lin Ga = p.lin[unit == 0 ? LIN_G1 : LIN_LAYER0 + 4 * l + 0];
  const Lin Gb = p.lin[unit == 0 ? LIN_G2 : LIN_LAYER0 + 4 * l + 1];
  float* ja = p.jact + (size_t)j * p.jstride + (size_t)unit * p.junit;
  const float* h = p.act + (size_t)(unit == 0 ? 1 : 1 + 2 * l) * p.stage_stride;
  {
    JF_PROF(0);
    tt_colsum_put<8>(h, r0, n, tid, red);
    __syncthreads();
    JF_PROF(1);
    if (!hp) {
      const float sum = tt_colsum_get<8>(red, col);
      const float mean = sum / (float)n, ssum = sum * p.sum_scale;
      if (unit == 0) { pool[col] = ssum; pool[H + col] = mean; }       // (sum, mean) in the stem, epic.py:373
      else { pool[col] = mean; pool[H + col] = ssum; }                 // (mean, sum, global) in the layers, :164-171
      ja[col] = pool[col]; ja[H + col] = pool[H + col];
      if (unit > 0 && col < Z) {
        const float g = p.jact[(size_t)j * p.jstride + (size_t)(unit - 1) * p.junit + p.LDP + p.Hp + col];   // previous unit's output
        pool[2 * H + col] = g; ja[2 * H + col] = g;
      }
    }
  }
  __syncthreads();
  JF_PROF(2);
  {   // fc_g1 / fc_global1.  The phase is bound by the number of memory instructions (the co-resident CTAs of an SM issue them through
      // one LSU), so a thread takes 4 output columns with one 16-byte weight load per k and the 8 warps split K; the partial
      // sums meet in shared memory.
    const int K = 2 * H + (unit > 0 ? Z : 0);
    const int c4 = tid & 31, kg = tid >> 5;
    const int kc = (K + 7) >> 3, k0 = kg * kc, k1 = (k0 + kc < K) ? k0 + kc : K;
    const float4* w = reinterpret_cast<const float4*>(Ga.Wt + (size_t)Ga.m_off * Ga.ldo) + c4;
    const int ld4 = Ga.ldo >> 2;
    float4 acc[2];
    acc[0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[1] = acc[0];
    int k = k0;
    for (; k + 4 <= k1; k += 4) {
      float4 wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[q] = __ldg(w + (size_t)(k + q) * ld4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x = pool[k + q];
        acc[q & 1].x = fmaf(wv[q].x, x, acc[q & 1].x); acc[q & 1].y = fmaf(wv[q].y, x, acc[q & 1].y);
        acc[q & 1].z = fmaf(wv[q].z, x, acc[q & 1].z); acc[q & 1].w = fmaf(wv[q].w, x, acc[q & 1].w);
      }
    }
    for (; k < k1; ++k) {
      const float4 wv = __ldg(w + (size_t)k * ld4);
      const float x = pool[k];
      acc[0].x = fmaf(wv.x, x, acc[0].x); acc[0].y = fmaf(wv.y, x, acc[0].y); acc[0].z = fmaf(wv.z, x, acc[0].z); acc[0].w = fmaf(wv.w, x, acc[0].w);
    }
    *reinterpret_cast<float4*>(red + kg * TT_H + c4 * 4) = make_float4(acc[0].x + acc[1].x, acc[0].y + acc[1].y, acc[0].z + acc[1].z, acc[0].w + acc[1].w);
    __syncthreads();
    if (!hp) {
      const float g1 = tt_lrelu(tt_colsum_get<8>(red, col) + p.beff[(size_t)j * p.bstride + Ga.bias_off + col], p.slope);
      g1s[col] = g1;
      ja[p.LDP + col] = g1;
    }
  }
  __syncthreads();
  JF_PROF(3);
  {   // fc_g2 / fc_global2: warp w handles outputs z = w, w + 8, ...; lanes split K
    const int warp = tid >> 5, lane = tid & 31;
    for (int z = warp; z < Z; z += 8) {
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) a = fmaf(__ldg(Gb.Wt + (size_t)(Gb.m_off + lane + 32 * q) * Gb.ldo + z), g1s[lane + 32 * q], a);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
      if (lane == 0) {
        a += p.beff[(size_t)j * p.bstride + Gb.bias_off + z];
        if (unit > 0) a += pool[2 * H + z];                              // residual, epic.py:184-186
        a = tt_lrelu(a, p.slope);
        gs[z] = a;
        ja[p.LDP + p.Hp + z] = a;
      }
    }
  }
  __syncthreads();
  JF_PROF(4);
  if (unit > 0 && !hp) {   // effective bias of fc_local1: + W_glob . g   (the broadcast global vector, epic.py:189-196)
    const Lin La = p.lin[LIN_LAYER0 + 4 * l + 2];
    float a = p.beff[(size_t)j * p.bstride + La.bias_off + col];
    float wz[32];
#pragma unroll
    for (int z = 0; z < 32; ++z) wz[z] = z < Z ? __ldg(La.Wt + (size_t)(La.g_off + z) * La.ldo + col) : 0.f;
#pragma unroll
    for (int z = 0; z < 32; ++z) a = fmaf(wz[z], z < Z ? gs[z] : 0.f, a);
    p.beff[(size_t)j * p.bstride + La.bias_off + col] = a;
  }
  JF_PROF(5);
  if (prof && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) for (int q = 0; q < 6; ++q) prof[(blockIdx.x ? 8 : 0) + q] = pc[q];
}

// head: v = lrelu(fc_l3(h_L)); flow-matching target, squared error, gradient seed (losses.py:61-62, :75-76).
// One thread per row (its 128 hidden features stream through L1), fc_l3's weights in shared memory.
__global__ void __launch_bounds__(128) tt_head_kernel(const TtCommon p) {
  __shared__ float w3[TT_H * 8];
  __shared__ float red[4];
  const int rows = *p.n_total;
  const Lin L3 = p.lin[p.n_lin - 1];
  const int F = p.F;
  for (int i = threadIdx.x; i < TT_H * F; i += blockDim.x) w3[i] = L3.Wt[(size_t)(L3.m_off + i / F) * L3.ldo + (i % F)];
  __syncthreads();
  const float* hL = p.act + (size_t)(1 + 2 * p.L) * p.stage_stride;
  const float inv_n = 1.f / (float)rows;
  float part = 0.f;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    const int j = p.rowjet[r];
    const int pidx = p.ridx[(size_t)j * p.N + (r - p.rowoff[j])];
    float v[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) v[f] = 0.f;
    const float4* hr = reinterpret_cast<const float4*>(hL + (size_t)r * TT_H);
    for (int c4 = 0; c4 < TT_H / 4; ++c4) {
      const float4 hv = hr[c4];
#pragma unroll
      for (int f = 0; f < 8; ++f)
        if (f < F) v[f] += hv.x * w3[(c4 * 4 + 0) * F + f] + hv.y * w3[(c4 * 4 + 1) * F + f] + hv.z * w3[(c4 * 4 + 2) * F + f] +
                           hv.w * w3[(c4 * 4 + 3) * F + f];
    }
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      if (f >= F) break;
      const float o = tt_lrelu(v[f] + p.beff[(size_t)j * p.bstride + L3.bias_off + f], p.slope);
      const size_t gi = ((size_t)j * p.N + pidx) * F + f;
      if (p.loss_kind >= 0) {
        const float x = p.x_in[gi], z = p.noise0[gi];
        float u;
        if (p.loss_kind == PFM_LOSS_FM_OT) u = (1.f - p.sigma) * z - x;
        else if (p.loss_kind == PFM_LOSS_CFM) u = z - x;
        else u = z;
        const float d = o - u;
        part += d * d;
        p.dpre3[(size_t)r * F + f] = 2.f * d * inv_n * tt_dlrelu(o, p.slope);
      } else {
        p.x_out[gi] = o;
        p.dpre3[(size_t)r * F + f] = tt_dlrelu(o, p.slope);      // the backward entry multiplies by the incoming gradient
      }
    }
  }
  if (p.loss_kind >= 0) {
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) part += __shfl_xor_sync(0xffffffffu, part, sft);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(p.loss_acc, (red[0] + red[1]) + (red[2] + red[3]));
  }
}

// head backward: dz2 of the last layer (or of fc_l2 when there are no layers) = (dpre3 . W3) * lrelu'(h_L) -> dact[1 + 2L]
// thread = (row, 4 columns)
__global__ void __launch_bounds__(256) tt_head_bwd_kernel(const TtCommon p) {
  __shared__ float w3[TT_H * 8];
  const int rows = *p.n_total;
  const Lin L3 = p.lin[p.n_lin - 1];
  const int F = p.F;
  for (int i = threadIdx.x; i < TT_H * F; i += blockDim.x) w3[i] = L3.Wt[(size_t)(L3.m_off + i / F) * L3.ldo + (i % F)];
  __syncthreads();
  float* out = p.dact + (size_t)(1 + 2 * p.L) * p.stage_stride;
  const uint32_t* sg = p.sgn + (size_t)(1 + 2 * p.L) * p.sgn_stride;
  // two rows per iteration: the (row -> dpre3, sign word) loads of both are in flight together
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)rows * 32; i += 2 * stride) {
    const int c = (int)(i & 31) * 4;
    const int r0 = (int)(i >> 5);
    const bool two = i + stride < (long long)rows * 32;
    const int r1 = two ? (int)((i + stride) >> 5) : r0;
    float d0[8], d1[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      d0[f] = f < F ? p.dpre3[(size_t)r0 * F + f] : 0.f;
      d1[f] = f < F ? p.dpre3[(size_t)r1 * F + f] : 0.f;
    }
    const uint32_t e0 = sg[(size_t)r0 * 4 + (c >> 5)] >> (c & 31), e1 = sg[(size_t)r1 * 4 + (c >> 5)] >> (c & 31);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
    for (int f = 0; f < 8; ++f) {
      if (f < F) {
        const float w0 = w3[(c + 0) * F + f], w1 = w3[(c + 1) * F + f], w2 = w3[(c + 2) * F + f], w3v = w3[(c + 3) * F + f];
        a.x = fmaf(w0, d0[f], a.x); a.y = fmaf(w1, d0[f], a.y); a.z = fmaf(w2, d0[f], a.z); a.w = fmaf(w3v, d0[f], a.w);
        b.x = fmaf(w0, d1[f], b.x); b.y = fmaf(w1, d1[f], b.y); b.z = fmaf(w2, d1[f], b.z); b.w = fmaf(w3v, d1[f], b.w);
      }
    }
    a.x *= (e0 & 1u) ? 1.f : p.slope; a.y *= (e0 & 2u) ? 1.f : p.slope; a.z *= (e0 & 4u) ? 1.f : p.slope; a.w *= (e0 & 8u) ? 1.f : p.slope;
    b.x *= (e1 & 1u) ? 1.f : p.slope; b.y *= (e1 & 2u) ? 1.f : p.slope; b.z *= (e1 & 4u) ? 1.f : p.slope; b.w *= (e1 & 8u) ? 1.f : p.slope;
    *reinterpret_cast<float4*>(out + (size_t)r0 * TT_H + c) = a;
    if (two) *reinterpret_cast<float4*>(out + (size_t)r1 * TT_H + c) = b;
  }
}

// Backward of one per-jet unit (mirrors epic_train.cu::global_backward and the code around it).  One CTA of 256 threads per jet.
//   unit l+1: db1 / db2 = per-jet sums of the pre-activation gradients of fc_local1 / fc_local2 (-> dbeff),
//             dG = carry + W_glob^T db1, then the global MLP backward; bc[j] = pooled gradient broadcast (overwritten)
//   unit 0  : stem global MLP backward from the carry; bc[j] += broadcast (both units pool h0); also the head's bias gradient
__global__ void __launch_bounds__(256, 4) tt_jet_bwd_kernel(const TtCommon p, int unit, int with_head) {
  __shared__ float db1[TT_H];
  __shared__ float pg1[TT_H];
  __shared__ float pg2[32];
  __shared__ float dG[32];
  __shared__ float din[2 * TT_H + 32];
  __shared__ __align__(16) float red[2 * 4 * TT_H];
  const int j = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, col = tid & 127, hp = tid >> 7;
  const int H = p.H, Z = p.Z;
  const int n = p.n_real[j], r0 = p.rowoff[j];
  const int l = unit - 1;
  const Lin Ga = p.lin[unit == 0 ? LIN_G1 : LIN_LAYER0 + 4 * l + 0];
  const Lin Gb = p.lin[unit == 0 ? LIN_G2 : LIN_LAYER0 + 4 * l + 1];
  const float* ja = p.jact + (size_t)j * p.jstride + (size_t)unit * p.junit;
  float* dbe = p.dbeff + (size_t)j * p.bstride;
  if (with_head) {
    const Lin L3 = p.lin[p.n_lin - 1];
    for (int f = warp; f < p.F; f += 8) {
      float a = 0.f;
      for (int r = lane; r < n; r += 32) a += p.dpre3[(size_t)(r0 + r) * p.F + f];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
      if (lane == 0) dbe[L3.bias_off + f] = a;
    }
  }
  if (unit > 0) {
    const Lin La = p.lin[LIN_LAYER0 + 4 * l + 2], Lb = p.lin[LIN_LAYER0 + 4 * l + 3];
    // threads [0,128): sums of dz1 (stage 2+2l); threads [128,256): sums of dz2 (stage 3+2l)
    tt_colsum_put<4>(p.dact + (size_t)(2 + 2 * l + hp) * p.stage_stride, r0, n, tid & 127, red + hp * 4 * TT_H);
    __syncthreads();
    const float sc = tt_colsum_get<4>(red + hp * 4 * TT_H, col);
    if (!hp) { db1[col] = sc; dbe[La.bias_off + col] = sc; }
    else dbe[Lb.bias_off + col] = sc;
    __syncthreads();
    for (int z = warp; z < Z; z += 8) {          // gradient w.r.t. the unit's new global vector: carry + W_glob^T . db1
      const float* w = La.Wt + (size_t)(La.g_off + z) * La.ldo;
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) a = fmaf(__ldg(w + lane + 32 * q), db1[lane + 32 * q], a);
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
      if (lane == 0) dG[z] = a + p.dGc[(size_t)j * p.Zp + z];
    }
  } else if (tid < Z) {
    dG[tid] = p.dGc[(size_t)j * p.Zp + tid];
  }
  __syncthreads();
  if (tid < Z) {
    const float g = ja[p.LDP + p.Hp + tid];
    const float v = dG[tid] * tt_dlrelu(g, p.slope);
    pg2[tid] = v;
    dbe[Gb.bias_off + tid] = v;
  }
  __syncthreads();
  if (!hp) {
    float a = 0.f;
    float wz[32];
#pragma unroll
    for (int z = 0; z < 32; ++z) wz[z] = z < Z ? __ldg(Gb.Wr + (size_t)z * Gb.ldr + col) : 0.f;
#pragma unroll
    for (int z = 0; z < 32; ++z) a = fmaf(wz[z], z < Z ? pg2[z] : 0.f, a);
    const float v = a * tt_dlrelu(ja[p.LDP + col], p.slope);
    pg1[col] = v;
    dbe[Ga.bias_off + col] = v;
  }
  __syncthreads();
  {   // din[k] = sum_o W_a[o][m_off + k] pg1[o]: thread = 4 consecutive k (one 16-byte load per weight row) x one third of the
      // 128 output rows; the three partial rows meet in shared memory (memory-instruction bound like fc_global1's forward)
    const int n4 = (Ga.m_len + 3) >> 2;                    // <= 72 (m_len <= 2 H + 32: tt_enabled caps the latent width at 32)
    const int og = tid / 72, k4 = tid - og * 72;
    if (og < 3 && k4 < n4) {
      const int o0 = og * 43, o1 = (o0 + 43 < H) ? o0 + 43 : H;
      const float4* wk = reinterpret_cast<const float4*>(Ga.Wr) + k4;
      const int ld4 = Ga.ldr >> 2;
      float4 acc[2];
      acc[0] = make_float4(0.f, 0.f, 0.f, 0.f); acc[1] = acc[0];
      int o = o0;
      for (; o + 4 <= o1; o += 4) {
        float4 wv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wv[q] = __ldg(wk + (size_t)(o + q) * ld4);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x = pg1[o + q];
          acc[q & 1].x = fmaf(wv[q].x, x, acc[q & 1].x); acc[q & 1].y = fmaf(wv[q].y, x, acc[q & 1].y);
          acc[q & 1].z = fmaf(wv[q].z, x, acc[q & 1].z); acc[q & 1].w = fmaf(wv[q].w, x, acc[q & 1].w);
        }
      }
      for (; o < o1; ++o) {
        const float4 wv = __ldg(wk + (size_t)o * ld4);
        const float x = pg1[o];
        acc[0].x = fmaf(wv.x, x, acc[0].x); acc[0].y = fmaf(wv.y, x, acc[0].y); acc[0].z = fmaf(wv.z, x, acc[0].z); acc[0].w = fmaf(wv.w, x, acc[0].w);
      }
      *reinterpret_cast<float4*>(red + og * 288 + k4 * 4) = make_float4(acc[0].x + acc[1].x, acc[0].y + acc[1].y, acc[0].z + acc[1].z, acc[0].w + acc[1].w);
    }
    __syncthreads();
    for (int k = tid; k < Ga.m_len; k += 256) din[k] = (red[k] + red[288 + k]) + red[576 + k];
  }
  __syncthreads();
  if (!hp) {
    const int o_mean = unit == 0 ? H : 0, o_sum = unit == 0 ? 0 : H;
    const float v = din[o_mean + col] / (float)n + p.sum_scale * din[o_sum + col];
    float* b = p.bc + (size_t)j * TT_H + col;
    *b = unit == 0 ? *b + v : v;
    if (unit > 0 && col < Z) p.dGc[(size_t)j * p.Zp + col] = din[2 * H + col] + pg2[col];     // fc_global1's global columns + residual
  }
}

// stem tail: per-jet sums of the fc_l1 / fc_l2 pre-activation gradients, gradient w.r.t. the per-particle input columns
__global__ void __launch_bounds__(256) tt_stem_bwd_kernel(const TtCommon p) {
  const int j = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, col = tid & 127, hp = tid >> 7;
  const int n = p.n_real[j], r0 = p.rowoff[j];
  const Lin L1 = p.lin[LIN_L1], L2 = p.lin[LIN_L2];
  float* dbe = p.dbeff + (size_t)j * p.bstride;
  __shared__ __align__(16) float red[2 * 4 * TT_H];
  tt_colsum_put<4>(p.dact + (size_t)hp * p.stage_stride, r0, n, tid & 127, red + hp * 4 * TT_H);
  __syncthreads();
  const float sc = tt_colsum_get<4>(red + hp * 4 * TT_H, col);
  dbe[(hp ? L2.bias_off : L1.bias_off) + col] = sc;
  if (p.dxs) {
    for (int r = warp; r < n; r += 8) {
      const float4 d = __ldcg(reinterpret_cast<const float4*>(p.dact + (size_t)(r0 + r) * TT_H + lane * 4));
      for (int c = 0; c < p.Kx; ++c) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(L1.Wt + (size_t)(L1.m_off + p.xin_off + c) * L1.ldo + lane * 4));
        float a = w.x * d.x + w.y * d.y + w.z * d.z + w.w * d.w;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) a += __shfl_xor_sync(0xffffffffu, a, sft);
        if (lane == 0) p.dxs[(size_t)(r0 + r) * p.Kx + c] = a;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
bool tt_enabled(const pfm_epic* h) {
  static const int force_simt = getenv("PFM_TRAIN_SIMT") ? atoi(getenv("PFM_TRAIN_SIMT")) : 0;
  const pfm_epic_cfg& c = h->cfg;
  return !force_simt && h->train_mode != PFM_TRAIN_CUDA_CORES && c.hid == TT_H && c.latent <= 32 && c.feats <= 8;
}

static int tt_pack(pfm_epic* h, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  const int n_gemm = 1 + 2 * c.layers;
  const size_t bytes = (size_t)n_gemm * 2 * 2 * TT_IMG;
  const size_t aux = sizeof(TtImgSrc) * n_gemm;
  if (h->tt_bytes < bytes + aux) {
    if (h->tt_store) cudaFree(h->tt_store);
    h->tt_store = nullptr; h->tt_bytes = 0;
    PFM_CUDA_CHECK(cudaMalloc(&h->tt_store, bytes + aux));
    h->tt_bytes = bytes + aux;
    std::vector<TtImgSrc> src(n_gemm);
    auto mk = [&](int lin_idx) { const Lin& L = h->lin_host[lin_idx]; TtImgSrc s; s.Wt = L.Wt; s.ldo = L.ldo; s.k0 = L.m_off; return s; };
    src[0] = mk(LIN_L2);
    for (int l = 0; l < c.layers; ++l) { src[1 + 2 * l] = mk(LIN_LAYER0 + 4 * l + 2); src[2 + 2 * l] = mk(LIN_LAYER0 + 4 * l + 3); }
    PFM_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(h->tt_store) + bytes, src.data(), aux, cudaMemcpyHostToDevice, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));          // src is a host temporary (pointers into the handle's weight store: once per allocation)
  }
  tt_pack_kernel<<<n_gemm * 2 * 8, 256, 0, st>>>(reinterpret_cast<const TtImgSrc*>(reinterpret_cast<uint8_t*>(h->tt_store) + bytes),
                                              reinterpret_cast<uint8_t*>(h->tt_store));
  PFM_CUDA_CHECK(cudaGetLastError());
  h->tt_dirty = false;
  h->last_launches++;
  return PFM_OK;
}

static int tt_workspace(pfm_epic* h, int B, int N) {
  const int Zp = (h->cfg.latent + 3) & ~3;
  const size_t rows = (size_t)B * N;
  const size_t stages = 2 + 2 * (size_t)h->cfg.layers;
  const size_t need = (size_t)B * h->bstride + (size_t)B * TT_H + (size_t)B * Zp + stages * rows * 4 + rows + 64;
  if (h->tt_ws_cap < need) {
    if (h->tt_ws) cudaFree(h->tt_ws);
    h->tt_ws = nullptr; h->tt_ws_cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(&h->tt_ws, need * sizeof(float)));
    h->tt_ws_cap = need;
  }
  return PFM_OK;
}

static void tt_common(pfm_epic* h, int B, int N, int Kx, int xin_off, const TrainLayout& lay, TtCommon* p) {
  const pfm_epic_cfg& c = h->cfg;
  memset(p, 0, sizeof(*p));
  const int Zp = (c.latent + 3) & ~3;
  p->lin = h->lin_dev; p->n_lin = h->n_lin; p->L = c.layers; p->H = c.hid; p->Z = c.latent; p->F = c.feats; p->Kx = Kx;
  p->xin_off = xin_off; p->N = N; p->B = B; p->bstride = h->bstride; p->sum_scale = c.sum_scale; p->slope = c.neg_slope;
  p->n_real = h->plan.n_real; p->ridx = h->plan.ridx; p->rowoff = h->plan.rowoff; p->n_total = h->plan.n_total;
  p->act = h->act; p->dact = h->dact; p->stage_stride = lay.stage_stride;
  p->yact = h->yact; p->jact = h->jact; p->junit = lay.junit; p->jstride = lay.jstride; p->LDP = lay.LDP; p->Hp = lay.Hp; p->Zp = lay.Zp;
  p->dpre3 = h->dpre3; p->dbeff = h->dbeff; p->loss_acc = h->loss_acc;
  float* w = h->tt_ws;
  p->beff = w; w += (size_t)B * h->bstride;
  p->bc = w; w += (size_t)B * TT_H;
  p->dGc = w; w += (size_t)B * Zp;
  p->sgn = reinterpret_cast<uint32_t*>(w); p->sgn_stride = (size_t)B * N * 4; w += (2 + 2 * (size_t)c.layers) * p->sgn_stride;
  p->rowjet = reinterpret_cast<const int*>(w);
}

// Debug (PFM_TT_CHECK=1): recompute a rowlin launch on CUDA cores in fp32 from the raw fp32 weights and report the largest deviation.
__global__ void rowlin_check_kernel(const RowLinP p, const float* __restrict__ Wt, int ldo, int k0, int transposed, float* __restrict__ maxerr) {
  const int rows = *p.n_total;
  const int r = blockIdx.x, o = threadIdx.x;
  if (r >= rows) return;
  const int jet = p.rowjet[r];
  float a = 0.f;
  for (int k = 0; k < 128; ++k) {
    const float x = p.X[(size_t)r * 128 + k];
    // forward: W[o][k] = Wt[(k0 + k) * ldo + o];  transposed: contraction over the linear's outputs: W[k][o]
    const float w = transposed ? Wt[(size_t)(k0 + o) * ldo + k] : Wt[(size_t)(k0 + k) * ldo + o];
    a = fmaf(x, w, a);
  }
  if (p.bias) a += p.bias[(size_t)jet * p.bias_ld + o];
  if (p.R) a += p.R[(size_t)r * 128 + o];
  if (p.bc) a += p.bc[(size_t)jet * 128 + o];
  if (p.act) a = tt_lrelu(a, p.slope);
  if (p.E) a *= ((p.E[(size_t)r * 4 + (o >> 5)] >> (o & 31)) & 1u) ? 1.f : p.slope;
  const float d = fabsf(a - p.Y[(size_t)r * 128 + o]);
  atomicMax(reinterpret_cast<int*>(maxerr), __float_as_int(d));
  atomicMax(reinterpret_cast<int*>(maxerr + 1), __float_as_int(fabsf(a)));
}

static int tt_rowlin(pfm_epic* h, RowLinP& q, int gemm, int transposed, const TtCommon& c, int grid, cudaStream_t st) {
  q.Wimg = reinterpret_cast<const uint8_t*>(h->tt_store) + (size_t)(2 * gemm + (transposed ? 1 : 0)) * 2 * TT_IMG;
  q.rowjet = c.rowjet; q.n_total = c.n_total; q.slope = c.slope;
  static const bool do_prof = getenv("PFM_TT_PROF") != nullptr;

  static long long* dprof = nullptr;
  if (do_prof) {
    if (!dprof) PFM_CUDA_CHECK(cudaMalloc(&dprof, 64));
    q.prof = dprof;
  }
  {
    const int fl = (q.bias ? RL2_BIAS : 0) | (q.R ? RL2_RES : 0) | (q.bc ? RL2_BC : 0) | (q.act ? RL2_ACT : 0) | (q.E ? RL2_E : 0) |
                   (q.sgn_out ? RL2_SGN : 0);
    const int smem2 = (int)sizeof(RowLin2Smem) + 1024;
    static bool attr2 = false;
    if (!attr2) {                                            // every instantiation the switch below can launch
      void (*ks[5])(const RowLinP) = {rowlin2_tc_kernel<(RL2_BIAS | RL2_RES | RL2_ACT | RL2_SGN)>, rowlin2_tc_kernel<(RL2_BIAS | RL2_ACT | RL2_SGN)>,
                                      rowlin2_tc_kernel<RL2_E>, rowlin2_tc_kernel<(RL2_RES | RL2_BC | RL2_E)>, rowlin2_tc_kernel<(RL2_RES | RL2_E)>};
      for (auto k : ks) PFM_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
      attr2 = true;
    }
    cudaError_t lerr = cudaSuccess;
    auto launch = [&](void (*kernel)(const RowLinP)) { lerr = launch_pdl(kernel, dim3((unsigned)grid), dim3(RL2_THREADS), (size_t)smem2, st, q); };
    switch (fl) {
      case RL2_BIAS | RL2_RES | RL2_ACT | RL2_SGN: launch(rowlin2_tc_kernel<(RL2_BIAS | RL2_RES | RL2_ACT | RL2_SGN)>); break;   // fc_l2, fc_local2
      case RL2_BIAS | RL2_ACT | RL2_SGN:           launch(rowlin2_tc_kernel<(RL2_BIAS | RL2_ACT | RL2_SGN)>); break;             // fc_local1
      case RL2_E:                                  launch(rowlin2_tc_kernel<RL2_E>); break;                                      // dz2 . W2
      case RL2_RES | RL2_BC | RL2_E:               launch(rowlin2_tc_kernel<(RL2_RES | RL2_BC | RL2_E)>); break;                 // dz1 . W1 + residual + pooling
      case RL2_RES | RL2_E:                        launch(rowlin2_tc_kernel<(RL2_RES | RL2_E)>); break;                          // stem
      default: set_error("rowlin pass with operand set %d is not instantiated", fl); return PFM_ERR_UNSUPPORTED;
    }
    PFM_CUDA_CHECK(lerr);
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  h->last_launches++;
  if (do_prof) {
    long long hp[8];
    PFM_CUDA_CHECK(cudaMemcpyAsync(hp, dprof, 64, cudaMemcpyDeviceToHost, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[pfm tt prof] gemm %2d %s  converters: wait_raw %lld  wait_mma %lld  convert %lld  barrier+load %lld  issue %lld | epilogue: wait_mma %lld  process %lld\n",
            gemm, transposed ? "bwd" : "fwd", hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6]);
  }
  static const bool check = getenv("PFM_TT_CHECK") != nullptr;
  if (check) {
    static float* dm = nullptr;
    if (!dm) PFM_CUDA_CHECK(cudaMalloc(&dm, 8));
    PFM_CUDA_CHECK(cudaMemsetAsync(dm, 0, 8, st));
    const int lin_idx = gemm == 0 ? LIN_L2 : (LIN_LAYER0 + 4 * ((gemm - 1) >> 1) + 2 + ((gemm - 1) & 1));
    const Lin& L = h->lin_host[lin_idx];
    rowlin_check_kernel<<<c.B * c.N, 128, 0, st>>>(q, L.Wt, L.ldo, L.m_off, transposed, dm);
    float hm[2];
    PFM_CUDA_CHECK(cudaMemcpyAsync(hm, dm, 8, cudaMemcpyDeviceToHost, st));
    PFM_CUDA_CHECK(cudaStreamSynchronize(st));
    fprintf(stderr, "[pfm tt check] gemm %2d %s max|err| %.3e  max|ref| %.3e\n", gemm, transposed ? "bwd" : "fwd", hm[0], hm[1]);
  }
  return PFM_OK;
}

// The training plan of this path: rows of the packed particles (prefix sum of the multiplicities) and the row -> jet map.
// (plan_count_kernel has filled n_real / ridx.)
int tt_plan(pfm_epic* h, int B, int N, cudaStream_t st) {
  int rc = tt_workspace(h, B, N);
  if (rc != PFM_OK) return rc;
  const pfm_epic_cfg& c = h->cfg;
  const int Zp = (c.latent + 3) & ~3;
  int* rowjet = reinterpret_cast<int*>(h->tt_ws + (size_t)B * h->bstride + (size_t)B * TT_H + (size_t)B * Zp +
                                       (2 + 2 * (size_t)c.layers) * (size_t)B * N * 4);
  tt_rows_kernel<<<1, 1024, 0, st>>>(h->plan.n_real, B, h->plan.rowoff, h->plan.n_total);
  tt_rowjet_kernel<<<(B + 7) / 8, 256, 0, st>>>(h->plan.n_real, h->plan.rowoff, B, rowjet);
  PFM_CUDA_CHECK(cudaGetLastError());
  h->last_launches += 2;
  return PFM_OK;
}

// output of the plain forward before the head scatters the real particles into it: 0 at padding, NaN for a jet without
// particles (its mean pooling is 0/0 in the reference, epic.py:161-163, and NaN * mask stays NaN)
__global__ void tt_out_init_kernel(float* __restrict__ out, const int* __restrict__ n_real, int NF) {
  const float v = n_real[blockIdx.x] == 0 ? __int_as_float(0x7fc00000) : 0.f;
  float* o = out + (size_t)blockIdx.x * NF;
  for (int i = threadIdx.x; i < NF; i += blockDim.x) o[i] = v;
}

// forward with saved activations (+ fused flow-matching loss): fills act / yact / jact / dpre3 / loss_acc like
// epic_simt.cu's TRAIN instantiation (tt_plan has run)
int tt_train_forward(pfm_epic* h, const TrainFwdArgs& a, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  int rc;
  if (!h->tt_store || h->tt_dirty) { rc = tt_pack(h, st); if (rc != PFM_OK) return rc; }
  TtCommon p;
  tt_common(h, a.B, a.N, a.Kx, a.xin_off, a.lay, &p);
  p.x_in = a.x_in; p.x_out = a.x_out; p.tjet = a.t; p.noise0 = a.noise0; p.noise1 = a.noise1; p.loss_kind = a.loss_kind; p.sigma = a.sigma;
  const size_t SS = a.lay.stage_stride, GS = p.sgn_stride;
  const int grid_rows = 4 * h->sm_count;
  const int grid_gemm = h->sm_count;
  const int jet_ctas = a.B;
  tt_beff_kernel<<<a.B, 256, 0, st>>>(h->tbias, a.tbias_per_jet, a.has_cbias ? h->cbias : nullptr, h->bstride, p.beff);
  if (a.loss_kind < 0) tt_out_init_kernel<<<a.B, 128, 0, st>>>(a.x_out, h->plan.n_real, a.N * c.feats);
  tt_stem_kernel<<<(a.B * a.N + 15) / 16 < 8 * grid_rows ? (a.B * a.N + 15) / 16 : 8 * grid_rows, 256, 0, st>>>(p);
  h->last_launches += 2;
  {   // fc_l2: h0 = lrelu(h1 . W^T + b + h1)     (epic.py:364-367)
    RowLinP q; memset(&q, 0, sizeof(q));
    q.X = h->act; q.R = h->act; q.bias = p.beff + h->lin_host[LIN_L2].bias_off; q.bias_ld = h->bstride; q.act = 1; q.Y = h->act + SS;
    q.sgn_out = p.sgn + GS;
    if ((rc = tt_rowlin(h, q, 0, 0, p, grid_gemm, st)) != PFM_OK) return rc;
  }
  static const bool jprof = getenv("PFM_TT_PROF") != nullptr;
  static long long* djp = nullptr;
  if (jprof && !djp) PFM_CUDA_CHECK(cudaMalloc(&djp, 128));
  tt_jet_fwd_kernel<<<jet_ctas, 256, 0, st>>>(p, 0, nullptr);
  h->last_launches++;
  for (int l = 0; l < c.layers; ++l) {
    tt_jet_fwd_kernel<<<jet_ctas, 256, 0, st>>>(p, l + 1, jprof ? djp : nullptr);
    if (jprof) {
      long long hp[16];
      PFM_CUDA_CHECK(cudaMemcpyAsync(hp, djp, 128, cudaMemcpyDeviceToHost, st));
      PFM_CUDA_CHECK(cudaStreamSynchronize(st));
      fprintf(stderr, "[pfm tt prof] jet fwd unit %d  first CTA: head %lld colsum %lld pool %lld g1 %lld g2 %lld bias %lld | last CTA: %lld %lld %lld %lld %lld %lld\n", l + 1,
              hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[8], hp[9], hp[10], hp[11], hp[12], hp[13]);
    }
    h->last_launches++;
    RowLinP q; memset(&q, 0, sizeof(q));      // fc_local1: u = lrelu(h . W1^T + beff1[jet])     (epic.py:194-196)
    q.X = h->act + (size_t)(1 + 2 * l) * SS; q.bias = p.beff + h->lin_host[LIN_LAYER0 + 4 * l + 2].bias_off; q.bias_ld = h->bstride;
    q.act = 1; q.Y = h->act + (size_t)(2 + 2 * l) * SS; q.sgn_out = p.sgn + (size_t)(2 + 2 * l) * GS;
    if ((rc = tt_rowlin(h, q, 1 + 2 * l, 0, p, grid_gemm, st)) != PFM_OK) return rc;
    memset(&q, 0, sizeof(q));                 // fc_local2: h' = lrelu(u . W2^T + beff2[jet] + h)  (epic.py:198-200)
    q.X = h->act + (size_t)(2 + 2 * l) * SS; q.R = h->act + (size_t)(1 + 2 * l) * SS;
    q.bias = p.beff + h->lin_host[LIN_LAYER0 + 4 * l + 3].bias_off; q.bias_ld = h->bstride; q.act = 1; q.Y = h->act + (size_t)(3 + 2 * l) * SS;
    q.sgn_out = p.sgn + (size_t)(3 + 2 * l) * GS;
    if ((rc = tt_rowlin(h, q, 2 + 2 * l, 0, p, grid_gemm, st)) != PFM_OK) return rc;
  }
  tt_head_kernel<<<grid_rows, 128, 0, st>>>(p);
  h->last_launches++;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// backward: fills dact / dbeff (/ dxs) like epic_train.cu::epic_bwd_kernel.  Every launch is a plain GEMM pass: the masked
// gradient entering a layer is produced by the epilogue of the pass before it.
int tt_train_backward(pfm_epic* h, const TrainBwdArgs& a, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  int rc;
  if (!h->train_tc) {          // the saved forward came from the CUDA-core kernels: this path's row map and sign bits are missing
    if ((rc = tt_plan(h, a.B, a.N, st)) != PFM_OK) return rc;
  }
  if (!h->tt_store || h->tt_dirty) { rc = tt_pack(h, st); if (rc != PFM_OK) return rc; }
  TtCommon p;
  tt_common(h, a.B, a.N, a.Kx, a.xin_off, a.lay, &p);
  p.dxs = a.want_dx ? h->dxs : nullptr;
  const size_t SS = a.lay.stage_stride, GS = p.sgn_stride;
  const int Zp = a.lay.Zp;
  const int grid_rows = 4 * h->sm_count;
  const int grid_gemm = h->sm_count;
  const int jet_ctas = a.B;
  if (!h->train_tc) {
    tt_sign_kernel<<<dim3(2 * h->sm_count, 2 + 2 * c.layers), 256, 0, st>>>(h->act, SS, p.sgn, GS, p.n_total);
    h->last_launches++;
  }
  PFM_CUDA_CHECK(cudaMemsetAsync(p.bc, 0, sizeof(float) * ((size_t)a.B * TT_H + (size_t)a.B * Zp), st));      // bc and dGc are adjacent
  tt_head_bwd_kernel<<<grid_rows, 256, 0, st>>>(p);            // dz2 of the last layer -> dact[1 + 2L]
  h->last_launches++;
  for (int l = c.layers - 1; l >= 0; --l) {
    const bool first = l == c.layers - 1;
    RowLinP q; memset(&q, 0, sizeof(q));
    // dz1 = (dz2 . W2) * lrelu'(u_l) -> dact[2+2l]
    q.X = h->dact + (size_t)(3 + 2 * l) * SS; q.E = p.sgn + (size_t)(2 + 2 * l) * GS; q.Y = h->dact + (size_t)(2 + 2 * l) * SS;
    if ((rc = tt_rowlin(h, q, 2 + 2 * l, 1, p, grid_gemm, st)) != PFM_OK) return rc;
    tt_jet_bwd_kernel<<<jet_ctas, 256, 0, st>>>(p, l + 1, first ? 1 : 0);
    h->last_launches++;
    if (l == 0) {
      tt_jet_bwd_kernel<<<jet_ctas, 256, 0, st>>>(p, 0, 0);
      h->last_launches++;
    }
    // gradient entering the layer below (pre-activation of its fc_local2, or of fc_l2 for l == 0):
    //   (dz2 [residual] + dz1 . W1(main) + bc[jet] [pooling]) * lrelu'(h_l) -> dact[1+2l]
    memset(&q, 0, sizeof(q));
    q.X = h->dact + (size_t)(2 + 2 * l) * SS; q.R = h->dact + (size_t)(3 + 2 * l) * SS; q.bc = p.bc; q.E = p.sgn + (size_t)(1 + 2 * l) * GS;
    q.Y = h->dact + (size_t)(1 + 2 * l) * SS;
    if ((rc = tt_rowlin(h, q, 1 + 2 * l, 1, p, grid_gemm, st)) != PFM_OK) return rc;
  }
  if (c.layers == 0) {
    tt_jet_bwd_kernel<<<jet_ctas, 256, 0, st>>>(p, 0, 1);
    h->last_launches++;
  }
  {   // stem: dz_l1 = (dz_l2 . W_l2 + dz_l2) * lrelu'(h1) -> dact[0]
    RowLinP q; memset(&q, 0, sizeof(q));
    q.X = h->dact + SS; q.R = h->dact + SS; q.E = p.sgn; q.Y = h->dact;
    if ((rc = tt_rowlin(h, q, 0, 1, p, grid_gemm, st)) != PFM_OK) return rc;
  }
  tt_stem_bwd_kernel<<<a.B, 256, 0, st>>>(p);
  h->last_launches++;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // namespace pfm
