// Weight gradients on the tensor cores:  dW[o][col0 + c] += sum_r Y[r][o] * X[r][c]   (same job table as
// epic_train.cu::xty_kernel, which it replaces; torch autograd forms these as addmm / mm_backward in the reference).
//
// The contraction runs over ROWS (particles or jets), i.e. over the slow index of both row-major operands, so both
// are MN-major tcgen05 operands: a stage of 64 rows is converted to bf16 and stored as [row][64-column swizzled
// 128-byte lines] -- the layout of a K-major tile read through an MN-major descriptor, as the pooling MMA of epic_tc.cu
// does.  fp32 accuracy is kept by the 3-term split  x = hi + lo (two bf16):  X^T Y ~ Xh^T Yh + Xh^T Yl + Xl^T Yh  (the
// dropped Xl^T Yl term is 2^-16 relative), three tcgen05.mma M=128 N=128 K=16 per 16 rows, fp32 accumulation in TMEM.
// D[m = c][n = o]: TMEM lane = input column c, so the atomics of a warp hit 32 consecutive floats of one dW row.
// One CTA = one 128 x 128 tile of dW over a chunk of rows (two CTAs per SM); double-buffered 32-row stages, the MMAs of stage s overlap the
// global loads + conversion of stage s+1.  Bound: HBM/L2 (each operand element is read once per 128-wide tile).
#include <cstdlib>
#include <type_traits>

#include "pfm_internal.cuh"
#include "tc_ptx.cuh"

namespace pfm {

using namespace tc;

static constexpr int XS_ROWS = 32;                  // rows per stage (64 KB of stage buffers: two CTAs per SM)
static constexpr int XS_TILE = XS_ROWS * 128 * 2;   // bytes of one bf16 [64 x 128] operand tile

struct XtyTcSmem {
  uint8_t buf[2][4][XS_TILE];                       // [stage buffer][Xh, Xl, Yh, Yl]
  uint64_t mbar[2];
  uint32_t tmem;
};

// 4 consecutive columns of row `r` (zero beyond `rows` / `width`) -> hi / lo bf16 at the swizzled position
__device__ __forceinline__ float4 xty_load4(const float* __restrict__ base, int ld, int r, int rows, int c, int width, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (r < rows && c < width) {
    const float* p = base + (size_t)r * ld + c;
    if (vec && c + 4 <= width) {
      v = __ldg(reinterpret_cast<const float4*>(p));
    } else {
      v.x = __ldg(p);
      if (c + 1 < width) v.y = __ldg(p + 1);
      if (c + 2 < width) v.z = __ldg(p + 2);
      if (c + 3 < width) v.w = __ldg(p + 3);
    }
  }
  return v;
}
__device__ __forceinline__ void xty_store4(uint8_t* hi, uint8_t* lo, int r, int c, float4 v) {
  const uint32_t off = sw128_offset(r, c, XS_ROWS * 128);
  uint2 h, l;
  split_bf16x4(v, h, l);
  *reinterpret_cast<uint2*>(hi + off) = h;
  *reinterpret_cast<uint2*>(lo + off) = l;
}

__global__ void __launch_bounds__(256, 2) xty_tc_kernel(const XtyJob* __restrict__ jobs, int n_jobs, const int* __restrict__ n_total,
                                                        const XtyJob single, int use_single, int min_chunk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  XtyTcSmem& s = *reinterpret_cast<XtyTcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  XtyJob J;
  int t_local;
  if (use_single) {
    J = single; t_local = blockIdx.y;
  } else {
    int jb = 0;
    const int tile = blockIdx.y;
    for (int i = 1; i < n_jobs; ++i)
      if (jobs[i].tile0 <= tile) jb = i;
    J = jobs[jb]; t_local = tile - J.tile0;
  }
  const int rows = J.rows >= 0 ? J.rows : *n_total;
  int chunk = (rows + (int)gridDim.x - 1) / (int)gridDim.x;
  chunk = (chunk + XS_ROWS - 1) / XS_ROWS * XS_ROWS;
  if (chunk < min_chunk) chunk = min_chunk;      // per-jet jobs (rows = jets): a few CTAs of 8 stages, not one CTA (+ a 128 x 128 atomic epilogue) per stage
  const int r_begin = blockIdx.x * chunk;
  if (r_begin >= rows) return;                     // uniform for the block, before any barrier / allocation
  const int r_end = min(rows, r_begin + chunk);
  const int o0 = (t_local / J.tiles_k) * 128, k0 = (t_local % J.tiles_k) * 128;
  const int wx = J.K - k0, wy = J.out - o0;        // valid columns of this tile (may exceed 128)
  const float* Xb = J.X + k0;
  const float* Yb = J.Y + o0;
  const bool vx = (J.ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(Xb) & 15) == 0;
  const bool vy = (J.ldy & 3) == 0 && (reinterpret_cast<uintptr_t>(Yb) & 15) == 0;

  if (warp == 0) tmem_alloc(&s.tmem, 128);
  if (tid == 0) { mbar_init(&s.mbar[0], 1); mbar_init(&s.mbar[1], 1); fence_barrier_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem;
  const uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);
  const int n_st = (r_end - r_begin + XS_ROWS - 1) / XS_ROWS;
  constexpr int NLD = XS_ROWS * 32 / 256;          // 16-byte loads per thread and operand
  // register prefetch two stages deep (a CTA's stage time is load latency + conversion: with one stage in flight it was
  // ~2.2 k cycles per 32 rows): stage st + 2 is requested as soon as stage st has been converted
  // FAST: both operands are full, 16-byte aligned 128-column tiles (the particle-row jobs, i.e. all the traffic): plain
  // vector loads with one row predicate instead of xty_load4's per-column guards
  auto run = [&](auto fast_tag) {
  constexpr bool FAST = decltype(fast_tag)::value;
  float4 xa[NLD], ya[NLD], xb[NLD], yb[NLD];
  auto fetch = [&](int st, float4 (&xv)[NLD], float4 (&yv)[NLD]) {       // 2 NLD independent 16-byte loads per thread
    const int r0 = r_begin + st * XS_ROWS;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + 256 * i, r = idx >> 5, c = (idx & 31) * 4;
      if (FAST) {
        const bool ok = r0 + r < r_end;
        xv[i] = ok ? __ldg(reinterpret_cast<const float4*>(Xb + (size_t)(r0 + r) * J.ldx + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        yv[i] = ok ? __ldg(reinterpret_cast<const float4*>(Yb + (size_t)(r0 + r) * J.ldy + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        xv[i] = xty_load4(Xb, J.ldx, r0 + r, r_end, c, wx < 128 ? wx : 128, vx);
        yv[i] = xty_load4(Yb, J.ldy, r0 + r, r_end, c, wy < 128 ? wy : 128, vy);
      }
    }
  };
  auto stage = [&](int st, float4 (&xv)[NLD], float4 (&yv)[NLD]) {
    const int b = st & 1;
    if (st >= 2) mbar_wait(&s.mbar[b], (uint32_t)(((st >> 1) - 1) & 1));     // the MMAs that read this buffer are done
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int idx = tid + 256 * i, r = idx >> 5, c = (idx & 31) * 4;
      xty_store4(s.buf[b][0], s.buf[b][1], r, c, xv[i]);
      xty_store4(s.buf[b][2], s.buf[b][3], r, c, yv[i]);
    }
    if (st + 2 < n_st) fetch(st + 2, xv, yv);      // flies during the barrier, the MMAs and the whole next stage
    fence_proxy_async();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t xh = desc_mnmajor(smem_u32(s.buf[b][0]), XS_ROWS * 128, 1024), xl = desc_mnmajor(smem_u32(s.buf[b][1]), XS_ROWS * 128, 1024);
        const uint64_t yh = desc_mnmajor(smem_u32(s.buf[b][2]), XS_ROWS * 128, 1024), yl = desc_mnmajor(smem_u32(s.buf[b][3]), XS_ROWS * 128, 1024);
#pragma unroll
        for (int k = 0; k < XS_ROWS / 16; ++k) {   // 16 rows per MMA: +2 KB in the MN-major view
          const uint64_t d = (uint64_t)(k * 128);
          mma_ss(tm, xh + d, yh + d, idesc, (st | k) ? 1u : 0u);
          mma_ss(tm, xh + d, yl + d, idesc, 1u);
          mma_ss(tm, xl + d, yh + d, idesc, 1u);
        }
        mma_commit(&s.mbar[b]);
      }
      __syncwarp();
    }
  };
  fetch(0, xa, ya);
  if (n_st > 1) fetch(1, xb, yb);
  for (int st = 0; st < n_st; st += 2) {
    stage(st, xa, ya);
    if (st + 1 < n_st) stage(st + 1, xb, yb);
  }
  };
  if (wx >= 128 && wy >= 128 && vx && vy) run(std::true_type{});
  else run(std::false_type{});
  {
    const int last = n_st - 1;
    mbar_wait(&s.mbar[last & 1], (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }
  // epilogue: lane = input column c (TMEM lane), 64 output columns per warp half
  {
    const int q = warp & 3, hf = warp >> 2;
    const int c = q * 32 + lane;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t v[32];
      tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * 64 + j * 32), v);
      tmem_wait_ld();
      if (c < wx) {
        float* dst = J.dW + J.col0 + k0 + c;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int o = hf * 64 + j * 32 + i;
          if (o < wy) atomicAdd(dst + (size_t)(o0 + o) * J.ldw, __uint_as_float(v[i]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

// Row chunks.  A launch holds 4-5 tiles that run over the particle rows (the rest are cheap per-jet tiles); a CTA streams its
// chunk at the latency-bound rate of one 32-row stage per ~2 k cycles, so the launch wants about two such CTAs per SM, but
// every extra CTA adds 16 K atomics onto the same dW tile.  Measured at 85 k rows (B = 1024 jets): 4096 rows per CTA 1.824 ms
// per training step, 2048: 1.775, 1792: 1.763, 1536: 1.764, 1280: 1.785, 512: 1.996.  Smaller batches keep ~48 chunks.
static int xty_tc_grid_x(int max_rows, int tiles, int sm_count) {
  (void)tiles; (void)sm_count;
  static const int forced = getenv("PFM_XTY_CHUNK") ? atoi(getenv("PFM_XTY_CHUNK")) : 0;
  int rows_per_cta = forced;
  if (rows_per_cta <= 0) {
    rows_per_cta = ((max_rows + 47) / 48 + XS_ROWS - 1) / XS_ROWS * XS_ROWS;
    rows_per_cta = rows_per_cta < 256 ? 256 : (rows_per_cta > 1792 ? 1792 : rows_per_cta);
  }
  const int gx = (max_rows + rows_per_cta - 1) / rows_per_cta;
  return gx < 1 ? 1 : gx;
}

static int xty_min_chunk() {
  static const int v = getenv("PFM_XTY_MINCHUNK") ? atoi(getenv("PFM_XTY_MINCHUNK")) : 8 * XS_ROWS;
  return v;
}

static int xty_tc_prepare() {
  static bool done = false;
  if (!done) {
    PFM_CUDA_CHECK(cudaFuncSetAttribute(xty_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(XtyTcSmem) + 1024));
    done = true;
  }
  return PFM_OK;
}

int xty_tc_launch(const XtyJob* jobs_dev, int n_jobs, int tiles, const int* n_total, int max_rows, int sm_count, cudaStream_t st) {
  if (n_jobs <= 0 || tiles <= 0 || max_rows <= 0) return PFM_OK;
  int rc = xty_tc_prepare();
  if (rc != PFM_OK) return rc;
  XtyJob dummy;
  memset(&dummy, 0, sizeof(dummy));
  dim3 grid((unsigned)xty_tc_grid_x(max_rows, tiles, sm_count), (unsigned)tiles);
  xty_tc_kernel<<<grid, 256, sizeof(XtyTcSmem) + 1024, st>>>(jobs_dev, n_jobs, n_total, dummy, 0, xty_min_chunk());
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int xty_tc_launch_one(const float* Y, int ldy, const float* X, int ldx, float* dW, int ldw, int out, int K, int col0, int rows,
                      int sm_count, cudaStream_t st) {
  if (rows <= 0 || out <= 0 || K <= 0) return PFM_OK;
  int rc = xty_tc_prepare();
  if (rc != PFM_OK) return rc;
  XtyJob J;
  J.Y = Y; J.X = X; J.dW = dW; J.ldy = ldy; J.ldx = ldx; J.ldw = ldw; J.out = out; J.K = K; J.col0 = col0; J.rows = rows; J.tile0 = 0;
  J.tiles_o = (out + 127) / 128; J.tiles_k = (K + 127) / 128;
  const int tiles = J.tiles_o * J.tiles_k;
  dim3 grid((unsigned)xty_tc_grid_x(rows, tiles, sm_count), (unsigned)tiles);
  xty_tc_kernel<<<grid, 256, sizeof(XtyTcSmem) + 1024, st>>>(nullptr, 0, nullptr, J, 1, xty_min_chunk());
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // namespace pfm
