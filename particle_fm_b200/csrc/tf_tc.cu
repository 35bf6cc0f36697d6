// bf16 tcgen05 version of the droid transformers' fused linear:
//     Y[rows, N] = [R +] act( LN(X[rows, K]) . W^T + bias [+ per-jet bias] )
// One CTA owns a block of 128 rows.  Warps 0-3 load the rows (coalesced fp32 reads), apply the LayerNorm in registers
// (warp-shuffle statistics), and write them ONCE as the bf16 A operand (K-major SWIZZLE_128B tile, resident for all
// column tiles); warp 4 streams the pre-swizzled weight image through a ring of 16 KB [128 n x 64 k] blocks with bulk
// async copies; warp 5 issues tcgen05.mma (M=128, N=128, K=16, fp32 accumulators in TMEM, two accumulators so that the
// epilogue of column tile t overlaps the MMAs of tile t+1); warps 0-3 then run the epilogue from TMEM (thread = row):
// bias, per-jet bias, activation, residual, 128-byte row-segment stores.
#include "pfm_internal.cuh"
#include "tc_ptx.cuh"
#include "tf_internal.cuh"

namespace pfm {

using namespace tc;

static constexpr int TT_THREADS = 192;
static constexpr uint32_t TT_BLK = 16384;       // one [128 x 64] bf16 block

struct TtSmemTail {
  uint64_t full[4], empty[4], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(TT_THREADS, 2) tf_linear_tc_kernel(const LinArgs a, int n_slots) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KB = a.K >> 6;                                   // k blocks of the A tile
  uint8_t* Atile = base;                                     // KB blocks of 16 KB
  uint8_t* ring = base + (size_t)KB * TT_BLK;                // n_slots blocks of 16 KB
  TtSmemTail& t = *reinterpret_cast<TtSmemTail*>(ring + (size_t)n_slots * TT_BLK);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * 128;
  const int n_tiles = a.N >> 7;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(&t.full[i], 1); mbar_init(&t.empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t.acc_full[i], 1); mbar_init(&t.acc_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&t.tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = t.tmem_base;

  if (warp < 4) {
    // ================= load rows, LayerNorm, bf16 A tile =================
    const int nchunks = a.K >> 3;                            // 16-byte (8 x bf16) chunks per row
    constexpr int RU = 4;                                    // rows in flight per warp: the global loads of 4 rows overlap
#pragma unroll 1
    for (int rr0 = 0; rr0 < 32; rr0 += RU) {
      float v[RU][2][8];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int row = row0 + warp * 32 + rr0 + u;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int ch = p * 32 + lane;
#pragma unroll
          for (int e = 0; e < 8; ++e) v[u][p][e] = 0.f;
          if (ch < nchunks && row < a.rows) {
            const float4* src = reinterpret_cast<const float4*>(a.X + (size_t)row * a.ldx + ch * 8);
            const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
            v[u][p][0] = x0.x; v[u][p][1] = x0.y; v[u][p][2] = x0.z; v[u][p][3] = x0.w;
            v[u][p][4] = x1.x; v[u][p][5] = x1.y; v[u][p][6] = x1.z; v[u][p][7] = x1.w;
          }
        }
      }
      if (a.ln_g) {
        float s[RU], q[RU];
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          s[u] = 0.f;
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int e = 0; e < 8; ++e) s[u] += v[u][p][e];
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
          for (int u = 0; u < RU; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], sh);
#pragma unroll
        for (int u = 0; u < RU; ++u) {
          s[u] /= (float)a.K;                                // mean
          q[u] = 0.f;
#pragma unroll
          for (int p = 0; p < 2; ++p)
            if (p * 32 + lane < nchunks)
#pragma unroll
              for (int e = 0; e < 8; ++e) { const float d = v[u][p][e] - s[u]; q[u] = fmaf(d, d, q[u]); }
        }
#pragma unroll
        for (int sh = 16; sh > 0; sh >>= 1)
#pragma unroll
          for (int u = 0; u < RU; ++u) q[u] += __shfl_xor_sync(0xffffffffu, q[u], sh);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int ch = p * 32 + lane;
          if (ch < nchunks) {
            const float4* gp = reinterpret_cast<const float4*>(a.ln_g + ch * 8);
            const float4* bp = reinterpret_cast<const float4*>(a.ln_b + ch * 8);
            const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1), b0 = __ldg(bp), b1 = __ldg(bp + 1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int u = 0; u < RU; ++u) {
              const float rstd = rsqrtf(q[u] / (float)a.K + a.eps);
              const bool live = row0 + warp * 32 + rr0 + u < a.rows;
#pragma unroll
              for (int e = 0; e < 8; ++e) v[u][p][e] = live ? (v[u][p][e] - s[u]) * rstd * g[e] + bb[e] : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int r = warp * 32 + rr0 + u;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
          const int ch = p * 32 + lane;
          if (ch < nchunks) {
            uint4 pk;
            pk.x = pack_bf16x2(v[u][p][0], v[u][p][1]); pk.y = pack_bf16x2(v[u][p][2], v[u][p][3]);
            pk.z = pack_bf16x2(v[u][p][4], v[u][p][5]); pk.w = pack_bf16x2(v[u][p][6], v[u][p][7]);
            uint8_t* dst = Atile + (size_t)(ch >> 3) * TT_BLK + (r >> 3) * 1024 + (r & 7) * 128 + (((ch & 7) ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = pk;
          }
        }
      }
    }
    fence_proxy_async();
  }
  __syncthreads();          // the A tile is complete and visible to the tensor core

  if (warp == 4) {
    // ================= weight producer =================
    uint32_t it = 0;
    for (int nt = 0; nt < n_tiles; ++nt)
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const uint32_t slot = it % n_slots, round = it / n_slots;
        mbar_wait(&t.empty[slot], (round & 1) ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&t.full[slot], TT_BLK);
          bulk_copy_g2s(ring + (size_t)slot * TT_BLK, a.img + ((size_t)nt * a.img_kblocks + a.kb0 + kb) * TT_BLK, TT_BLK, &t.full[slot]);
        }
        __syncwarp();
      }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    const uint64_t adesc = desc_kmajor(smem_u32(Atile));
    const uint64_t rdesc = desc_kmajor(smem_u32(ring));
    uint32_t it = 0;
    for (int nt = 0; nt < n_tiles; ++nt) {
      const int buf = nt & 1;
      mbar_wait(&t.acc_empty[buf], ((nt >> 1) & 1) ^ 1);          // the epilogue of tile nt-2 has drained this accumulator
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb, ++it) {
        const uint32_t slot = it % n_slots, round = it / n_slots;
        mbar_wait(&t.full[slot], round & 1);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ss(tm + buf * 128, adesc + (uint64_t)(kb * (TT_BLK >> 4) + ks * 2), rdesc + (uint64_t)(slot * (TT_BLK >> 4) + ks * 2), idesc,
                   (kb | ks) ? 1u : 0u);
          mma_commit(&t.empty[slot]);
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(&t.acc_full[buf]);
      __syncwarp();
    }
  } else if (warp < 4) {
    // ================= epilogue: thread = row = TMEM lane =================
    const int r = tid, row = row0 + r;
    const bool live = row < a.rows;
    const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
    const float* jbrow = (a.jb && live) ? a.jb + (size_t)(a.rowjet ? a.rowjet[row] : row) * a.jb_stride : nullptr;
    float* stg = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(&t) + 256) + warp * (32 * 17);   // [32 rows][16 + 1]
    for (int nt = 0; nt < n_tiles; ++nt) {
      const int buf = nt & 1;
      mbar_wait(&t.acc_full[buf], (nt >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(lane_base + buf * 128 + c * 32, v);
        tmem_wait_ld();
        // 16 columns at a time through a warp-private shared-memory tile: the residual is READ and the result WRITTEN as
        // 64-byte row segments (two rows per instruction) instead of one 16-byte piece per lane and row, and the per-thread
        // bias loads of a half are all in flight together
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int n0 = nt * 128 + c * 32 + half * 16;
          float4 bq[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) bq[q] = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + n0 + q * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (jbrow) {
            float4 jv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) jv[q] = __ldg(reinterpret_cast<const float4*>(jbrow + n0 + q * 4));
#pragma unroll
            for (int q = 0; q < 4; ++q) { bq[q].x += jv[q].x; bq[q].y += jv[q].y; bq[q].z += jv[q].z; bq[q].w += jv[q].w; }
          }
          if (a.R) {
#pragma unroll
            for (int it = 0; it < 16; ++it) {
              const int rr = it * 2 + (lane >> 4), cc = lane & 15;
              const int grow = row0 + warp * 32 + rr;
              stg[rr * 17 + cc] = grow < a.rows ? a.R[(size_t)grow * a.ldr + n0 + cc] : 0.f;
            }
            __syncwarp();
          }
          float o[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float badd[4] = {bq[q].x, bq[q].y, bq[q].z, bq[q].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x = __uint_as_float(v[half * 16 + q * 4 + e]) + badd[e];
              if (a.act) x = x > 0.f ? x : x * a.slope;
              if (a.R) x += stg[lane * 17 + q * 4 + e];
              o[q * 4 + e] = x;
            }
          }
#pragma unroll
          for (int e = 0; e < 16; ++e) stg[lane * 17 + e] = o[e];
          __syncwarp();
#pragma unroll
          for (int it = 0; it < 16; ++it) {
            const int rr = it * 2 + (lane >> 4), cc = lane & 15;
            const int grow = row0 + warp * 32 + rr;
            if (grow < a.rows) a.Y[(size_t)grow * a.ldy + n0 + cc] = stg[rr * 17 + cc];
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(&t.acc_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 256);
}

__global__ void tf_pack_bf16_kernel(const float* __restrict__ Wt, int in, int out, int ldo, uint8_t* __restrict__ img, int kblocks) {
  const int n_tiles = (out + 127) / 128;
  const long long total = (long long)n_tiles * kblocks * 128 * 64;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int n = (int)(idx % 128);                   // consecutive threads -> consecutive n: coalesced reads of the k-major copy
  const int k = (int)((idx / 128) % 64);
  const int blk = (int)(idx / (128 * 64));
  const int nt = blk / kblocks, kb = blk - nt * kblocks;
  const int gn = nt * 128 + n, gk = kb * 64 + k;
  const float w = (gn < out && gk < in) ? Wt[(size_t)gk * ldo + gn] : 0.f;
  *reinterpret_cast<__nv_bfloat16*>(img + (size_t)blk * TT_BLK + sw128_offset(n, k, 0)) = __float2bfloat16(w);
}

bool tf_tc_linear_supported(const LinArgs& a) {
  return a.img != nullptr && a.K >= 64 && a.K <= 512 && (a.K & 63) == 0 && a.N >= 128 && (a.N & 127) == 0 && (a.ldx & 3) == 0 &&
         (a.ldy & 3) == 0 && (!a.R || (a.ldr & 3) == 0) && (!a.jb || (a.jb_stride & 3) == 0);
}

int tf_tc_linear(const LinArgs& a, int max_smem, cudaStream_t st) {
  const int KB = a.K >> 6;
  int n_slots = KB <= 4 ? 2 : 4;            // K <= 256: 96 KB per CTA -> two CTAs per SM overlap load / MMA / epilogue
  const size_t tail = 256 + 4 * 32 * 17 * sizeof(float) + 1024;      // barriers, epilogue staging tiles, alignment slack
  size_t smem = (size_t)(KB + n_slots) * TT_BLK + tail;
  if ((int)smem > max_smem) { n_slots = 2; smem = (size_t)(KB + n_slots) * TT_BLK + tail; }
  if ((int)smem > max_smem) { set_error("bf16 transformer linear: K=%d does not fit shared memory", a.K); return PFM_ERR_UNSUPPORTED; }
  static bool attr_set = false;
  if (!attr_set) {
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set = true;
  }
  tf_linear_tc_kernel<<<(a.rows + 127) / 128, TT_THREADS, smem, st>>>(a, n_slots);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int tf_tc_pack(const float* Wt, int in, int out, int ldo, uint8_t* img, int kblocks, cudaStream_t st) {
  const int n_tiles = (out + 127) / 128;
  const long long total = (long long)n_tiles * kblocks * 128 * 64;
  tf_pack_bf16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(Wt, in, out, ldo, img, kblocks);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // namespace pfm
