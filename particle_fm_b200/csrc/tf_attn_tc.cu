// Masked self attention of the droid transformers on tcgen05 (bf16 mode):   A = softmax(Q K^T / sqrt(dh)) V
// per (jet, head) over the jet's REAL particles only (padding skipped), head dim 16.
//
// One CTA walks (jet, head) units.  Per unit:
//   loaders (warps 0-3, thread = key / query row) gather K, V and the query tile from the packed [rows, 3D] fp32 QKV
//     activations and write them as bf16 MMA operands into shared memory (K-major SWIZZLE_128B):
//       Q  [128 queries x 16]  A operand          K  [<=256 keys x 16]  B operand of S = Q K^T
//       Vt [16 dims x <=256 keys]                 B operand of O = P V   (keys are the contraction dim)
//   MMA warp (warp 4):  S[128 x Npad] = Q K^T          ONE tcgen05.mma (M=128, N=Npad, K=16), fp32 in TMEM
//   softmax (warps 0-3, thread = query row = TMEM lane): row max, p = exp2(s - max), row sum, P -> bf16 IN PLACE in TMEM
//   MMA warp:            O[128 x 16] = P V               Npad/16 TS-MMAs (A = P from TMEM), accumulator in TMEM cols 128..143
//   softmax threads:     A[row][head*16 ..] = O / sum     (fp32, 64 bytes per query)
// Jets with more than 128 particles take several query tiles against the same K / Vt.  256 TMEM columns and ~57 KB of
// shared memory per CTA, so two or three CTAs share an SM and overlap each other's serial phases.
#include "pfm_internal.cuh"
#include "tc_ptx.cuh"
#include "tf_internal.cuh"

namespace pfm {

using namespace tc;

static constexpr int AT_THREADS = 160;
static constexpr int AT_DH = 16;
static constexpr int AT_KMAX = 256;          // keys per jet handled by this kernel

struct AtSmem {
  alignas(1024) uint8_t Q[128 * 128];        // [128 rows][128 B]: 16 bf16 used per row (first K step)
  alignas(1024) uint8_t K[AT_KMAX * 128];    // [256 keys][128 B]
  alignas(1024) uint8_t Vt[4 * 2048];        // [16 dims][256 keys]: 4 blocks (64 keys) of 16 rows x 128 B
  uint64_t ops_ready, s_full, p_ready, o_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(AT_THREADS) tf_attn_tc_kernel(const float* __restrict__ QKV, int ld, int D, int heads,
                                                                const int* __restrict__ n_real, const int* __restrict__ rowoff,
                                                                float* __restrict__ A, int lda, float scale, int n_units) {
  extern __shared__ uint8_t smem_raw[];
  AtSmem& s = *reinterpret_cast<AtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&s.ops_ready, 128); mbar_init(&s.s_full, 1); mbar_init(&s.p_ready, 128); mbar_init(&s.o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(&s.tmem_base, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = s.tmem_base;
  const float sl2 = scale * 1.4426950408889634f;      // scores in log2 units
  uint32_t ph = 0;                                    // uses of every barrier so far (all four advance together)

  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int jet = unit / heads, head = unit - jet * heads;
    const int n = n_real[jet], r0 = rowoff[jet];
    if (n == 0) continue;                             // uniform over the CTA
    const int npad = (n + 15) & ~15;
    const int m_tiles = (n + 127) >> 7;
    for (int mt = 0; mt < m_tiles; ++mt, ++ph) {
      if (warp < 4) {
        // ---------------- operands ----------------
        if (mt == 0) {
          for (int key = tid; key < npad; key += 128) {
            float kf[16], vf[16];
            if (key < n) {
              const float4* src = reinterpret_cast<const float4*>(QKV + (size_t)(r0 + key) * ld + head * AT_DH);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 kv = __ldg(src + (D >> 2) + q), vv = __ldg(src + (D >> 1) + q);
                kf[q * 4 + 0] = kv.x; kf[q * 4 + 1] = kv.y; kf[q * 4 + 2] = kv.z; kf[q * 4 + 3] = kv.w;
                vf[q * 4 + 0] = vv.x; vf[q * 4 + 1] = vv.y; vf[q * 4 + 2] = vv.z; vf[q * 4 + 3] = vv.w;
              }
            } else {
#pragma unroll
              for (int d = 0; d < 16; ++d) { kf[d] = 0.f; vf[d] = 0.f; }
            }
            uint8_t* krow = s.K + (key >> 3) * 1024 + (key & 7) * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint4 pk;
              pk.x = pack_bf16x2(kf[c * 8 + 0], kf[c * 8 + 1]); pk.y = pack_bf16x2(kf[c * 8 + 2], kf[c * 8 + 3]);
              pk.z = pack_bf16x2(kf[c * 8 + 4], kf[c * 8 + 5]); pk.w = pack_bf16x2(kf[c * 8 + 6], kf[c * 8 + 7]);
              *reinterpret_cast<uint4*>(krow + ((c ^ (key & 7)) << 4)) = pk;
            }
#pragma unroll
            for (int d = 0; d < 16; ++d)
              *reinterpret_cast<__nv_bfloat16*>(s.Vt + sw128_offset(d, key, 2048)) = __float2bfloat16(vf[d]);
          }
        }
        {
          const int qrow = mt * 128 + tid;
          float qf[16];
          if (qrow < n) {
            const float4* src = reinterpret_cast<const float4*>(QKV + (size_t)(r0 + qrow) * ld + head * AT_DH);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 v = __ldg(src + q);
              qf[q * 4 + 0] = v.x * sl2; qf[q * 4 + 1] = v.y * sl2; qf[q * 4 + 2] = v.z * sl2; qf[q * 4 + 3] = v.w * sl2;
            }
          } else {
#pragma unroll
            for (int d = 0; d < 16; ++d) qf[d] = 0.f;
          }
          uint8_t* qr = s.Q + (tid >> 3) * 1024 + (tid & 7) * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 pk;
            pk.x = pack_bf16x2(qf[c * 8 + 0], qf[c * 8 + 1]); pk.y = pack_bf16x2(qf[c * 8 + 2], qf[c * 8 + 3]);
            pk.z = pack_bf16x2(qf[c * 8 + 4], qf[c * 8 + 5]); pk.w = pack_bf16x2(qf[c * 8 + 6], qf[c * 8 + 7]);
            *reinterpret_cast<uint4*>(qr + ((c ^ (tid & 7)) << 4)) = pk;
          }
        }
        fence_proxy_async();
        mbar_arrive(&s.ops_ready);
        // ---------------- softmax over the row of S ----------------
        mbar_wait(&s.s_full, ph & 1);
        tc_fence_after();
        const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
        float mx = -INFINITY;
        for (int c0 = 0; c0 < npad; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(lane_base + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i < n) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        float l = 0.f;
        for (int c0 = 0; c0 < npad; c0 += 32) {
          uint32_t v[32], pk[16];
          tmem_ld32(lane_base + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = c0 + i < n ? exp2f(__uint_as_float(v[i]) - mx) : 0.f;
            const float p1 = c0 + i + 1 < n ? exp2f(__uint_as_float(v[i + 1]) - mx) : 0.f;
            const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
            // the row sum uses the ROUNDED probabilities, so that O / l is exactly a weighted mean of the values
            l += __bfloat162float(pb.x) + __bfloat162float(pb.y);
            pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&pb);
          }
          tmem_st16(lane_base + (c0 >> 1), pk);       // P columns [c0/2, c0/2+16) overlap only S columns already read
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&s.p_ready);
        // ---------------- output ----------------
        mbar_wait(&s.o_full, ph & 1);
        tc_fence_after();
        uint32_t o[16];
        tmem_ld16(lane_base + 128, o);
        tmem_wait_ld();
        tc_fence_before();
        const int qrow = mt * 128 + tid;
        if (qrow < n) {
          const float inv = 1.f / l;
          float4* dst = reinterpret_cast<float4*>(A + (size_t)(r0 + qrow) * lda + head * AT_DH);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            dst[q] = make_float4(__uint_as_float(o[q * 4]) * inv, __uint_as_float(o[q * 4 + 1]) * inv,
                                 __uint_as_float(o[q * 4 + 2]) * inv, __uint_as_float(o[q * 4 + 3]) * inv);
        }
      } else {
        // ---------------- MMA issuer ----------------
        const uint64_t qdesc = desc_kmajor(smem_u32(s.Q)), kdesc = desc_kmajor(smem_u32(s.K)), vdesc = desc_kmajor(smem_u32(s.Vt));
        mbar_wait(&s.ops_ready, ph & 1);
        tc_fence_after();
        if (elect_one()) {
          mma_ss(tm, qdesc, kdesc, make_idesc_bf16(128, npad, 0, 0), 0u);           // S = Q K^T  (one K=16 step)
          mma_commit(&s.s_full);
        }
        __syncwarp();
        mbar_wait(&s.p_ready, ph & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t idesc_pv = make_idesc_bf16(128, 16, 0, 0);
          for (int k = 0; k < (npad >> 4); ++k)                                     // O += P[:, 16k..16k+16) . V[16k..16k+16, :]
            mma_ts(tm + 128, tm + (uint32_t)k * 8u, vdesc + (uint64_t)((k >> 2) * 128 + (k & 3) * 2), idesc_pv, k ? 1u : 0u);
          mma_commit(&s.o_full);
        }
        __syncwarp();
      }
    }
    // K / Vt / Q of this unit must not be overwritten while a straggler still reads TMEM results computed from them: the
    // o_full wait above orders every thread after the last MMA of the unit, and the next unit's MMAs wait for ops_ready.
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 256);
}

int tf_attn_tc(const float* QKV, int ld, int D, int heads, const int* n_real, const int* rowoff, float* A, int lda, float scale,
               int B, int N, int sm_count, cudaStream_t st) {
  if (D / heads != AT_DH || N > AT_KMAX) { set_error("tensor-core attention needs head dim 16 and <= %d particles", AT_KMAX); return PFM_ERR_UNSUPPORTED; }
  const int smem = (int)sizeof(AtSmem) + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tf_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  const int n_units = B * heads;
  const int grid = n_units < 2 * sm_count ? n_units : 2 * sm_count;
  tf_attn_tc_kernel<<<grid, AT_THREADS, smem, st>>>(QKV, ld, D, heads, n_real, rowoff, A, lda, scale, n_units);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // namespace pfm
