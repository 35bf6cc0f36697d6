// C ABI of libpfm_b200.so: handle management, weight repacking, jet packing plan, hoisted bias
// tables, dispatch to the fp32 CUDA-core or bf16 tcgen05 kernels.  See include/pfm_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cstdlib>

#include "pfm_internal.cuh"

namespace pfm {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// weight repack:  Wt[k*ldo + o] = W[o*in + k]   (k-major fp32 copy, zero in the o-padding)
// ---------------------------------------------------------------------------------------------
__global__ void transpose_weight_kernel(const float* __restrict__ W, float* __restrict__ Wt, int out, int in,
                                        int ldo) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int total = in * ldo;
  if (idx >= total) return;
  int k = idx / ldo, o = idx - k * ldo;
  Wt[idx] = (o < out) ? W[(size_t)o * in + k] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// hoisted bias tables
//   tbias[row][boff + o] = b[o] + sum_k W[o][t_off + k] * code[row][k]   (+ input-time columns of fc_l1)
//   cbias[jet][boff + o] =        sum_k W[o][c_off + k] * cond[jet][k]
// grid = (rows, n_lin), block = 128 threads striding over the outputs
// ---------------------------------------------------------------------------------------------
__global__ void tbias_kernel(const Lin* __restrict__ lin, const float* __restrict__ code, int t_dim,
                             const float* __restrict__ code_in, int t_in, float* __restrict__ tbias, int bstride) {
  const Lin L = lin[blockIdx.y];
  const int row = blockIdx.x;
  for (int o = threadIdx.x; o < L.out; o += blockDim.x) {
    float acc = L.b[o];
    if (L.t_len > 0) {
      const float* c = code + (size_t)row * t_dim;
      for (int k = 0; k < L.t_len; ++k) acc = fmaf(L.Wt[(size_t)(L.t_off + k) * L.ldo + o], c[k], acc);
    }
    if (blockIdx.y == 0 && t_in > 0) {   // add_time_to_input: first t_in columns of fc_l1's main block
      const float* c = code_in + (size_t)row * t_in;
      for (int k = 0; k < t_in; ++k) acc = fmaf(L.Wt[(size_t)(L.m_off + k) * L.ldo + o], c[k], acc);
    }
    tbias[(size_t)row * bstride + L.bias_off + o] = acc;
  }
}

// The same table for many rows (training: one row per jet): a CTA handles TB_ROWS rows of one linear, so a thread loads the
// time columns of its output's weights once per TB_ROWS rows (32 independent loads, one latency) and reads the codes from
// shared memory.  Per (row, output) the accumulation order is tbias_kernel's (bias, then k ascending): identical bits.
static constexpr int TB_ROWS = 8;
static constexpr int TB_KMAX = 64;         // time-code width handled here (2 * frequencies = 32 in the shipped configs)

__global__ void __launch_bounds__(128, 8) tbias_rows_kernel(const Lin* __restrict__ lin, const float* __restrict__ code, int t_dim,
                                                         const float* __restrict__ code_in, int t_in, float* __restrict__ tbias,
                                                         int bstride, int rows) {
  __shared__ __align__(16) float sc[TB_ROWS][TB_KMAX];
  const Lin L = lin[blockIdx.y];
  const int row0 = blockIdx.x * TB_ROWS;
  const int nr = rows - row0 < TB_ROWS ? rows - row0 : TB_ROWS;
  const int T = L.t_len;                                   // <= TB_KMAX (checked by the caller)
  const bool t4 = (T & 3) == 0;
  for (int i = threadIdx.x; i < TB_ROWS * TB_KMAX; i += blockDim.x) {
    const int r = i / TB_KMAX, k = i % TB_KMAX;
    sc[r][k] = (r < nr && k < T) ? code[(size_t)(row0 + r) * t_dim + k] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < L.out; o += blockDim.x) {
    float acc[TB_ROWS];
    const float b = L.b[o];
#pragma unroll
    for (int r = 0; r < TB_ROWS; ++r) acc[r] = b;
    for (int k0 = 0; k0 < T; k0 += 32) {
      float w[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) w[k] = k0 + k < T ? __ldg(L.Wt + (size_t)(L.t_off + k0 + k) * L.ldo + o) : 0.f;
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        if (k0 + k < T) {
#pragma unroll
          for (int r = 0; r < TB_ROWS; ++r) {
            const float4 c = *reinterpret_cast<const float4*>(&sc[r][k0 + k]);
            acc[r] = fmaf(w[k], c.x, acc[r]);
            if (t4) {                                       // T a multiple of 4 (every shipped config): no per-term guard
              acc[r] = fmaf(w[k + 1], c.y, acc[r]);
              acc[r] = fmaf(w[k + 2], c.z, acc[r]);
              acc[r] = fmaf(w[k + 3], c.w, acc[r]);
            } else {
              if (k0 + k + 1 < T) acc[r] = fmaf(w[k + 1], c.y, acc[r]);
              if (k0 + k + 2 < T) acc[r] = fmaf(w[k + 2], c.z, acc[r]);
              if (k0 + k + 3 < T) acc[r] = fmaf(w[k + 3], c.w, acc[r]);
            }
          }
        }
      }
    }
    for (int r = 0; r < nr; ++r) {
      float a = acc[r];
      if (blockIdx.y == 0 && t_in > 0) {   // add_time_to_input: first t_in columns of fc_l1's main block
        const float* c = code_in + (size_t)(row0 + r) * t_in;
        for (int k = 0; k < t_in; ++k) a = fmaf(L.Wt[(size_t)(L.m_off + k) * L.ldo + o], c[k], a);
      }
      tbias[(size_t)(row0 + r) * bstride + L.bias_off + o] = a;
    }
  }
}

__global__ void cbias_kernel(const Lin* __restrict__ lin, const float* __restrict__ cond, int cond_dim,
                             float* __restrict__ cbias, int bstride) {
  const Lin L = lin[blockIdx.y];
  const int jet = blockIdx.x;
  const float* c = cond + (size_t)jet * cond_dim;
  for (int o = threadIdx.x; o < L.out; o += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < L.c_len; ++k) acc = fmaf(L.Wt[(size_t)(L.c_off + k) * L.ldo + o], c[k], acc);
    cbias[(size_t)jet * bstride + L.bias_off + o] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// plan: count + compact the real particles of every jet (one warp per jet), then pack consecutive
// jets greedily into groups of <= R_cap rows and <= J_cap jets (one CTA work item each).
// ---------------------------------------------------------------------------------------------
__global__ void plan_count_kernel(const float* __restrict__ mask, int B, int N, int* __restrict__ n_real,
                                  uint16_t* __restrict__ ridx) {
  int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (jet >= B) return;
  int count = 0;
  for (int p0 = 0; p0 < N; p0 += 32) {
    int p = p0 + lane;
    bool real = p < N && (mask == nullptr || mask[(size_t)jet * N + p] != 0.f);
    unsigned bal = __ballot_sync(0xffffffffu, real);
    if (real) ridx[(size_t)jet * N + count + __popc(bal & ((1u << lane) - 1))] = (uint16_t)p;
    count += __popc(bal);
  }
  if (lane == 0) n_real[jet] = count;
}

__global__ void plan_group_kernel(const int* __restrict__ n_real, int B, int R_cap, int J_cap,
                                  int2* __restrict__ groups, int* __restrict__ n_groups, int* __restrict__ counter,
                                  int* __restrict__ rowoff, int* __restrict__ n_total) {
  extern __shared__ int s_n[];
  for (int i = threadIdx.x; i < B; i += blockDim.x) s_n[i] = n_real[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int g = 0, first = 0, rows = 0, cnt = 0, total = 0;
    for (int j = 0; j < B; ++j) {
      int n = s_n[j];
      rowoff[j] = total;
      total += n;
      if (cnt > 0 && (rows + n > R_cap || cnt >= J_cap)) {
        groups[g++] = make_int2(first, cnt);
        first = j; rows = 0; cnt = 0;
      }
      rows += n; cnt += 1;
    }
    if (cnt > 0) groups[g++] = make_int2(first, cnt);
    *n_groups = g;
    *counter = 0;
    *n_total = total;
  }
}

// Inference plan: bin-pack the jets into groups of <= R_cap rows / <= J_cap jets (best-fit decreasing on a histogram of
// the multiplicities).  The fused kernels are latency-bound per group, so their time is proportional to the NUMBER of
// groups: packing jets of mixed sizes (JetNet-150: 15..150 particles, 256-row groups) raises the fill from ~80 % (greedy
// over consecutive jets) to ~97 %.  Deterministic (no atomics in the ordering).  One block of PP_WARPS warps: every warp
// packs a contiguous range of the jets on its own (its own histogram in shared memory, stable counting sort with
// warp-match ranks, best-fit loop), then the per-warp group lists are compacted into one -- the price is at most one
// partly filled group per warp.
static constexpr int PP_WARPS = 8;

__global__ void __launch_bounds__(PP_WARPS * 32) plan_pack_kernel(const int* __restrict__ n_real, int B, int R_cap, int J_cap,
                                                                  int2* __restrict__ groups, int2* __restrict__ groups_tmp,
                                                                  int* __restrict__ n_groups, int* __restrict__ counter,
                                                                  int* __restrict__ jetmap, int* __restrict__ order) {
  extern __shared__ int sp[];
  __shared__ int g_of[PP_WARPS + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int stride = R_cap + 1;
  int* cnt = sp + warp * 3 * stride;   // [R_cap + 1] jets per multiplicity
  int* start = cnt + stride;           // [R_cap + 1] first slot of the bucket in this warp's part of `order` (descending multiplicity)
  int* cur = cnt + 2 * stride;         // [R_cap + 1] used entries of the bucket
  const int per = (B + PP_WARPS - 1) / PP_WARPS;
  const int seg0 = min(B, warp * per), seg1 = min(B, seg0 + per);
  const int nseg = seg1 - seg0;
  for (int i = lane; i <= R_cap; i += 32) { cnt[i] = 0; cur[i] = 0; }
  __syncwarp();
  for (int i = seg0 + lane; i < seg1; i += 32) atomicAdd(&cnt[min(n_real[i], R_cap)], 1);
  __syncwarp();
  if (lane == 0) {
    int pos = 0;
    for (int n = R_cap; n >= 0; --n) { start[n] = pos; pos += cnt[n]; }
  }
  __syncwarp();
  // stable counting sort: jets of equal multiplicity keep their batch order (rank inside a block of 32 by warp match)
  for (int i0 = seg0; i0 < seg1; i0 += 32) {
    const int i = i0 + lane;
    const bool valid = i < seg1;
    const int n = valid ? min(n_real[i], R_cap) : -1 - lane;            // invalid lanes: unique keys
    const unsigned same = __match_any_sync(0xffffffffu, n);
    const int leader = __ffs(same) - 1;
    const int rank = __popc(same & ((1u << lane) - 1u));
    int base = 0;
    if (valid && lane == leader) { base = cur[n]; cur[n] = base + __popc(same); }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) order[seg0 + start[n] + base + rank] = i;
    __syncwarp();
  }
  for (int n = lane; n <= R_cap; n += 32) cur[n] = 0;
  __syncwarp();
  // largest multiplicity n in [1, lim] with an unused jet, or 0
  auto find_le = [&](int lim) -> int {
    for (int base = lim; base >= 1; base -= 32) {
      const int n = base - lane;
      const bool ok = n >= 1 && cnt[n] - cur[n] > 0;
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (bal) return base - (__ffs(bal) - 1);
    }
    return 0;
  };
  int pos = 0, g = 0, top = R_cap;
  int remaining = nseg - cnt[0], zero_left = cnt[0];
  while (remaining > 0) {
    int n = find_le(top);
    top = n;
    const int first = pos;
    int rows = 0, jets = 0;
    while (n > 0) {
      if (lane == 0) { jetmap[seg0 + pos] = order[seg0 + start[n] + cur[n]]; cur[n]++; }
      __syncwarp();
      ++pos; rows += n; ++jets; --remaining;
      if (jets >= J_cap || remaining == 0) break;
      const int lim = (R_cap - rows) < top ? (R_cap - rows) : top;
      n = lim >= 1 ? find_le(lim) : 0;
    }
    while (jets < J_cap && zero_left > 0) {        // empty jets ride along in spare jet slots
      if (lane == 0) { jetmap[seg0 + pos] = order[seg0 + start[0] + cur[0]]; cur[0]++; }
      __syncwarp();
      ++pos; ++jets; --zero_left;
    }
    if (lane == 0) groups_tmp[seg0 + g] = make_int2(seg0 + first, jets);
    ++g;
  }
  while (zero_left > 0) {
    const int first = pos;
    int jets = 0;
    while (jets < J_cap && zero_left > 0) {
      if (lane == 0) { jetmap[seg0 + pos] = order[seg0 + start[0] + cur[0]]; cur[0]++; }
      __syncwarp();
      ++pos; ++jets; --zero_left;
    }
    if (lane == 0) groups_tmp[seg0 + g] = make_int2(seg0 + first, jets);
    ++g;
  }
  if (lane == 0) g_of[warp + 1] = g;
  __syncthreads();
  if (tid == 0) {
    g_of[0] = 0;
    for (int w = 0; w < PP_WARPS; ++w) g_of[w + 1] += g_of[w];
    *n_groups = g_of[PP_WARPS];
    *counter = 0;
  }
  __syncthreads();
  for (int w = 0; w < PP_WARPS; ++w) {
    const int s0 = min(B, w * per), cntw = g_of[w + 1] - g_of[w];
    for (int i = tid; i < cntw; i += blockDim.x) groups[g_of[w] + i] = groups_tmp[s0 + i];
  }
}

__global__ void rowmajor_main_kernel(const float* __restrict__ W, float* __restrict__ Wr, int out, int in, int m_off,
                                     int m_len, int ldr) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= out * ldr) return;
  int o = idx / ldr, k = idx - o * ldr;
  Wr[idx] = k < m_len ? W[(size_t)o * in + m_off + k] : 0.f;
}

static int ensure_plan(pfm_epic* h, int B, int N) {
  Plan& p = h->plan;
  if (B > p.capB) {
    if (p.n_real) cudaFree(p.n_real);
    if (p.groups) cudaFree(p.groups);
    if (p.rowoff) cudaFree(p.rowoff);
    if (p.jetmap) cudaFree(p.jetmap);
    if (p.order) cudaFree(p.order);
    if (p.groups_tmp) cudaFree(p.groups_tmp);
    PFM_CUDA_CHECK(cudaMalloc(&p.groups_tmp, sizeof(int2) * B));
    PFM_CUDA_CHECK(cudaMalloc(&p.jetmap, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&p.order, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&p.n_real, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&p.groups, sizeof(int2) * B));
    PFM_CUDA_CHECK(cudaMalloc(&p.rowoff, sizeof(int) * B));
    p.capB = B;
  }
  if ((long long)B * N > p.capBN) {
    if (p.ridx) cudaFree(p.ridx);
    PFM_CUDA_CHECK(cudaMalloc(&p.ridx, sizeof(uint16_t) * (size_t)B * N));
    p.capBN = B * N;
  }
  if (!p.n_groups) {
    PFM_CUDA_CHECK(cudaMalloc(&p.n_groups, sizeof(int)));
    PFM_CUDA_CHECK(cudaMalloc(&p.counter, sizeof(int)));
    PFM_CUDA_CHECK(cudaMalloc(&p.n_total, sizeof(int)));
  }
  return PFM_OK;
}

static int ensure_floats(float** buf, size_t* cap, size_t need) {
  if (need > *cap) {
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(buf, sizeof(float) * need));
    *cap = need;
  }
  return PFM_OK;
}

static const int kMaxJetsPerCall = 12000;   // training: plan_group_kernel stages n_real in 48 KB of shared memory
static const int kMaxJetsPerLaunch = 1 << 20;   // sampling / forward: jets per fused launch (larger requests run in slices)

// Everything one hot-path call needs besides the kernel itself.
static int run_chunked(pfm_epic* h, const float* t_code, int t_rows, bool per_jet_t, const float* t_code_in,
                       int t_in, const float* x_in, float* x_out, const float* mask, const float* cond,
                       int B, int N, int Kx, int xin_off, int n_evals, int solver, int n_steps,
                       const float* dt, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  if (!h->weights_set) { set_error("weights not set (call pfm_epic_set_weights first)"); return PFM_ERR_STATE; }
  if (B <= 0 || N <= 0) { set_error("B and N must be positive (B=%d N=%d)", B, N); return PFM_ERR_INVALID; }
  if (N > 65535) { set_error("N=%d exceeds 65535", N); return PFM_ERR_INVALID; }
  const int cond_dim = c.global_cond_dim > c.local_cond_dim ? c.global_cond_dim : c.local_cond_dim;
  if (cond_dim > 0 && cond == nullptr) { set_error("cond is NULL but the net is conditioned"); return PFM_ERR_INVALID; }
  const bool any_t = (c.t_local_cat || c.t_global_cat) && c.t_dim > 0;
  if ((any_t && t_code == nullptr) || (t_in > 0 && t_code_in == nullptr)) {
    set_error("time code is NULL but the net takes a time code"); return PFM_ERR_INVALID;
  }
  cudaError_t e0 = cudaSetDevice(h->device);
  if (e0 != cudaSuccess) { set_error("cudaSetDevice(%d): %s", h->device, cudaGetErrorString(e0)); return PFM_ERR_CUDA; }
  // this call rewrites the handle-wide plan buffers (n_real, ridx, groups, counter): a pfm_epic_forward_train still waiting
  // for its pfm_epic_backward is invalidated here, so that a stale backward fails with PFM_ERR_STATE instead of running
  // on a foreign plan
  h->train_B = 0;
  h->last_launches = 0;
  h->last_groups_host = 0;
  h->ev_used = 0;

  int R_cap = 0, J_cap = 0, rc;
  if (h->precision == PFM_PREC_BF16) {
    rc = tc_supported(h, N);
    if (rc != PFM_OK) return rc;
    rc = tc_plan_caps(h, N, &R_cap, &J_cap);
  } else {
    rc = simt_plan_caps(h, N, &R_cap, &J_cap);
  }
  if (rc != PFM_OK) return rc;

  // time bias table: shared by all jets (rows = evaluations) or one row per jet
  const int trows = per_jet_t ? B : (t_rows > 0 ? t_rows : 1);
  rc = ensure_floats(&h->tbias, &h->tbias_cap, (size_t)trows * h->bstride);
  if (rc != PFM_OK) return rc;
  tbias_kernel<<<dim3(trows, h->n_lin), 128, 0, st>>>(h->lin_dev, t_code, c.t_dim, t_code_in, t_in, h->tbias,
                                                     h->bstride);
  h->last_launches++;
  if (cond_dim > 0) {
    rc = ensure_floats(&h->cbias, &h->cbias_cap, (size_t)B * h->bstride);
    if (rc != PFM_OK) return rc;
    cbias_kernel<<<dim3(B, h->n_lin), 128, 0, st>>>(h->lin_dev, cond, cond_dim, h->cbias, h->bstride);
    h->last_launches++;
  }
  PFM_CUDA_CHECK(cudaGetLastError());

  for (int b0 = 0; b0 < B; b0 += kMaxJetsPerLaunch) {
    const int nb = (B - b0 < kMaxJetsPerLaunch) ? (B - b0) : kMaxJetsPerLaunch;
    rc = ensure_plan(h, nb, N);
    if (rc != PFM_OK) return rc;
    const float* mk = mask ? mask + (size_t)b0 * N : nullptr;
    plan_count_kernel<<<(nb + 7) / 8, 256, 0, st>>>(mk, nb, N, h->plan.n_real, h->plan.ridx);
    const size_t pp_smem = sizeof(int) * 3 * (size_t)(R_cap + 1) * PP_WARPS;
    if (pp_smem > 48 * 1024) PFM_CUDA_CHECK(cudaFuncSetAttribute(plan_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pp_smem));
    plan_pack_kernel<<<1, PP_WARPS * 32, pp_smem, st>>>(
        h->plan.n_real, nb, R_cap, J_cap, h->plan.groups, h->plan.groups_tmp, h->plan.n_groups, h->plan.counter, h->plan.jetmap,
        h->plan.order);
    h->last_launches += 2;
    PFM_CUDA_CHECK(cudaGetLastError());
    RunArgs a;
    a.jetmap = h->plan.jetmap;
    a.x_in = x_in + (size_t)b0 * N * Kx;
    a.x_out = x_out + (size_t)b0 * N * c.feats;
    a.B = nb; a.N = N; a.Kx = Kx; a.xin_off = xin_off;
    a.n_evals = n_evals; a.solver = solver; a.n_steps = n_steps; a.dt = dt;
    a.tbias_per_jet = per_jet_t ? 1 : 0;
    a.has_cbias = cond_dim > 0;
    a.step_kind = h->step_kind; a.coef = h->step_coef; a.noise = h->step_noise;
    a.noise_step_stride = (long long)B * N * c.feats; a.jet0 = b0;
    // per-jet tables are indexed by the jet index inside this chunk
    float* tb_save = h->tbias; float* cb_save = h->cbias;
    if (per_jet_t) h->tbias += (size_t)b0 * h->bstride;
    if (cond_dim > 0) h->cbias += (size_t)b0 * h->bstride;
    cudaEvent_t e_start = nullptr, e_stop = nullptr;
    if (h->timing) {
      while ((int)h->ev_pool.size() < h->ev_used + 2) {
        cudaEvent_t e;
        PFM_CUDA_CHECK(cudaEventCreate(&e));
        h->ev_pool.push_back(e);
      }
      e_start = h->ev_pool[h->ev_used]; e_stop = h->ev_pool[h->ev_used + 1];
      h->ev_used += 2;
      PFM_CUDA_CHECK(cudaEventRecord(e_start, st));
    }
    rc = (h->precision == PFM_PREC_BF16) ? tc_run(h, a, st) : simt_run(h, a, st);
    if (h->timing && rc == PFM_OK) PFM_CUDA_CHECK(cudaEventRecord(e_stop, st));
    h->tbias = tb_save; h->cbias = cb_save;
    if (rc != PFM_OK) return rc;
    h->last_launches++;
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// ---------------------------------------------------------------------------------------------
// training: forward with saved activations (+ fused flow-matching loss), backward, weight gradients
// ---------------------------------------------------------------------------------------------
__global__ void loss_finalize_kernel(const float* __restrict__ acc, const int* __restrict__ n_total,
                                     const int* __restrict__ n_real, int B, float* __restrict__ loss) {
  // A jet without particles makes the reference's loss NaN (its pooled mean is 0/0 and NaN * mask stays NaN in the sum):
  // the packed kernels never visit such a jet, so the NaN is put back here.
  __shared__ int empty;
  if (threadIdx.x == 0) empty = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < B; j += blockDim.x)
    if (n_real[j] == 0) empty = 1;
  __syncthreads();
  if (threadIdx.x == 0)
    *loss = empty ? __int_as_float(0x7fc00000) : *acc / (float)(*n_total);      // sum((v-u)^2) / sum(mask)   (losses.py:76,130,341)
}

// dpre3[row][f] (holds leaky_relu'(pre3)) *= grad_out[jet][particle][f]; one warp per jet
__global__ void seed_kernel(float* __restrict__ dpre3, const float* __restrict__ grad_out, const int* __restrict__ n_real,
                            const uint16_t* __restrict__ ridx, const int* __restrict__ rowoff, int B, int N, int F) {
  const int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (jet >= B) return;
  const int n = n_real[jet], r0 = rowoff[jet];
  for (int i = lane; i < n * F; i += 32) {
    const int r = i / F, f = i - r * F;
    const int part = ridx[(size_t)jet * N + r];
    dpre3[(size_t)(r0 + r) * F + f] *= grad_out[((size_t)jet * N + part) * F + f];
  }
}

// grad_x[jet][particle][:] = dxs[row][:] for real particles, 0 for padding
__global__ void scatter_dx_kernel(const float* __restrict__ dxs, float* __restrict__ grad_x, const int* __restrict__ n_real,
                                  const uint16_t* __restrict__ ridx, const int* __restrict__ rowoff, int B, int N, int Kx) {
  const int jet = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (jet >= B) return;
  const int n = n_real[jet], r0 = rowoff[jet];
  float* dst = grad_x + (size_t)jet * N * Kx;
  for (int i = lane; i < N * Kx; i += 32) dst[i] = 0.f;
  __syncwarp();
  for (int i = lane; i < n * Kx; i += 32) {
    const int r = i / Kx, c = i - r * Kx;
    dst[(size_t)ridx[(size_t)jet * N + r] * Kx + c] = dxs[(size_t)(r0 + r) * Kx + c];
  }
}

static size_t grad_floats(const pfm_epic* h) {
  size_t n = 0;
  for (const Lin& L : h->lin_host) n += (size_t)L.out * L.in + L.out;
  return n;
}

static constexpr size_t kPinSlot = 64 * 1024;       // bytes per staging slot

int upload_table(pfm_epic* h, int slot, void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes + sizeof(void*) > kPinSlot) {               // too large for the staging slot: plain (pageable) copy
    PFM_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return PFM_OK;
  }
  if (!h->pin_stage) {
    PFM_CUDA_CHECK(cudaHostAlloc(&h->pin_stage, 3 * kPinSlot, cudaHostAllocDefault));
    h->pin_cap = 3 * kPinSlot;
  }
  uint8_t* stage = reinterpret_cast<uint8_t*>(h->pin_stage) + (size_t)slot * kPinSlot;
  void** tag = reinterpret_cast<void**>(stage + kPinSlot - sizeof(void*));
  if (h->pin_used[slot] == bytes && *tag == dst && memcmp(stage, src, bytes) == 0) return PFM_OK;    // device copy is current
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(st, &cs);
  // the slot may still be the source of the previous copy: wait for that copy only (an event, not the stream)
  if (h->pin_ev_set[slot] && cs == cudaStreamCaptureStatusNone) PFM_CUDA_CHECK(cudaEventSynchronize(h->pin_ev[slot]));
  memcpy(stage, src, bytes);
  *tag = dst;
  h->pin_used[slot] = bytes;
  PFM_CUDA_CHECK(cudaMemcpyAsync(dst, stage, bytes, cudaMemcpyHostToDevice, st));
  if (cs == cudaStreamCaptureStatusNone) {
    if (!h->pin_ev[slot]) PFM_CUDA_CHECK(cudaEventCreateWithFlags(&h->pin_ev[slot], cudaEventDisableTiming));
    PFM_CUDA_CHECK(cudaEventRecord(h->pin_ev[slot], st));
    h->pin_ev_set[slot] = true;
  }
  return PFM_OK;
}

// CTA groups of the fused CUDA-core training kernels (needed by their backward when the forward ran on the tensor-core path)
int train_plan_groups(pfm_epic* h, int B, const TrainLayout& lay, cudaStream_t st) {
  plan_group_kernel<<<1, 1024, sizeof(int) * B, st>>>(h->plan.n_real, B, lay.R_cap, lay.J_cap, h->plan.groups, h->plan.n_groups,
                                                     h->plan.counter, h->plan.rowoff, h->plan.n_total);
  PFM_CUDA_CHECK(cudaGetLastError());
  h->last_launches++;
  return PFM_OK;
}

// forward half shared by pfm_epic_loss_fwd_bwd and pfm_epic_forward_train
static int train_forward_common(pfm_epic* h, const float* t_code, int t_rows, const float* t_code_in, int t_in,
                                const float* x, float* out, const float* t_jet, const float* noise0, const float* noise1,
                                int loss_kind, float sigma, const float* mask, const float* cond, int B, int N, int Kx,
                                int xin_off, TrainLayout* lay, cudaStream_t st) {
  const pfm_epic_cfg& c = h->cfg;
  if (!h->weights_set) { set_error("weights not set (call pfm_epic_set_weights first)"); return PFM_ERR_STATE; }
  if (B <= 0 || N <= 0 || N > 65535) { set_error("bad batch shape B=%d N=%d", B, N); return PFM_ERR_INVALID; }
  if (B > kMaxJetsPerCall) { set_error("training batch of %d jets exceeds %d per call", B, kMaxJetsPerCall); return PFM_ERR_INVALID; }
  const int cond_dim = c.global_cond_dim > c.local_cond_dim ? c.global_cond_dim : c.local_cond_dim;
  if (cond_dim > 0 && cond == nullptr) { set_error("cond is NULL but the net is conditioned"); return PFM_ERR_INVALID; }
  const bool any_t = (c.t_local_cat || c.t_global_cat) && c.t_dim > 0;
  if ((any_t && !t_code) || (t_in > 0 && !t_code_in)) { set_error("time code is NULL but the net takes a time code"); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  h->last_launches = 0; h->ev_used = 0; h->train_B = 0;
  int rc = train_layout(h, B, N, lay);
  if (rc != PFM_OK) return rc;
  const bool per_jet = t_rows == B && B > 1;
  const int trows = per_jet ? B : 1;
  rc = ensure_floats(&h->tbias, &h->tbias_cap, (size_t)trows * h->bstride);
  if (rc != PFM_OK) return rc;
  if (trows >= 2 * TB_ROWS && c.t_dim <= TB_KMAX)
    tbias_rows_kernel<<<dim3((trows + TB_ROWS - 1) / TB_ROWS, h->n_lin), 128, 0, st>>>(h->lin_dev, t_code, c.t_dim, t_code_in, t_in, h->tbias,
                                                                                       h->bstride, trows);
  else
    tbias_kernel<<<dim3(trows, h->n_lin), 128, 0, st>>>(h->lin_dev, t_code, c.t_dim, t_code_in, t_in, h->tbias, h->bstride);
  h->last_launches++;
  if (cond_dim > 0) {
    rc = ensure_floats(&h->cbias, &h->cbias_cap, (size_t)B * h->bstride);
    if (rc != PFM_OK) return rc;
    cbias_kernel<<<dim3(B, h->n_lin), 128, 0, st>>>(h->lin_dev, cond, cond_dim, h->cbias, h->bstride);
    h->last_launches++;
  }
  rc = ensure_plan(h, B, N);
  if (rc != PFM_OK) return rc;
  plan_count_kernel<<<(B + 7) / 8, 256, 0, st>>>(mask, B, N, h->plan.n_real, h->plan.ridx);
  h->train_tc = tt_enabled(h);
  if (h->train_tc) {           // rows only: prefix sum of the multiplicities + row -> jet map (no CTA groups)
    if ((rc = tt_plan(h, B, N, st)) != PFM_OK) return rc;
    h->last_launches += 1;
  } else {
    plan_group_kernel<<<1, 1024, sizeof(int) * B, st>>>(h->plan.n_real, B, lay->R_cap, lay->J_cap, h->plan.groups, h->plan.n_groups,
                                                       h->plan.counter, h->plan.rowoff, h->plan.n_total);
    h->last_launches += 2;
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  const size_t rows = (size_t)B * N;
  const size_t stages = 2 + 2 * (size_t)c.layers;
  if ((rc = ensure_floats(&h->act, &h->act_cap, stages * lay->stage_stride)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->dact, &h->dact_cap, stages * lay->stage_stride)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->yact, &h->yact_cap, rows * Kx)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->jact, &h->jact_cap, (size_t)B * lay->jstride)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->dpre3, &h->dpre3_cap, rows * c.feats)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->dbeff, &h->dbeff_cap, (size_t)B * h->bstride)) != PFM_OK) return rc;
  if ((rc = ensure_floats(&h->dxs, &h->dxs_cap, rows * Kx)) != PFM_OK) return rc;
  if (!h->loss_acc) {
    PFM_CUDA_CHECK(cudaMalloc(&h->loss_acc, sizeof(float)));
    PFM_CUDA_CHECK(cudaMalloc(&h->ones, sizeof(float)));
    const float one = 1.f;
    PFM_CUDA_CHECK(cudaMemcpy(h->ones, &one, sizeof(float), cudaMemcpyHostToDevice));
  }
  PFM_CUDA_CHECK(cudaMemsetAsync(h->loss_acc, 0, sizeof(float), st));
  TrainFwdArgs a;
  a.x_in = x; a.x_out = out; a.t = t_jet; a.noise0 = noise0; a.noise1 = noise1; a.loss_kind = loss_kind; a.sigma = sigma;
  a.B = B; a.N = N; a.Kx = Kx; a.xin_off = xin_off; a.has_cbias = cond_dim > 0; a.tbias_per_jet = per_jet ? 1 : 0;
  a.lay = *lay;
  if (h->train_tc) {
    rc = tt_train_forward(h, a, st);
    if (rc != PFM_OK) return rc;
  } else {
    rc = simt_train_forward(h, a, st);
    if (rc != PFM_OK) return rc;
    h->last_launches++;
  }
  h->train_B = B; h->train_N = N; h->train_Kx = Kx; h->train_xin_off = xin_off;
  return PFM_OK;
}

}  // namespace pfm

using namespace pfm;

// =============================================================================================
// exported C ABI
// =============================================================================================
extern "C" {

int pfm_version(void) { return PFM_VERSION; }

const char* pfm_last_error(void) { return g_err; }

static void linear_layout(const pfm_epic_cfg& c, int i, int n_lin, Lin* L) {
  const int H = c.hid, Z = c.latent;
  const int tl = c.t_local_cat ? c.t_dim : 0, tg = c.t_global_cat ? c.t_dim : 0;
  const int cl = c.local_cond_dim, cg = c.global_cond_dim;
  memset(L, 0, sizeof(Lin));
  bool local;
  int main_len, g_len = 0, out;
  if (i == LIN_L1) { local = true; main_len = c.input_dim; out = H; }
  else if (i == LIN_L2) { local = true; main_len = H; out = H; }
  else if (i == LIN_G1) { local = false; main_len = 2 * H; out = H; }
  else if (i == LIN_G2) { local = false; main_len = H; out = Z; }
  else if (i == n_lin - 1) { local = true; main_len = H; out = c.feats; }
  else {
    int r = (i - LIN_LAYER0) & 3;
    if (r == 0) { local = false; main_len = 2 * H + Z; out = H; }
    else if (r == 1) { local = false; main_len = H; out = Z; }
    else if (r == 2) { local = true; main_len = H; g_len = Z; out = H; }
    else { local = true; main_len = H; out = H; }
  }
  L->out = out;
  L->t_off = 0; L->t_len = local ? tl : tg;
  L->m_off = L->t_len; L->m_len = main_len;
  L->g_off = L->m_off + main_len; L->g_len = g_len;
  L->c_off = L->g_off + g_len; L->c_len = local ? cl : cg;
  L->in = L->c_off + L->c_len;
  L->ldo = round_up(out, 4);
}

int pfm_epic_create(const pfm_epic_cfg* cfg, int device, pfm_epic** out) {
  if (!cfg || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  *out = nullptr;
  const pfm_epic_cfg& c = *cfg;
  if (c.feats <= 0 || c.input_dim <= 0 || c.hid <= 0 || c.latent <= 0 || c.layers < 0 || c.t_dim < 0) {
    set_error("invalid dims: feats=%d input_dim=%d hid=%d latent=%d layers=%d t_dim=%d", c.feats, c.input_dim,
              c.hid, c.latent, c.layers, c.t_dim);
    return PFM_ERR_INVALID;
  }
  if (c.local_cond_dim != 0 && c.local_cond_dim != c.global_cond_dim && c.global_cond_dim != 0) {
    set_error("local_cond_dim (%d) must be 0 or equal to global_cond_dim (%d): the reference feeds ONE cond tensor "
              "to both (epic.py:347-357)", c.local_cond_dim, c.global_cond_dim);
    return PFM_ERR_INVALID;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error("no CUDA device available (%s); libpfm_b200 has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    return PFM_ERR_CUDA;
  }
  if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaSetDevice(device));
  pfm_epic* h = new pfm_epic();
  h->cfg = c;
  h->device = device;
  h->n_lin = 4 + 4 * c.layers + 1;
  h->precision = PFM_PREC_FP32;
  h->weights_set = false;
  h->lin_dev = nullptr; h->wt_store = nullptr; h->b_store = nullptr; h->tc_store = nullptr; h->tc_bytes = 0; h->tc_dirty = false;
  h->wn_rows = nullptr; h->wn_total_rows = 0; h->wn_goff = nullptr; h->wn_ptrs = nullptr;
  h->wr_store = nullptr; h->wr_floats = 0;
  h->act = nullptr; h->act_cap = 0; h->dact = nullptr; h->dact_cap = 0; h->yact = nullptr; h->yact_cap = 0;
  h->jact = nullptr; h->jact_cap = 0; h->dpre3 = nullptr; h->dpre3_cap = 0;
  h->dbeff = nullptr; h->dbeff_cap = 0; h->dxs = nullptr; h->dxs_cap = 0; h->loss_acc = nullptr; h->ones = nullptr;
  h->grad_chunks = 0;
  h->hs_spill = nullptr; h->hs_spill_cap = 0; h->dh_spill = nullptr; h->dh_spill_cap = 0;
  h->jobs_dev = nullptr; h->jobs_cap = 0; h->train_B = 0; h->train_N = 0; h->train_Kx = 0; h->train_xin_off = 0;
  h->tbias = nullptr; h->tbias_cap = 0; h->cbias = nullptr; h->cbias_cap = 0;
  memset(&h->plan, 0, sizeof(h->plan));
  h->last_launches = 0; h->last_groups_host = 0;
  h->timing = false; h->ev_used = 0;
  h->step_kind = 0; h->step_coef = nullptr; h->step_noise = nullptr;
  h->pin_stage = nullptr; h->pin_cap = 0; h->pin_used[0] = h->pin_used[1] = h->pin_used[2] = 0;
  for (int i = 0; i < 3; ++i) { h->pin_ev[i] = nullptr; h->pin_ev_set[i] = false; }
  h->tt_store = nullptr; h->tt_bytes = 0; h->tt_dirty = true; h->tt_ws = nullptr; h->tt_ws_cap = 0; h->train_tc = false; h->train_mode = PFM_TRAIN_AUTO;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  h->lin_host.resize(h->n_lin);
  size_t wt = 0, bf = 0, wr = 0;
  int boff = 0;
  for (int i = 0; i < h->n_lin; ++i) {
    Lin& L = h->lin_host[i];
    linear_layout(c, i, h->n_lin, &L);
    L.bias_off = boff;
    boff += L.ldo;
    wt += (size_t)(L.in + 8) * L.ldo;   // +8 slack rows: kernels may read (never use) up to 3 rows past a block
    bf += L.ldo;
    L.ldr = round_up(L.m_len, 4);
    wr += (size_t)(L.out + 8) * L.ldr;
  }
  h->wr_floats = wr + 1024;
  h->bstride = boff;
  h->wt_floats = wt + 1024;
  h->b_floats = bf;
  cudaError_t e1 = cudaMalloc(&h->wt_store, sizeof(float) * h->wt_floats);
  cudaError_t e2 = cudaMalloc(&h->b_store, sizeof(float) * h->b_floats);
  cudaError_t e3 = cudaMalloc(&h->lin_dev, sizeof(Lin) * h->n_lin);
  cudaError_t e4 = cudaMalloc(&h->wr_store, sizeof(float) * h->wr_floats);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
    set_error("cudaMalloc failed for packed weights");
    pfm_epic_destroy(h);
    return PFM_ERR_CUDA;
  }
  cudaMemset(h->wt_store, 0, sizeof(float) * h->wt_floats);
  cudaMemset(h->b_store, 0, sizeof(float) * h->b_floats);
  cudaMemset(h->wr_store, 0, sizeof(float) * h->wr_floats);
  size_t wo = 0, bo = 0, ro = 0;
  for (int i = 0; i < h->n_lin; ++i) {
    Lin& L = h->lin_host[i];
    L.Wt = h->wt_store + wo;
    L.b = h->b_store + bo;
    L.Wr = h->wr_store + ro;
    wo += (size_t)(L.in + 8) * L.ldo;
    bo += L.ldo;
    ro += (size_t)(L.out + 8) * L.ldr;
  }
  cudaMemcpy(h->lin_dev, h->lin_host.data(), sizeof(Lin) * h->n_lin, cudaMemcpyHostToDevice);
  *out = h;
  return PFM_OK;
}

void pfm_epic_destroy(pfm_epic* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->lin_dev) cudaFree(h->lin_dev);
  if (h->wn_rows) cudaFree(h->wn_rows);
  if (h->wn_goff) cudaFree(h->wn_goff);
  if (h->wn_ptrs) cudaFree(h->wn_ptrs);
  if (h->wt_store) cudaFree(h->wt_store);
  if (h->b_store) cudaFree(h->b_store);
  if (h->tc_store) cudaFree(h->tc_store);
  if (h->tt_store) cudaFree(h->tt_store);
  if (h->tt_ws) cudaFree(h->tt_ws);
  if (h->pin_stage) cudaFreeHost(h->pin_stage);
  for (int i = 0; i < 3; ++i) if (h->pin_ev[i]) cudaEventDestroy(h->pin_ev[i]);
  if (h->wr_store) cudaFree(h->wr_store);
  if (h->act) cudaFree(h->act);
  if (h->dact) cudaFree(h->dact);
  if (h->yact) cudaFree(h->yact);
  if (h->dxs) cudaFree(h->dxs);
  if (h->ones) cudaFree(h->ones);
  if (h->jobs_dev) cudaFree(h->jobs_dev);
  if (h->hs_spill) cudaFree(h->hs_spill);
  if (h->dh_spill) cudaFree(h->dh_spill);
  if (h->jact) cudaFree(h->jact);
  if (h->dpre3) cudaFree(h->dpre3);
  if (h->dbeff) cudaFree(h->dbeff);
  if (h->loss_acc) cudaFree(h->loss_acc);
  if (h->tbias) cudaFree(h->tbias);
  if (h->cbias) cudaFree(h->cbias);
  if (h->plan.n_real) cudaFree(h->plan.n_real);
  if (h->plan.ridx) cudaFree(h->plan.ridx);
  if (h->plan.groups) cudaFree(h->plan.groups);
  if (h->plan.n_groups) cudaFree(h->plan.n_groups);
  if (h->plan.counter) cudaFree(h->plan.counter);
  if (h->plan.rowoff) cudaFree(h->plan.rowoff);
  if (h->plan.jetmap) cudaFree(h->plan.jetmap);
  if (h->plan.order) cudaFree(h->plan.order);
  if (h->plan.groups_tmp) cudaFree(h->plan.groups_tmp);
  if (h->plan.n_total) cudaFree(h->plan.n_total);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->grad_ev) cudaEventDestroy(e);
  delete h;
}

int pfm_epic_num_linears(const pfm_epic* h) { return h ? h->n_lin : PFM_ERR_INVALID; }

int pfm_epic_linear_shape(const pfm_epic* h, int i, int32_t* out_features, int32_t* in_features) {
  if (!h || i < 0 || i >= h->n_lin) { set_error("linear index out of range"); return PFM_ERR_INVALID; }
  if (out_features) *out_features = h->lin_host[i].out;
  if (in_features) *in_features = h->lin_host[i].in;
  return PFM_OK;
}

int pfm_epic_set_weights(pfm_epic* h, const float* const* weights, const float* const* biases, int n,
                         void* stream) {
  if (!h || !weights || !biases) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (n != h->n_lin) { set_error("expected %d linears, got %d", h->n_lin, n); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  for (int i = 0; i < n; ++i) {
    const Lin& L = h->lin_host[i];
    if (!weights[i] || !biases[i]) { set_error("null weight/bias pointer for linear %d", i); return PFM_ERR_INVALID; }
    int total = L.in * L.ldo;
    transpose_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(weights[i], const_cast<float*>(L.Wt), L.out, L.in,
                                                                 L.ldo);
    rowmajor_main_kernel<<<(L.out * L.ldr + 255) / 256, 256, 0, st>>>(weights[i], const_cast<float*>(L.Wr), L.out, L.in,
                                                                      L.m_off, L.m_len, L.ldr);
    PFM_CUDA_CHECK(cudaMemcpyAsync(const_cast<float*>(L.b), biases[i], sizeof(float) * L.out,
                                   cudaMemcpyDeviceToDevice, st));
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  h->weights_set = true;
  h->tc_dirty = true; h->tt_dirty = true;
  if (h->precision == PFM_PREC_BF16) {
    int rc = tc_pack_weights(h, st);
    if (rc != PFM_OK) return rc;
  }
  return PFM_OK;
}

// ---- weight-norm fold and its backward, one launch each for all linears ------------------------------------------
// (the reference re-creates `weight = g * v / ||v||` in a forward pre-hook of every linear, epic.py:66-81 with torch's old
// nn.utils.weight_norm, and autograd differentiates it; here: block = one output row of one linear)
static int wn_tables(pfm_epic* h) {
  if (h->wn_rows) return PFM_OK;
  std::vector<int2> rows;
  std::vector<long long> goff(2 * (size_t)h->n_lin);
  long long off = 0;
  for (int i = 0; i < h->n_lin; ++i) {
    const Lin& L = h->lin_host[i];
    for (int o = 0; o < L.out; ++o) rows.push_back(make_int2(i, o));
    goff[2 * i] = off; off += (long long)L.out * L.in;
    goff[2 * i + 1] = off; off += L.out;
  }
  h->wn_total_rows = (int)rows.size();
  PFM_CUDA_CHECK(cudaMalloc(&h->wn_rows, sizeof(int2) * rows.size()));
  PFM_CUDA_CHECK(cudaMalloc(&h->wn_goff, sizeof(long long) * goff.size()));
  PFM_CUDA_CHECK(cudaMalloc(&h->wn_ptrs, sizeof(float*) * 8 * (size_t)h->n_lin));
  PFM_CUDA_CHECK(cudaMemcpy(h->wn_rows, rows.data(), sizeof(int2) * rows.size(), cudaMemcpyHostToDevice));
  PFM_CUDA_CHECK(cudaMemcpy(h->wn_goff, goff.data(), sizeof(long long) * goff.size(), cudaMemcpyHostToDevice));
  return PFM_OK;
}

__device__ __forceinline__ float block_sum_128(float v, float* red) {
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  return red[0] + red[1] + red[2] + red[3];
}

// ptrs: [v_0..v_{n-1} | g_0.. | b_0..]   (g_i == NULL: plain linear, v_i is the weight itself)
__global__ void __launch_bounds__(128) wn_fold_kernel(const Lin* __restrict__ lins, const int2* __restrict__ rows,
                                                      const float* const* __restrict__ ptrs, int n) {
  __shared__ float red[4];
  const int2 lr = rows[blockIdx.x];
  const Lin L = lins[lr.x];
  const int o = lr.y;
  const float* v = ptrs[lr.x] + (size_t)o * L.in;
  const float* g = ptrs[n + lr.x];
  float scale = 1.f;
  if (g) {
    float ss = 0.f;
    for (int k = threadIdx.x; k < L.in; k += 128) { const float x = v[k]; ss = fmaf(x, x, ss); }
    ss = block_sum_128(ss, red);
    scale = g[o] / sqrtf(ss);
  }
  float* Wt = const_cast<float*>(L.Wt);
  float* Wr = const_cast<float*>(L.Wr);
  for (int k = threadIdx.x; k < L.in; k += 128) {
    const float w = v[k] * scale;
    Wt[(size_t)k * L.ldo + o] = w;
    const int km = k - L.m_off;
    if (km >= 0 && km < L.m_len) Wr[(size_t)o * L.ldr + km] = w;
  }
  if (threadIdx.x == 0) const_cast<float*>(L.b)[o] = ptrs[2 * n + lr.x][o];
}

// ptrs: [v | g | dv | dg | db];  dW = grad_flat[goff[2i] + o*in + k],  db = grad_flat[goff[2i+1] + o];  all scaled by *scale
__global__ void __launch_bounds__(128) wn_bwd_kernel(const Lin* __restrict__ lins, const int2* __restrict__ rows,
                                                     const long long* __restrict__ goff, const float* __restrict__ grad_flat,
                                                     const float* __restrict__ scale_ptr, const float* const* __restrict__ ptrs, int n) {
  __shared__ float red[4];
  const int2 lr = rows[blockIdx.x];
  const int in = lins[lr.x].in, o = lr.y;
  const float s = scale_ptr ? *scale_ptr : 1.f;
  const float* v = ptrs[lr.x] + (size_t)o * in;
  const float* g = ptrs[n + lr.x];
  float* dv = const_cast<float*>(ptrs[2 * n + lr.x]) + (size_t)o * in;
  float* dg = const_cast<float*>(ptrs[3 * n + lr.x]);
  float* db = const_cast<float*>(ptrs[4 * n + lr.x]);
  const float* dw = grad_flat + goff[2 * lr.x] + (size_t)o * in;
  if (g) {
    float ss = 0.f, dot = 0.f;
    for (int k = threadIdx.x; k < in; k += 128) { const float x = v[k]; ss = fmaf(x, x, ss); dot = fmaf(dw[k], x, dot); }
    ss = block_sum_128(ss, red);
    dot = block_sum_128(dot, red);
    const float nrm = sqrtf(ss), gv = g[o];
    const float a = s * gv / nrm, b = dot / ss;
    for (int k = threadIdx.x; k < in; k += 128) dv[k] = a * (dw[k] - v[k] * b);
    if (threadIdx.x == 0) dg[o] = s * dot / nrm;
  } else {
    for (int k = threadIdx.x; k < in; k += 128) dv[k] = s * dw[k];
  }
  if (threadIdx.x == 0) db[o] = s * grad_flat[goff[2 * lr.x + 1] + o];
}

int pfm_epic_set_params(pfm_epic* h, const float* const* v, const float* const* g, const float* const* b, int n, void* stream) {
  if (!h || !v || !g || !b) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (n != h->n_lin) { set_error("expected %d linears, got %d", h->n_lin, n); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = wn_tables(h);
  if (rc != PFM_OK) return rc;
  std::vector<const float*> tab(3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    if (!v[i] || !b[i]) { set_error("null weight/bias pointer for linear %d", i); return PFM_ERR_INVALID; }
    tab[i] = v[i]; tab[n + i] = g[i]; tab[2 * n + i] = b[i];
  }
  // pageable source: staged by the runtime before the call returns
  rc = upload_table(h, 0, h->wn_ptrs, tab.data(), sizeof(float*) * tab.size(), st);
  if (rc != PFM_OK) return rc;
  wn_fold_kernel<<<h->wn_total_rows, 128, 0, st>>>(h->lin_dev, h->wn_rows, h->wn_ptrs, n);
  PFM_CUDA_CHECK(cudaGetLastError());
  h->weights_set = true;
  h->tc_dirty = true; h->tt_dirty = true;      // the bf16 / hi-lo images are repacked lazily by the next call that needs them
  return PFM_OK;
}

int pfm_epic_param_grads(pfm_epic* h, const float* grad_flat, const float* scale, const float* const* v, const float* const* g,
                         float* const* dv, float* const* dg, float* const* db, int n, void* stream) {
  if (!h || !grad_flat || !v || !g || !dv || !dg || !db) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (n != h->n_lin) { set_error("expected %d linears, got %d", h->n_lin, n); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  int rc = wn_tables(h);
  if (rc != PFM_OK) return rc;
  std::vector<const float*> tab(5 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    if (!v[i] || !dv[i] || !db[i] || (g[i] && !dg[i])) { set_error("null pointer for linear %d", i); return PFM_ERR_INVALID; }
    tab[i] = v[i]; tab[n + i] = g[i]; tab[2 * n + i] = dv[i]; tab[3 * n + i] = dg[i]; tab[4 * n + i] = db[i];
  }
  rc = upload_table(h, 1, h->wn_ptrs + 3 * (size_t)n, tab.data(), sizeof(float*) * tab.size(), st);
  if (rc != PFM_OK) return rc;
  wn_bwd_kernel<<<h->wn_total_rows, 128, 0, st>>>(h->lin_dev, h->wn_rows, h->wn_goff, grad_flat, scale, h->wn_ptrs + 3 * (size_t)n, n);
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

int pfm_epic_set_precision(pfm_epic* h, int precision) {
  if (!h) { set_error("null handle"); return PFM_ERR_INVALID; }
  if (precision != PFM_PREC_FP32 && precision != PFM_PREC_BF16) { set_error("unknown precision %d", precision); return PFM_ERR_INVALID; }
  if (precision == PFM_PREC_BF16) {
    int rc = tc_supported(h, 1);
    if (rc != PFM_OK) return rc;
  }
  int old = h->precision;
  h->precision = precision;
  // The bf16 weight images are (re)packed lazily by the next forward / sample ON ITS OWN STREAM (tc_run packs when
  // tc_store is NULL): packing here on the legacy stream would not be ordered against a non-blocking consumer stream.
  (void)old;
  return PFM_OK;
}

// test hook: copy n floats of an internal training array to the host (which: 0 act, 1 dact, 2 dbeff, 3 jact, 4 dpre3, 5 yact)
int pfm_epic_debug_copy(pfm_epic* h, int which, float* host, long long n) {
  if (!h || !host || which < 0 || which > 5) { set_error("bad argument"); return PFM_ERR_INVALID; }
  const float* src = which == 0 ? h->act : which == 1 ? h->dact : which == 2 ? h->dbeff : which == 3 ? h->jact : which == 4 ? h->dpre3 : h->yact;
  if (!src) { set_error("no such array"); return PFM_ERR_STATE; }
  PFM_CUDA_CHECK(cudaDeviceSynchronize());
  PFM_CUDA_CHECK(cudaMemcpy(host, src, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
  return PFM_OK;
}

int pfm_epic_set_train_mode(pfm_epic* h, int mode) {
  if (!h || (mode != PFM_TRAIN_AUTO && mode != PFM_TRAIN_CUDA_CORES)) { set_error("bad train mode %d", mode); return PFM_ERR_INVALID; }
  h->train_mode = mode;
  return PFM_OK;
}

int pfm_epic_forward(pfm_epic* h, const float* t_code, int t_rows, const float* x, const float* mask,
                     const float* cond, float* out, int B, int N, void* stream) {
  if (!h || !x || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (t_rows != 1 && t_rows != B) { set_error("t_rows must be 1 or B (got %d, B=%d)", t_rows, B); return PFM_ERR_INVALID; }
  const bool per_jet = (t_rows == B) && B > 1;
  return run_chunked(h, t_code, per_jet ? B : 1, per_jet, nullptr, 0, x, out, mask, cond, B, N, h->cfg.input_dim, 0,
                     1, -1, 0, nullptr, (cudaStream_t)stream);
}

int pfm_epic_sample(pfm_epic* h, float* x_inout, const float* mask, const float* cond, const float* t_codes,
                    const float* t_codes_in, const float* dt, int solver, int n_steps, int B, int N, void* stream) {
  if (!h || !x_inout || !dt) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (solver != PFM_SOLVER_EULER && solver != PFM_SOLVER_MIDPOINT) { set_error("unknown solver %d", solver); return PFM_ERR_INVALID; }
  if (n_steps <= 0) { set_error("n_steps must be positive"); return PFM_ERR_INVALID; }
  const int t_in = h->cfg.input_dim - h->cfg.feats;
  if (t_in < 0) { set_error("input_dim < feats"); return PFM_ERR_INVALID; }
  if (t_in > 0 && !t_codes_in) { set_error("input_dim > feats needs t_codes_in (add_time_to_input)"); return PFM_ERR_INVALID; }
  const int n_evals = n_steps * (solver == PFM_SOLVER_MIDPOINT ? 2 : 1);
  return run_chunked(h, t_codes, n_evals, false, t_codes_in, t_in, x_inout, x_inout, mask, cond, B, N, h->cfg.feats,
                     t_in, n_evals, solver, n_steps, dt, (cudaStream_t)stream);
}

int pfm_epic_sample_diffusion(pfm_epic* h, float* x_inout, const float* mask, const float* cond, const float* t_codes,
                              const float* t_codes_in, const float* coef, const float* noise, const float* dt, int step_kind,
                              int solver, int n_steps, int B, int N, void* stream) {
  if (!h || !x_inout || !coef) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (step_kind != PFM_STEP_PF_ODE && step_kind != PFM_STEP_DDIM && step_kind != PFM_STEP_EM) { set_error("unknown step kind %d", step_kind); return PFM_ERR_INVALID; }
  if (step_kind == PFM_STEP_PF_ODE) {
    if (solver != PFM_SOLVER_EULER && solver != PFM_SOLVER_MIDPOINT) { set_error("unknown solver %d", solver); return PFM_ERR_INVALID; }
    if (!dt) { set_error("the probability-flow ODE needs the step sizes dt"); return PFM_ERR_INVALID; }
  } else {
    solver = PFM_SOLVER_EULER;       // one network evaluation per step
  }
  if (step_kind == PFM_STEP_EM && !noise) { set_error("the Euler-Maruyama sampler needs the per-step noise"); return PFM_ERR_INVALID; }
  if (n_steps <= 0) { set_error("n_steps must be positive"); return PFM_ERR_INVALID; }
  if (h->precision != PFM_PREC_FP32) { set_error("the diffusion samplers run on the fp32 path (PFM_PREC_FP32)"); return PFM_ERR_UNSUPPORTED; }
  const int t_in = h->cfg.input_dim - h->cfg.feats;
  if (t_in < 0) { set_error("input_dim < feats"); return PFM_ERR_INVALID; }
  if (t_in > 0 && !t_codes_in) { set_error("input_dim > feats needs t_codes_in (add_time_to_input)"); return PFM_ERR_INVALID; }
  const int n_evals = n_steps * (solver == PFM_SOLVER_MIDPOINT ? 2 : 1);
  h->step_kind = step_kind; h->step_coef = coef; h->step_noise = noise;
  const int rc = run_chunked(h, t_codes, n_evals, false, t_codes_in, t_in, x_inout, x_inout, mask, cond, B, N, h->cfg.feats,
                             t_in, n_evals, solver, n_steps, dt, (cudaStream_t)stream);
  h->step_kind = 0; h->step_coef = nullptr; h->step_noise = nullptr;
  return rc;
}

int pfm_epic_last_launches(const pfm_epic* h) { return h ? h->last_launches : PFM_ERR_INVALID; }

int pfm_epic_last_groups(const pfm_epic* h) {
  if (!h || !h->plan.n_groups) return PFM_ERR_INVALID;
  int g = 0;
  cudaSetDevice(h->device);
  if (cudaMemcpy(&g, h->plan.n_groups, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return PFM_ERR_CUDA;
  return g;
}

int pfm_epic_set_timing(pfm_epic* h, int enable) {
  if (!h) { set_error("null handle"); return PFM_ERR_INVALID; }
  h->timing = enable != 0;
  h->ev_used = 0;
  return PFM_OK;
}

float pfm_epic_last_kernel_ms(pfm_epic* h) {
  if (!h || !h->timing || h->ev_used < 2) return -1.f;
  float total = 0.f;
  for (int i = 0; i + 1 < h->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventSynchronize(h->ev_pool[i + 1]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]) != cudaSuccess) return -1.f;
    total += ms;
  }
  return total;
}

int pfm_epic_grad_chunks(const pfm_epic* h) { return h ? h->grad_chunks : PFM_ERR_INVALID; }

int pfm_epic_grad_chunk_range(const pfm_epic* h, int i, long long* offset, long long* count) {
  if (!h || i < 0 || i >= h->grad_chunks) { set_error("gradient chunk index out of range"); return PFM_ERR_INVALID; }
  if (offset) *offset = h->grad_chunk_off[i];
  if (count) *count = h->grad_chunk_off[i + 1] - h->grad_chunk_off[i];
  return PFM_OK;
}

int pfm_epic_stream_wait_grad_chunk(pfm_epic* h, int i, void* stream) {
  if (!h || i < 0 || i >= h->grad_chunks) { set_error("gradient chunk index out of range"); return PFM_ERR_INVALID; }
  PFM_CUDA_CHECK(cudaStreamWaitEvent((cudaStream_t)stream, h->grad_ev[i], 0));
  return PFM_OK;
}

long long pfm_epic_grad_size(const pfm_epic* h) { return h ? (long long)grad_floats(h) : (long long)PFM_ERR_INVALID; }

int pfm_epic_loss_fwd_bwd(pfm_epic* h, const float* x1, const float* t, const float* t_code, const float* t_code_in,
                          const float* noise0, const float* noise1, const float* mask, const float* cond, int loss_kind,
                          float sigma, float* loss_out, float* grad_flat, int B, int N, void* stream) {
  if (!h || !x1 || !t || !noise0 || !loss_out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (loss_kind != PFM_LOSS_FM_OT && loss_kind != PFM_LOSS_CFM && loss_kind != PFM_LOSS_DROID) {
    set_error("unknown loss kind %d", loss_kind); return PFM_ERR_INVALID;
  }
  if (loss_kind == PFM_LOSS_CFM && !noise1) { set_error("the CFM loss needs noise1"); return PFM_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  const pfm_epic_cfg& c = h->cfg;
  const int t_in = c.input_dim - c.feats;
  if (t_in < 0) { set_error("input_dim < feats"); return PFM_ERR_INVALID; }
  TrainLayout lay;
  int rc = train_forward_common(h, t_code, B, t_code_in, t_in, x1, nullptr, t, noise0, noise1, loss_kind, sigma, mask, cond, B, N,
                                c.feats, t_in, &lay, st);
  if (rc != PFM_OK) return rc;
  loss_finalize_kernel<<<1, 256, 0, st>>>(h->loss_acc, h->plan.n_total, h->plan.n_real, B, loss_out);
  h->last_launches++;
  if (!grad_flat) return PFM_OK;
  PFM_CUDA_CHECK(cudaMemsetAsync(grad_flat, 0, sizeof(float) * grad_floats(h), st));
  TrainBwdArgs b;
  b.B = B; b.N = N; b.Kx = c.feats; b.xin_off = t_in;
  b.t_code = t_code; b.t_ld = (B > 1) ? c.t_dim : 0; b.t_code_in = t_code_in; b.t_in = t_in;
  b.cond = cond; b.cond_dim = c.global_cond_dim > c.local_cond_dim ? c.global_cond_dim : c.local_cond_dim;
  b.grad_flat = grad_flat; b.want_dx = false; b.lay = lay;
  return train_backward(h, b, st);
}

int pfm_epic_forward_train(pfm_epic* h, const float* t_code, int t_rows, const float* x, const float* mask,
                           const float* cond, float* out, int B, int N, void* stream) {
  if (!h || !x || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (t_rows != 1 && t_rows != B) { set_error("t_rows must be 1 or B (got %d, B=%d)", t_rows, B); return PFM_ERR_INVALID; }
  TrainLayout lay;
  return train_forward_common(h, t_code, t_rows, nullptr, 0, x, out, nullptr, nullptr, nullptr, -1, 0.f, mask, cond, B, N,
                              h->cfg.input_dim, 0, &lay, (cudaStream_t)stream);
}

int pfm_epic_backward(pfm_epic* h, const float* t_code, int t_rows, const float* cond, const float* grad_out, float* grad_x,
                      float* grad_flat, int B, int N, void* stream) {
  if (!h || !grad_out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (h->train_B != B || h->train_N != N || h->train_Kx != h->cfg.input_dim) {
    set_error("pfm_epic_backward: no matching pfm_epic_forward_train (saved B=%d N=%d, asked B=%d N=%d)", h->train_B,
              h->train_N, B, N);
    return PFM_ERR_STATE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const pfm_epic_cfg& c = h->cfg;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  TrainLayout lay;
  int rc = train_layout(h, B, N, &lay);
  if (rc != PFM_OK) return rc;
  seed_kernel<<<(B + 7) / 8, 256, 0, st>>>(h->dpre3, grad_out, h->plan.n_real, h->plan.ridx, h->plan.rowoff, B, N, c.feats);
  h->last_launches++;
  if (grad_flat) PFM_CUDA_CHECK(cudaMemsetAsync(grad_flat, 0, sizeof(float) * grad_floats(h), st));
  TrainBwdArgs b;
  b.B = B; b.N = N; b.Kx = c.input_dim; b.xin_off = 0;
  b.t_code = t_code; b.t_ld = (t_rows == B && B > 1) ? c.t_dim : 0; b.t_code_in = nullptr; b.t_in = 0;
  b.cond = cond; b.cond_dim = c.global_cond_dim > c.local_cond_dim ? c.global_cond_dim : c.local_cond_dim;
  b.grad_flat = grad_flat; b.want_dx = grad_x != nullptr; b.lay = lay;
  rc = train_backward(h, b, st);
  if (rc != PFM_OK) return rc;
  if (grad_x) {
    scatter_dx_kernel<<<(B + 7) / 8, 256, 0, st>>>(h->dxs, grad_x, h->plan.n_real, h->plan.ridx, h->plan.rowoff, B, N, c.input_dim);
    h->last_launches++;
  }
  h->train_B = 0;            // the saved dpre3 has been consumed
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // extern "C"
