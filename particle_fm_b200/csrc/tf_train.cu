// Training of the PC-Droid-style set transformers (SURVEY 8 rows a10 / a11 under a7-a9 / a12), fp32 CUDA cores.
//
// The reference trains these networks through torch autograd over the eager modules
// (particle_fm/models/components/droid_transformer.py:211-284, 331-344, 386-397, 529-548, 696-711 under
// losses.py:38-77 / 101-136 / 308-342).  Two facts of that code shape this file:
//   * the networks do NOT mask their output and the losses sum (v - u)^2 over every slot, so padded particles act as
//     queries and contribute to the loss and to the gradients.  Training therefore runs on the DENSE [B*N] rows
//     (row = jet*N + particle); only the KEYS of the attention are restricted to the real particles, exactly like the
//     reference's kv_mask.  (Sampling keeps skipping padding: there padded slots never reach a real one.)
//   * every op is per row except the attention and the per-jet context tables, so one evaluation is a short program of
//     fused linears  Y = [R +] act(LN(X) . W^T + b [+ table[jet]]).  The forward records that program on a tape
//     (tensors live in one arena, nothing is overwritten); the backward replays it in reverse with four kernels per
//     linear: activation gradient + column sums (bias and per-jet table gradients), dX = dY . W through the same row-block
//     GEMM kernel as the forward, LayerNorm backward (which also re-creates the normalised input), and the weight
//     gradient dW = dY^T . LN(X) through epic_train.cu's xty tiles.  Gradients land in ONE flat buffer in state_dict order.
#include <cstring>

#include "pfm_internal.cuh"
#include "simt_common.cuh"
#include "tf_internal.cuh"

namespace pfm {

// ---------------------------------------------------------------------------------------------
// tape
// ---------------------------------------------------------------------------------------------
struct TTensor { size_t off; int rows, w; bool grad; };
struct AttnDesc {
  int Q, q_col, KV, k_col, v_col, O, LSE;     // tensor ids and column offsets
  int q_per_jet, k_per_jet;                   // rows per jet of the query / key tensors
  int keys;                                   // 0: the jet's real particles (n_real / ridx);  > 0: that many keys, all real
};
struct TOp {
  int kind;                                   // 0 linear, 1 attention, 2 token broadcast
  int X, Y, R, JB, jets;                      // tensor ids (-1: none); jets: 0 none, 1 row/N, 2 row/ntok
  int K, k0, act, rows; bool use_bias;
  const TfLN* ln; const TfLinear* L;
  AttnDesc ad;
};

struct TfTape {
  std::vector<TTensor> T;
  std::vector<TOp> ops;
  size_t floats = 0;
  float *data = nullptr, *grad = nullptr; size_t cap = 0;
  float *dxn = nullptr, *xn = nullptr; size_t scratch_cap = 0;
  float* acc = nullptr;                       // [0] sum of squared errors, [1] sum(mask)
  int B = 0, N = 0;
  // plan of the dense training layout: real-particle lists (attention keys) and row -> jet maps
  int *n_real = nullptr, *rowjet = nullptr, *tokjet = nullptr; uint16_t* ridx = nullptr; int planB = 0, planBN = 0;
  int ctxin = -1, xs = -1, out = -1, u = -1;
  bool live = false;
  const float* cond = nullptr;
};

void tf_tape_destroy(pfm_tf* h) {
  if (!h->tape) return;
  for (float* p : {h->tape->data, h->tape->grad, h->tape->dxn, h->tape->xn, h->tape->acc})
    if (p) cudaFree(p);
  for (void* p : {(void*)h->tape->n_real, (void*)h->tape->rowjet, (void*)h->tape->tokjet, (void*)h->tape->ridx})
    if (p) cudaFree(p);
  delete h->tape;
  h->tape = nullptr;
}

static int new_tensor(TfTape& tp, int rows, int w, bool grad = true) {
  TTensor t;
  t.off = tp.floats; t.rows = rows; t.w = w; t.grad = grad;
  tp.floats += ((size_t)rows * w + 63) & ~(size_t)63;
  tp.T.push_back(t);
  return (int)tp.T.size() - 1;
}

static void op_linear(TfTape& tp, int X, int K, const TfLN* ln, const TfLinear& L, int k0, bool use_bias, int JB, int jets, int R, int Y,
                      int act, int rows) {
  TOp o;
  memset(&o, 0, sizeof(o));
  o.kind = 0; o.X = X; o.Y = Y; o.R = R; o.JB = JB; o.jets = jets; o.K = K; o.k0 = k0; o.act = act; o.rows = rows; o.use_bias = use_bias;
  o.ln = ln; o.L = &L;
  tp.ops.push_back(o);
}

// The program of one evaluation (same dataflow as tf_simt.cu::tf_eval, nothing in place).
static void tape_build(pfm_tf* h, TfTape& tp, int B, int N) {
  const pfm_tf_cfg& c = h->cfg;
  const int D = c.model_dim, T = c.t_dim, C = c.cond_dim, F = c.feats, CO = c.ctxt_out, EH = c.embd_hddn, DHd = c.dense_hddn;
  const int hmax = EH > DHd ? EH : DHd;
  const int t_in = c.add_time_to_input ? T : 0;
  const int rows = B * N;
  tp.T.clear(); tp.ops.clear(); tp.floats = 0; tp.B = B; tp.N = N;
  tp.ctxin = new_tensor(tp, B, T + C, false);
  const int c1 = new_tensor(tp, B, EH), ctx = new_tensor(tp, B, CO);
  op_linear(tp, tp.ctxin, T + C, nullptr, h->ctxt.l1, 0, true, -1, 0, -1, c1, 1, B);
  op_linear(tp, c1, EH, &h->ctxt.ln, h->ctxt.l2, 0, true, -1, 0, -1, ctx, 0, B);
  const int n_tab = 2 + (int)h->layers.size();
  std::vector<int> tab(n_tab);
  for (int i = 0; i < n_tab; ++i) tab[i] = new_tensor(tp, B, hmax);
  op_linear(tp, ctx, CO, nullptr, h->node.l1, t_in + F, true, -1, 0, -1, tab[0], 0, B);
  if (t_in > 0) op_linear(tp, tp.ctxin, T, nullptr, h->node.l1, 0, false, -1, 0, tab[0], tab[0], 0, B);
  op_linear(tp, ctx, CO, nullptr, h->outp.l1, D, true, -1, 0, -1, tab[1], 0, B);
  for (size_t l = 0; l < h->layers.size(); ++l) op_linear(tp, ctx, CO, nullptr, h->layers[l].dense.l1, D, true, -1, 0, -1, tab[2 + l], 0, B);
  // dense network on rows: l1 over the main block of columns + per-jet table, LayerNorm, l2 (+ residual)
  auto dense = [&](const TfDense& d, const TfLN* pre, int X, int K, int k0, int table, int jets, int R, int nrows) -> int {
    const int H1 = new_tensor(tp, nrows, d.l1.out);
    op_linear(tp, X, K, pre, d.l1, k0, false, table, jets, -1, H1, 1, nrows);
    const int Y = new_tensor(tp, nrows, d.l2.out);
    op_linear(tp, H1, d.l1.out, &d.ln, d.l2, 0, true, -1, 0, R, Y, 0, nrows);
    return Y;
  };
  tp.xs = new_tensor(tp, rows, F, false);
  int hcur = dense(h->node, nullptr, tp.xs, F, t_in, tab[0], 1, -1, rows);
  auto attention = [&](int Q, int q_col, int KV, int k_col, int v_col, int q_per_jet, int k_per_jet, int keys, int qrows) -> int {
    TOp o;
    memset(&o, 0, sizeof(o));
    o.kind = 1;
    o.ad.Q = Q; o.ad.q_col = q_col; o.ad.KV = KV; o.ad.k_col = k_col; o.ad.v_col = v_col;
    o.ad.O = new_tensor(tp, qrows, D);
    o.ad.LSE = new_tensor(tp, qrows, c.num_heads, false);
    o.ad.q_per_jet = q_per_jet; o.ad.k_per_jet = k_per_jet; o.ad.keys = keys;
    tp.ops.push_back(o);
    return o.ad.O;
  };
  if (c.kind == 0) {
    for (int l = 0; l < c.num_layers; ++l) {
      const TfLayer& Ly = h->layers[l];
      const int QKV = new_tensor(tp, rows, 3 * D);
      op_linear(tp, hcur, D, &Ly.n1, Ly.qkv_or_q, 0, true, -1, 0, -1, QKV, 0, rows);
      const int A = attention(QKV, 0, QKV, D, 2 * D, N, N, 0, rows);
      const int hmid = new_tensor(tp, rows, D);
      op_linear(tp, A, D, &Ly.mha_ln, Ly.out, 0, true, -1, 0, hcur, hmid, 0, rows);
      hcur = dense(Ly.dense, &Ly.n2, hmid, D, 0, tab[2 + l], 1, hmid, rows);
    }
  } else {
    const int nt = c.num_tokens, TR = B * nt;
    int tok = new_tensor(tp, TR, D);
    {
      TOp o;
      memset(&o, 0, sizeof(o));
      o.kind = 2; o.Y = tok; o.rows = TR;
      tp.ops.push_back(o);
    }
    for (int l = 0; l < c.num_layers; ++l) {
      const TfLayer& Fr = h->layers[l];
      const TfLayer& To = h->layers[c.num_layers + l];
      // tokens <- sequence (keys: the real particles)
      const int tq = new_tensor(tp, TR, D), skv = new_tensor(tp, rows, 2 * D);
      op_linear(tp, tok, D, &Fr.n1, Fr.qkv_or_q, 0, true, -1, 0, -1, tq, 0, TR);
      op_linear(tp, hcur, D, &Fr.n0, Fr.kv, 0, true, -1, 0, -1, skv, 0, rows);
      const int tA = attention(tq, 0, skv, 0, D, nt, N, 0, TR);
      const int tmid = new_tensor(tp, TR, D);
      op_linear(tp, tA, D, &Fr.mha_ln, Fr.out, 0, true, -1, 0, tok, tmid, 0, TR);
      tok = dense(Fr.dense, &Fr.n2, tmid, D, 0, tab[2 + l], 2, tmid, TR);
      // sequence <- tokens (no mask)
      const int sq = new_tensor(tp, rows, D), tkv = new_tensor(tp, TR, 2 * D);
      op_linear(tp, hcur, D, &To.n1, To.qkv_or_q, 0, true, -1, 0, -1, sq, 0, rows);
      op_linear(tp, tok, D, &To.n0, To.kv, 0, true, -1, 0, -1, tkv, 0, TR);
      const int sA = attention(sq, 0, tkv, 0, D, N, nt, nt, rows);
      const int hmid = new_tensor(tp, rows, D);
      op_linear(tp, sA, D, &To.mha_ln, To.out, 0, true, -1, 0, hcur, hmid, 0, rows);
      hcur = dense(To.dense, &To.n2, hmid, D, 0, tab[2 + c.num_layers + l], 1, hmid, rows);
    }
  }
  const TfLN* fin = c.kind == 0 ? &h->final_norm : nullptr;
  tp.out = dense(h->outp, fin, hcur, D, 0, tab[1], 1, -1, rows);
  tp.u = new_tensor(tp, rows, F, false);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__global__ void tt_ctxin_kernel(const float* __restrict__ t_code, int t_stride, int T, const float* __restrict__ cond, int C, int rows,
                                float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = T + C;
  if (idx >= rows * W) return;
  const int r = idx / W, c = idx - r * W;
  out[idx] = c < T ? t_code[(size_t)r * t_stride + c] : cond[(size_t)r * C + (c - T)];
}

__global__ void tt_tokens_kernel(const float* __restrict__ tok0, int ntokD, int total, float* __restrict__ tok) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < total) tok[idx] = tok0[idx % ntokD];
}
// d tok0[i] = sum_b d tok[b*ntokD + i]
__global__ void tt_tokens_bwd_kernel(const float* __restrict__ dtok, int ntokD, int B, float* __restrict__ dtok0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntokD) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += dtok[(size_t)b * ntokD + i];
  atomicAdd(dtok0 + i, s);
}

__global__ void tt_axpy_kernel(float* __restrict__ y, const float* __restrict__ x, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// dPre = dY * act'(Y) (in place) and column sums of dPre over a chunk of rows: bias gradient (split into two state_dict
// parts at column `split`) and per-jet table gradient  djb[jet(row)][o] += dPre[row][o].
static constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(256) tt_colsum_kernel(float* __restrict__ dY, int ld, const float* __restrict__ Yact, float slope,
                                                        int rows, int N, float* __restrict__ gb0, float* __restrict__ gb1, int split,
                                                        float* __restrict__ djb, int jb_stride, int rows_per_jet) {
  const int r0 = blockIdx.x * CS_ROWS;
  const int r1 = min(rows, r0 + CS_ROWS);
  for (int o = threadIdx.x; o < N; o += blockDim.x) {
    float tot = 0.f, seg = 0.f;
    int jet = djb ? r0 / rows_per_jet : 0;
    for (int r = r0; r < r1; ++r) {
      float g = dY[(size_t)r * ld + o];
      if (Yact) {
        if (!(Yact[(size_t)r * ld + o] > 0.f)) g *= slope;
        dY[(size_t)r * ld + o] = g;
      }
      tot += g;
      if (djb) {
        const int j = r / rows_per_jet;
        if (j != jet) { atomicAdd(djb + (size_t)jet * jb_stride + o, seg); seg = 0.f; jet = j; }
        seg += g;
      }
    }
    if (djb) atomicAdd(djb + (size_t)jet * jb_stride + o, seg);
    if (gb0) atomicAdd(o < split ? gb0 + o : gb1 + (o - split), tot);
  }
}

// LayerNorm backward, one warp per row (K <= 512):  xhat = (x - mean) rstd,  dxhat = dXn g,
//   dX += rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat));   Xn = xhat g + b (re-created for the weight gradient);
//   dg += sum_rows dXn xhat,  dbeta += sum_rows dXn.
__global__ void __launch_bounds__(256) tt_ln_bwd_kernel(const float* __restrict__ dXn, int ldd, const float* __restrict__ X, int ldx,
                                                        int K, const float* __restrict__ g, const float* __restrict__ b, float eps,
                                                        int rows, float* __restrict__ dX, int ldg, float* __restrict__ Xn, int ldn,
                                                        float* __restrict__ dg, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float ag[16], ab[16], gv[16], bv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    ag[i] = 0.f; ab[i] = 0.f;
    gv[i] = c < K ? g[c] : 0.f; bv[i] = c < K ? b[c] : 0.f;
  }
  const float invK = 1.f / (float)K;
  for (int row = blockIdx.x * wpb + warp; row < rows; row += gridDim.x * wpb) {
    float xv[16], dv[16];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      xv[i] = c < K ? X[(size_t)row * ldx + c] : 0.f;
      dv[i] = (c < K && dXn) ? dXn[(size_t)row * ldd + c] : 0.f;
      s += xv[i];
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
    const float mean = s * invK;
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (lane + 32 * i < K) { const float d = xv[i] - mean; v = fmaf(d, d, v); }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sh);
    const float rstd = rsqrtf(v * invK + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < K) {
        const float xh = (xv[i] - mean) * rstd;
        const float dxh = dv[i] * gv[i];
        s1 += dxh; s2 = fmaf(dxh, xh, s2);
        ag[i] = fmaf(dv[i], xh, ag[i]); ab[i] += dv[i];
        xv[i] = xh;
      }
    }
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, sh); s2 += __shfl_xor_sync(0xffffffffu, s2, sh); }
    s1 *= invK; s2 *= invK;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;
      if (c < K) {
        if (Xn) Xn[(size_t)row * ldn + c] = fmaf(xv[i], gv[i], bv[i]);
        if (dX) dX[(size_t)row * ldg + c] += rstd * (dv[i] * gv[i] - s1 - xv[i] * s2);
      }
    }
  }
  if (dg) {                                        // block-level sum of the warps' partial gain / shift gradients, then one
    __shared__ float red[8][512];                  // atomic per column and block
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 16; ++i) red[warp][lane + 32 * i] = pass == 0 ? ag[i] : ab[i];
      __syncthreads();
      for (int c = threadIdx.x; c < K; c += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < wpb; ++w) t += red[w][c];
        atomicAdd((pass == 0 ? dg : dbeta) + c, t);
      }
    }
  }
}

// dX[row][c] += sum_o dY[row][o] W[o][k0 + c]   for a handful of rows or a column offset the row-block GEMM cannot address
__global__ void tt_dx_small_kernel(const float* __restrict__ dY, int ldy, int out, const float* __restrict__ W, int ldw, int k0, int K,
                                   int rows, float* __restrict__ dX, int ldx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * K) return;
  const int r = idx / K, c = idx - r * K;
  const float* dy = dY + (size_t)r * ldy;
  const float* w = W + k0 + c;
  float a0 = 0.f, a1 = 0.f;
  int o = 0;
  for (; o + 1 < out; o += 2) { a0 = fmaf(dy[o], w[(size_t)o * ldw], a0); a1 = fmaf(dy[o + 1], w[(size_t)(o + 1) * ldw], a1); }
  if (o < out) a0 = fmaf(dy[o], w[(size_t)o * ldw], a0);
  dX[(size_t)r * ldx + c] += a0 + a1;
}

// ---- attention ---------------------------------------------------------------------------------
struct AttnArgs {
  const float* Q; int ldq; const float* K; const float* V; int ldkv;
  float* O; int ldo; float* lse; int heads;
  const float* dO; float* dQ; float* dK; float* dV;       // backward only (same leading dimensions as Q / K,V)
  int q_per_jet, k_per_jet, keys;
  const int* n_real; const uint16_t* ridx; int ridx_stride;
  float scale;
};

// forward with the log-sum-exp saved per (query row, head): grid (B, heads), thread = query, K/V of the head in smem
template <int DH>
__global__ void __launch_bounds__(128) tt_attn_fwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int jet = blockIdx.x, head = blockIdx.y;
  const int nk = a.keys > 0 ? a.keys : a.n_real[jet];
  float* Ks = sm;
  float* Vs = sm + (size_t)nk * DH;
  for (int i = threadIdx.x; i < nk * (DH / 4); i += blockDim.x) {
    const int t = i / (DH / 4), d4 = i - t * (DH / 4);
    const int krow = jet * a.k_per_jet + (a.keys > 0 ? t : (int)a.ridx[(size_t)jet * a.ridx_stride + t]);
    *reinterpret_cast<float4*>(Ks + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(a.K + (size_t)krow * a.ldkv + head * DH + d4 * 4);
    *reinterpret_cast<float4*>(Vs + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(a.V + (size_t)krow * a.ldkv + head * DH + d4 * 4);
  }
  __syncthreads();
  const float sl2 = a.scale * 1.4426950408889634f;
  for (int t = threadIdx.x; t < a.q_per_jet; t += blockDim.x) {
    const size_t qrow = (size_t)jet * a.q_per_jet + t;
    float q[DH], o[DH];
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      const float4 v = *reinterpret_cast<const float4*>(a.Q + qrow * a.ldq + head * DH + d4 * 4);
      q[d4 * 4 + 0] = v.x * sl2; q[d4 * 4 + 1] = v.y * sl2; q[d4 * 4 + 2] = v.z * sl2; q[d4 * 4 + 3] = v.w * sl2;
    }
#pragma unroll
    for (int d = 0; d < DH; ++d) o[d] = 0.f;
    float m = -INFINITY, l = 0.f;
    for (int k = 0; k < nk; ++k) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) s = fmaf(q[d], Ks[k * DH + d], s);
      const float mn = fmaxf(m, s);
      const float corr = exp2f(m - mn), p = exp2f(s - mn);
      l = fmaf(l, corr, p);
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] = fmaf(p, Vs[k * DH + d], o[d] * corr);
      m = mn;
    }
    const float inv = 1.f / l;        // no keys: 0/0 = NaN, like a softmax over an empty key set in the reference
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4)
      *reinterpret_cast<float4*>(a.O + qrow * a.ldo + head * DH + d4 * 4) =
          make_float4(o[d4 * 4] * inv, o[d4 * 4 + 1] * inv, o[d4 * 4 + 2] * inv, o[d4 * 4 + 3] * inv);
    a.lse[qrow * a.heads + head] = m + log2f(l);
  }
}

// backward: P = exp2(s - lse), D_i = dO_i . O_i, dS = P (dO_i . V_j - D_i);  pass A (thread = query) forms dQ, pass B
// (thread = key) re-forms the scores and accumulates dK, dV -- no atomics, every output element has one owner.
template <int DH>
__global__ void __launch_bounds__(128) tt_attn_bwd_kernel(const AttnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int jet = blockIdx.x, head = blockIdx.y;
  const int nk = a.keys > 0 ? a.keys : a.n_real[jet];
  const int nq = a.q_per_jet;
  float* Ks = sm;
  float* Vs = Ks + (size_t)nk * DH;
  float* Qs = Vs + (size_t)nk * DH;       // pre-scaled by scale*log2(e)
  float* dOs = Qs + (size_t)nq * DH;
  float* Ls = dOs + (size_t)nq * DH;
  float* Ds = Ls + nq;
  const float sl2 = a.scale * 1.4426950408889634f;
  for (int i = threadIdx.x; i < nk * (DH / 4); i += blockDim.x) {
    const int t = i / (DH / 4), d4 = i - t * (DH / 4);
    const int krow = jet * a.k_per_jet + (a.keys > 0 ? t : (int)a.ridx[(size_t)jet * a.ridx_stride + t]);
    *reinterpret_cast<float4*>(Ks + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(a.K + (size_t)krow * a.ldkv + head * DH + d4 * 4);
    *reinterpret_cast<float4*>(Vs + t * DH + d4 * 4) = *reinterpret_cast<const float4*>(a.V + (size_t)krow * a.ldkv + head * DH + d4 * 4);
  }
  for (int t = threadIdx.x; t < nq; t += blockDim.x) {
    const size_t qrow = (size_t)jet * nq + t;
    float dd = 0.f;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      const float4 qv = *reinterpret_cast<const float4*>(a.Q + qrow * a.ldq + head * DH + d4 * 4);
      const float4 gv = *reinterpret_cast<const float4*>(a.dO + qrow * a.ldo + head * DH + d4 * 4);
      const float4 ov = *reinterpret_cast<const float4*>(a.O + qrow * a.ldo + head * DH + d4 * 4);
      *reinterpret_cast<float4*>(Qs + t * DH + d4 * 4) = make_float4(qv.x * sl2, qv.y * sl2, qv.z * sl2, qv.w * sl2);
      *reinterpret_cast<float4*>(dOs + t * DH + d4 * 4) = gv;
      dd = fmaf(gv.x, ov.x, dd); dd = fmaf(gv.y, ov.y, dd); dd = fmaf(gv.z, ov.z, dd); dd = fmaf(gv.w, ov.w, dd);
    }
    Ls[t] = a.lse[qrow * a.heads + head];
    Ds[t] = dd;
  }
  __syncthreads();
  // pass A: dQ
  for (int t = threadIdx.x; t < nq; t += blockDim.x) {
    float q[DH], g[DH], dq[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { q[d] = Qs[t * DH + d]; g[d] = dOs[t * DH + d]; dq[d] = 0.f; }
    const float lse = Ls[t], dd = Ds[t];
    for (int k = 0; k < nk; ++k) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(q[d], Ks[k * DH + d], s); dp = fmaf(g[d], Vs[k * DH + d], dp); }
      const float ds = exp2f(s - lse) * (dp - dd);
#pragma unroll
      for (int d = 0; d < DH; ++d) dq[d] = fmaf(ds, Ks[k * DH + d], dq[d]);
    }
    const size_t qrow = (size_t)jet * nq + t;
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4)
      *reinterpret_cast<float4*>(a.dQ + qrow * a.ldq + head * DH + d4 * 4) =
          make_float4(dq[d4 * 4] * a.scale, dq[d4 * 4 + 1] * a.scale, dq[d4 * 4 + 2] * a.scale, dq[d4 * 4 + 3] * a.scale);
  }
  // pass B: dK, dV   (Qs holds q * scale * log2e: dK = scale * sum dS q  =  sum dS Qs / log2e)
  for (int k = threadIdx.x; k < nk; k += blockDim.x) {
    float kk[DH], vv[DH], dk[DH], dv[DH];
#pragma unroll
    for (int d = 0; d < DH; ++d) { kk[d] = Ks[k * DH + d]; vv[d] = Vs[k * DH + d]; dk[d] = 0.f; dv[d] = 0.f; }
    for (int t = 0; t < nq; ++t) {
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) { s = fmaf(Qs[t * DH + d], kk[d], s); dp = fmaf(dOs[t * DH + d], vv[d], dp); }
      const float p = exp2f(s - Ls[t]);
      const float ds = p * (dp - Ds[t]);
#pragma unroll
      for (int d = 0; d < DH; ++d) { dv[d] = fmaf(p, dOs[t * DH + d], dv[d]); dk[d] = fmaf(ds, Qs[t * DH + d], dk[d]); }
    }
    const int krow = jet * a.k_per_jet + (a.keys > 0 ? k : (int)a.ridx[(size_t)jet * a.ridx_stride + k]);
    const float kscale = 0.6931471805599453f;      // 1 / log2(e)
#pragma unroll
    for (int d4 = 0; d4 < DH / 4; ++d4) {
      *reinterpret_cast<float4*>(a.dK + (size_t)krow * a.ldkv + head * DH + d4 * 4) =
          make_float4(dk[d4 * 4] * kscale, dk[d4 * 4 + 1] * kscale, dk[d4 * 4 + 2] * kscale, dk[d4 * 4 + 3] * kscale);
      *reinterpret_cast<float4*>(a.dV + (size_t)krow * a.ldkv + head * DH + d4 * 4) =
          make_float4(dv[d4 * 4], dv[d4 * 4 + 1], dv[d4 * 4 + 2], dv[d4 * 4 + 3]);
    }
  }
}

// ---- loss ---------------------------------------------------------------------------------------
// flow-matching interpolation (losses.py:56-62 FM-OT, :115-119 CFM, :320-326 droid) on the dense rows; t per jet
__global__ void tt_interp_kernel(const float* __restrict__ x1, const float* __restrict__ t, const float* __restrict__ n0,
                                 const float* __restrict__ n1, const float* __restrict__ mask, int kind, float sigma, int N, int F,
                                 size_t total, float* __restrict__ y, float* __restrict__ u) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t row = i / F;
  const float tt = t[row / N], m = mask[row];
  const float x = x1[i], z = n0[i];
  float yy, uu;
  if (kind == PFM_LOSS_FM_OT) {
    yy = __fadd_rn(__fmul_rn(1.f - tt, x), __fmul_rn(__fadd_rn(sigma, __fmul_rn(1.f - sigma, tt)), z));
    uu = __fmul_rn(__fadd_rn(__fmul_rn(1.f - sigma, z), -x), m);
  } else if (kind == PFM_LOSS_CFM) {
    yy = __fadd_rn(__fadd_rn(__fmul_rn(1.f - tt, x), __fmul_rn(tt, z)), __fmul_rn(sigma, n1[i]));
    uu = __fmul_rn(__fadd_rn(z, -x), m);
  } else {
    yy = __fadd_rn(x, __fmul_rn(tt, z));
    uu = __fmul_rn(z, m);
  }
  y[i] = yy;
  u[i] = uu;
}

__global__ void tt_sum_kernel(const float* __restrict__ v, size_t n, float* __restrict__ out) {
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += v[i];
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// loss = sum (v - u)^2 / sum(mask) over EVERY slot (the reference does not mask v);  dv = 2 (v - u) / sum(mask)
__global__ void tt_seed_kernel(const float* __restrict__ v, const float* __restrict__ u, size_t n, float* __restrict__ acc,
                               float* __restrict__ dv) {
  float s = 0.f;
  const float inv = 1.f / acc[1];
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = v[i] - u[i];
    s = fmaf(d, d, s);
    if (dv) dv[i] = 2.f * d * inv;
  }
#pragma unroll
  for (int sh = 16; sh > 0; sh >>= 1) s += __shfl_xor_sync(0xffffffffu, s, sh);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc, s);
}
__global__ void tt_loss_out_kernel(const float* __restrict__ acc, float* __restrict__ loss) { *loss = acc[0] / acc[1]; }

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int tape_ensure(pfm_tf* h, int B, int N) {
  if (!h->tape) h->tape = new TfTape();
  TfTape& tp = *h->tape;
  if (tp.B != B || tp.N != N || tp.ops.empty()) tape_build(h, tp, B, N);
  if (tp.floats > tp.cap) {
    if (tp.data) cudaFree(tp.data);
    if (tp.grad) cudaFree(tp.grad);
    tp.data = tp.grad = nullptr; tp.cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(&tp.data, sizeof(float) * tp.floats));
    PFM_CUDA_CHECK(cudaMalloc(&tp.grad, sizeof(float) * tp.floats));
    tp.cap = tp.floats;
  }
  const pfm_tf_cfg& c = h->cfg;
  int wmax = c.model_dim;
  for (int w : {c.embd_hddn, c.dense_hddn, c.t_dim + c.cond_dim, c.ctxt_out}) wmax = w > wmax ? w : wmax;
  const size_t need = (size_t)B * N * wmax + 64;
  if (need > tp.scratch_cap) {
    if (tp.dxn) cudaFree(tp.dxn);
    if (tp.xn) cudaFree(tp.xn);
    tp.dxn = tp.xn = nullptr; tp.scratch_cap = 0;
    PFM_CUDA_CHECK(cudaMalloc(&tp.dxn, sizeof(float) * need));
    PFM_CUDA_CHECK(cudaMalloc(&tp.xn, sizeof(float) * need));
    tp.scratch_cap = need;
  }
  if (!tp.acc) PFM_CUDA_CHECK(cudaMalloc(&tp.acc, sizeof(float) * 2));
  return PFM_OK;
}

static inline float* tdata(TfTape& tp, int id) { return tp.data + tp.T[id].off; }
static inline float* tgrad(TfTape& tp, int id) { return tp.T[id].grad ? tp.grad + tp.T[id].off : nullptr; }

template <int DH>
static int attn_launch(pfm_tf* h, const AttnArgs& a, int B, bool bwd, int nk_max, cudaStream_t st) {
  const size_t smem = sizeof(float) * (bwd ? (2 * (size_t)nk_max * DH + 2 * (size_t)a.q_per_jet * DH + 2 * (size_t)a.q_per_jet)
                                           : 2 * (size_t)nk_max * DH);
  if ((int)smem > h->max_smem) { set_error("attention (training): %zu B of shared memory needed for %d keys / %d queries", smem, nk_max, a.q_per_jet); return PFM_ERR_UNSUPPORTED; }
  if (bwd) {
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tt_attn_bwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    tt_attn_bwd_kernel<DH><<<dim3(B, a.heads), 128, smem, st>>>(a);
  } else {
    PFM_CUDA_CHECK(cudaFuncSetAttribute(tt_attn_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
    tt_attn_fwd_kernel<DH><<<dim3(B, a.heads), 128, smem, st>>>(a);
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

static int run_attention(pfm_tf* h, TfTape& tp, const AttnDesc& d, bool bwd, cudaStream_t st) {
  const pfm_tf_cfg& c = h->cfg;
  const int dh = c.model_dim / c.num_heads;
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = tdata(tp, d.Q) + d.q_col; a.ldq = tp.T[d.Q].w;
  a.K = tdata(tp, d.KV) + d.k_col; a.V = tdata(tp, d.KV) + d.v_col; a.ldkv = tp.T[d.KV].w;
  a.O = tdata(tp, d.O); a.ldo = tp.T[d.O].w; a.lse = tdata(tp, d.LSE); a.heads = c.num_heads;
  if (bwd) {
    a.dO = tgrad(tp, d.O); a.dQ = tgrad(tp, d.Q) + d.q_col; a.dK = tgrad(tp, d.KV) + d.k_col; a.dV = tgrad(tp, d.KV) + d.v_col;
  }
  a.q_per_jet = d.q_per_jet; a.k_per_jet = d.k_per_jet; a.keys = d.keys;
  a.n_real = tp.n_real; a.ridx = tp.ridx; a.ridx_stride = tp.N;
  a.scale = 1.f / sqrtf((float)dh);
  const int nk_max = d.keys > 0 ? d.keys : tp.N;
  h->last_launches++;
  if (dh == 16) return attn_launch<16>(h, a, tp.B, bwd, nk_max, st);
  if (dh == 8) return attn_launch<8>(h, a, tp.B, bwd, nk_max, st);
  if (dh == 4) return attn_launch<4>(h, a, tp.B, bwd, nk_max, st);
  set_error("attention (training): head dim %d not supported (4, 8, 16)", dh);
  return PFM_ERR_UNSUPPORTED;
}

static void fill_lin(pfm_tf* h, TfTape& tp, const TOp& o, LinArgs& a) {
  memset(&a, 0, sizeof(a));
  a.X = tdata(tp, o.X); a.ldx = tp.T[o.X].w; a.K = o.K;
  a.ln_g = o.ln ? o.ln->g : nullptr; a.ln_b = o.ln ? o.ln->b : nullptr;
  a.Wt = o.L->Wt + (size_t)o.k0 * o.L->ldo; a.ldo = o.L->ldo; a.N = o.L->out;
  a.bias = o.use_bias ? o.L->b : nullptr;
  if (o.JB >= 0) { a.jb = tdata(tp, o.JB); a.jb_stride = tp.T[o.JB].w; a.rowjet = o.jets == 2 ? tp.tokjet : tp.rowjet; }
  if (o.R >= 0) { a.R = tdata(tp, o.R); a.ldr = tp.T[o.R].w; }
  a.Y = tdata(tp, o.Y); a.ldy = tp.T[o.Y].w;
  a.act = o.act; a.slope = h->cfg.neg_slope; a.eps = h->cfg.ln_eps;
  a.rows = o.rows;
  a.img = o.L->img; a.img_kblocks = o.L->kblocks; a.kb0 = o.k0 / 64;
}

__global__ void tt_rowjet_kernel(int* __restrict__ rowjet, int rows, int per) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < rows) rowjet[i] = i / per;
}

// dense row maps of the training layout (row -> jet) and the real-particle lists for the attention keys
static int train_plan(pfm_tf* h, TfTape& tp, const float* mask, int B, int N, cudaStream_t st) {
  const int rows = B * N, TR = B * h->cfg.num_tokens;
  if (B > tp.planB) {
    for (void* p : {(void*)tp.n_real, (void*)tp.tokjet}) if (p) cudaFree(p);
    tp.n_real = tp.tokjet = nullptr; tp.planB = 0;
    PFM_CUDA_CHECK(cudaMalloc(&tp.n_real, sizeof(int) * B));
    PFM_CUDA_CHECK(cudaMalloc(&tp.tokjet, sizeof(int) * TR));
    tp.planB = B;
  }
  if (rows > tp.planBN) {
    for (void* p : {(void*)tp.ridx, (void*)tp.rowjet}) if (p) cudaFree(p);
    tp.ridx = nullptr; tp.rowjet = nullptr; tp.planBN = 0;
    PFM_CUDA_CHECK(cudaMalloc(&tp.ridx, sizeof(uint16_t) * rows));
    PFM_CUDA_CHECK(cudaMalloc(&tp.rowjet, sizeof(int) * rows));
    tp.planBN = rows;
  }
  plan_count_kernel<<<(B + 7) / 8, 256, 0, st>>>(mask, B, N, tp.n_real, tp.ridx);
  tt_rowjet_kernel<<<(rows + 255) / 256, 256, 0, st>>>(tp.rowjet, rows, N);
  tt_rowjet_kernel<<<(TR + 255) / 256, 256, 0, st>>>(tp.tokjet, TR, h->cfg.num_tokens);
  h->last_launches += 3;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

static int tape_forward(pfm_tf* h, TfTape& tp, cudaStream_t st) {
  int rc;
  for (const TOp& o : tp.ops) {
    if (o.kind == 0) {
      LinArgs a;
      fill_lin(h, tp, o, a);
      if ((rc = tf_launch_linear(h, a, o.L->ldo, (o.k0 % 64) == 0, st)) != PFM_OK) return rc;      // bf16 mode: tcgen05 linears
    } else if (o.kind == 1) {
      if ((rc = run_attention(h, tp, o.ad, false, st)) != PFM_OK) return rc;
    } else {
      const int total = o.rows * h->cfg.model_dim;
      tt_tokens_kernel<<<(total + 255) / 256, 256, 0, st>>>(h->tok0, h->cfg.num_tokens * h->cfg.model_dim, total, tdata(tp, o.Y));
      h->last_launches++;
    }
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

// dXout[rows, K] = [R +] dY[rows, N] . Wrow[:, k0 : k0+K]   (fp32 row-block GEMM, or tcgen05 bf16 in PFM_PREC_BF16 mode;
// the tensor-core kernel holds a [128 x <=512] A tile, so wider dY (the fused QKV projection: 768) goes in slices)
static int dx_linear(pfm_tf* h, const TfLinear& L, const float* dY, int ldy, int N, int k0, int K, const float* R, int ldr, float* out,
                     int ldo, int rows, cudaStream_t st) {
  LinArgs a;
  memset(&a, 0, sizeof(a));
  a.X = dY; a.ldx = ldy; a.K = N;
  a.Wt = L.Wrow + k0; a.ldo = L.ldw; a.N = K;
  a.R = R; a.ldr = ldr; a.Y = out; a.ldy = ldo; a.rows = rows;
  const bool tc = h->precision == PFM_PREC_BF16 && L.img_bwd && k0 == 0 && rows >= 256 && (K & 127) == 0 && (N & 63) == 0 &&
                  (ldy & 3) == 0 && (ldo & 3) == 0 && (!R || (ldr & 3) == 0);
  if (!tc) return tf_launch_linear(h, a, L.ldw, false, st);
  for (int n0 = 0; n0 < N; n0 += 512) {
    LinArgs b = a;
    b.X = dY + n0; b.K = N - n0 < 512 ? N - n0 : 512;
    b.img = L.img_bwd; b.img_kblocks = L.kblocks_bwd; b.kb0 = n0 / 64;
    if (n0 > 0) { b.R = out; b.ldr = ldo; }
    int rc = tf_launch_linear(h, b, L.ldw, true, st);
    if (rc != PFM_OK) return rc;
  }
  return PFM_OK;
}

static int linear_backward(pfm_tf* h, TfTape& tp, const TOp& o, float* flat, cudaStream_t st) {
  const TfLinear& L = *o.L;
  const int N = L.out, rows = o.rows;
  if (rows <= 0) return PFM_OK;
  float* dY = tgrad(tp, o.Y);
  const int ldy = tp.T[o.Y].w;
  int rc;
  // 1. activation gradient (in place) + bias / per-jet table gradients
  {
    float* gb0 = o.use_bias ? flat + L.gb_off[0] : nullptr;
    float* gb1 = (o.use_bias && L.n_parts > 1) ? flat + L.gb_off[1] : gb0;
    const int split = L.n_parts > 1 ? L.split : N;
    float* djb = o.JB >= 0 ? tgrad(tp, o.JB) : nullptr;
    const int per = o.jets == 2 ? h->cfg.num_tokens : tp.N;
    if (o.act || gb0 || djb) {
      tt_colsum_kernel<<<(rows + CS_ROWS - 1) / CS_ROWS, 256, 0, st>>>(dY, ldy, o.act ? tdata(tp, o.Y) : nullptr, h->cfg.neg_slope, rows, N,
                                                                      gb0, gb1, split, djb, djb ? tp.T[o.JB].w : 0, per);
      h->last_launches++;
    }
  }
  // 2. residual branch
  if (o.R >= 0 && o.R != o.Y && tgrad(tp, o.R)) {
    if (tp.T[o.R].w != ldy) { set_error("internal: residual width mismatch"); return PFM_ERR_INVALID; }
    const size_t n = (size_t)rows * ldy;
    tt_axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(tgrad(tp, o.R), dY, n);
    h->last_launches++;
  }
  // 3. input gradient
  float* dX = tgrad(tp, o.X);
  const float* Xsrc = tdata(tp, o.X);
  int ldxs = tp.T[o.X].w;
  if (o.ln) {
    if (!dX) { set_error("internal: LayerNorm input without gradient"); return PFM_ERR_INVALID; }
    if ((rc = dx_linear(h, L, dY, ldy, N, o.k0, o.K, nullptr, 0, tp.dxn, o.K, rows, st)) != PFM_OK) return rc;
    const int blocks = (rows + 7) / 8 < 2 * h->sm_count ? (rows + 7) / 8 : 2 * h->sm_count;
    tt_ln_bwd_kernel<<<blocks, 256, 0, st>>>(tp.dxn, o.K, Xsrc, ldxs, o.K, o.ln->g, o.ln->b, h->cfg.ln_eps, rows, dX, ldxs, tp.xn, o.K,
                                             flat + o.ln->gg_off, flat + o.ln->gb_off);
    h->last_launches++;
    Xsrc = tp.xn; ldxs = o.K;
  } else if (dX) {
    if ((o.k0 & 3) == 0 && rows >= 64) {
      if ((rc = dx_linear(h, L, dY, ldy, N, o.k0, o.K, dX, tp.T[o.X].w, dX, tp.T[o.X].w, rows, st)) != PFM_OK) return rc;
    } else {
      const int total = rows * o.K;
      tt_dx_small_kernel<<<(total + 127) / 128, 128, 0, st>>>(dY, ldy, N, L.Wrow, L.ldw, o.k0, o.K, rows, dX, tp.T[o.X].w);
      h->last_launches++;
    }
  }
  // 4. weight gradient  dW[o][k0 + c] += sum_r dPre[r][o] Xn[r][c]
  for (int p = 0; p < (L.n_parts > 1 ? 2 : 1); ++p) {
    const int o0 = p == 0 ? 0 : L.split;
    const int o1 = (L.n_parts > 1 && p == 0) ? L.split : N;
    rc = xty_use_simt() ? xty_launch_one(dY + o0, ldy, Xsrc, ldxs, flat + L.gw_off[p], L.in, o1 - o0, o.K, o.k0, rows, st)
                        : xty_tc_launch_one(dY + o0, ldy, Xsrc, ldxs, flat + L.gw_off[p], L.in, o1 - o0, o.K, o.k0, rows, h->sm_count, st);
    if (rc != PFM_OK) return rc;
    h->last_launches++;
  }
  return PFM_OK;
}

static int tape_backward(pfm_tf* h, TfTape& tp, float* flat, cudaStream_t st) {
  int rc;
  for (size_t i = tp.ops.size(); i-- > 0;) {
    const TOp& o = tp.ops[i];
    if (o.kind == 0) {
      if ((rc = linear_backward(h, tp, o, flat, st)) != PFM_OK) return rc;
    } else if (o.kind == 1) {
      if ((rc = run_attention(h, tp, o.ad, true, st)) != PFM_OK) return rc;
    } else {
      const int ntokD = h->cfg.num_tokens * h->cfg.model_dim;
      tt_tokens_bwd_kernel<<<(ntokD + 127) / 128, 128, 0, st>>>(tgrad(tp, o.Y), ntokD, tp.B, flat + h->tok0_goff);
      h->last_launches++;
    }
  }
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

static int train_check(pfm_tf* h, const float* t_code, const float* cond, const float* mask, int B, int N) {
  if (!h->weights_set) { set_error("weights not set (call pfm_tf_set_weights first)"); return PFM_ERR_STATE; }
  if (B <= 0 || N <= 0 || N > 65535) { set_error("bad batch shape B=%d N=%d", B, N); return PFM_ERR_INVALID; }
  if (!t_code) { set_error("time code is NULL"); return PFM_ERR_INVALID; }
  if (!mask) { set_error("mask is NULL: the droid transformers need a mask (droid_transformer.py:537)"); return PFM_ERR_INVALID; }
  if (h->cfg.cond_dim > 0 && !cond) { set_error("cond is NULL but the net is conditioned"); return PFM_ERR_INVALID; }
  if ((long long)B * N > (1ll << 30)) { set_error("batch too large"); return PFM_ERR_INVALID; }
  if (h->cfg.embd_hddn > 512 || h->cfg.dense_hddn > 512 || h->cfg.model_dim > 512) { set_error("training: widths above 512 are not supported"); return PFM_ERR_UNSUPPORTED; }
  return PFM_OK;
}

// plan + context input + forward of the tape on x (dense [B*N, F]); leaves the result in the tape's output tensor
static int train_forward(pfm_tf* h, const float* t_code, int t_rows, const float* x_dense, const float* mask, const float* cond, int B,
                         int N, cudaStream_t st) {
  int rc;
  if ((rc = tape_ensure(h, B, N)) != PFM_OK) return rc;
  TfTape& tp = *h->tape;
  if ((rc = train_plan(h, tp, mask, B, N, st)) != PFM_OK) return rc;
  const pfm_tf_cfg& c = h->cfg;
  const int W = c.t_dim + c.cond_dim;
  tt_ctxin_kernel<<<(B * W + 255) / 256, 256, 0, st>>>(t_code, t_rows == 1 ? 0 : c.t_dim, c.t_dim, cond, c.cond_dim, B, tdata(tp, tp.ctxin));
  h->last_launches++;
  if (x_dense) PFM_CUDA_CHECK(cudaMemcpyAsync(tdata(tp, tp.xs), x_dense, sizeof(float) * (size_t)B * N * c.feats, cudaMemcpyDeviceToDevice, st));
  if ((rc = tape_forward(h, tp, st)) != PFM_OK) return rc;
  tp.live = true;
  return PFM_OK;
}

}  // namespace pfm

using namespace pfm;

extern "C" {

int64_t pfm_tf_grad_size(const pfm_tf* h) { return h ? (int64_t)h->grad_floats : (int64_t)PFM_ERR_INVALID; }

int pfm_tf_forward_train(pfm_tf* h, const float* t_code, int t_rows, const float* x, const float* mask, const float* cond, float* out,
                         int B, int N, void* stream) {
  if (!h || !x || !out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (t_rows != 1 && t_rows != B) { set_error("t_rows must be 1 or B"); return PFM_ERR_INVALID; }
  int rc = train_check(h, t_code, cond, mask, B, N);
  if (rc != PFM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  h->last_launches = 0;
  if ((rc = train_forward(h, t_code, t_rows, x, mask, cond, B, N, st)) != PFM_OK) return rc;
  TfTape& tp = *h->tape;
  PFM_CUDA_CHECK(cudaMemcpyAsync(out, tdata(tp, tp.out), sizeof(float) * (size_t)B * N * h->cfg.feats, cudaMemcpyDeviceToDevice, st));
  return PFM_OK;
}

int pfm_tf_backward(pfm_tf* h, const float* dout, float* grad_flat, void* stream) {
  if (!h || !dout || !grad_flat) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (!h->tape || !h->tape->live) { set_error("pfm_tf_backward without a saved forward (call pfm_tf_forward_train first)"); return PFM_ERR_STATE; }
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  TfTape& tp = *h->tape;
  h->last_launches = 0;
  PFM_CUDA_CHECK(cudaMemsetAsync(tp.grad, 0, sizeof(float) * tp.floats, st));
  PFM_CUDA_CHECK(cudaMemsetAsync(grad_flat, 0, sizeof(float) * h->grad_floats, st));
  PFM_CUDA_CHECK(cudaMemcpyAsync(tgrad(tp, tp.out), dout, sizeof(float) * (size_t)tp.B * tp.N * h->cfg.feats, cudaMemcpyDeviceToDevice, st));
  int rc = tape_backward(h, tp, grad_flat, st);
  tp.live = false;
  return rc;
}

int pfm_tf_loss_fwd_bwd(pfm_tf* h, const float* x1, const float* t, const float* t_code, const float* noise0, const float* noise1,
                        const float* mask, const float* cond, int loss_kind, float sigma, float* loss_out, float* grad_flat, int B, int N,
                        void* stream) {
  if (!h || !x1 || !t || !noise0 || !loss_out) { set_error("null argument"); return PFM_ERR_INVALID; }
  if (loss_kind != PFM_LOSS_FM_OT && loss_kind != PFM_LOSS_CFM && loss_kind != PFM_LOSS_DROID) { set_error("unknown loss kind %d", loss_kind); return PFM_ERR_INVALID; }
  if (loss_kind == PFM_LOSS_CFM && !noise1) { set_error("the CFM loss needs the second noise tensor"); return PFM_ERR_INVALID; }
  int rc = train_check(h, t_code, cond, mask, B, N);
  if (rc != PFM_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  PFM_CUDA_CHECK(cudaSetDevice(h->device));
  h->last_launches = 0;
  if ((rc = tape_ensure(h, B, N)) != PFM_OK) return rc;
  TfTape& tp = *h->tape;
  const int F = h->cfg.feats;
  const size_t total = (size_t)B * N * F;
  tt_interp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x1, t, noise0, noise1, mask, loss_kind, sigma, N, F, total,
                                                                   tdata(tp, tp.xs), tdata(tp, tp.u));
  PFM_CUDA_CHECK(cudaMemsetAsync(tp.acc, 0, sizeof(float) * 2, st));
  tt_sum_kernel<<<h->sm_count, 256, 0, st>>>(mask, (size_t)B * N, tp.acc + 1);
  h->last_launches += 2;
  if ((rc = train_forward(h, t_code, B, nullptr, mask, cond, B, N, st)) != PFM_OK) return rc;
  if (grad_flat) {
    PFM_CUDA_CHECK(cudaMemsetAsync(tp.grad, 0, sizeof(float) * tp.floats, st));
    PFM_CUDA_CHECK(cudaMemsetAsync(grad_flat, 0, sizeof(float) * h->grad_floats, st));
  }
  tt_seed_kernel<<<2 * h->sm_count, 256, 0, st>>>(tdata(tp, tp.out), tdata(tp, tp.u), total, tp.acc, grad_flat ? tgrad(tp, tp.out) : nullptr);
  tt_loss_out_kernel<<<1, 1, 0, st>>>(tp.acc, loss_out);
  h->last_launches += 2;
  if (grad_flat && (rc = tape_backward(h, tp, grad_flat, st)) != PFM_OK) return rc;
  tp.live = false;
  PFM_CUDA_CHECK(cudaGetLastError());
  return PFM_OK;
}

}  // extern "C"
