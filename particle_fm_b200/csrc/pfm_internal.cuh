// Internal declarations shared by the translation units of libpfm_b200.so (not part of the ABI).
#pragma once
#include <cstring>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/pfm_b200.h"

namespace pfm {

// One weight-normed linear of the encoder after folding, split by what its input columns multiply
// (concat orders: epic.py:361,365,376,379 stem; :180-196 layers; :388 head).
//   [ time | main | (global, fc_local1 only) | cond ]
struct Lin {
  int out, in;
  int t_off, t_len;   // time-code columns          -> hoisted into the per-evaluation bias table
  int m_off, m_len;   // per-particle (local linears) or per-jet vector (global linears) columns
  int g_off, g_len;   // fc_local1: columns multiplying the broadcast global vector
  int c_off, c_len;   // conditioning columns       -> hoisted into the per-jet bias table
  int bias_off;       // offset of this linear inside one bias-table row
  int ldo;            // leading dimension of the k-major (transposed) fp32 copy, round_up(out, 4)
  const float* Wt;    // k-major fp32 copy  Wt[k*ldo + o] = W[o][k],  k in [0, in)  (+ slack rows)
  const float* b;     // [out]
  const float* Wr;    // row-major copy of the main block  Wr[o*ldr + k] = W[o][m_off + k]  (backward: dX = dY . W)
  int ldr;            // round_up(m_len, 4)
};

enum LinIdx { LIN_L1 = 0, LIN_L2 = 1, LIN_G1 = 2, LIN_G2 = 3, LIN_LAYER0 = 4 };
// layer l: LIN_LAYER0 + 4*l + {0: fc_global1, 1: fc_global2, 2: fc_local1, 3: fc_local2}; last: fc_l3

struct Plan {          // device buffers describing how jets are packed into CTA work groups
  int* n_real;         // [B]     real particles per jet
  uint16_t* ridx;      // [B*N]   ridx[b*N + r] = particle index of the r-th real particle of jet b
  int2* groups;        // [B]     (first jet, number of jets) of every group
  int2* groups_tmp;    // [B]     per-warp group lists of the inference plan before compaction
  int* n_groups;       // [1]
  int* counter;        // [1]     dynamic work counter for persistent CTAs
  int* jetmap;         // [B]     inference plan: jet handled at position i (groups are ranges of POSITIONS; jets are
                       //         bin-packed by multiplicity, not taken in batch order)
  int* order;          // [B]     scratch: jets sorted by multiplicity
  int* rowoff;         // [B]     first packed row of every jet (exclusive prefix sum of n_real)
  int* n_total;        // [1]     total number of real particles (= sum(mask), the loss denominator)
  int capB, capBN;
};

}  // namespace pfm

struct pfm_epic {
  pfm_epic_cfg cfg;
  int device;
  int n_lin;
  int precision;
  bool weights_set;
  int sm_count;
  int max_smem_optin;
  std::vector<pfm::Lin> lin_host;   // host copy (device pointers inside)
  pfm::Lin* lin_dev;                // device copy of the descriptors
  float* wt_store;                  // all k-major fp32 copies
  float* b_store;                   // all biases
  float* wr_store;                  // row-major main blocks (training backward)
  size_t wt_floats, b_floats, wr_floats;
  int bstride;                      // floats per bias-table row
  // bf16 tensor-core path: pre-swizzled shared-memory images of the H x H per-particle weights
  void* tc_store;
  bool tc_dirty;            // the fp32 weights changed after the bf16 images were packed: tc_run repacks on its own stream
  size_t tc_bytes;
  // workspaces (grown on demand)
  float* tbias; size_t tbias_cap;   // [rows, bstride]  b + W_t . time_code
  float* cbias; size_t cbias_cap;   // [B, bstride]     W_c . cond
  pfm::Plan plan;
  int last_launches, last_groups_host;
  // training workspaces (grown on demand; see epic_train.cu for the layouts)
  float* act; size_t act_cap;       // saved per-particle activations      [(2+2L) stages, rows, Hp]
  float* dact; size_t dact_cap;     // per-particle pre-activation grads   [(2+2L) stages, rows, Hp]
  float* yact; size_t yact_cap;     // per-particle network input          [rows, Kx]
  float* jact; size_t jact_cap;     // saved per-jet vectors               [B, L+1 units, LDP + Hp + Zp]
  float* dpre3; size_t dpre3_cap;   // gradient at the head pre-activation [rows, F]
  float* dbeff; size_t dbeff_cap;   // gradient of the per-jet effective biases [B, bstride]
  float* dxs; size_t dxs_cap;       // gradient w.r.t. the per-particle input   [rows, Kx]
  std::vector<cudaEvent_t> grad_ev;      // recorded after every chunk of weight-gradient jobs of the last backward
  std::vector<long long> grad_chunk_off; // [chunks + 1] offsets of the chunks in the flat gradient
  int grad_chunks;
  float* hs_spill; size_t hs_spill_cap;   // spill slabs of the fp32 kernels (hidden features of jets too large for smem)
  float* dh_spill; size_t dh_spill_cap;
  float* loss_acc;                  // [1] sum of squared errors
  float* ones;                      // [1] = 1.0f (the "input" of a bias in the weight-gradient jobs)
  void* jobs_dev; size_t jobs_cap;  // device copy of the weight-gradient job table
  int train_B, train_N, train_Kx, train_xin_off;   // shape of the saved forward (0 = none)
  // weight-norm fold / backward inside the library (pfm_epic_set_params / pfm_epic_param_grads)
  int2* wn_rows;                    // [wn_total_rows] (linear, output row) of every block of the fold kernels
  int wn_total_rows;
  long long* wn_goff;               // [2 n_lin] offsets of W_i and b_i in the flat gradient
  const float** wn_ptrs;            // [8 n_lin] device copy of the caller's pointer tables
  // tensor-core training path (epic_train_tc.cu): hi / lo bf16 images of the 128 x 128 per-particle weights and their
  // transposes, workspace (effective biases, per-jet broadcast / carry vectors, dh, row -> jet map)
  void* tt_store; size_t tt_bytes; bool tt_dirty;
  float* tt_ws; size_t tt_ws_cap;
  bool train_tc;                    // the saved forward was produced by the tensor-core path
  int train_mode;                   // pfm_train_mode requested by the host
  // diffusion step program of the next sampling call (pfm_epic_sample_diffusion; 0 = plain flow-matching ODE)
  int step_kind; const float* step_coef; const float* step_noise;
  // pinned host staging of the small tables uploaded per call (pointer tables of the weight-norm kernels, weight-gradient job
  // table): async copies from PINNED memory are legal inside CUDA-graph capture, pageable ones are not; an unchanged table is
  // not uploaded again
  void* pin_stage; size_t pin_cap; size_t pin_used[3]; cudaEvent_t pin_ev[3]; bool pin_ev_set[3];
  bool timing;
  std::vector<cudaEvent_t> ev_pool;   // start/stop pairs of the main kernel, one pair per chunk of a call
  int ev_used;
};

namespace pfm {

void set_error(const char* fmt, ...);
#define PFM_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      pfm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return PFM_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

struct RunArgs {
  const float* x_in;      // forward: [B,N,input_dim]; sample: [B,N,feats] (in/out)
  float* x_out;           // forward: [B,N,feats];     sample: same buffer as x_in
  int B, N;
  int Kx;                 // per-particle input columns read from x_in (input_dim or feats)
  int xin_off;            // first row of fc_l1's main block used for those columns
  int n_evals;            // network evaluations (1 for forward)
  int solver;             // -1: forward only; else pfm_solver
  int n_steps;
  const float* dt;        // [n_steps] device
  int tbias_per_jet;      // 0: bias-table row = evaluation index; 1: row = jet index
  bool has_cbias;
  const int* jetmap;      // position -> jet (nullptr: identity)
  int step_kind;          // 0: dx/dt = v (flow matching); else pfm_step_kind (diffusion samplers, fp32 path only)
  const float* coef;      // [n_evals, 4] schedule values of the step program
  const float* noise;     // PFM_STEP_EM: [n_steps, B_total, N, feats]
  long long noise_step_stride;   // B_total * N * feats
  int jet0;               // index of this launch's first jet inside the whole request (noise addressing)
};

// fp32 CUDA-core path (epic_simt.cu)
int simt_plan_caps(const pfm_epic* h, int N, int* R_cap, int* J_cap);
int simt_run(pfm_epic* h, const RunArgs& a, cudaStream_t st);
// training, fp32 CUDA cores (forward with saved activations: epic_simt.cu; backward: epic_train.cu)
struct TrainLayout {       // strides of the saved-activation arrays, shared by forward and backward
  size_t stage_stride;     // floats between two stages of act / dact  (= rows_cap * Hp)
  int Hp;                  // row stride of act / dact
  int LDP, Zp;             // per-jet unit = [pool input (LDP) | g1 (Hp) | g (Zp)]
  int junit, jstride;      // floats per unit / per jet
  int R_cap, J_cap;        // group capacity both kernels are planned for
};
int train_layout(const pfm_epic* h, int B, int N, TrainLayout* lay);
struct TrainFwdArgs {
  const float* x_in;       // loss_kind < 0: network input [B,N,Kx];  else the data x1 [B,N,F]
  float* x_out;            // loss_kind < 0: network output [B,N,F]
  const float* t; const float* noise0; const float* noise1; int loss_kind; float sigma;
  int B, N, Kx, xin_off; bool has_cbias; int tbias_per_jet;
  TrainLayout lay;
};
int simt_train_forward(pfm_epic* h, const TrainFwdArgs& a, cudaStream_t st);
struct TrainBwdArgs {
  int B, N, Kx, xin_off;
  const float* t_code; int t_ld;        // [B or 1, t_dim]; t_ld = 0 when one row serves every jet
  const float* t_code_in; int t_in;     // hoisted add_time_to_input columns (loss path) or NULL
  const float* cond; int cond_dim;
  float* grad_flat;                     // [sum_i out_i*in_i + out_i]  (W_0 | b_0 | W_1 | b_1 ...)
  bool want_dx;
  TrainLayout lay;
};
int train_backward(pfm_epic* h, const TrainBwdArgs& a, cudaStream_t st);
// weight-gradient product dW[o*ldw + col0 + c] += sum_r Y[r*ldy + o] * X[r*ldx + c]  (epic_train.cu::xty_kernel)
struct XtyJob {
  const float* Y; const float* X; float* dW;
  int ldy, ldx, ldw, out, K, col0;
  int rows;            // >= 0: fixed row count;  < 0: *n_total (particle rows)
  int tile0;           // first tile index of this job in grid.y
  int tiles_o, tiles_k;
};
bool xty_use_simt();
int xty_tc_launch(const XtyJob* jobs_dev, int n_jobs, int tiles, const int* n_total, int max_rows, int sm_count, cudaStream_t st);
int xty_tc_launch_one(const float* Y, int ldy, const float* X, int ldx, float* dW, int ldw, int out, int K, int col0, int rows,
                      int sm_count, cudaStream_t st);
int xty_launch_one(const float* Y, int ldy, const float* X, int ldx, float* dW, int ldw, int out, int K, int col0, int rows,
                   cudaStream_t st);
int simt_caps_for_train(const pfm_epic* h, int N, int* R_cap, int* J_cap, int* TC, int* RB, int* KC);
// training with the per-particle GEMMs on tcgen05, fp32-accurate 3-term bf16 split (epic_train_tc.cu); hid == 128
bool tt_enabled(const pfm_epic* h);
int tt_plan(pfm_epic* h, int B, int N, cudaStream_t st);
int train_plan_groups(pfm_epic* h, int B, const TrainLayout& lay, cudaStream_t st);
// upload `bytes` of a small host table to `dst` through slot `slot` (0..2) of the handle's pinned staging buffer; skipped when the
// slot already holds the same bytes for the same destination
int upload_table(pfm_epic* h, int slot, void* dst, const void* src, size_t bytes, cudaStream_t st);
int tt_train_forward(pfm_epic* h, const TrainFwdArgs& a, cudaStream_t st);
int tt_train_backward(pfm_epic* h, const TrainBwdArgs& a, cudaStream_t st);
// bf16 tcgen05 path (epic_tc.cu)
int tc_supported(const pfm_epic* h, int N);
int tc_plan_caps(const pfm_epic* h, int N, int* R_cap, int* J_cap);
int tc_pack_weights(pfm_epic* h, cudaStream_t st);
int tc_run(pfm_epic* h, const RunArgs& a, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): a kernel launched through launch_pdl may be scheduled while the previous kernel
// of the stream is still draining; it must call pdl_wait() before it touches global memory.  In a kernel launched the ordinary way pdl_wait() is a no-op.  Works under stream capture.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Early trigger: only for kernels whose whole grid is resident at once (one wave).  A multi-wave kernel that triggers early
// lets its successor's CTAs take the SM slots its own remaining CTAs need.  Measured on the training step (1024 jets): only
// the GEMM passes launched this way 1.69 ms; the per-jet kernels and xty_tc as well 1.79-1.80 ms (with or without their own
// early trigger), so only rowlin2_tc_kernel uses it.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

}  // namespace pfm
