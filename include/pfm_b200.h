/*
 * pfm_b200.h -- C ABI of the B200-native hot path of ewencedr/particle_fm.
 *
 * One shared library (particle_fm_b200/lib/libpfm_b200.so, hand-written CUDA for sm_100a) that a
 * host binds through FFI (ctypes in this repo).  No torch / C++ types cross the
 * boundary: plain pointers, sizes and a CUDA stream handle.
 *
 * Each entry point names the reference interface it replaces (paths under the reference tree):
 *   pfm_epic_forward   <- CNF.forward                 particle_fm/models/flow_matching_module.py:191-204
 *                          EPiC_encoder.forward        particle_fm/models/components/epic.py:304-391
 *                          EPiC_layer.forward          particle_fm/models/components/epic.py:85-203
 *   pfm_epic_sample    <- CNF.decode (euler/midpoint)  flow_matching_module.py:245-287
 *                          + torchdyn fixed-step loop (third party; restated in oracle/ode_oracle.py)
 *                          + ode_wrapper.forward       flow_matching_module.py:62-71
 *   pfm_epic_loss_*    <- FlowMatchingLoss / ConditionalFlowMatchingLoss / DroidLoss .forward
 *                          particle_fm/models/components/losses.py:38-77, :101-136, :308-342
 *                          and the autograd backward of the network the reference gets from torch
 *
 * Conventions
 *   - every function returns 0 (PFM_OK) or a negative pfm_status; the message of the last failure
 *     on the calling thread is pfm_last_error().  No C++ exception crosses the ABI.
 *   - all data pointers are CALLER-OWNED DEVICE memory, contiguous fp32 unless stated, valid until
 *     the stream reaches the call.  The handle owns only its packed weights and workspaces.
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *   - a handle is bound to one device and is not thread-safe; one handle per rank.
 *   - there is no CPU fallback: without a CUDA device pfm_epic_create fails with PFM_ERR_CUDA.
 */
#ifndef PFM_B200_H
#define PFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFM_VERSION 100

typedef enum {
  PFM_OK = 0,
  PFM_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  PFM_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
  PFM_ERR_STATE = -3,       /* e.g. weights not set */
  PFM_ERR_UNSUPPORTED = -4  /* shape not supported by the selected precision path */
} pfm_status;

/* Arithmetic of the per-particle contractions.
 * FP32: CUDA-core fp32 FMA everywhere (strict mode; per-step parity 1e-3 and far below).
 * BF16: tcgen05 tensor-core GEMMs with bf16 operands, fp32 accumulation in TMEM, fp32 residual
 *       stream, fp32 pooling inputs rounded to bf16 (per-step parity 2e-2).  Requires hid == 128. */
typedef enum { PFM_PREC_FP32 = 0, PFM_PREC_BF16 = 1 } pfm_precision;

typedef enum { PFM_SOLVER_EULER = 0, PFM_SOLVER_MIDPOINT = 1 } pfm_solver;

/* loss kinds of losses.py: FM-OT :38-77, CFM :101-136, droid :308-342 */
typedef enum { PFM_LOSS_FM_OT = 0, PFM_LOSS_CFM = 1, PFM_LOSS_DROID = 2 } pfm_loss_kind;

/* Resolved constructor arguments of EPiC_encoder (epic.py:226-243). */
typedef struct {
  int32_t feats;           /* output features per particle                      (feats)        */
  int32_t input_dim;       /* per-particle input width                          (input_dim)    */
  int32_t hid;             /* hidden width of the per-particle MLPs             (hid_d)        */
  int32_t latent;          /* width of the per-jet global vector                (latent)       */
  int32_t layers;          /* number of EPiC layers                             (equiv_layers) */
  int32_t t_dim;           /* width of the time code = 2*frequencies                             */
  int32_t t_local_cat;     /* time code concatenated to every per-particle linear               */
  int32_t t_global_cat;    /* time code concatenated to every per-jet linear                    */
  int32_t global_cond_dim; /* conditioning width on per-jet linears (0 = none)                  */
  int32_t local_cond_dim;  /* conditioning width on per-particle linears (0 or global_cond_dim) */
  float sum_scale;         /* factor on the sum pooling (1e-2)                                   */
  float neg_slope;         /* leaky_relu slope (0.01, F.leaky_relu default)                      */
} pfm_epic_cfg;

typedef struct pfm_epic pfm_epic;

int pfm_version(void);
const char* pfm_last_error(void);

/* Create / destroy a network handle on CUDA device `device`. */
int pfm_epic_create(const pfm_epic_cfg* cfg, int device, pfm_epic** out);
void pfm_epic_destroy(pfm_epic* h);

/* Number of linears (4 + 4*layers + 1) and, for linear i in state_dict order
 * (fc_l1, fc_l2, fc_g1, fc_g2, nn_list.{l}.fc_global1, .fc_global2, .fc_local1, .fc_local2, fc_l3),
 * its (out, in) shape as the reference constructs it (epic.py:66-81, :262-300). */
int pfm_epic_num_linears(const pfm_epic* h);
int pfm_epic_linear_shape(const pfm_epic* h, int i, int32_t* out_features, int32_t* in_features);

/* Hand over the FOLDED weights (W = g*v/||v||, row-major [out,in]) and biases of all linears, in
 * the order above, as arrays of n device pointers (the pointer arrays themselves are host memory).
 * The library repacks into its own layouts and keeps the packed copy; call again whenever the
 * parameters change (optimizer step, EMA swap). */
int pfm_epic_set_weights(pfm_epic* h, const float* const* weights, const float* const* biases, int n,
                         void* stream);
/* The same with the weight-norm fold done by the library: v[i] = weight_v [out,in], g[i] = weight_g [out] (NULL for a plain
 * linear: v[i] is then the weight itself), b[i] = bias; one launch for all linears (replaces the forward pre-hook of
 * torch's nn.utils.weight_norm, epic.py:66-81).  pfm_epic_param_grads maps the flat folded-weight gradient of the training
 * calls onto d(weight_v), d(weight_g), d(bias) (what autograd does through torch._weight_norm in the reference), scaled by
 * *scale (one device float, e.g. the upstream gradient of the loss; NULL = 1). */
int pfm_epic_set_params(pfm_epic* h, const float* const* v, const float* const* g, const float* const* b, int n,
                        void* stream);
int pfm_epic_param_grads(pfm_epic* h, const float* grad_flat, const float* scale, const float* const* v,
                         const float* const* g, float* const* dv, float* const* dg, float* const* db, int n, void* stream);

int pfm_epic_set_precision(pfm_epic* h, int precision /* pfm_precision */);

/* One evaluation of the vector field (CNF.forward after the time embedding).
 *   t_code  [t_rows, t_dim]   time code; t_rows == 1 (one time for the whole batch, sampling) or
 *                             t_rows == B (one time per jet, training).  May be NULL iff t_dim == 0.
 *   x       [B, N, input_dim] per-particle input (already holds the time code if add_time_to_input)
 *   mask    [B, N]            non-zero = real particle; NULL = all real
 *   cond    [B, cond_dim]     NULL iff global_cond_dim == 0 and local_cond_dim == 0
 *   out     [B, N, feats]     = leaky_relu(fc_l3(...)) * mask; a jet without real particles is NaN
 *                             in every entry, as in the reference (0/0 mean, epic.py:161,370)   */
int pfm_epic_forward(pfm_epic* h, const float* t_code, int t_rows, const float* x, const float* mask,
                     const float* cond, float* out, int B, int N, void* stream);

/* Fixed-step integration of dx/dt = v(t, x) from t = 1 to t = 0 with the state resident on chip:
 * ONE launch for all steps.
 *   x_inout    [B, N, feats]     in: initial noise (already multiplied by mask); out: end point
 *   t_codes    [n_evals, t_dim]  time code of every network evaluation, in evaluation order
 *                                (n_evals = n_steps for Euler, 2*n_steps for midpoint)
 *   t_codes_in [n_evals, t_in]   extra per-particle input columns prepended to x at every
 *                                evaluation (add_time_to_input: t_in = input_dim - feats); else NULL
 *   dt         [n_steps]         host-computed fp32 step sizes of the reversed-time grid
 * Euler:    x <- x + dt*(-v(t_k, x));   midpoint: x <- x + dt*(-v(t_k+dt/2, x + 0.5*dt*(-v(t_k, x)))) */
int pfm_epic_sample(pfm_epic* h, float* x_inout, const float* mask, const float* cond,
                    const float* t_codes, const float* t_codes_in, const float* dt, int solver,
                    int n_steps, int B, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training (fp32 path).  Gradients are returned w.r.t. the FOLDED weights and the biases in one
 * flat buffer  [W_0 (out_0 x in_0, row-major) | b_0 | W_1 | b_1 | ...]  of pfm_epic_grad_size()
 * floats, linears in the order of pfm_epic_set_weights; the host maps dW onto weight_g / weight_v
 * (W = g*v/||v||).  A flat buffer is what a data-parallel all-reduce wants.
 * --------------------------------------------------------------------------------------------- */
long long pfm_epic_grad_size(const pfm_epic* h);

/* Arithmetic of the training kernels (results are fp32-accurate in both modes: loss 1e-5, gradients 1e-4 against autograd).
 * PFM_TRAIN_AUTO (default): with hid == 128 the 128 x 128 per-particle linears and their transposes run on tcgen05 tensor
 * cores with the 3-term bf16 split (x = hi + lo; the dropped lo*lo term is 2^-16 relative), fp32 accumulation in TMEM,
 * as a program of kernels over the packed real particles of the batch (csrc/epic_train_tc.cu); other widths use the
 * fused fp32 CUDA-core kernels.  PFM_TRAIN_CUDA_CORES forces the latter (A/B runs, strict comparisons). */
typedef enum { PFM_TRAIN_AUTO = 0, PFM_TRAIN_CUDA_CORES = 1 } pfm_train_mode;
int pfm_epic_set_train_mode(pfm_epic* h, int mode /* pfm_train_mode */);
/* Test hook: synchronise the device and copy n floats of a training array of the last forward / backward to HOST memory.
 * which: 0 saved post-activations act[stage][row][hid_p], 1 pre-activation gradients dact (same shape), 2 per-jet
 * effective-bias gradients [B][bias row], 3 saved per-jet vectors, 4 head gradient seed [row][feats], 5 network input
 * [row][Kx]; rows = packed real particles in batch order.  Used by the parity tests to tell apart arithmetic error from
 * a leaky_relu kink taken on different sides by two fp32 evaluation orders. */
int pfm_epic_debug_copy(pfm_epic* h, int which, float* host, long long n);

/* Data-parallel overlap: the weight gradients of the last backward are produced in pfm_epic_grad_chunks() consecutive
 * slices of the flat buffer (offset / count in floats).  pfm_epic_stream_wait_grad_chunk makes `stream` wait until slice
 * i is complete, so the host can enqueue the all-reduce of slice i on a side stream while later slices are still being
 * computed (the reference gets this overlap from DDP's bucketed reducer, configs/trainer/ddp.yaml:4-9). */
int pfm_epic_grad_chunks(const pfm_epic* h);
int pfm_epic_grad_chunk_range(const pfm_epic* h, int i, long long* offset, long long* count);
int pfm_epic_stream_wait_grad_chunk(pfm_epic* h, int i, void* stream);

/* Fused flow-matching training step:  losses.py:38-77 (FM-OT), :101-136 (CFM), :308-342 (droid)
 * + the autograd backward of the network.  The random draws are the caller's (the reference draws
 * t on the CPU generator and the noise on the device):
 *   x1      [B, N, feats]   data            t       [B]          per-jet time
 *   t_code  [B, t_dim]      time code of t  t_code_in [B, input_dim-feats]  (add_time_to_input) or NULL
 *   noise0  [B, N, feats]   z / x0          noise1  [B, N, feats] (CFM's extra epsilon) or NULL
 *   mask    [B, N] or NULL  cond [B, cond_dim] or NULL
 * Computes  y, u_t  by the loss kind's interpolation, v = net(t, y),
 *   loss_out[0] = sum((v - u_t)^2) / sum(mask)            (device scalar)
 * and, if grad_flat != NULL, d loss / d(folded weights, biases) into grad_flat (overwritten). */
int pfm_epic_loss_fwd_bwd(pfm_epic* h, const float* x1, const float* t, const float* t_code,
                          const float* t_code_in, const float* noise0, const float* noise1,
                          const float* mask, const float* cond, int loss_kind, float sigma,
                          float* loss_out, float* grad_flat, int B, int N, void* stream);

/* Generic differentiable evaluation (autograd of CNF.forward / EPiC_encoder.forward for any loss):
 * pfm_epic_forward_train = pfm_epic_forward that also keeps the activations in the handle;
 * pfm_epic_backward consumes them once: given grad_out = dL/d out [B,N,feats] it writes
 * dL/dx into grad_x [B,N,input_dim] (optional) and dL/d(weights) into grad_flat (optional). */
int pfm_epic_forward_train(pfm_epic* h, const float* t_code, int t_rows, const float* x,
                           const float* mask, const float* cond, float* out, int B, int N, void* stream);
int pfm_epic_backward(pfm_epic* h, const float* t_code, int t_rows, const float* cond,
                      const float* grad_out, float* grad_x, float* grad_flat, int B, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PC-Droid-style set transformers (fp32 path):
 *   pfm_tf_forward  <- FullTransformerEncoder.forward     components/droid_transformer.py:529-548
 *                      FullCrossAttentionEncoder.forward   components/droid_transformer.py:696-711
 *                      behind CNF.forward                  flow_matching_module.py:148-161, :191-204
 *   pfm_tf_sample   <- CNF.decode (euler / midpoint) with those networks, state resident on the device
 * Padding is skipped (only keys are masked in the reference, every other op is per token); the padded
 * slots of the result are 0 (the reference leaves unmasked values there that its callers multiply away).
 * --------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t kind;              /* 0 = FullTransformerEncoder, 1 = FullCrossAttentionEncoder              */
  int32_t feats;             /* particle features = outp_dim                                            */
  int32_t t_dim;             /* width of the time code = 2*frequencies                                  */
  int32_t cond_dim;          /* global_cond_dim (context = [time code, cond])                           */
  int32_t add_time_to_input; /* time code also concatenated to every particle (True in the YAMLs)       */
  int32_t model_dim, num_layers, num_heads;
  int32_t ctxt_out;          /* ctxt_embd_config.outp_dim (64)                                          */
  int32_t embd_hddn;         /* hidden width of the node / ctxt / outp embedders (2*model_dim)          */
  int32_t dense_hddn;        /* hidden width of the per-layer feed-forward nets                         */
  int32_t num_tokens;        /* learned global tokens of the cross-attention encoder (4)                */
  float neg_slope;           /* nn.LeakyReLU(0.1), get_act("lrlu")                                      */
  float ln_eps;              /* nn.LayerNorm eps (1e-5)                                                 */
} pfm_tf_cfg;

typedef struct pfm_tf pfm_tf;

int pfm_tf_create(const pfm_tf_cfg* cfg, int device, pfm_tf** out);
void pfm_tf_destroy(pfm_tf* h);
/* Parameter tensors in the canonical order = the reference's state_dict order of CNF.net restricted to
 * parameters (ctxt_emdb, te.* / cae.*, node_embd, outp_embd); matrices are row-major [rows, cols],
 * vectors have cols == 1.  pfm_tf_set_weights takes n device pointers in that order. */
int pfm_tf_num_params(const pfm_tf* h);
int pfm_tf_param_shape(const pfm_tf* h, int i, int32_t* rows, int32_t* cols);
int pfm_tf_set_weights(pfm_tf* h, const float* const* params, int n, void* stream);
/* x [B,N,feats] (WITHOUT the time columns: the library hoists them), t_code [1|B, t_dim], mask [B,N] or
 * NULL, cond [B,cond_dim] or NULL, out [B,N,feats]. */
int pfm_tf_forward(pfm_tf* h, const float* t_code, int t_rows, const float* x, const float* mask,
                   const float* cond, float* out, int B, int N, void* stream);
int pfm_tf_sample(pfm_tf* h, float* x_inout, const float* mask, const float* cond, const float* t_codes,
                  const float* dt, int solver, int n_steps, int B, int N, void* stream);
/* PFM_PREC_BF16: the large linears (K multiple of 64, N multiple of 128) run on tcgen05 tensor cores with bf16
 * operands / fp32 accumulation; LayerNorm statistics, attention, biases and residuals stay fp32. */
int pfm_tf_set_precision(pfm_tf* h, int precision /* pfm_precision */);
int pfm_tf_last_launches(const pfm_tf* h);

/* ---- training of the droid transformers (replaces torch autograd over droid_transformer.py:211-284, 331-344,
 * 386-397, 529-548, 696-711).  The reference neither masks the networks' output nor the loss terms of padded slots
 * (losses.py:74-76, 130, 339-341 sum over every slot), so training runs on all B*N rows with only the attention KEYS
 * restricted to real particles; mask is required.  fp32 whatever pfm_tf_set_precision says.
 *   pfm_tf_grad_size      floats of the flat gradient: the parameters of pfm_tf_param_shape, in that order, each row-major
 *   pfm_tf_forward_train  pfm_tf_forward on the dense rows (padded slots get the reference's values, not 0) + saved tape
 *   pfm_tf_backward       d(out) [B,N,feats] -> grad_flat, consumes the tape of the last pfm_tf_forward_train
 *   pfm_tf_loss_fwd_bwd   FlowMatchingLoss / ConditionalFlowMatchingLoss / DroidLoss forward (losses.py:38-77, 101-136,
 *                         308-342) with the given draws (t [B], t_code [B,t_dim], noise0/noise1 [B,N,feats]) and, when
 *                         grad_flat != NULL, its backward; loss_out: one device float. */
int64_t pfm_tf_grad_size(const pfm_tf* h);
int pfm_tf_forward_train(pfm_tf* h, const float* t_code, int t_rows, const float* x, const float* mask,
                         const float* cond, float* out, int B, int N, void* stream);
int pfm_tf_backward(pfm_tf* h, const float* dout, float* grad_flat, void* stream);
int pfm_tf_loss_fwd_bwd(pfm_tf* h, const float* x1, const float* t, const float* t_code, const float* noise0,
                        const float* noise1, const float* mask, const float* cond, int loss_kind, float sigma,
                        float* loss_out, float* grad_flat, int B, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused optimizer step over FLAT fp32 buffers of n elements (all device memory): global-norm gradient clipping
 * (torch.nn.utils.clip_grad_norm_, Lightning gradient_clip_val), AdamW (torch.optim.AdamW, configs/model/flow_matching.yaml:3-7)
 * and, if ema != NULL, the EMA callback's update ema -= (ema - w) * (1 - ema_decay) (callbacks/ema.py:73-81), in two launches.
 *   step       1-based step count (bias corrections 1 - beta^step); step <= 0: the count is kept ON THE DEVICE in
 *              workspace[2] (as an int, start it at 0) and incremented by the call -- for replay from a CUDA graph
 *   max_norm   <= 0: no clipping
 *   workspace  3 device words: [0] scratch (sum of squares), [1] receives the total gradient norm before clipping,
 *              [2] the device-side step count */
int pfm_clip_adamw(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                   float beta2, float eps, float weight_decay, float max_norm, int step, float* ema, float ema_decay,
                   float* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Diffusion family on the same network (loss_type="diffusion"): the samplers of
 *   components/solver.py:23-143 (ddim_sampler, euler_maruyama_sampler) and the probability-flow ODE of
 *   flow_matching_module.py:62-69 (ode_wrapper, loss_type == "diffusion") behind CNF.decode :279-287, :304-327,
 * as per-evaluation update rules of the SAME single-launch integrator (state resident on chip).
 *   coef [n_evals, 4]  host-computed fp32 schedule values of every network evaluation (device memory):
 *     PFM_STEP_PF_ODE : {beta(t), noise_rate(t), -, -}   drift f = -0.5*beta*(x - v/noise_rate), stepped with `solver`
 *                       (Euler / midpoint in reversed time, dt as in pfm_epic_sample)
 *     PFM_STEP_DDIM   : {signal_rate, noise_rate, next_signal_rate, next_noise_rate}
 *                       pred = (x - nr*v)/sr;  x <- nsr*pred + nnr*v;  the result is pred of the last step
 *     PFM_STEP_EM     : {beta, noise_rate, delta_t, sqrt(beta*delta_t)}
 *                       x += 0.5*beta*(x + 2*(-v/nr))*delta_t;  x += sqrt(beta*delta_t) * noise[step]
 *   noise [n_steps, B, N, feats]  (PFM_STEP_EM only) the caller's per-step normal draws (the reference draws
 *                       torch.randn_like(x_t) on the device at every step, solver.py:131)
 * fp32 path only (PFM_PREC_FP32). */
typedef enum { PFM_STEP_PF_ODE = 1, PFM_STEP_DDIM = 2, PFM_STEP_EM = 3 } pfm_step_kind;
int pfm_epic_sample_diffusion(pfm_epic* h, float* x_inout, const float* mask, const float* cond, const float* t_codes,
                              const float* t_codes_in, const float* coef, const float* noise, const float* dt, int step_kind,
                              int solver, int n_steps, int B, int N, void* stream);

/* ---------------------------------------------------------------------------------------------
 * generate_data post-processing on the device  <- particle_fm/utils/data_generation.py:105-123
 *   out[b,n,f] = post(x[b,n,f]) (* mask[b,n] if mask != NULL), with
 *   post(v) = v * scale[f] + shift[f]        (inverse_normalize_tensor, data/components/utils.py:183-200;
 *                                             scale = std/sigma, shift = mean; two roundings like the eager ops)
 *             then 1 - exp(.) on column log_col (log_pt, :116-117; log_col < 0: none);  scale == shift == NULL: identity.
 *   first_only_col >= 0: on that column the affine part is applied to particle 0 of every jet only -- what the
 *             reference's pt_standardization branch does (:111-113 passes the 2-D slice batch[..., 2] to
 *             inverse_normalize_tensor, whose ``tensor[..., 0]`` then indexes particles); -1 otherwise.
 * x, mask: device memory.  scale / shift: HOST arrays of F floats (passed by value to the kernel).  out: device memory
 * OR pinned (page-locked) host memory -- the kernel writes the final values straight into the caller's host buffer,
 * there is no separate device -> host copy.  F <= PFM_POST_MAX_FEATS. */
#define PFM_POST_MAX_FEATS 16
int pfm_postprocess(const float* x, const float* mask, float* out, long long B, int N, int F, const float* scale,
                    const float* shift, int log_col, int first_only_col, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CFM-OT mini-batch coupling  <- ConditionalFlowMatchingOTLoss.forward, losses.py:165-189
 * pfm_ot_assign: for every jet k the exact optimal assignment between the N noise points x0[k] and the N data points
 *   x1[k] under the squared Euclidean cost (what POT's ot.emd returns for uniform marginals, as a permutation):
 *   sigma[k, i] = j.  cost[k] (optional, double) = sum_i |x0[k,i] - x1[k,sigma(i)]|^2.  One warp per jet, N <= 320.
 * pfm_ot_gather: the resampled pairs of losses.py:183-189 for host-drawn row picks pick[k, m] in [0, N):
 *   x0p[k,m] = x0[k,pick], x1p[k,m] = x1[k,sigma[pick]], mask_ot[k,m] = mask[k,sigma[pick]] (mask NULL = ones). */
int pfm_ot_assign(const float* x0, const float* x1, int B, int N, int F, int32_t* sigma, double* cost, void* stream);
int pfm_ot_gather(const float* x0, const float* x1, const float* mask, const int32_t* sigma, const int32_t* pick, int B, int N,
                  int F, float* x0p, float* x1p, float* mask_ot, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Jet-feature flow  <- CNF of particle_fm/models/flow_matching_no_sets.py:41-93 around
 *                      small_cond_MLP_model, components/mlp.py:24-68 (blocks of nn.Linear + activation; every block's
 *                      input is torch.cat([t, x, cond]))
 * A chain of n_linears dense layers on rows (one row = one event).  For linear i: out_widths[i] outputs,
 * concat[i] != 0: its input is [time code | previous output | cond] (K = t_dim + prev + cond_dim), else the previous
 * output; act[i] != 0: the activation follows it.  The last linear maps back to `features`.
 * pfm_mlp_sample integrates dx/dt = v(t, x, cond) from t = 1 to 0 exactly like pfm_epic_sample (one launch, the state
 * of a row tile resident in shared memory, weights streamed from L2); pfm_mlp_forward is one evaluation. */
typedef enum { PFM_ACT_ELU = 0, PFM_ACT_TANH = 1, PFM_ACT_RELU = 2, PFM_ACT_LEAKY_RELU = 3, PFM_ACT_SILU = 4 } pfm_act;
typedef struct {
  int32_t features;   /* row width in and out                         */
  int32_t t_dim;      /* width of the time code (2 * freqs)           */
  int32_t cond_dim;   /* width of cond (1: m_jj)                      */
  int32_t n_linears;
  int32_t act;        /* pfm_act                                      */
} pfm_mlp_cfg;
typedef struct pfm_mlp pfm_mlp;
int pfm_mlp_create(const pfm_mlp_cfg* cfg, const int32_t* out_widths, const int32_t* concat, const int32_t* act, int device,
                   pfm_mlp** out);
void pfm_mlp_destroy(pfm_mlp* h);
int pfm_mlp_linear_shape(const pfm_mlp* h, int i, int32_t* out_features, int32_t* in_features);
/* weights[i]: row-major [out, in] fp32 (nn.Linear.weight), biases[i]: [out]; device pointers, host pointer arrays */
int pfm_mlp_set_weights(pfm_mlp* h, const float* const* weights, const float* const* biases, int n, void* stream);
/* t_code [1|B, t_dim], x [B, features], cond [B, cond_dim] -> out [B, features] */
int pfm_mlp_forward(pfm_mlp* h, const float* t_code, int t_rows, const float* x, const float* cond, float* out, int B,
                    void* stream);
/* x_inout [B, features]; t_codes [n_evals, t_dim]; dt [n_steps] (see pfm_epic_sample) */
int pfm_mlp_sample(pfm_mlp* h, float* x_inout, const float* cond, const float* t_codes, const float* dt, int solver,
                   int n_steps, int B, void* stream);

/* Introspection for tests / bench: kernels launched by the last call on this handle and the
 * number of CTA work groups the last plan produced. */
int pfm_epic_last_launches(const pfm_epic* h);
int pfm_epic_last_groups(const pfm_epic* h);

/* Optional device-side timing of the dominant (fused network/integrator) kernel: when enabled, CUDA
 * events are recorded on the launching stream around that kernel only.  pfm_epic_last_kernel_ms
 * synchronises on the stop event and returns the summed duration of the last call (ms), < 0 if none. */
int pfm_epic_set_timing(pfm_epic* h, int enable);
float pfm_epic_last_kernel_ms(pfm_epic* h);

#ifdef __cplusplus
}
#endif
#endif /* PFM_B200_H */
