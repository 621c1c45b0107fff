/* kvae_kalman.h — C ABI of the B200-native Kalman filter / RTS smoother / ELBO / adjoint.
 *
 * This is the drop-in boundary for the hot path of rodrigo-paganini/kalman-vae.  The reference has
 * no FFI of its own (it is pure PyTorch); every entry point below names the reference interface
 * whose arithmetic it replaces (paths relative to the reference checkout):
 *
 *   kvae_kf_filter_smooth_fwd  <- KalmanFilter.filter   kvae/kalman/kalman_filter.py:107-201
 *                                 KalmanFilter.smooth   kvae/kalman/kalman_filter.py:240-279
 *                                 (filter_step :31-104, smooth_step :204-237) including the mixing
 *                                 DynamicsParameter.compute_step      kvae/kalman/dyn_param.py:58-60
 *                                 SwitchingDynamicsParameter.compute_batch
 *                                                                     kvae/kalman/switch_dyn_param.py:82-86
 *   kvae_kf_elbo_fwd           <- KalmanFilter.elbo     kvae/kalman/kalman_filter.py:305-401
 *                                 (_safe_cholesky :282-302, first attempt: jitter*I added)
 *   kvae_kf_bwd                <- loss.backward() through all of the above
 *                                 (kvae/train/train.py:53; the reference relies on autograd)
 *
 * Conventions
 *   - all tensors are contiguous fp32 DEVICE buffers in the reference's batch-major layouts
 *     ([B,T,...], kalman_filter.py:193-201); base pointers must be 16-byte aligned.
 *     Means are [B,T,n] (the reference's trailing singleton dimension does not change the layout).
 *   - nothing is allocated, no host synchronisation happens; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream) of device `device`
 *     (-1 = current device).  Every entry point is re-entrant (autograd calls backward from
 *     another thread than forward).
 *   - return value: 0 ok; <0 invalid argument / unsupported shape (see kvae_last_error());
 *     >0 a cudaError_t raised by the launch.
 *   - `info` is a device int32 status word the kernels OR bits into; the caller zeroes it and decides when to
 *     read it:  KVAE_INFO_PIVOT   a pivot of the filter / smoother / R / Sigma0 factorisations was not positive
 *                                 (the reference raises torch.linalg.LinAlgError there),
 *               KVAE_INFO_CHOL_S  chol(sym(Sigma_smooth) + jitter I) failed for some (b,t)      } _safe_cholesky
 *               KVAE_INFO_CHOL_Q  chol(sym(Q_t) + jitter_q I) failed for some (b,t)              } would retry, 10x
 *               KVAE_INFO_PEER    the data-parallel exchange gave up waiting for a peer (gradients not written).
 */
#ifndef KVAE_KALMAN_H
#define KVAE_KALMAN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVAE_ABI_VERSION 6
#define KVAE_FLAG_SMOOTH_ONLY 1  /* kvae_dims.flags: forward entry skips the filter sweep (states given) */
#define KVAE_FLAG_ELBO_ONLY 2    /* kvae_dims.flags: kvae_kf_bwd differentiates the ELBO alone, see kvae_grads */
#define KVAE_FLAG_WITH_ELBO 4    /* kvae_dims.flags: kvae_kf_bwd also EVALUATES the ELBO (fused value + adjoint), see below */
#define KVAE_FLAG_RAW_SUMS 8     /* with KVAE_FLAG_WITH_ELBO: leave every gradient un-normalised (data-parallel callers) */
#define KVAE_INFO_PIVOT 1
#define KVAE_INFO_PEER 2
#define KVAE_INFO_CHOL_S 4
#define KVAE_INFO_CHOL_Q 8

typedef struct kvae_dims {
  int32_t B;          /* sequences in this call (the per-rank shard)               */
  int32_t T;          /* time steps                                               */
  int32_t n;          /* z_dim  (state)                                           */
  int32_t p;          /* a_dim  (observation)                                     */
  int32_t m;          /* u_dim  (control)                                         */
  int32_t K;          /* number of base matrices mixed by alpha                   */
  int32_t q_per_mode; /* 1: Q is [K,n,n], Q_t = sum_k alpha_k Q_k (switching)     */
                      /* 0: Q is [n,n] fixed (lstm, KalmanFilter.Q buffer)        */
  int32_t c_shared;   /* 1: C_t = C[0] (switching); 0: C_t = sum_k alpha_k C_k    */
  int32_t lanes;      /* lanes of a warp that own one sequence; 0 = library picks */
  int32_t flags;      /* KVAE_FLAG_*                                              */
} kvae_dims;

/* problem inputs shared by all entry points */
typedef struct kvae_inputs {
  const float* Y;      /* [B,T,p] observations a_t                                  */
  const float* U;      /* [B,T,m] controls; NULL = all zeros (model.py:149-150)     */
  const float* mask;   /* [B,T] 1 = observed, 0 = missing; NULL = all ones          */
  const float* alpha;  /* [B,T,K] mixture weights                                   */
  const float* A;      /* [K,n,n]                                                   */
  const float* Bm;     /* [K,n,m]                                                   */
  const float* C;      /* [K,p,n]                                                   */
  const float* Q;      /* [K,n,n] or [n,n] (see q_per_mode)                         */
  const float* R;      /* [p,p]                                                     */
  const float* mu0;    /* [n]                                                       */
  const float* Sigma0; /* [n,n]                                                     */
  const float* mu_init;    /* [B,n]   optional per-sequence initial belief (NULL = mu0)     */
  const float* Sigma_init; /* [B,n,n] optional per-sequence initial belief (NULL = Sigma0)  */
  /* Explicit per-step matrices for the FORWARD entry (all NULL = mix from alpha).  With A_dense given
   * the kernel reads A_t,B_t,C_t (and Q_t if Q_dense) from these [B,T,...] tensors and ignores alpha
   * (which may be NULL): the per-step forms KalmanFilter.filter_step kalman_filter.py:31-104 and
   * .smooth_step :204-237 are T=1 / T=2 calls of this kind. */
  const float* A_dense;    /* [B,T,n,n] */
  const float* B_dense;    /* [B,T,n,m] */
  const float* C_dense;    /* [B,T,p,n] */
  const float* Q_dense;    /* [B,T,n,n] or NULL */
} kvae_inputs;

/* the six state tensors (written by the forward pass, read by ELBO / backward) */
typedef struct kvae_states {
  float* mus_filt;      /* [B,T,n]   */
  float* Sigmas_filt;   /* [B,T,n,n] */
  float* mus_pred;      /* [B,T,n]   */
  float* Sigmas_pred;   /* [B,T,n,n] */
  float* mus_smooth;    /* [B,T,n]   NULL in the forward pass = filter only */
  float* Sigmas_smooth; /* [B,T,n,n] */
  /* optional, kvae_kf_mask_partials_count(d) floats: the forward entry writes the per-CTA sums of the mask there
   * (kalman_filter.py:392 normalises the ELBO by sum(mask)); kvae_kf_bwd with KVAE_FLAG_WITH_ELBO (without RAW_SUMS) that
   * finds it non-NULL applies 1/max(sum mask,1) inside its sweep instead of re-scaling dY/dalpha/dU afterwards.  It must
   * come from a forward call on the SAME mask and dims (incl. lanes).  NULL = not used. */
  float* mask_partials;
  /* optional outputs of the FORWARD entry ([B,T,p] each; NULL = not wanted): the observation-space projections
   * KVAE.impute forms after smoothing, a_filtered = C_t mu_{t|t} (kvae/model/model.py:287-288) and
   * a_imputed = C_t mu_{t|T} (:280-281), emitted by the filter / smoother sweeps themselves (C_t is in registers there).
   * a_smooth needs mus_smooth / Sigmas_smooth (the smoother sweep); both are ignored by the ELBO / backward entries. */
  float* a_filt;
  float* a_smooth;
} kvae_states;

int kvae_abi_version(void);
const char* kvae_last_error(void); /* thread-local, never NULL */

/* 1 if (n,p,m,K,q_per_mode,c_shared,lanes) is instantiated, else 0 */
int kvae_supported(const kvae_dims* d);
/* lanes per sequence the library would use for this problem size */
int kvae_pick_lanes(const kvae_dims* d);

/* number of floats of kvae_states.mask_partials for this problem (0 if the shape is not instantiated) */
size_t kvae_kf_mask_partials_count(const kvae_dims* d);

/* Forward recursion.  A_list [B,T,n,n], B_list [B,T,n,m], C_list [B,T,p,n] may each be NULL
 * (not materialised). */
int kvae_kf_filter_smooth_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                              float* A_list, float* B_list, float* C_list,
                              int32_t* info, int device, void* stream);

/* Filter sweep with the LSTM "dynamics parameter network" in the loop: what
 * KalmanFilter.filter does with DynamicsParameter when observations are missing -- alpha_t =
 * softmax(head(LSTM_step(y_for_dyn_{t-1}))), y_for_dyn_t = mask_t y_t + (1 - mask_t) C_t mu_{t|t-1}
 * (kalman_filter.py:142,159,183-185; dyn_param.py:50-56) -- as ONE launch instead of a cuDNN step + ~10 ops + a filter
 * launch per time step.  in->alpha is ignored; alpha_out [B,T,K] receives the weights (= dyn_params.state_seq).
 * Weights in torch.nn.LSTM layout (1 layer, gate order i,f,g,o), hidden <= 52; lstm variant (q_per_mode = c_shared = 0),
 * d->lanes = 0 or n; st->mus_smooth / Sigmas_smooth are not written (run the forward entry with
 * KVAE_FLAG_SMOOTH_ONLY afterwards).  Forward only. */
typedef struct kvae_lstm {
  const float* w_ih;    /* [4*hidden, p]      lstm.weight_ih_l0 */
  const float* w_hh;    /* [4*hidden, hidden] lstm.weight_hh_l0 */
  const float* b_ih;    /* [4*hidden] */
  const float* b_hh;    /* [4*hidden] */
  const float* w_head;  /* [K, hidden]        head_w.weight */
  const float* b_head;  /* [K] */
  const float* h0;      /* [B, hidden] initial state or NULL (zeros) */
  const float* c0;
  float* h_out;         /* [B, hidden] final state or NULL */
  float* c_out;
  int32_t hidden;
} kvae_lstm;
int kvae_kf_filter_lstm_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                            float* A_list, float* B_list, float* C_list, const kvae_lstm* lstm, float* alpha_out,
                            int32_t* info, int device, void* stream);

/* ELBO of the smoothed posterior with the reparameterised sample z = mu_s + chol(Sigma_s + jitter I) eps.
 * terms[8] (device, fp32): [0] sum log p(z_t|z_{t-1})  [1] sum mask*log p(y_t|z_t)  [2] sum log p(z_0)
 *   [3] sum entropy  [4] sum(mask)  [5] elbo = ([0]+[1]+[2]+[3]) / max([4],1)  [6] 1/max([4],1)  [7] 0
 * workspace: kvae_kf_elbo_workspace_bytes(d) bytes of device scratch. */
size_t kvae_kf_elbo_workspace_bytes(const kvae_dims* d);
int kvae_kf_elbo_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                     const float* eps, float jitter, float* terms, void* workspace,
                     int32_t* info, int device, void* stream);

/* The rungs of KalmanFilter._safe_cholesky (kalman_filter.py:282-302) beyond the first attempt.  The reference runs one
 * ladder per call of _safe_cholesky, i.e. one for Sigma_smooth (:348) and an independent one for Q (:364-365): jitter
 * 1e-6, 1e-5, .. 1e-2 for the WHOLE batch as soon as any matrix fails, then L = diag(sqrt(clamp(diag(sym(X)), 1e-6))).
 * The kernels report which family failed (KVAE_INFO_CHOL_S / _Q); the caller re-launches with the next rung:
 *   jitter (argument of the entry point) is added to sym(Sigma_smooth), jitter_q to sym(Q_t);
 *   diag_smooth / diag_q != 0 select the diagonal fallback for that family (its jitter is then ignored); the adjoint
 *   differentiates the fallback as autograd would (only unclamped diagonal entries carry gradient).
 * NULL = { jitter_q = jitter, no fallback }: what kvae_kf_elbo_fwd / kvae_kf_bwd do. */
typedef struct kvae_chol_opts {
  float jitter_q;
  int32_t diag_smooth;
  int32_t diag_q;
} kvae_chol_opts;
int kvae_kf_elbo_fwd_ex(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                        const float* eps, float jitter, const kvae_chol_opts* opts, float* terms, void* workspace,
                        int32_t* info, int device, void* stream);

/* Cotangents of the nine `smooth` outputs (each may be NULL = zero). */
typedef struct kvae_cotangents {
  const float* mus_smooth;    const float* Sigmas_smooth;
  const float* mus_filt;      const float* Sigmas_filt;
  const float* mus_pred;      const float* Sigmas_pred;
  const float* A_list;        const float* B_list;        const float* C_list;
} kvae_cotangents;

typedef struct kvae_grads {
  float* dY;      /* [B,T,p]                                              */
  float* dU;      /* [B,T,m]  may be NULL                                 */
  float* dalpha;  /* [B,T,K]                                              */
  float* dA;      /* [K,n,n]                                              */
  float* dBm;     /* [K,n,m]                                              */
  float* dC;      /* [K,p,n]  (c_shared: only dC[0] is non-zero)          */
  float* dQ;      /* [K,n,n]  q_per_mode only; else may be NULL           */
  /* KVAE_FLAG_ELBO_ONLY: gradient of g_elbo*elbo with (mus_smooth, Sigmas_smooth) treated as free inputs
   * (KalmanFilter.elbo called with states that are not this library's own smoothed states): dmus/dSigmas
   * receive d elbo / d mu, d elbo / d Sigma; dY, dU, dalpha, dA.. hold the ELBO's direct terms only. */
  float* dmus;    /* [B,T,n]   */
  float* dSigmas; /* [B,T,n,n] */
} kvae_grads;

/* Explicit adjoint (reverse-time) pass: gradient of
 *     g_elbo * elbo  +  sum_i <cot_i, smooth_output_i>
 * with respect to Y, U, alpha, A, Bm, C, Q.  g_elbo is a DEVICE scalar (NULL = 0: no ELBO term;
 * then eps/terms may be NULL too); `terms` is the array written by kvae_kf_elbo_fwd (its [6] is
 * the global normaliser, which under data parallelism the caller overwrites with the all-reduced
 * value).  `cot` may be NULL.  workspace: kvae_kf_bwd_workspace_bytes(d) bytes. */
size_t kvae_kf_bwd_workspace_bytes(const kvae_dims* d);
int kvae_kf_bwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                const float* eps, float jitter, const float* g_elbo, float* terms,
                const kvae_cotangents* cot, const kvae_grads* grads, void* workspace,
                int32_t* info, int device, void* stream);

/* kvae_kf_bwd with the _safe_cholesky rung given explicitly (see kvae_chol_opts) */
int kvae_kf_bwd_ex(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                   const float* eps, float jitter, const kvae_chol_opts* opts, const float* g_elbo, float* terms,
                   const kvae_cotangents* cot, const kvae_grads* grads, void* workspace,
                   int32_t* info, int device, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Data parallelism (one process per GPU, batch sharded): the single exchange of the training step, i.e. the
 * sum over ranks of [dA | dB | dC | dQ | 5 ELBO sums] and the GLOBAL normalisation 1/max(sum mask, 1)
 * (kalman_filter.py:392-400), done by this library's own kernels over NVLink peer memory (CUDA IPC) instead of
 * an NCCL all-reduce followed by scaling kernels.  The reference has no multi-device path.
 *
 *   kvae_dp_create   allocates this rank's exchange buffer on `device` and returns its IPC handle
 *                    (kvae_dp_handle_bytes() bytes) for the caller to all-gather (any transport);
 *                    nfloats = K*n*n + K*n*m + K*p*n (+ K*n*n if q_per_mode).
 *   kvae_dp_connect  maps the peers' buffers; `handles` = world handles in rank order.
 *   kvae_dp_finalize call after kvae_kf_bwd(KVAE_FLAG_WITH_ELBO | KVAE_FLAG_RAW_SUMS) on the same stream, on EVERY rank
 *                    the same number of times: afterwards grads->dA/dBm/dC/dQ hold the normalised GLOBAL parameter
 *                    gradients (bit-identical on all ranks: rank-ordered fp64 sums), terms[0..7] the global ELBO
 *                    terms, and this rank's dY/dalpha/dU are scaled by the global normaliser.  Two launches, no host
 *                    synchronisation, capturable in a CUDA graph.  info is set to 2 if a peer never arrives (~20 s).
 */
typedef struct kvae_dp_comm kvae_dp_comm;
const char* kvae_dp_last_error(void);
size_t kvae_dp_handle_bytes(void);
int kvae_dp_create(int device, int rank, int world, size_t nfloats, kvae_dp_comm** out, void* handle_out);
int kvae_dp_connect(kvae_dp_comm* c, const void* handles);
int kvae_dp_destroy(kvae_dp_comm* c);
int kvae_dp_finalize(const kvae_dims* d, kvae_dp_comm* c, const kvae_grads* g, float* terms, int32_t* info,
                     void* stream);
/* The fused form of the two calls above: kvae_kf_bwd (d->flags must hold KVAE_FLAG_WITH_ELBO | KVAE_FLAG_RAW_SUMS) whose
 * final kernel ALSO does what kvae_dp_finalize does -- local reduction of the per-CTA partials, push to every rank,
 * global normalisation, scaling of dY/dalpha/dU -- so a data-parallel step has the same three launches as a single-GPU
 * step (k_filter_smooth, k_bwd, k_bwd_final_dp) and no NCCL call.  Same calling discipline as kvae_dp_finalize (every
 * rank, same number of times, same communicator; do not interleave the two forms on one communicator within a step). */
int kvae_kf_bwd_dp(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st,
                   const float* eps, float jitter, const float* g_elbo, float* terms,
                   const kvae_grads* grads, void* workspace, int32_t* info, int device, void* stream,
                   kvae_dp_comm* comm);

/* ---------------------------------------------------------------------------------------------------------
 * SKVAE regime sampler:  SwitchingDynamicsParameter.compute_batch
 * kvae/kalman/switch_dyn_param.py:51-79 — the T-step Gumbel-softmax Markov chain over K regimes and its
 * log q / log p bookkeeping — as one forward and one explicit-adjoint launch (the reference loops over t in Python
 * with ~12 small ops per step and differentiates them with autograd).  The bi-GRU that produces the logits
 * (MarkovVariationalRegimePosterior :113-129) stays in PyTorch/cuDNN.
 *   logits [B,T,K,K] (slice t=0 is never read, :67), init_logits [B,K], gumbel [B,T,K] = -log(Exp(1)) noise drawn by the
 *   caller (torch.nn.functional.gumbel_softmax draws it internally), trans [K,K] = StickyRegimePrior.transition_matrix.
 *   Outputs y_seq [B,T,K] (= state_seq, the alpha of the Kalman kernels), log_q [B,T], log_p [B,T].
 *   hard != 0: straight-through one-hot samples (eval mode, :52 `hard=not is_training`).
 * The backward entry returns the gradient of  <g_y, y_seq> + <g_logq, log_q> + <g_logp, log_p>  (each g_* may be NULL)
 * with respect to logits (d_logits [B,T,K,K], slice t=0 zero) and init_logits (d_init [B,K]).  K in 2..8. */
typedef struct kvae_regime_dims {
  int32_t B, T, K;
  int32_t hard;
  float tau;          /* Gumbel-softmax temperature (switch_dyn_param.py:15: 0.5) */
} kvae_regime_dims;
const char* kvae_regime_last_error(void);
int kvae_regime_supported(int K);
int kvae_regime_sample_fwd(const kvae_regime_dims* d, const float* logits, const float* init_logits,
                           const float* gumbel, const float* trans, float* y_seq, float* log_q, float* log_p,
                           int device, void* stream);
int kvae_regime_sample_bwd(const kvae_regime_dims* d, const float* logits, const float* init_logits,
                           const float* gumbel, const float* trans, const float* y_seq, const float* g_y,
                           const float* g_logq, const float* g_logp, float* d_logits, float* d_init,
                           int device, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * VAE-side reductions feeding the same loss:  vae_loss  kvae/vae/losses.py:62-111 and
 * KVAE.reparameterize  kvae/model/model.py:81-84.  Frames x [frames, pixels] (= [B*T, C*H*W]), reconstruction means /
 * logits x_mu (same shape), latents a, a_mu, a_var [frames, a_dim], mask [frames] or NULL (= ones).
 *   out8: [0] vae_elbo = scale_reconstruction * recon + beta * reg   [1] recon = sum(m log p(x|a)) / denom
 *         [2] reg = (sum m log p(a) - sum m log q(a|x)) / denom       [3] 1/denom, denom = max(sum m, 1)   [4..7] raw sums
 *   bernoulli != 0: log p(x|a) = -BCEWithLogits(x_mu, x) (x_var ignored); else Gaussian with the scalar variance x_var.
 * The backward entry returns the gradient of g3[0]*vae_elbo + g3[1]*recon + g3[2]*reg (g3: 3 device floats) with
 * respect to x_mu, a, a_mu, a_var.  workspace: kvae_vae_loss_workspace_bytes(d) bytes (forward only). */
typedef struct kvae_vae_dims {
  int32_t frames, pixels, a_dim, bernoulli;
  float x_var, scale_reconstruction, beta;
} kvae_vae_dims;
const char* kvae_vae_last_error(void);
size_t kvae_vae_loss_workspace_bytes(const kvae_vae_dims* d);
int kvae_vae_loss_fwd(const kvae_vae_dims* d, const float* x, const float* x_mu, const float* a, const float* a_mu,
                      const float* a_var, const float* mask, float* out8, void* workspace, int device, void* stream);
int kvae_vae_loss_bwd(const kvae_vae_dims* d, const float* x, const float* x_mu, const float* a, const float* a_mu,
                      const float* a_var, const float* mask, const float* g3, const float* out8, float* d_x_mu, float* d_a,
                      float* d_a_mu, float* d_a_var, int device, void* stream);
/* a = mu + eps * sqrt(var + 1e-6); backward: d_mu = g (the caller's own), d_var = g * eps / (2 sqrt(var + 1e-6)) */
int kvae_vae_reparam_fwd(const float* mu, const float* var, const float* eps, long n, float* a, int device, void* stream);
int kvae_vae_reparam_bwd(const float* var, const float* eps, const float* g, long n, float* d_var, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KVAE_KALMAN_H */
