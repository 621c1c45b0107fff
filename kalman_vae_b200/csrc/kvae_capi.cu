// kvae_capi.cu — the extern "C" boundary declared in include/kvae_kalman.h.
#include <cstdio>
#include <cstring>
#include "kvae_configs.h"
#include "kvae_ops.h"

namespace {
thread_local char g_err[256] = "";
int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (dev >= 0) { cudaGetDevice(&prev); if (prev != dev) { cudaSetDevice(dev); switched = true; } }
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};
bool variant_ok(const kvae_dims& d) { return (d.q_per_mode != 0) == (d.c_shared != 0); }

// lanes per sequence.  Measured on B200 (profiles/r02_*):
//  * n = 4, u_dim = 4, T % 4 == 0 and at least 12 288 sequences: ONE lane, i.e. the thread-per-sequence kernels with
//    TMA-staged streams (csrc/kvae_seq.cuh): B = 262 144, T = 20: forward 0.72 ms vs 0.84 ms, adjoint 1.93 vs 2.31 ms.
//    Crossover measured in profiles/r02_lanes_crossover.log (T = 20, fused step): B = 8 192: 125 us (4 lanes) vs 181 us
//    (1 lane); 12 288: 219 vs 193; 16 384: 238 vs 203; 32 768: 465 vs 394 (24 576, where the one-lane grid is 1.3 waves,
//    is the exception: 360 vs 380).
//  * smaller batches are latency bound (a sequence is a chain of 4 T dependent steps): both families need ~2 000 cycles
//    per step there; one row per lane (L = n) has more warps to overlap (cfg2, B = 8 192: adjoint 90 vs 130 us).
//  * n = 16: L = 16 (fewer spills).  n = 8: smallest count that still gives >= 8 warps per SM.
int pick_lanes(const kvae_dims& d) {
  const int n = d.n;
  if (n > 0 && (n & (n - 1)) != 0)   // z_dim not a power of two: its largest power-of-two divisor (see kvae_shape.cu)
    return (n % 16 == 0) ? 16 : (n % 8 == 0) ? 8 : (n % 4 == 0) ? 4 : (n % 2 == 0) ? 2 : 1;
  if (n == 4 && d.m == 4 && d.T % 4 == 0 && d.B >= 12288 && !(d.flags & KVAE_FLAG_SMOOTH_ONLY)) return 1;
  if (n <= 4) return n < 1 ? 1 : n;
  if (n != 8) return n;
  const long want_threads = 148L * 8 * 32;
  return ((long)d.B * 4 >= want_threads) ? 4 : 8;
}
}  // namespace


extern "C" {

int kvae_abi_version(void) { return KVAE_ABI_VERSION; }
const char* kvae_last_error(void) { return g_err; }

int kvae_pick_lanes(const kvae_dims* d) { return d ? pick_lanes(*d) : 0; }

int kvae_supported(const kvae_dims* d) {
  if (!d || !variant_ok(*d)) return 0;
  const int lanes = d->lanes ? d->lanes : pick_lanes(*d);
#define X(n_, p_, m_, k_) \
  if (d->n == n_ && d->p == p_ && d->m == m_ && d->K == k_) return kvae::ShapeOps<n_, p_, m_, k_>::lanes_ok(lanes) ? 1 : 0;
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return 0;
}

static int check_common(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, kvae_dims* dd) {
  if (!d || !in || !st) return fail(-1, "null argument");
  if (d->B <= 0 || d->T <= 0) return fail(-1, "B and T must be positive");
  if (!variant_ok(*d)) return fail(-2, "unsupported variant: q_per_mode and c_shared must both be 0 (lstm) or 1 (switching)");
  *dd = *d;
  if (dd->lanes == 0) dd->lanes = pick_lanes(*dd);
  if (!kvae_supported(dd)) return fail(-2, "unsupported (n,p,m,K,lanes): see kvae_configs.h");
  if (!in->Y || !in->A || !in->Bm || !in->C || !in->Q || !in->R || !in->mu0 || !in->Sigma0)
    return fail(-1, "null input tensor");
  if (!in->alpha && !in->A_dense) return fail(-1, "alpha is null and no explicit A_dense/B_dense/C_dense given");
  if (in->A_dense && (!in->B_dense || !in->C_dense)) return fail(-1, "A_dense needs B_dense and C_dense");
  if (!st->mus_filt || !st->Sigmas_filt || !st->mus_pred || !st->Sigmas_pred) return fail(-1, "null state tensor");
  return 0;
}

int kvae_kf_filter_smooth_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, float* A_list,
                              float* B_list, float* C_list, int32_t* info, int device, void* stream) {
  kvae_dims dd;
  if (int rc = check_common(d, in, st, &dd)) return rc;
  if (!info) return fail(-1, "info must not be null");
  if ((st->mus_smooth == nullptr) != (st->Sigmas_smooth == nullptr)) return fail(-1, "mus_smooth/Sigmas_smooth: both or neither");
  if ((dd.flags & KVAE_FLAG_SMOOTH_ONLY) && !st->mus_smooth) return fail(-1, "smooth-only call without smoothed-state buffers");
  DeviceGuard guard(device);
#define X(n_, p_, m_, k_)                                                                                   \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) {                                              \
    int rc = kvae::ShapeOps<n_, p_, m_, k_>::fwd(dd, *in, *st, A_list, B_list, C_list, info, (cudaStream_t)stream); \
    if (rc > 0) return fail(rc, cudaGetErrorString((cudaError_t)rc));                                       \
    if (rc < 0) return fail(rc, "forward launch rejected");                                                 \
    return 0;                                                                                               \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return fail(-2, "unsupported shape");
}

size_t kvae_kf_mask_partials_count(const kvae_dims* d) {
  if (!d) return 0;
  kvae_dims dd = *d;
  if (dd.lanes == 0) dd.lanes = pick_lanes(dd);
  if (!kvae_supported(&dd)) return 0;
#define X(n_, p_, m_, k_) \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) return (size_t)kvae::ShapeOps<n_, p_, m_, k_>::fwd_grid(dd);
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return 0;
}

int kvae_kf_filter_lstm_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, float* A_list, float* B_list,
                            float* C_list, const kvae_lstm* lstm, float* alpha_out, int32_t* info, int device, void* stream) {
  if (!d || !in || !st || !lstm || !alpha_out || !info) return fail(-1, "null argument");
  if (d->B <= 0 || d->T <= 0) return fail(-1, "B and T must be positive");
  if (d->q_per_mode || d->c_shared) return fail(-2, "the LSTM dynamics network belongs to the lstm variant (q_per_mode = c_shared = 0)");
  if (!in->Y || !in->A || !in->Bm || !in->C || !in->Q || !in->R || !in->mu0 || !in->Sigma0) return fail(-1, "null input tensor");
  if (in->A_dense) return fail(-2, "explicit per-step matrices cannot be combined with the in-kernel dynamics network");
  if (!st->mus_filt || !st->Sigmas_filt || !st->mus_pred || !st->Sigmas_pred) return fail(-1, "null state tensor");
  if (!lstm->w_ih || !lstm->w_hh || !lstm->b_ih || !lstm->b_hh || !lstm->w_head || !lstm->b_head) return fail(-1, "null LSTM weight");
  if (lstm->hidden < 1 || lstm->hidden > 52) return fail(-2, "hidden size must be in 1..52");
  kvae_dims dd = *d;
  if (dd.lanes == 0) dd.lanes = dd.n;
  DeviceGuard guard(device);
#define X(n_, p_, m_, k_)                                                                                   \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) {                                              \
    int rc = kvae::ShapeOps<n_, p_, m_, k_>::fwd_lstm(dd, *in, *st, A_list, B_list, C_list, *lstm, alpha_out, info, (cudaStream_t)stream); \
    if (rc > 0) return fail(rc, cudaGetErrorString((cudaError_t)rc));                                       \
    if (rc < 0) return fail(-2, "shape / lane count not instantiated for the in-kernel LSTM (K > 1, n <= 8, lanes = n)"); \
    return 0;                                                                                               \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return fail(-2, "unsupported shape");
}

size_t kvae_kf_elbo_workspace_bytes(const kvae_dims* d) {
  if (!d) return 0;
  kvae_dims dd = *d;
  if (dd.lanes == 0) dd.lanes = pick_lanes(dd);
#define X(n_, p_, m_, k_) \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) return kvae::ShapeOps<n_, p_, m_, k_>::elbo_ws(dd);
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return 0;
}

int kvae_kf_elbo_fwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                     float* terms, void* workspace, int32_t* info, int device, void* stream) {
  return kvae_kf_elbo_fwd_ex(d, in, st, eps, jitter, nullptr, terms, workspace, info, device, stream);
}

int kvae_kf_elbo_fwd_ex(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                        const kvae_chol_opts* opts, float* terms, void* workspace, int32_t* info, int device, void* stream) {
  const float jitter_q = opts ? opts->jitter_q : jitter;
  const int chol_diag = opts ? ((opts->diag_smooth ? 1 : 0) | (opts->diag_q ? 2 : 0)) : 0;
  kvae_dims dd;
  if (int rc = check_common(d, in, st, &dd)) return rc;
  if (!info || !eps || !terms || !workspace) return fail(-1, "null argument");
  if (!st->mus_smooth || !st->Sigmas_smooth) return fail(-1, "smoothed states required");
  if (in->A_dense) return fail(-2, "explicit per-step matrices are supported by the forward entry only");
  DeviceGuard guard(device);
#define X(n_, p_, m_, k_)                                                                                   \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) {                                              \
    int rc = kvae::ShapeOps<n_, p_, m_, k_>::elbo(dd, *in, *st, eps, jitter, jitter_q, chol_diag, terms, workspace, info, (cudaStream_t)stream); \
    if (rc > 0) return fail(rc, cudaGetErrorString((cudaError_t)rc));                                       \
    if (rc < 0) return fail(rc, "elbo launch rejected");                                                    \
    return 0;                                                                                               \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return fail(-2, "unsupported shape");
}

static size_t bwd_ws_for(const kvae_dims& dd) {
#define X(n_, p_, m_, k_) \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) return kvae::ShapeOps<n_, p_, m_, k_>::bwd_ws(dd);
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return 0;
}
size_t kvae_kf_bwd_workspace_bytes(const kvae_dims* d) {
  if (!d) return 0;
  kvae_dims dd = *d;
  if (dd.lanes != 0) return bwd_ws_for(dd);
  // library-picked lanes: a call with dense cotangents runs on the lane-group kernels even where the thread-per-sequence
  // kernels are the default (see bwd_impl), so the workspace covers both
  dd.lanes = pick_lanes(dd);
  size_t ws = bwd_ws_for(dd);
  if (dd.lanes == 1 && dd.n > 1) {
    dd.lanes = dd.n;
    const size_t w2 = bwd_ws_for(dd);
    if (w2 > ws) ws = w2;
  }
  return ws;
}

static int bwd_impl(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                    const float* g_elbo, float* terms, const kvae_cotangents* cot, const kvae_grads* grads,
                    void* workspace, int32_t* info, int device, void* stream, kvae_dp_comm* comm,
                    const kvae_chol_opts* opts = nullptr) {
  const float jitter_q = opts ? opts->jitter_q : jitter;
  const int chol_diag = opts ? ((opts->diag_smooth ? 1 : 0) | (opts->diag_q ? 2 : 0)) : 0;
  kvae_dims dd;
  if (int rc = check_common(d, in, st, &dd)) return rc;
  if (!info || !grads || !workspace) return fail(-1, "null argument");
  if (!st->mus_smooth || !st->Sigmas_smooth) return fail(-1, "smoothed states required");
  if (g_elbo && (!eps || !terms)) return fail(-1, "g_elbo given without eps/terms");
  if (!grads->dY || !grads->dalpha || !grads->dA || !grads->dBm || !grads->dC) return fail(-1, "null gradient buffer");
  if (dd.q_per_mode && !grads->dQ) return fail(-1, "dQ required when q_per_mode");
  if ((dd.flags & KVAE_FLAG_ELBO_ONLY) && (!grads->dmus || !grads->dSigmas || !g_elbo)) return fail(-1, "ELBO_ONLY needs dmus, dSigmas and g_elbo");
  if (in->A_dense) return fail(-2, "explicit per-step matrices are supported by the forward entry only");
  if (dd.flags & KVAE_FLAG_WITH_ELBO) {
    if (!g_elbo) return fail(-1, "WITH_ELBO needs g_elbo, eps and terms");
    const bool any_cot = cot && (cot->mus_smooth || cot->Sigmas_smooth || cot->mus_filt || cot->Sigmas_filt || cot->mus_pred ||
                                 cot->Sigmas_pred || cot->A_list || cot->B_list || cot->C_list);
    if (any_cot) return fail(-2, "WITH_ELBO cannot be combined with dense cotangents (they must not be normalised)");
    if ((dd.flags & KVAE_FLAG_ELBO_ONLY) && !(dd.flags & KVAE_FLAG_RAW_SUMS))
      return fail(-2, "WITH_ELBO|ELBO_ONLY needs RAW_SUMS (dmus/dSigmas are not covered by the fused normalisation)");
  } else if (dd.flags & KVAE_FLAG_RAW_SUMS) {
    return fail(-2, "RAW_SUMS without WITH_ELBO");
  }
  {   // dense cotangents are not covered by the thread-per-sequence adjoint: with library-picked lanes use lane groups
    const bool any_cot = cot && (cot->mus_smooth || cot->Sigmas_smooth || cot->mus_filt || cot->Sigmas_filt || cot->mus_pred ||
                                 cot->Sigmas_pred || cot->A_list || cot->B_list || cot->C_list);
    if (d->lanes == 0 && dd.lanes == 1 && dd.n > 1 && (any_cot || (dd.flags & KVAE_FLAG_ELBO_ONLY))) dd.lanes = dd.n;
  }
  kvae::DpView view;
  if (comm) {
    if ((dd.flags & (KVAE_FLAG_WITH_ELBO | KVAE_FLAG_RAW_SUMS)) != (KVAE_FLAG_WITH_ELBO | KVAE_FLAG_RAW_SUMS) ||
        (dd.flags & KVAE_FLAG_ELBO_ONLY))
      return fail(-2, "kvae_kf_bwd_dp needs flags WITH_ELBO | RAW_SUMS (and not ELBO_ONLY)");
    if (!kvae::kvae_dp_get_view(comm, &view)) return fail(-1, "kvae_kf_bwd_dp: communicator not connected");
  }
  DeviceGuard guard(device);
  kvae::BwdExtra x{eps, jitter, jitter_q, chol_diag, g_elbo, terms, cot, grads, workspace, comm ? &view : nullptr};
#define X(n_, p_, m_, k_)                                                                                   \
  if (dd.n == n_ && dd.p == p_ && dd.m == m_ && dd.K == k_) {                                              \
    int rc = kvae::ShapeOps<n_, p_, m_, k_>::bwd(dd, *in, *st, x, info, (cudaStream_t)stream);             \
    if (rc > 0) return fail(rc, cudaGetErrorString((cudaError_t)rc));                                       \
    if (rc < 0) return fail(rc, "backward launch rejected");                                                \
    return 0;                                                                                               \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return fail(-2, "unsupported shape");
}

int kvae_kf_bwd(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                const float* g_elbo, float* terms, const kvae_cotangents* cot, const kvae_grads* grads,
                void* workspace, int32_t* info, int device, void* stream) {
  return bwd_impl(d, in, st, eps, jitter, g_elbo, terms, cot, grads, workspace, info, device, stream, nullptr);
}

int kvae_kf_bwd_ex(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                   const kvae_chol_opts* opts, const float* g_elbo, float* terms, const kvae_cotangents* cot,
                   const kvae_grads* grads, void* workspace, int32_t* info, int device, void* stream) {
  return bwd_impl(d, in, st, eps, jitter, g_elbo, terms, cot, grads, workspace, info, device, stream, nullptr, opts);
}

int kvae_kf_bwd_dp(const kvae_dims* d, const kvae_inputs* in, const kvae_states* st, const float* eps, float jitter,
                   const float* g_elbo, float* terms, const kvae_grads* grads, void* workspace, int32_t* info,
                   int device, void* stream, kvae_dp_comm* comm) {
  if (!comm) return fail(-1, "kvae_kf_bwd_dp: null communicator");
  return bwd_impl(d, in, st, eps, jitter, g_elbo, terms, nullptr, grads, workspace, info, device, stream, comm);
}

}  // extern "C"
