// kvae_shape.cu — compiled once per shape with -DKV_N= -DKV_P= -DKV_M= -DKV_K= ; defines
// ShapeOps<KV_N,KV_P,KV_M,KV_K> (lane-count and dynamics-variant dispatch).
#include "kvae_ops.h"
#include "kvae_seq_bwd.cuh"

namespace kvae {

namespace {
constexpr int N = KV_N, P = KV_P, M = KV_M, K = KV_K;

// lane counts instantiated for this z_dim (rows per lane R = N / L kept <= 4 for N >= 8).  A lane count must be a power
// of two that divides z_dim: a z_dim that is not a power of two (shapes built on demand, kalman_vae_b200/build.py
// build_shape_lib) runs on its largest power-of-two divisor (z_dim odd: one lane per sequence).
#define KV_N_POW2 ((KV_N & (KV_N - 1)) == 0)
#if !KV_N_POW2
#define KV_G ((KV_N % 16 == 0) ? 16 : (KV_N % 8 == 0) ? 8 : (KV_N % 4 == 0) ? 4 : (KV_N % 2 == 0) ? 2 : 1)
#define KV_FOR_EACH_L(X) X(KV_G)
#define KV_L_OK(l) ((l) == KV_G)
#elif KV_N <= 4
#define KV_FOR_EACH_L(X) X(1) X(2) X(KV_N)
#define KV_L_OK(l) ((l) == 1 || (l) == 2 || (l) == KV_N)
#elif KV_N == 8
#define KV_FOR_EACH_L(X) X(4) X(8)
#define KV_L_OK(l) ((l) == 4 || (l) == 8)
#else
#define KV_FOR_EACH_L(X) X(KV_N / 2) X(KV_N)
#define KV_L_OK(l) ((l) == KV_N / 2 || (l) == KV_N)
#endif

Args make_args(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, int32_t* info) {
  Args a{};
  a.B = d.B; a.T = d.T;
  a.Y = in.Y; a.U = in.U; a.mask = in.mask; a.alpha = in.alpha;
  a.mu_f = st.mus_filt; a.Sig_f = st.Sigmas_filt; a.mu_p = st.mus_pred; a.Sig_p = st.Sigmas_pred;
  a.mu_s = st.mus_smooth; a.Sig_s = st.Sigmas_smooth;
  a.mu_init = in.mu_init; a.Sig_init = in.Sigma_init;
  a.dA = in.A_dense; a.dB = in.B_dense; a.dC = in.C_dense; a.dQ = in.Q_dense;
  a.smooth_only = (d.flags & KVAE_FLAG_SMOOTH_ONLY) ? 1 : 0;
  a.mask_part = nullptr;
  a.a_filt = st.a_filt; a.a_smooth = st.a_smooth;
  a.jitter_q = 1e-6f;
  a.chol_diag = 0;
  a.info = info;
  return a;
}
BasePtrs make_base(const kvae_inputs& in) { return BasePtrs{in.A, in.Bm, in.C, in.Q, in.R, in.mu0, in.Sigma0}; }
}  // namespace

template <> bool ShapeOps<N, P, M, K>::lanes_ok(int lanes) { return KV_L_OK(lanes); }

// thread-per-sequence kernels (csrc/kvae_seq.cuh): lanes == 1, z_dim = u_dim = 4, T % 4 == 0
constexpr bool SEQ_SHAPE = (N == 4 && M == 4);
static bool seq_dims_ok(const kvae_dims& d) { return SEQ_SHAPE && d.lanes == 1 && d.T % 4 == 0 && !(d.flags & KVAE_FLAG_SMOOTH_ONLY); }

template <> int ShapeOps<N, P, M, K>::fwd_grid(const kvae_dims& d) {
  if (seq_dims_ok(d)) return seq_grid(d.B);
#define X(l) \
  if (d.lanes == (l)) { if constexpr (N % (l) == 0) return fwd_grid_of<Cfg<N, P, M, K, (l), false, false>>(d.B); }
  KV_FOR_EACH_L(X)
#undef X
  return 0;
}

template <>
int ShapeOps<N, P, M, K>::fwd(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, float* A_list,
                              float* B_list, float* C_list, int32_t* info, cudaStream_t s) {
  Args a = make_args(d, in, st, info);
  a.A_list = A_list; a.B_list = B_list; a.C_list = C_list;
  if (!a.smooth_only && !in.A_dense) a.mask_part = st.mask_partials;
  const BasePtrs bp = make_base(in);
  const int smooth = (st.mus_smooth != nullptr) ? 1 : 0;
  const bool sw = d.q_per_mode != 0;
  if constexpr (SEQ_SHAPE) {
    if (seq_dims_ok(d) && seq_eligible(a)) {
      return sw ? launch_seq_fwd<Cfg<N, P, M, K, 1, true, true>>(a, bp, smooth, s)
                : launch_seq_fwd<Cfg<N, P, M, K, 1, false, false>>(a, bp, smooth, s);
    }
  }
#define X(l)                                                                                   \
  if (d.lanes == (l)) {                                                                        \
    if constexpr (N % (l) == 0) {                                                              \
      return sw ? launch_fwd<Cfg<N, P, M, K, (l), true, true>>(a, bp, smooth, s)               \
                : launch_fwd<Cfg<N, P, M, K, (l), false, false>>(a, bp, smooth, s);            \
    }                                                                                          \
  }
  KV_FOR_EACH_L(X)
#undef X
  return -3;
}

// LSTM dynamics in the filter loop: lstm variant, widest lane count, shapes with K > 1 and n <= 8
template <>
int ShapeOps<N, P, M, K>::fwd_lstm(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, float* A_list, float* B_list,
                                   float* C_list, const kvae_lstm& lw, float* alpha_out, int32_t* info, cudaStream_t s) {
#if KV_K > 1 && KV_N <= 8 && KV_N_POW2
  constexpr int LW = N;   // one row per lane
  if (d.lanes != LW || d.q_per_mode || lw.hidden > 52 || lw.hidden < 1) return -3;
  Args a = make_args(d, in, st, info);
  a.alpha = nullptr;
  a.A_list = A_list; a.B_list = B_list; a.C_list = C_list;
  const BasePtrs bp = make_base(in);
  LstmPtrs p{lw.w_ih, lw.w_hh, lw.b_ih, lw.b_hh, lw.w_head, lw.b_head, lw.h0, lw.c0, lw.h_out, lw.c_out, alpha_out, lw.hidden};
  return launch_fwd_lstm<Cfg<N, P, M, K, LW, false, false>>(a, bp, p, s);
#else
  (void)d; (void)in; (void)st; (void)A_list; (void)B_list; (void)C_list; (void)lw; (void)alpha_out; (void)info; (void)s;
  return -3;
#endif
}

template <> size_t ShapeOps<N, P, M, K>::elbo_ws(const kvae_dims& d) {
#define X(l) \
  if (d.lanes == (l)) { if constexpr (N % (l) == 0) return elbo_ws_bytes<Cfg<N, P, M, K, (l), false, false>>(d.B, d.T); }
  KV_FOR_EACH_L(X)
#undef X
  return 0;
}

template <>
int ShapeOps<N, P, M, K>::elbo(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, const float* eps,
                               float jitter, float jitter_q, int chol_diag, float* terms, void* ws, int32_t* info, cudaStream_t s) {
  Args a = make_args(d, in, st, info);
  a.eps = eps;
  a.jitter_q = jitter_q;
  a.chol_diag = chol_diag;
  const BasePtrs bp = make_base(in);
  const bool sw = d.q_per_mode != 0;
#define X(l)                                                                                     \
  if (d.lanes == (l)) {                                                                          \
    if constexpr (N % (l) == 0) {                                                                \
      return sw ? launch_elbo<Cfg<N, P, M, K, (l), true, true>>(a, bp, jitter, terms, ws, s)     \
                : launch_elbo<Cfg<N, P, M, K, (l), false, false>>(a, bp, jitter, terms, ws, s);  \
    }                                                                                            \
  }
  KV_FOR_EACH_L(X)
#undef X
  return -3;
}

template <> size_t ShapeOps<N, P, M, K>::bwd_ws(const kvae_dims& d) {
  const bool sw = d.q_per_mode != 0;
  size_t seq_ws = 0;   // a lanes == 1 call may run on either kernel family (dense cotangents -> lane groups): take the larger
  if constexpr (SEQ_SHAPE) {
    if (seq_dims_ok(d)) seq_ws = sw ? seq_bwd_ws_bytes<Cfg<N, P, M, K, 1, true, true>>(d.B, d.T)
                                    : seq_bwd_ws_bytes<Cfg<N, P, M, K, 1, false, false>>(d.B, d.T);
  }
#define X(l)                                                                                     \
  if (d.lanes == (l)) {                                                                          \
    if constexpr (N % (l) == 0) {                                                                \
      const size_t lg = sw ? bwd_ws_bytes<Cfg<N, P, M, K, (l), true, true>>(d.B, d.T)           \
                           : bwd_ws_bytes<Cfg<N, P, M, K, (l), false, false>>(d.B, d.T);        \
      return lg > seq_ws ? lg : seq_ws;                                                          \
    }                                                                                            \
  }
  KV_FOR_EACH_L(X)
#undef X
  return 0;
}

template <>
int ShapeOps<N, P, M, K>::bwd(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, const BwdExtra& x,
                              int32_t* info, cudaStream_t s) {
  Args a = make_args(d, in, st, info);
  a.eps = x.eps;
  a.jitter_q = x.jitter_q;
  a.chol_diag = x.chol_diag;
  const BasePtrs bp = make_base(in);
  BwdArgs w{};
  if (x.cot) {
    w.c_mu_s = x.cot->mus_smooth; w.c_Sig_s = x.cot->Sigmas_smooth; w.c_mu_f = x.cot->mus_filt;
    w.c_Sig_f = x.cot->Sigmas_filt; w.c_mu_p = x.cot->mus_pred; w.c_Sig_p = x.cot->Sigmas_pred;
    w.c_A = x.cot->A_list; w.c_B = x.cot->B_list; w.c_C = x.cot->C_list;
  }
  w.dY = x.grads->dY; w.dU = x.grads->dU; w.dalpha = x.grads->dalpha;
  w.jitter = x.jitter;
  w.elbo_only = (d.flags & KVAE_FLAG_ELBO_ONLY) ? 1 : 0;
  w.with_elbo = (d.flags & KVAE_FLAG_WITH_ELBO) ? 1 : 0;
  w.raw_sums = (d.flags & KVAE_FLAG_RAW_SUMS) ? 1 : 0;
  w.terms_out = x.terms;
  if (w.with_elbo && !w.raw_sums && !w.elbo_only && st.mask_partials) {
    w.mask_part = st.mask_partials;
    w.n_mask_part = fwd_grid(d);
  }
  w.e_dSig = x.grads->dSigmas; w.e_dmu = x.grads->dmus;
  GradPtrs gp{x.grads->dA, x.grads->dBm, x.grads->dC, x.grads->dQ};
  const bool sw = d.q_per_mode != 0;
  if constexpr (SEQ_SHAPE) {
    if (seq_dims_ok(d) && seq_bwd_eligible(a, w, x.g_elbo)) {
      return sw ? launch_seq_bwd<Cfg<N, P, M, K, 1, true, true>>(a, w, bp, x.g_elbo, x.terms, x.workspace, gp, s, x.dp)
                : launch_seq_bwd<Cfg<N, P, M, K, 1, false, false>>(a, w, bp, x.g_elbo, x.terms, x.workspace, gp, s, x.dp);
    }
  }
#define X(l)                                                                                             \
  if (d.lanes == (l)) {                                                                                  \
    if constexpr (N % (l) == 0) {                                                                        \
      return sw ? launch_bwd<Cfg<N, P, M, K, (l), true, true>>(a, w, bp, x.g_elbo, x.terms, x.workspace, gp, s, x.dp)    \
                : launch_bwd<Cfg<N, P, M, K, (l), false, false>>(a, w, bp, x.g_elbo, x.terms, x.workspace, gp, s, x.dp); \
    }                                                                                                    \
  }
  KV_FOR_EACH_L(X)
#undef X
  return -3;
}

}  // namespace kvae
