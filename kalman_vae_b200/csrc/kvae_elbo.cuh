// kvae_elbo.cuh — A.3: ELBO of the smoothed posterior (kalman_filter.py:305-401), one sequence per
// lane group, forward in time.  z_t = mu_s + chol(sym(Sigma_s)+jitter I) eps_t is the reference's
// rsample (:348-351) with the standard-normal draw eps supplied by the caller.
#pragma once
#include "kvae_fwd.cuh"

namespace kvae {

#define KV_LOG2PI 1.8378770664093453f

// own rows of sym(X) (+ jitter on the diagonal)
template <class C>
KV_FN void sym_jitter_rows(const Group<C::L, C::R>& g, const float (&X)[C::R][C::N], TileRef buf, float jitter,
                           float (&out)[C::R][C::N]) {
  constexpr int N = C::N, R = C::R;
  auto X_v = publish<C::MEM, C::L, R, N>(g, X, buf);
  float Xt[R][N];
  tr_rows<R, N>(X_v, g.row0(), Xt);
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j)
    out[r][j] = 0.5f * (X[r][j] + Xt[r][j]) + ((g.row0() + r == j) ? jitter : 0.f);
}

// sum over own rows of log(diag) then over the group
template <class C> KV_FN float logdet_half(const Group<C::L, C::R>& g, const float (&dg)[C::R]) {
  float s = 0.f;
  KV_UNROLL for (int r = 0; r < C::R; ++r) s += logf(dg[r]);
  return g.allreduce1(s);
}

// jitters and fallback switches of the two ELBO factorisations (see Args::jitter_q / chol_diag)
struct CholOpt {
  float js, jq;
  int diag_s, diag_q;
};
KV_FN CholOpt chol_opt(const Args& a, float jitter_s) {
  return CholOpt{jitter_s, a.jitter_q, a.chol_diag & 1, (a.chol_diag >> 1) & 1};
}

// constants of the ELBO that do not depend on t
template <class C> struct ElboConst {
  float LR[C::P][C::P], invdR[C::P], logdetR;          // chol(R)                         (:373)
  float LQ[C::R][C::N], invdQ[C::N], logdetQ;          // chol(sym(Q)+jitter I), fixed Q  (:364-365)
};

// returns the status bits of what failed (0 = fine): chol(R) -> KV_INFO_PIVOT, chol(sym(Q)+jitter) -> KV_INFO_CHOL_Q
template <class C>
KV_FN int elbo_const(const Group<C::L, C::R>& g, const float* base, TileRef xbuf, const CholOpt& co, ElboConst<C>& ec) {
  constexpr int N = C::N, P = C::P, R = C::R;
  int bad = 0;
  float Rm[P][P];
  KV_UNROLL for (int a = 0; a < P; ++a) KV_UNROLL for (int b = 0; b < P; ++b) Rm[a][b] = base[Base<C>::oR + a * P + b];
  if (!chol_small<P>(Rm, ec.LR, ec.invdR)) bad |= KV_INFO_PIVOT;
  ec.logdetR = 0.f;
  KV_UNROLL for (int a = 0; a < P; ++a) ec.logdetR += logf(ec.LR[a][a]);
  if constexpr (!C::QPM) {
    float Q[R][N], Qs[R][N], dg[R];
    copy_rows<C, N, Base<C>::ldQ>(base + Base<C>::oQ, g.row0(), Q);
    sym_jitter_rows<C>(g, Q, xbuf, co.diag_q ? 0.f : co.jq, Qs);
    unsigned clq;
    if (!chol_dist_opt<C::L, R>(g, Qs, ec.LQ, ec.invdQ, dg, co.diag_q, clq)) bad |= KV_INFO_CHOL_Q;
    ec.logdetQ = logdet_half<C>(g, dg);
  }
  return bad;
}

// Per-step pieces of the ELBO that the adjoint recomputes as well.
template <class C> struct ElboStep {
  float Ls[C::R][C::N];   // own rows of chol(sym(Sigma_s)+jI)
  float invd[C::N];
  float dg[C::R];
  float z_own[C::R];
  float z[C::N];          // replicated sample
  unsigned clamped;       // diagonal fallback: bit j = diagonal entry j of sym(Sigma_s) was clamped
};

// z_t from the lane's rows of Sigma_s and entries of mu_s (xbuf: [N x N] tile, vbuf: N-vector slot)
template <class C>
KV_FN bool elbo_sample_rows(const Group<C::L, C::R>& g, TileRef xbuf, TileRef vbuf, const float (&Ss)[C::R][C::N],
                            const float (&mus)[C::R], const CholOpt& co, const float (&eps)[C::N], ElboStep<C>& es) {
  constexpr int N = C::N, R = C::R;
  float Sj[R][N];
  sym_jitter_rows<C>(g, Ss, xbuf, co.diag_s ? 0.f : co.js, Sj);        // (:287, :293)
  const bool ok = chol_dist_opt<C::L, R>(g, Sj, es.Ls, es.invd, es.dg, co.diag_s, es.clamped);
  KV_UNROLL for (int r = 0; r < R; ++r) {
    float s = 0.f;
    KV_UNROLL for (int q = 0; q < N; ++q) s = fmaf(es.Ls[r][q], eps[q], s);
    es.z_own[r] = mus[r] + s;                                                    // (:349-351)
  }
  allgather<C::MEM, C::L, R>(g, es.z_own, vbuf, es.z);
  return ok;
}
// same, loading Sigma_s / mu_s at index bt
template <class C>
KV_FN bool elbo_sample_t(const Args& a, const Group<C::L, C::R>& g, TileRef xbuf, TileRef vbuf, long bt, const CholOpt& co,
                         const float (&eps)[C::N], ElboStep<C>& es) {
  constexpr int N = C::N, R = C::R;
  const int row0 = g.row0();
  float Ss[R][N], mus[R];
  KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_s + (bt * N + row0 + r) * N, Ss[r]);
  load_row<R>(a.mu_s + bt * N + row0, mus);
  return elbo_sample_rows<C>(g, xbuf, vbuf, Ss, mus, co, eps, es);
}

// ELBO terms of time steps [t0, t1) of sequence b (the steps are independent given z_{t0-1}, which is
// recomputed at the chunk start, so a sequence can be cut into chunks owned by different lane groups).
// acc: [0] transition  [1] emission  [2] init  [3] entropy  [4] sum(mask).  zbuf (nullable): z_t is stored
// there ([B,T,N]) for the adjoint.
template <class C>
KV_FN void elbo_sweep(const Args& a, const float* base, const FTiles<C>& tl, const Group<C::L, C::R>& g, int b, bool active,
                      float jitter, int t0, int t1, float* zbuf, double (&acc)[5]) {
  constexpr int N = C::N, P = C::P, M = C::M, R = C::R, L = C::L;
  constexpr bool MEM = C::MEM;
  const int row0 = g.row0();
  const int T = a.T;
  const CholOpt co = chol_opt(a, jitter);
  ElboConst<C> ec;
  int bad = elbo_const<C>(g, base, tl.nn(0), co, ec);
  bool ok = true, ok_s = true, ok_q = true;
  typename view_of<MEM, L, R, N>::type LQc_v = publish<MEM, L, R, N>(g, ec.LQ, tl.nn(3));
  float zprev[N];
  KV_UNROLL for (int j = 0; j < N; ++j) zprev[j] = 0.f;
  double s_tr = 0.0, s_em = 0.0, s_in = 0.0, s_en = 0.0, s_m = 0.0;
  if (t0 > 0) {
    const long btp = (long)b * T + (t0 - 1);
    float epsp[N];
    load_row<N>(a.eps + btp * N, epsp);
    ElboStep<C> esp;
    ok_s = elbo_sample_t<C>(a, g, tl.nn(0), tl.vec(0), btp, co, epsp, esp) && ok_s;
    KV_UNROLL for (int j = 0; j < N; ++j) zprev[j] = esp.z[j];
  }

  for (int t = t0; t < t1; ++t) {
    const long bt = (long)b * T + t;
    StepIn<C> in;
    load_step<C>(a, bt, in);
    float eps[N];
    load_row<N>(a.eps + bt * N, eps);
    ElboStep<C> es;
    ok_s = elbo_sample_t<C>(a, g, tl.nn(0), tl.vec(0), bt, co, eps, es) && ok_s;
    if (zbuf && active) store_row<R>(zbuf + bt * N + row0, es.z_own);

    // entropy = -log N(z; mu_s, Ls Ls^T) = 1/2 |eps|^2 + sum log Ls_ii + n/2 log 2pi            (:389)
    float e2 = 0.f;
    KV_UNROLL for (int j = 0; j < N; ++j) e2 = fmaf(eps[j], eps[j], e2);
    s_en += (double)(0.5f * e2 + logdet_half<C>(g, es.dg) + 0.5f * N * KV_LOG2PI);

    if (t == 0) {
      // log N(z_0; mu0, Sigma0)                                                                 (:380-381)
      float S0[R][N], L0[R][N], invd0[N], dg0[R];
      copy_rows<C, N>(base + Base<C>::oS0, row0, S0);
      ok = chol_dist<L, R>(g, S0, L0, invd0, dg0) && ok;
      auto L0_v = publish<MEM, L, R, N>(g, L0, tl.nn(1));
      float w[N];
      KV_UNROLL for (int j = 0; j < N; ++j) w[j] = es.z[j] - base[Base<C>::oMu0 + j];
      solve_vec_l<N>(w, L0_v, invd0);
      float q = 0.f;
      KV_UNROLL for (int j = 0; j < N; ++j) q = fmaf(w[j], w[j], q);
      s_in += (double)(-0.5f * (N * KV_LOG2PI + q) - logdet_half<C>(g, dg0));
    } else {
      // log N(z_t - A_t z_{t-1} - B_t u_t; 0, sym(Q_t)+jI)                                      (:353-369)
      float A[R][N], Bm[R][M];
      mix_A<C>(base, in.al, row0, A);
      mix_B<C>(base, in.al, row0, Bm);
      float x_own[R], x[N];
      KV_UNROLL for (int r = 0; r < R; ++r) {
        float s1 = 0.f, s2 = 0.f;
        KV_UNROLL for (int j = 0; j < N; ++j) s1 = fmaf(A[r][j], zprev[j], s1);
        KV_UNROLL for (int j = 0; j < M; ++j) s2 = fmaf(Bm[r][j], in.u[j], s2);
        x_own[r] = es.z_own[r] - (s1 + s2);
      }
      allgather<MEM, L, R>(g, x_own, tl.vec(1), x);
      float q = 0.f, ld;
      if constexpr (C::QPM) {
        float Q[R][N], Qs[R][N], LQ[R][N], invdQ[N], dgQ[R];
        mix_Q<C>(base, in.al, row0, Q);
        sym_jitter_rows<C>(g, Q, tl.nn(1), co.diag_q ? 0.f : co.jq, Qs);
        unsigned clq;
        ok_q = chol_dist_opt<L, R>(g, Qs, LQ, invdQ, dgQ, co.diag_q, clq) && ok_q;
        auto LQ_v = publish<MEM, L, R, N>(g, LQ, tl.nn(2));
        solve_vec_l<N>(x, LQ_v, invdQ);
        ld = logdet_half<C>(g, dgQ);
      } else {
        solve_vec_l<N>(x, LQc_v, ec.invdQ);
        ld = ec.logdetQ;
      }
      KV_UNROLL for (int j = 0; j < N; ++j) q = fmaf(x[j], x[j], q);
      s_tr += (double)(-0.5f * (N * KV_LOG2PI + q) - ld);
    }
    {
      // mask_t * log N(y_t - C_t z_t; 0, R)                                                     (:372-377)
      float Ct[R][P];
      mix_Ct<C>(base, in.al, row0, Ct);
      float e[P];
      KV_UNROLL for (int q = 0; q < P; ++q) {
        float s = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) s = fmaf(Ct[r][q], es.z_own[r], s);
        e[q] = s;
      }
      g.allreduce(e);
      KV_UNROLL for (int q = 0; q < P; ++q) e[q] = in.y[q] - e[q];
      RegView<P, P> LR_v{ec.LR};
      solve_vec_l<P>(e, LR_v, ec.invdR);
      float q2 = 0.f;
      KV_UNROLL for (int q = 0; q < P; ++q) q2 = fmaf(e[q], e[q], q2);
      s_em += (double)((-0.5f * (P * KV_LOG2PI + q2) - ec.logdetR) * in.m);
      s_m += (double)in.m;
    }
    KV_UNROLL for (int j = 0; j < N; ++j) zprev[j] = es.z[j];
  }
  if (active && g.lane == 0) {
    acc[0] += s_tr; acc[1] += s_em; acc[2] += s_in; acc[3] += s_en; acc[4] += s_m;
  }
  if (active) kv_info_or(a.info, bad | (ok ? 0 : KV_INFO_PIVOT) | (ok_s ? 0 : KV_INFO_CHOL_S) | (ok_q ? 0 : KV_INFO_CHOL_Q));
}

}  // namespace kvae
