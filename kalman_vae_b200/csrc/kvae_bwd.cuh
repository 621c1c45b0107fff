// kvae_bwd.cuh — explicit adjoint of the Kalman hot path (the reference relies on autograd,
// kvae/train/train.py:53).  One sequence per lane group, two sweeps:
//
//   sweep 3 (t = 0..T-1)  : A.3 ELBO adjoint + A.4 smoother adjoint.  Needs a one-step lookahead
//                           (z_{t+1}) and leaves (Sigma_f, mu_f, Sigma_p, mu_p)-bar in scratch.
//   sweep 4 (t = T-1..0)  : A.5 filter adjoint + A.0 mixing adjoint.
//
// K, S^-1, J, chol(Sigma_s), Q^-1 are recomputed from the saved public tensors.  Equation labels
// follow SURVEY.md Appendix A; oracle/adjoint.py is the same computation in torch.
#pragma once
#include "kvae_elbo.cuh"

namespace kvae {

struct BwdArgs {
  // dense cotangents of the nine smooth outputs (nullable)
  const float *c_mu_s, *c_Sig_s, *c_mu_f, *c_Sig_f, *c_mu_p, *c_Sig_p, *c_A, *c_B, *c_C;
  // outputs
  float *dY, *dU, *dalpha;          // dU nullable
  // scratch [B,T,...]
  float *w_Sig_f, *w_mu_f, *w_Sig_p, *w_mu_p;
  float *e_dSig, *e_dmu;            // elbo_only outputs [B,T,N,N], [B,T,N]
  float *terms_out;                 // with_elbo: the 8 ELBO terms (see include/kvae_kalman.h) are written here
  float c_elbo;                     // g_elbo / max(sum mask, 1); 0 = no ELBO term
  float jitter;
  int with_elbo;                    // 1: sweep 3 also accumulates the ELBO value sums (fused forward value: the
                                    //    separate ELBO kernel is not needed; c_elbo then excludes the normaliser)
  const float* mask_part;           // with_elbo: per-CTA mask sums written by the forward kernel (nullable)
  int n_mask_part;
  int raw_sums;                     // with_elbo: 1 = leave the gradients un-normalised (data parallel callers apply
                                    //    the GLOBAL 1/max(sum mask,1) after their all-reduce)
  int elbo_only;                    // 1: adjoint of the ELBO alone w.r.t. (mu, Sigma) given as the smoothed states:
                                    //    no smoother / filter adjoint; w_Sig_f / w_mu_f receive dSigma / dmu
};

// Per-warp tiles of the backward sweeps: 8 [n x n] + 3 [n x p] + 2 vector slots.
template <class C> using BTiles = TileSet<C::L, C::R, C::P, 7, 3, 2, C::MEM>;

// own entries of a replicated vector
template <class C> KV_FN void pick_own(const Group<C::L, C::R>& g, const float (&full)[C::N], float (&own)[C::R]) {
  if constexpr (C::L == 1) {
    KV_UNROLL for (int r = 0; r < C::R; ++r) own[r] = full[r];
  } else {
    KV_UNROLL for (int r = 0; r < C::R; ++r) {
      float v = 0.f;
      KV_UNROLL for (int j = 0; j < C::N; ++j) v = (j == g.row0() + r) ? full[j] : v;
      own[r] = v;
    }
  }
}
// own entries of X^T v for a fully visible X and replicated v
template <class C, class V> KV_FN void matTvec_own(const V& Xv, int row0, const float (&v)[C::N], float (&out)[C::R]) {
  KV_UNROLL for (int r = 0; r < C::R; ++r) {
    float s = 0.f;
    KV_UNROLL for (int k = 0; k < C::N; ++k) s = fmaf(Xv.at(k, row0 + r), v[k], s);
    out[r] = s;
  }
}
template <int R, int NC> KV_FN void add_rows(float (&dst)[R][NC], const float (&src)[R][NC]) {
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < NC; ++j) dst[r][j] += src[r][j];
}
template <int R, int NC> KV_FN void load_rows_opt(const float* p, long off, float (&dst)[R][NC]) {
  if (p) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<NC>(p + off + (long)r * NC, dst[r]); }
  else { KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < NC; ++j) dst[r][j] = 0.f; }
}
template <int R> KV_FN void load_vec_opt(const float* p, long off, float (&dst)[R]) {
  if (p) load_row<R>(p + off, dst);
  else { KV_UNROLL for (int r = 0; r < R; ++r) dst[r] = 0.f; }
}

// ---------------------------------------------------------------------------------------
// Parameter-gradient accumulators held in registers (A.0 adjoint):
//   dA_k += alpha_k Abar_t ... ; returns the partial <Abar_t, A_k> + ... for dalpha.
// ---------------------------------------------------------------------------------------
template <class C> struct GradAcc {
  static constexpr int K = C::K, R = C::R, N = C::N, M = C::M, P = C::P;
  static constexpr int KQ = C::QPM ? C::K : 0;
  static constexpr int oA = 0, oB = oA + K * R * N, oCt = oB + K * R * M, oQ = oCt + C::KC * R * P;
  static constexpr int count = oQ + KQ * R * N;
  // flat parameter layout of the reduced gradient: dA [K][N][N] | dB [K][N][M] | dC [K][P][N] | dQ [K][N][N]
  static constexpr int fA = 0, fB = K * N * N, fC = fB + K * N * M, fQ = fC + K * P * N;
  static constexpr int PSZ = fQ + KQ * N * N;
  // Small shapes: one register accumulator per (mode,row,col) element of the lane's rows (42 floats at n=4, L=4).
  // Large shapes (n=16: 392 floats per lane): DENSE — the lane stores its rows of the per-step cotangents
  // Abar_t/Bbar_t/Qbar_t(/Cbar_t^T) into dense scratch [B,T,..] (sweep 3 stores, sweep 4 adds) and a separate
  // contraction kernel forms dA_k = sum_{b,t} alpha_{b,t,k} Abar_{b,t} etc.; only a shared C keeps registers.
  // (A shared-memory atomic accumulator was measured 2.2x slower than even the spilling register version.)
  static constexpr bool DENSE = count > 128;
  static constexpr int nreg = DENSE ? (C::CSH ? R * P : 1) : count;
  float v[nreg];
  float *dnA, *dnB, *dnQ, *dnCt;   // DENSE: [B,T,N,N], [B,T,N,M], [B,T,N,N], [B,T,N,P] scratch
  bool on;                         // false for tail groups that mirror a valid sequence (they must not store)

  KV_FN void zero() {
    KV_UNROLL for (int i = 0; i < nreg; ++i) v[i] = 0.f;
  }
  // dal[k] += <Xbar, X_k> over own rows (caller all-reduces); acc_k += al[k] * Xbar  (or dense store / add)
  template <int COLS, int OFF, int MODES, int LD>
  KV_FN void one(const float* basek, float* dense, int row0, const float (&al)[K], const float (&Xb)[R][COLS], float (&dal)[K],
                 long bt, bool first) {
    KV_UNROLL for (int k = 0; k < MODES; ++k) {
      float s = 0.f;
      KV_UNROLL for (int r = 0; r < R; ++r) {
        float row[COLS];
        load_row<COLS>(basek + (k * N + row0 + r) * LD, row);
        s = kv_dot<COLS>(Xb[r], row, s);
        if constexpr (!DENSE) {
          KV_UNROLL for (int j = 0; j + 1 < COLS; j += 2)
            kv_fma2(v[OFF + (k * R + r) * COLS + j], v[OFF + (k * R + r) * COLS + j + 1], al[k], al[k], Xb[r][j], Xb[r][j + 1]);
          if constexpr (COLS % 2 == 1)
            v[OFF + (k * R + r) * COLS + COLS - 1] = fmaf(al[k], Xb[r][COLS - 1], v[OFF + (k * R + r) * COLS + COLS - 1]);
        }
      }
      dal[k] += s;
    }
    if constexpr (DENSE) {
      if (on) {
        KV_UNROLL for (int r = 0; r < R; ++r) {
          float* dst = dense + (bt * N + row0 + r) * COLS;
          if (first) store_row<COLS>(dst, Xb[r]);
          else {
            float cur[COLS];
            load_row<COLS>(dst, cur);
            KV_UNROLL for (int j = 0; j < COLS; ++j) cur[j] += Xb[r][j];
            store_row<COLS>(dst, cur);
          }
        }
      }
    } else { (void)dense; (void)bt; (void)first; }
  }
  KV_FN void addA(const float* base, int row0, const float (&al)[K], const float (&Xb)[R][N], float (&dal)[K], long bt, bool first) {
    one<N, oA, K, Base<C>::ldA>(base + Base<C>::oA, dnA, row0, al, Xb, dal, bt, first);
  }
  KV_FN void addB(const float* base, int row0, const float (&al)[K], const float (&Xb)[R][M], float (&dal)[K], long bt, bool first) {
    one<M, oB, K, Base<C>::ldB>(base + Base<C>::oB, dnB, row0, al, Xb, dal, bt, first);
  }
  KV_FN void addQ(const float* base, int row0, const float (&al)[K], const float (&Xb)[R][N], float (&dal)[K], long bt, bool first) {
    if constexpr (C::QPM) one<N, oQ, K, Base<C>::ldQ>(base + Base<C>::oQ, dnQ, row0, al, Xb, dal, bt, first);
  }
  KV_FN void addCt(const float* base, int row0, const float (&al)[K], const float (&Xb)[R][P], float (&dal)[K], long bt, bool first) {
    if constexpr (C::CSH) {
      constexpr int o = DENSE ? 0 : oCt;
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < P; ++j) v[o + r * P + j] += Xb[r][j];
    } else {
      one<P, oCt, K, Base<C>::ldCt>(base + Base<C>::oCt, dnCt, row0, al, Xb, dal, bt, first);
    }
  }
  // zero the dense A/B/Q cotangents of step index bt (t = 0 receives nothing from sweep 3)
  KV_FN void dense_zero_abq(int row0, long bt) {
    if constexpr (DENSE) {
      if (on) {
        float zn[N], zm[M];
        KV_UNROLL for (int j = 0; j < N; ++j) zn[j] = 0.f;
        KV_UNROLL for (int j = 0; j < M; ++j) zm[j] = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) {
          store_row<N>(dnA + (bt * N + row0 + r) * N, zn);
          store_row<M>(dnB + (bt * N + row0 + r) * M, zm);
          if constexpr (C::QPM) store_row<N>(dnQ + (bt * N + row0 + r) * N, zn);
        }
      }
    } else { (void)row0; (void)bt; }
  }
  // f(flat parameter index, value) for every register accumulator of the lane owning rows row0..
  template <class F> KV_FN void for_each(int row0, F&& f) const {
    if constexpr (!DENSE) {
      KV_UNROLL for (int k = 0; k < K; ++k) KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j)
        f(fA + (k * N + row0 + r) * N + j, v[oA + (k * R + r) * N + j]);
      KV_UNROLL for (int k = 0; k < K; ++k) KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < M; ++j)
        f(fB + (k * N + row0 + r) * M + j, v[oB + (k * R + r) * M + j]);
      KV_UNROLL for (int k = 0; k < C::KC; ++k) KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < P; ++j)
        f(fC + (k * P + j) * N + row0 + r, v[oCt + (k * R + r) * P + j]);
      if constexpr (C::QPM) {
        KV_UNROLL for (int k = 0; k < K; ++k) KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j)
          f(fQ + (k * N + row0 + r) * N + j, v[oQ + (k * R + r) * N + j]);
      }
    } else if constexpr (C::CSH) {
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < P; ++j) f(fC + j * N + row0 + r, v[r * P + j]);
    } else { (void)row0; (void)f; }
  }
  static constexpr int nreduce = DENSE ? (C::CSH ? R * P : 0) : count;   // register accumulators to reduce per CTA
};

// Everything sweep 3 reads from global memory for step t, fetched one step ahead (software prefetch):
// the sweep is latency bound at small batch, and these L2 round trips would otherwise sit on its critical path.
template <class C> struct S3In {
  StepIn<C> in;                       // y_t, u_t, alpha_t, mask_t
  float al1[C::K], u1[C::M];          // alpha_{t+1}, u_{t+1}
  float eps1[C::N];                   // eps_{t+1}
  float Ss1[C::R][C::N], ms1[C::R];   // Sigma_s, mu_s at t+1
  float Sf[C::R][C::N];               // Sigma_f at t
  float Sp1[C::R][C::N], mp1[C::R];   // Sigma_p, mu_p at t+1
};
template <class C> KV_FN void load_s3(const Args& a, long bt, bool has_next, bool has_elbo, int row0, S3In<C>& s) {
  constexpr int N = C::N, M = C::M, K = C::K, R = C::R;
  load_step<C>(a, bt, s.in);
  KV_UNROLL for (int k = 0; k < K; ++k) s.al1[k] = 0.f;
  KV_UNROLL for (int j = 0; j < M; ++j) s.u1[j] = 0.f;
  KV_UNROLL for (int j = 0; j < N; ++j) s.eps1[j] = 0.f;
  if (has_next) {
    load_row<K>(a.alpha + (bt + 1) * K, s.al1);
    if (a.U) load_row<M>(a.U + (bt + 1) * M, s.u1);
    if (has_elbo) load_row<N>(a.eps + (bt + 1) * N, s.eps1);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      load_row<N>(a.Sig_f + (bt * N + row0 + r) * N, s.Sf[r]);
      load_row<N>(a.Sig_p + ((bt + 1) * N + row0 + r) * N, s.Sp1[r]);
      load_row<N>(a.Sig_s + ((bt + 1) * N + row0 + r) * N, s.Ss1[r]);
    }
    load_row<R>(a.mu_s + (bt + 1) * N + row0, s.ms1);
    load_row<R>(a.mu_p + (bt + 1) * N + row0, s.mp1);
  }
}

// ---------------------------------------------------------------------------------------
// sweep 3: ELBO adjoint (A.3) + smoother adjoint (A.4), forward in time
// ---------------------------------------------------------------------------------------
template <class C>
KV_FN void bwd_sweep3(const Args& a, const BwdArgs& w, const float* base, const BTiles<C>& tl, const Group<C::L, C::R>& g,
                      int b, bool active, GradAcc<C>& acc, double (&el)[5]) {
  constexpr int N = C::N, P = C::P, M = C::M, R = C::R, L = C::L, K = C::K;
  constexpr bool MEM = C::MEM;
  const TileRef T0 = tl.nn(0), T1 = tl.nn(1), T2 = tl.nn(2), T3 = tl.nn(3), T4 = tl.nn(4), T5 = tl.nn(5), TP = tl.nn(6);
  const TileRef VB = tl.vec(0), VB2 = tl.vec(1);
  // elbo_sample / sym_jitter_rows use Tiles<C> offsets oX0 (= T0) and oV (remapped below)
  const int row0 = g.row0();
  const int T = a.T;
  const float c = w.c_elbo;
  const bool has_elbo = (c != 0.f);
  bool ok = true, ok_s = true, ok_q = true;
  const CholOpt co = chol_opt(a, w.jitter);
  int bad0 = 0;

  ElboConst<C> ec;
  if (has_elbo) bad0 = elbo_const<C>(g, base, T0, co, ec);
  else {
    KV_UNROLL for (int q = 0; q < P; ++q) { ec.invdR[q] = 0.f; KV_UNROLL for (int q2 = 0; q2 < P; ++q2) ec.LR[q][q2] = 0.f; }
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) ec.LQ[r][j] = 0.f;
    KV_UNROLL for (int j = 0; j < N; ++j) ec.invdQ[j] = 0.f;
  }
  typename view_of<MEM, L, R, N>::type LQc_v = publish<MEM, L, R, N>(g, ec.LQ, TP);

  // state carried over iterations
  ElboStep<C> es;                    // sample at t
  float xbar_own[R];                 // xbar_t (own entries), zero at t = 0
  float Ssb[R][N], msb[R];           // (Sigma_s, mu_s)-bar at t: cotangent + D-bar / d-bar from t-1
  float dal_carry[K];                // A/B/Q part of dalpha_t (computed one step ahead)
  KV_UNROLL for (int r = 0; r < R; ++r) xbar_own[r] = 0.f;
  KV_UNROLL for (int k = 0; k < K; ++k) dal_carry[k] = 0.f;
  {
    const long bt0 = (long)b * T;
    load_rows_opt<R, N>(w.c_Sig_s, (bt0 * N + row0) * N, Ssb);
    load_vec_opt<R>(w.c_mu_s, bt0 * N + row0, msb);
    if (has_elbo) {
      float eps0[N];
      load_row<N>(a.eps + bt0 * N, eps0);
      ok_s = elbo_sample_t<C>(a, g, T0, VB, bt0, co, eps0, es) && ok_s;
    }
    // (Sigma_p, mu_p)-bar at t = 0 get no smoother contribution
    if (active) {
      float z0[R][N], v0[R];
      load_rows_opt<R, N>(w.c_Sig_p, (bt0 * N + row0) * N, z0);
      load_vec_opt<R>(w.c_mu_p, bt0 * N + row0, v0);
      KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(w.w_Sig_p + (bt0 * N + row0 + r) * N, z0[r]);
      store_row<R>(w.w_mu_p + bt0 * N + row0, v0);
      if (w.dU && g.lane == 0) { float zu[M]; KV_UNROLL for (int j = 0; j < M; ++j) zu[j] = 0.f; store_row<M>(w.dU + bt0 * M, zu); }
    }
  }

  // software prefetch only for small states: for n = 16 the second copy of the per-step inputs would spill
  constexpr bool PF = (C::N <= 8);
  S3In<C> pf;
  if constexpr (PF) load_s3<C>(a, (long)b * T, T > 1, has_elbo, row0, pf);
  float eps_cur[N];    // eps_t (the previous step's eps_{t+1})
  KV_UNROLL for (int j = 0; j < N; ++j) eps_cur[j] = 0.f;
  if (has_elbo) load_row<N>(a.eps + (long)b * T * N, eps_cur);
  for (int t = 0; t < T; ++t) {
    const long bt = (long)b * T + t;
    const bool has_next = (t + 1 < T);
    if constexpr (!PF) load_s3<C>(a, bt, has_next, has_elbo, row0, pf);
    const S3In<C> cu = pf;                                   // this step's inputs (PF: fetched one step ago)
    const StepIn<C>& in = cu.in;
    float al1[K], u1[M];
    KV_UNROLL for (int k = 0; k < K; ++k) al1[k] = cu.al1[k];
    // u_{t+1} is first used deep inside the step; "+ 0" makes this the first consumer, so the scoreboard wait of its
    // (long finished) load happens here and not after the next prefetch has been issued on the same scoreboard
    KV_UNROLL for (int j = 0; j < M; ++j) u1[j] = cu.u1[j] + 0.0f;
    float A1[R][N];
    mix_A<C>(base, al1, row0, A1);
    // prefetch of the next step, issued after the mixing loads (see smoother_sweep)
    if constexpr (PF) { if (has_next) load_s3<C>(a, bt + 1, t + 2 < T, has_elbo, row0, pf); }
    float Ab[R][N], Bb[R][M], Qb[R][N], Ctb[R][P];   // (A,B,Q)-bar at t+1 and C^T-bar at t
    if (has_next) {
      load_rows_opt<R, N>(w.c_A, ((bt + 1) * N + row0) * N, Ab);
      load_rows_opt<R, M>(w.c_B, ((bt + 1) * N + row0) * M, Bb);
    } else {
      KV_UNROLL for (int r = 0; r < R; ++r) { KV_UNROLL for (int j = 0; j < N; ++j) Ab[r][j] = 0.f; KV_UNROLL for (int j = 0; j < M; ++j) Bb[r][j] = 0.f; }
    }
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Qb[r][j] = 0.f;
    if (w.c_C) {
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Ctb[r][q] = w.c_C[(bt * P + q) * N + row0 + r];
    } else {
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Ctb[r][q] = 0.f;
    }
    float dy[P];
    KV_UNROLL for (int q = 0; q < P; ++q) dy[q] = 0.f;
    float du1[M];
    KV_UNROLL for (int j = 0; j < M; ++j) du1[j] = 0.f;
    float xbar_next_own[R];
    KV_UNROLL for (int r = 0; r < R; ++r) xbar_next_own[r] = 0.f;
    ElboStep<C> es1;  // sample at t+1

    // ---------------------------------------------------------------- A.3 ELBO adjoint at t
    if (has_elbo) {
      float zbar_own[R];
      KV_UNROLL for (int r = 0; r < R; ++r) zbar_own[r] = xbar_own[r];
      float v_tr = 0.f, v_em = 0.f, v_in = 0.f, v_en = 0.f;   // this lane's share of the ELBO value terms of step t
      if (has_next) {
        ok_s = elbo_sample_rows<C>(g, T0, VB, cu.Ss1, cu.ms1, co, cu.eps1, es1) && ok_s;
        // x_{t+1} = z_{t+1} - A1 z_t - B1 u_{t+1};  q = Qj^-1 x;  xbar = -c q
        float B1[R][M];
        mix_B<C>(base, al1, row0, B1);
        float x_own[R], x[N];
        KV_UNROLL for (int r = 0; r < R; ++r) {
          float s1 = 0.f, s2 = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) s1 = fmaf(A1[r][j], es.z[j], s1);
          KV_UNROLL for (int j = 0; j < M; ++j) s2 = fmaf(B1[r][j], u1[j], s2);
          x_own[r] = es1.z_own[r] - (s1 + s2);
        }
        allgather<MEM, L, R>(g, x_own, VB2, x);
        float xbar[N];
        if constexpr (C::QPM) {
          float Q1[R][N], Qs[R][N], LQ[R][N], invdQ[N], dgQ[R];
          mix_Q<C>(base, al1, row0, Q1);
          sym_jitter_rows<C>(g, Q1, T1, co.diag_q ? 0.f : co.jq, Qs);
          unsigned clq;
          ok_q = chol_dist_opt<L, R>(g, Qs, LQ, invdQ, dgQ, co.diag_q, clq) && ok_q;
          auto LQ_v = publish<MEM, L, R, N>(g, LQ, T2);
          solve_vec_l<N>(x, LQ_v, invdQ);
          if (w.with_elbo) {   // log N(x; 0, Qj) = -1/2 (n log 2pi + |LQ^-1 x|^2) - sum log diag LQ   (own-lane log terms)
            float q2 = 0.f, ld = 0.f;
            KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(x[j], x[j], q2);
            KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(dgQ[r]);
            v_tr = (g.lane == 0 ? -0.5f * (N * KV_LOG2PI + q2) : 0.f) - ld;
          }
          solve_vec_lt<N>(x, LQ_v, invdQ);        // x := q
          // Qbar_{t+1} += c/2 (q q^T - Qj^-1)
          float Qi[R][N], q_own[R];
          KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Qi[r][j] = (row0 + r == j) ? 1.f : 0.f;
          solve_rows_llt<R, N>(Qi, LQ_v, invdQ);
          pick_own<C>(g, x, q_own);
          KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) {
            // diagonal fallback: only the unclamped diagonal entries of sym(Q_t) reach L (kalman_filter.py:298-302)
            const bool live = !co.diag_q || (row0 + r == j && !((clq >> j) & 1u));
            Qb[r][j] += live ? 0.5f * c * (q_own[r] * x[j] - Qi[r][j]) : 0.f;
          }
        } else {
          solve_vec_l<N>(x, LQc_v, ec.invdQ);
          if (w.with_elbo && g.lane == 0) {
            float q2 = 0.f;
            KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(x[j], x[j], q2);
            v_tr = -0.5f * (N * KV_LOG2PI + q2) - ec.logdetQ;
          }
          solve_vec_lt<N>(x, LQc_v, ec.invdQ);
        }
        KV_UNROLL for (int j = 0; j < N; ++j) xbar[j] = -c * x[j];
        pick_own<C>(g, xbar, xbar_next_own);
        // zbar_t += -A1^T xbar ; Abar_{t+1} += -xbar z_t^T ; Bbar_{t+1} += -xbar u^T ; ubar_{t+1} += -B1^T xbar
        {
          auto A1_v = publish<MEM, L, R, N>(g, A1, T1);
          float av[R];
          matTvec_own<C>(A1_v, row0, xbar, av);
          KV_UNROLL for (int r = 0; r < R; ++r) zbar_own[r] -= av[r];
        }
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) Ab[r][j] = fmaf(-xbar_next_own[r], es.z[j], Ab[r][j]);
          KV_UNROLL for (int j = 0; j < M; ++j) Bb[r][j] = fmaf(-xbar_next_own[r], u1[j], Bb[r][j]);
          KV_UNROLL for (int j = 0; j < M; ++j) du1[j] = fmaf(-B1[r][j], xbar_next_own[r], du1[j]);
        }
      }
      // emission at t
      {
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
        if (w.with_elbo && g.lane == 0) {
          float q2 = 0.f;
          KV_UNROLL for (int q = 0; q < P; ++q) q2 = fmaf(e[q], e[q], q2);
          v_em = (-0.5f * (P * KV_LOG2PI + q2) - ec.logdetR) * in.m;
        }
        solve_vec_lt<P>(e, LR_v, ec.invdR);       // R^-1 e
        KV_UNROLL for (int q = 0; q < P; ++q) { e[q] = -c * in.m * e[q]; dy[q] += e[q]; }   // ebar
        KV_UNROLL for (int r = 0; r < R; ++r) {
          float s = 0.f;
          KV_UNROLL for (int q = 0; q < P; ++q) {
            Ctb[r][q] = fmaf(-es.z_own[r], e[q], Ctb[r][q]);
            s = fmaf(Ct[r][q], e[q], s);
          }
          zbar_own[r] -= s;
        }
      }
      if (t == 0) {  // zbar_0 += -c Sigma0^-1 (z_0 - mu0)
        float S0[R][N], L0[R][N], invd0[N], dg0[R];
        copy_rows<C, N>(base + Base<C>::oS0, row0, S0);
        ok = chol_dist<L, R>(g, S0, L0, invd0, dg0) && ok;
        auto L0_v = publish<MEM, L, R, N>(g, L0, T1);
        float wv[N], w_own[R];
        KV_UNROLL for (int j = 0; j < N; ++j) wv[j] = es.z[j] - base[Base<C>::oMu0 + j];
        solve_vec_l<N>(wv, L0_v, invd0);
        if (w.with_elbo) {
          float q2 = 0.f, ld = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(wv[j], wv[j], q2);
          KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(dg0[r]);
          v_in = (g.lane == 0 ? -0.5f * (N * KV_LOG2PI + q2) : 0.f) - ld;
        }
        solve_vec_lt<N>(wv, L0_v, invd0);
        pick_own<C>(g, wv, w_own);
        KV_UNROLL for (int r = 0; r < R; ++r) zbar_own[r] = fmaf(-c, w_own[r], zbar_own[r]);
      }
      if (w.with_elbo) {   // entropy of step t, then add this step's terms to the lane's fp64 sums
        float ld = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(es.dg[r]);
        if (g.lane == 0) {
          float e2 = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) e2 = fmaf(eps_cur[j], eps_cur[j], e2);
          v_en = 0.5f * e2 + 0.5f * N * KV_LOG2PI;
        }
        v_en += ld;
        if (active) {
          el[0] += (double)v_tr; el[1] += (double)v_em; el[2] += (double)v_in; el[3] += (double)v_en;
          if (g.lane == 0) el[4] += (double)in.m;
        }
      }
      // mu_s-bar += zbar ; Sigma_s-bar += sym(Ls^-T Phi Ls^-1), Phi = tril_strict(v eps^T) + diag((v.eps + c)/2), v = Ls^T zbar
      KV_UNROLL for (int r = 0; r < R; ++r) msb[r] += zbar_own[r];
      {
        float zbar[N], eps[N];
        allgather<MEM, L, R>(g, zbar_own, VB2, zbar);
        KV_UNROLL for (int j = 0; j < N; ++j) eps[j] = eps_cur[j];
        auto Ls_v = publish<MEM, L, R, N>(g, es.Ls, T1);
        float v_own[R];
        matTvec_own<C>(Ls_v, row0, zbar, v_own);
        float Phi[R][N];
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) {
          const int i = row0 + r;
          const float ve = v_own[r] * eps[j];
          // (diagonal fallback: L depends on the unclamped diagonal entries of sym(Sigma_s) only)
          const bool live_d = !co.diag_s || !((es.clamped >> i) & 1u);
          Phi[r][j] = (j < i) ? (co.diag_s ? 0.f : ve) : ((j == i && live_d) ? 0.5f * (ve + c) : 0.f);
        }
        solve_rows_l<R, N>(Phi, Ls_v, es.invd);                       // Z = Phi Ls^-1
        float Zt[R][N];
        {
          auto Z_v = publish<MEM, L, R, N>(g, Phi, T2);
          tr_rows<R, N>(Z_v, row0, Zt);
        }
        solve_rows_l<R, N>(Zt, Ls_v, es.invd);                        // Y^T = Z^T Ls^-1
        float Yr[R][N];
        {
          auto Yt_v = publish<MEM, L, R, N>(g, Zt, T3);
          tr_rows<R, N>(Yt_v, row0, Yr);
        }
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Ssb[r][j] += 0.5f * (Zt[r][j] + Yr[r][j]);
      }
    }

    // ---------------------------------------------------------------- A.4 smoother adjoint at t
    float Sfb[R][N], mfb[R];
    load_rows_opt<R, N>(w.c_Sig_f, (bt * N + row0) * N, Sfb);
    load_vec_opt<R>(w.c_mu_f, bt * N + row0, mfb);
    float Ssb1[R][N], msb1[R];
    if (has_next && !w.elbo_only) {
      const float (&Sf)[R][N] = cu.Sf;
      const float (&Sp1)[R][N] = cu.Sp1;
      const float (&Ss1)[R][N] = cu.Ss1;
      const float (&ms1)[R] = cu.ms1;
      const float (&mp1)[R] = cu.mp1;
      float J[R][N], LU[R][N], invu[N];
      ok = smoother_gain<C>(g, T0, T1, Sf, A1, Sp1, J, LU, invu) && ok;        // A1_v in T0, LU_v in T1
      typename view_of<MEM, L, R, N>::type A1_v, LU_v;
      if constexpr (MEM) { A1_v = MemView<L, R, N>{T0.p, T0.g}; LU_v = MemView<L, R, N>{T1.p, T1.g}; }
      else { A1_v = RegView<N, N>{A1}; LU_v = RegView<N, N>{LU}; }
      float D[R][N], d_own[R], d[N];
      KV_UNROLL for (int r = 0; r < R; ++r) {
        KV_UNROLL for (int j = 0; j < N; ++j) D[r][j] = Ss1[r][j] - Sp1[r][j];
        d_own[r] = ms1[r] - mp1[r];
      }
      allgather<MEM, L, R>(g, d_own, VB, d);
      float Gs[R][N];
      {
        auto S_v = publish<MEM, L, R, N>(g, Ssb, T2);
        float St[R][N];
        tr_rows<R, N>(S_v, row0, St);
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Gs[r][j] = 0.5f * (Ssb[r][j] + St[r][j]);
      }
      add_rows<R, N>(Sfb, Gs);
      KV_UNROLL for (int r = 0; r < R; ++r) mfb[r] += msb[r];
      auto J_v = publish<MEM, L, R, N>(g, J, T3);
      float GJ[R][N];
      mm_RS<false>(Gs, J_v, GJ);                                  // Gs J
      float Jb[R][N];
      {
        auto D_v = publish<MEM, L, R, N>(g, D, T4);
        mm_RSt<false>(GJ, D_v, Jb);                               // Gs J D^T
        mm_RS<true>(GJ, D_v, Jb);                                 // + Gs^T J D  (Gs is exactly symmetric)
      }
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Jb[r][j] = fmaf(msb[r], d[j], Jb[r][j]);
      float Db[R][N], db[R];
      {
        auto GJ_v = publish<MEM, L, R, N>(g, GJ, T5);
        mm_StS<false, R, N, N>(J_v, row0, GJ_v, Db);              // D-bar = J^T Gs J
        float msb_full[N];
        allgather<MEM, L, R>(g, msb, VB2, msb_full);
        matTvec_own<C>(J_v, row0, msb_full, db);                  // d-bar = J^T mu_s-bar
      }
      load_rows_opt<R, N>(w.c_Sig_s, ((bt + 1) * N + row0) * N, Ssb1);
      load_vec_opt<R>(w.c_mu_s, (bt + 1) * N + row0, msb1);
      add_rows<R, N>(Ssb1, Db);
      KV_UNROLL for (int r = 0; r < R; ++r) msb1[r] += db[r];
      float Spb1[R][N], mpb1[R];
      load_rows_opt<R, N>(w.c_Sig_p, ((bt + 1) * N + row0) * N, Spb1);
      load_vec_opt<R>(w.c_mu_p, (bt + 1) * N + row0, mpb1);
      solve_rows_lut<R, N>(Jb, LU_v, invu);                       // W-bar = J-bar Sp1^-T   (in place)
      {
        auto Wb_v = publish<MEM, L, R, N>(g, Jb, T4);
        float JW[R][N];
        mm_StS<false, R, N, N>(J_v, row0, Wb_v, JW);              // J^T W-bar
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) Spb1[r][j] -= Db[r][j] + JW[r][j];
          mpb1[r] -= db[r];
        }
        mm_RS<true>(Jb, A1_v, Sfb);                               // Sigma_f-bar_t += W-bar A1
        auto Sf_v = publish<MEM, L, R, N>(g, Sf, T5);
        mm_StS<true, R, N, N>(Wb_v, row0, Sf_v, Ab);              // A-bar_{t+1} += W-bar^T Sigma_f
      }
      if (active) {
        KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(w.w_Sig_p + ((bt + 1) * N + row0 + r) * N, Spb1[r]);
        store_row<R>(w.w_mu_p + (bt + 1) * N + row0, mpb1);
      }
    } else {
      add_rows<R, N>(Sfb, Ssb);                                   // T-1: copied, not symmetrised
      KV_UNROLL for (int r = 0; r < R; ++r) mfb[r] += msb[r];
      KV_UNROLL for (int r = 0; r < R; ++r) { msb1[r] = 0.f; KV_UNROLL for (int j = 0; j < N; ++j) Ssb1[r][j] = 0.f; }
    }
    if (active) {
      KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(w.w_Sig_f + (bt * N + row0 + r) * N, Sfb[r]);
      store_row<R>(w.w_mu_f + bt * N + row0, mfb);
    }

    // ---------------------------------------------------------------- A.0 adjoint of this sweep's parts
    float dal_next[K], dal_c[K];
    KV_UNROLL for (int k = 0; k < K; ++k) { dal_next[k] = 0.f; dal_c[k] = 0.f; }
    if (has_next) {
      acc.addA(base, row0, al1, Ab, dal_next, bt + 1, true);
      acc.addB(base, row0, al1, Bb, dal_next, bt + 1, true);
      acc.addQ(base, row0, al1, Qb, dal_next, bt + 1, true);
    }
    if (t == 0) acc.dense_zero_abq(row0, bt);
    acc.addCt(base, row0, in.al, Ctb, dal_c, bt, true);
    float red[2 * K + M];
    KV_UNROLL for (int k = 0; k < K; ++k) { red[k] = dal_next[k]; red[K + k] = dal_c[k]; }
    KV_UNROLL for (int j = 0; j < M; ++j) red[2 * K + j] = du1[j];
    g.allreduce(red);
    if (active && g.lane == 0) {
      float da[K];
      KV_UNROLL for (int k = 0; k < K; ++k) da[k] = dal_carry[k] + red[K + k];
      store_row<K>(w.dalpha + bt * K, da);
      store_row<P>(w.dY + bt * P, dy);
      if (w.dU && has_next) {
        float duv[M];
        KV_UNROLL for (int j = 0; j < M; ++j) duv[j] = red[2 * K + j];
        store_row<M>(w.dU + (bt + 1) * M, duv);
      }
    }
    // carry
    KV_UNROLL for (int k = 0; k < K; ++k) dal_carry[k] = red[k];
    KV_UNROLL for (int r = 0; r < R; ++r) {
      xbar_own[r] = xbar_next_own[r];
      msb[r] = msb1[r];
      KV_UNROLL for (int j = 0; j < N; ++j) Ssb[r][j] = Ssb1[r][j];
    }
    if (has_elbo && has_next) es = es1;
    KV_UNROLL for (int j = 0; j < N; ++j) eps_cur[j] = cu.eps1[j];
  }
  if (active) kv_info_or(a.info, bad0 | (ok ? 0 : KV_INFO_PIVOT) | (ok_s ? 0 : KV_INFO_CHOL_S) | (ok_q ? 0 : KV_INFO_CHOL_Q));
}

// What sweep 4 reads from global memory for step t (prefetched one step ahead, as in sweep 3)
template <class C> struct S4In {
  StepIn<C> in;
  float Sfb[C::R][C::N], mfb[C::R], Spb[C::R][C::N], mpb[C::R];   // scratch left by sweep 3
  float Sp[C::R][C::N], mup[C::R];                               // Sigma_p, mu_p at t
  float Sprev[C::R][C::N], muprev[C::N];                         // filtered belief at t-1 (or the initial one)
};
template <class C>
KV_FN void load_s4(const Args& a, const BwdArgs& w, const float* base, int b, int t, int row0, S4In<C>& s) {
  constexpr int N = C::N, R = C::R;
  const long bt = (long)b * a.T + t;
  load_step<C>(a, bt, s.in);
  KV_UNROLL for (int r = 0; r < R; ++r) {
    load_row<N>(w.w_Sig_f + (bt * N + row0 + r) * N, s.Sfb[r]);
    load_row<N>(w.w_Sig_p + (bt * N + row0 + r) * N, s.Spb[r]);
    load_row<N>(a.Sig_p + (bt * N + row0 + r) * N, s.Sp[r]);
  }
  load_row<R>(w.w_mu_f + bt * N + row0, s.mfb);
  load_row<R>(w.w_mu_p + bt * N + row0, s.mpb);
  load_row<R>(a.mu_p + bt * N + row0, s.mup);
  if (t > 0) {
    KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_f + ((bt - 1) * N + row0 + r) * N, s.Sprev[r]);
    load_row<N>(a.mu_f + (bt - 1) * N, s.muprev);
  } else {
    if (a.Sig_init) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_init + ((long)b * N + row0 + r) * N, s.Sprev[r]); }
    else copy_rows<C, N>(base + Base<C>::oS0, row0, s.Sprev);
    if (a.mu_init) load_row<N>(a.mu_init + (long)b * N, s.muprev);
    else load_row<N>(base + Base<C>::oMu0, s.muprev);
  }
}

// ---------------------------------------------------------------------------------------
// sweep 4: filter adjoint (A.5) + mixing adjoint (A.0), backward in time
// ---------------------------------------------------------------------------------------
template <class C>
KV_FN void bwd_sweep4(const Args& a, const BwdArgs& w, const float* base, const BTiles<C>& tl, const Group<C::L, C::R>& g,
                      int b, bool active, GradAcc<C>& acc) {
  constexpr int N = C::N, P = C::P, M = C::M, R = C::R, L = C::L, K = C::K;
  constexpr bool MEM = C::MEM;
  const TileRef T0 = tl.nn(0), T1 = tl.nn(1), T2 = tl.nn(2), T3 = tl.nn(3), T4 = tl.nn(4), T5 = tl.nn(5);
  const TileRef CB = tl.np(0), KB = tl.np(1), PB = tl.np(2), VB = tl.vec(0);
  const int row0 = g.row0();
  const int T = a.T;
  bool ok = true;
  const float* Rm = base + Base<C>::oR;

  float Sf_carry[R][N], mf_carry[R];
  KV_UNROLL for (int r = 0; r < R; ++r) { mf_carry[r] = 0.f; KV_UNROLL for (int j = 0; j < N; ++j) Sf_carry[r][j] = 0.f; }

  constexpr bool PF = (C::N <= 8);
  S4In<C> pf;
  if constexpr (PF) load_s4<C>(a, w, base, b, T - 1, row0, pf);
  for (int t = T - 1; t >= 0; --t) {
    const long bt = (long)b * T + t;
    if constexpr (!PF) load_s4<C>(a, w, base, b, t, row0, pf);
    const S4In<C> cu = pf;                                   // this step's inputs (PF: fetched one step ago)
    const StepIn<C>& in = cu.in;
    // the partial dalpha_t / dY_t / dU_t left by sweep 3 are added to at the END of this step: fetch them now so that
    // the L2 round trip overlaps the step instead of stalling the (in-order) warp right before the stores
    float da_old[K], dy_old[P], du_old[M];
    KV_UNROLL for (int k = 0; k < K; ++k) da_old[k] = 0.f;
    KV_UNROLL for (int q = 0; q < P; ++q) dy_old[q] = 0.f;
    KV_UNROLL for (int j = 0; j < M; ++j) du_old[j] = 0.f;
    if (active && g.lane == 0) {
      load_row<K>(w.dalpha + bt * K, da_old);
      load_row<P>(w.dY + bt * P, dy_old);
      if (w.dU) load_row<M>(w.dU + bt * M, du_old);
    }
    float Sfb[R][N], mfb[R], Spb[R][N], mpb[R];
    KV_UNROLL for (int r = 0; r < R; ++r) {
      KV_UNROLL for (int j = 0; j < N; ++j) { Sfb[r][j] = cu.Sfb[r][j] + Sf_carry[r][j]; Spb[r][j] = cu.Spb[r][j]; }
      mfb[r] = cu.mfb[r] + mf_carry[r];
      mpb[r] = cu.mpb[r];
    }
    const float (&Sp)[R][N] = cu.Sp;
    const float (&mup)[R] = cu.mup;
    const float (&Sprev)[R][N] = cu.Sprev;
    const float (&muprev)[N] = cu.muprev;
    float A[R][N], Bm[R][M], Ct[R][P];
    mix_A<C>(base, in.al, row0, A);
    mix_B<C>(base, in.al, row0, Bm);
    mix_Ct<C>(base, in.al, row0, Ct);
    // prefetch of the next step, issued after the mixing loads (see smoother_sweep)
    if constexpr (PF) { if (t > 0) load_s4<C>(a, w, base, b, t - 1, row0, pf); }

    // recompute the gain
    auto Ct_v = publish<MEM, L, R, P>(g, Ct, CB);
    GainOut<C> go;
    ok = gain<C>(g, base, Sp, mup, Ct, Ct_v, in.y, in.m, go) && ok;
    float G[R][N];
    mm_RSt<false>(go.Kg, Ct_v, G);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) G[r][j] = ((row0 + r == j) ? 1.0f : 0.0f) - G[r][j];

    // Gf = sym(Sigma_f-bar)
    float Gf[R][N];
    {
      auto S_v = publish<MEM, L, R, N>(g, Sfb, T0);
      float St[R][N];
      tr_rows<R, N>(S_v, row0, St);
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Gf[r][j] = 0.5f * (Sfb[r][j] + St[r][j]);
    }
    auto G_v = publish<MEM, L, R, N>(g, G, T1);
    auto Sp_v = publish<MEM, L, R, N>(g, Sp, T2);
    float GfG[R][N];
    mm_RS<false>(Gf, G_v, GfG);                                   // Gf G
    float Gb[R][N];
    mm_RSt<false>(GfG, Sp_v, Gb);                                 // G-bar = Gf G Sp^T + Gf^T G Sp
    mm_RS<true>(GfG, Sp_v, Gb);
    {
      auto GfG_v = publish<MEM, L, R, N>(g, GfG, T3);
      mm_StS<true, R, N, N>(G_v, row0, GfG_v, Spb);               // Sp-bar' = Sp-bar + G^T Gf G
    }
    // K-bar = Gf K (R^T + R) - G-bar C^T + mu_f-bar r^T
    auto K_v = publish<MEM, L, R, P>(g, go.Kg, KB);
    float GfK[R][P];
    mm_RS<false>(Gf, K_v, GfK);
    float Kb[R][P];
    mm_RS<false>(Gb, Ct_v, Kb);                                   // G-bar C^T
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
      float s = 0.f;
      KV_UNROLL for (int q2 = 0; q2 < P; ++q2) s = fmaf(GfK[r][q2], Rm[q * P + q2] + Rm[q2 * P + q], s);
      Kb[r][q] = s - Kb[r][q] + mfb[r] * go.r[q];
    }
    // C^T-bar (own rows) : -G-bar^T K + P Sb^T + (Sp^T C^T) Sb + Sp^T P-bar - mu_p r-bar^T
    float Ctb[R][P];
    {
      auto Gb_v = publish<MEM, L, R, N>(g, Gb, T4);
      mm_StS<false, R, N, P>(Gb_v, row0, K_v, Ctb);
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Ctb[r][q] = -Ctb[r][q];
    }
    // r-bar = K^T mu_f-bar ; K0-bar^T K0 (both reduced over the group)
    float K0b[R][P];
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) K0b[r][q] = in.m * Kb[r][q];
    float red[P + P * P];
    KV_UNROLL for (int q = 0; q < P; ++q) {
      float s = 0.f;
      KV_UNROLL for (int r = 0; r < R; ++r) s = fmaf(go.Kg[r][q], mfb[r], s);
      red[q] = s;
      KV_UNROLL for (int q2 = 0; q2 < P; ++q2) {
        float s2 = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) s2 = fmaf(K0b[r][q], go.K0[r][q2], s2);
        red[P + q * P + q2] = s2;                                 // M1[q][q2] = (K0-bar^T K0)[q][q2]
      }
    }
    g.allreduce(red);
    float rb[P];
    KV_UNROLL for (int q = 0; q < P; ++q) rb[q] = red[q];
    // S-bar = sym(-S^-1 M1):  (S^-1 M1)^T = M1^T S^-1
    float Sb[P][P];
    {
      float M1t[P][P];
      KV_UNROLL for (int q = 0; q < P; ++q) KV_UNROLL for (int q2 = 0; q2 < P; ++q2) M1t[q][q2] = red[P + q2 * P + q];
      RegView<P, P> Lc_v{go.Lc};
      solve_rows_llt<P, P>(M1t, Lc_v, go.invd);                   // M1t[q2][q] = (S^-1 M1)[q][q2]
      KV_UNROLL for (int q = 0; q < P; ++q) KV_UNROLL for (int q2 = 0; q2 < P; ++q2) Sb[q][q2] = -0.5f * (M1t[q2][q] + M1t[q][q2]);
    }
    // P-bar = K0-bar S^-1
    float Pb[R][P];
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Pb[r][q] = K0b[r][q];
    {
      RegView<P, P> Lc_v{go.Lc};
      solve_rows_llt<R, P>(Pb, Lc_v, go.invd);
    }
    {
      float P2[R][P];
      mm_StS<false, R, N, P>(Sp_v, row0, Ct_v, P2);              // Sp^T C^T
      auto Pb_v = publish<MEM, L, R, P>(g, Pb, PB);
      mm_StS<true, R, N, P>(Sp_v, row0, Pb_v, Ctb);              // + Sp^T P-bar
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
        float s = Ctb[r][q];
        KV_UNROLL for (int q2 = 0; q2 < P; ++q2) {
          s = fmaf(go.Pm[r][q2], Sb[q][q2], s);                  // P Sb^T
          s = fmaf(P2[r][q2], Sb[q2][q], s);                     // (Sp^T C^T) Sb
        }
        Ctb[r][q] = s - mup[r] * rb[q];
      }
    }
    // Sp-bar' += C^T Sb C + P-bar C ;  mu_p-bar' = mu_p-bar + mu_f-bar - C^T r-bar
    {
      float T3m[R][P];
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
        float s = Pb[r][q];
        KV_UNROLL for (int q2 = 0; q2 < P; ++q2) s = fmaf(Ct[r][q2], Sb[q2][q], s);
        T3m[r][q] = s;
      }
      mm_RSt<true>(T3m, Ct_v, Spb);
    }
    KV_UNROLL for (int r = 0; r < R; ++r) {
      float s = 0.f;
      KV_UNROLL for (int q = 0; q < P; ++q) s = fmaf(Ct[r][q], rb[q], s);
      mpb[r] = mpb[r] + mfb[r] - s;
    }
    // A-bar = Sp-bar' (A Sprev^T) + Sp-bar'^T (A Sprev) + mu_p-bar' mu_prev^T
    float Ab[R][N];
    {
      auto Sv = publish<MEM, L, R, N>(g, Sprev, T0);
      float M1[R][N], M2[R][N];
      mm_RSt<false>(A, Sv, M1);
      mm_RS<false>(A, Sv, M2);
      auto M1_v = publish<MEM, L, R, N>(g, M1, T3);
      mm_RS<false>(Spb, M1_v, Ab);
      auto M2_v = publish<MEM, L, R, N>(g, M2, T4);
      auto Spb_v = publish<MEM, L, R, N>(g, Spb, T5);
      mm_StS<true, R, N, N>(Spb_v, row0, M2_v, Ab);
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Ab[r][j] = fmaf(mpb[r], muprev[j], Ab[r][j]);
    }
    float Bb[R][M], du[M];
    KV_UNROLL for (int j = 0; j < M; ++j) du[j] = 0.f;
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < M; ++j) {
      Bb[r][j] = mpb[r] * in.u[j];
      du[j] = fmaf(Bm[r][j], mpb[r], du[j]);
    }
    // direct cotangents of A_list/B_list at t = 0 are not covered by sweep 3 (which handles t+1 >= 1)
    if (t == 0) {
      float cA[R][N], cB[R][M];
      load_rows_opt<R, N>(w.c_A, (bt * N + row0) * N, cA);
      load_rows_opt<R, M>(w.c_B, (bt * N + row0) * M, cB);
      add_rows<R, N>(Ab, cA);
      add_rows<R, M>(Bb, cB);
    }
    // carry to t-1:  Sigma_f-bar += A^T Sp-bar' A ; mu_f-bar += A^T mu_p-bar'
    {
      auto A_v = publish<MEM, L, R, N>(g, A, T0);
      float T4m[R][N];
      mm_RS<false>(Spb, A_v, T4m);
      auto T4_v = publish<MEM, L, R, N>(g, T4m, T1);
      mm_StS<false, R, N, N>(A_v, row0, T4_v, Sf_carry);
      float mpb_full[N];
      allgather<MEM, L, R>(g, mpb, VB, mpb_full);
      matTvec_own<C>(A_v, row0, mpb_full, mf_carry);
    }
    // A.0: contract with the base matrices
    float dal[K];
    KV_UNROLL for (int k = 0; k < K; ++k) dal[k] = 0.f;
    acc.addA(base, row0, in.al, Ab, dal, bt, false);
    acc.addB(base, row0, in.al, Bb, dal, bt, false);
    acc.addQ(base, row0, in.al, Spb, dal, bt, false);
    acc.addCt(base, row0, in.al, Ctb, dal, bt, false);
    float red2[K + M];
    KV_UNROLL for (int k = 0; k < K; ++k) red2[k] = dal[k];
    KV_UNROLL for (int j = 0; j < M; ++j) red2[K + j] = du[j];
    g.allreduce(red2);
    if (active && g.lane == 0) {
      float da[K], dyv[P];
      KV_UNROLL for (int k = 0; k < K; ++k) da[k] = da_old[k] + red2[k];
      store_row<K>(w.dalpha + bt * K, da);
      KV_UNROLL for (int q = 0; q < P; ++q) dyv[q] = dy_old[q] + rb[q];
      store_row<P>(w.dY + bt * P, dyv);
      if (w.dU) {
        float duv[M];
        KV_UNROLL for (int j = 0; j < M; ++j) duv[j] = du_old[j] + red2[K + j];
        store_row<M>(w.dU + bt * M, duv);
      }
    }
  }
  if (!ok && active) kv_info_or(a.info, KV_INFO_PIVOT);
}

}  // namespace kvae
