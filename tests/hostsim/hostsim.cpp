// TEST TOOLING — host (g++) build of the device arithmetic headers with one lane per sequence.
// Lets the CPU-only build box check kvae_fwd/elbo/bwd.cuh against the oracle before any GPU
// time is spent.  Never loaded by the kalman_vae_b200 package (the product path is CUDA only).
#include <vector>
#include <cstring>
#include "../../kalman_vae_b200/csrc/kvae_configs.h"
#include "../../kalman_vae_b200/csrc/kvae_fwd.cuh"
#include "../../kalman_vae_b200/csrc/kvae_elbo.cuh"
#include "../../kalman_vae_b200/csrc/kvae_bwd.cuh"

using namespace kvae;

struct HostParams { const float *A, *Bm, *C, *Q, *R, *mu0, *S0; };

template <class C> static void run_fwd(const Args& a, const HostParams& hp, int smooth) {
  std::vector<float> base(Base<C>::total);
  for (int i = 0; i < Base<C>::total; ++i) base_fill<C>(base.data(), i, hp.A, hp.Bm, hp.C, hp.Q, hp.R, hp.mu0, hp.S0);
  std::vector<float> tiles(FTiles<C>::warp_total + 4);
  Group<C::L, C::R> g{0, 1u};
  FTiles<C> tl{tiles.data(), 0};
  for (int b = 0; b < a.B; ++b) {
    float Sig[C::R][C::N], mu[C::N], mu_own[C::R];
    alignas(16) float slot[InStage<C, true>::group_floats + 4];
    filter_sweep<C>(a, base.data(), tl, g, b, true, slot, Sig, mu, mu_own);
    if (smooth) smoother_sweep<C>(a, base.data(), tl, g, b, true, slot, Sig, mu_own);
  }
}

extern "C" int hostsim_fwd(int N, int P, int M, int K, int switching, int force_mem, int smooth, int B, int T,
                           const float* Y, const float* U, const float* mask, const float* alpha,
                           const float* A, const float* Bm, const float* C, const float* Q, const float* R,
                           const float* mu0, const float* S0,
                           float* mu_f, float* Sig_f, float* mu_p, float* Sig_p, float* A_list, float* B_list, float* C_list,
                           float* mu_s, float* Sig_s, int* info) {
  Args a{};
  a.B = B; a.T = T; a.Y = Y; a.U = U; a.mask = mask; a.alpha = alpha;
  a.mu_f = mu_f; a.Sig_f = Sig_f; a.mu_p = mu_p; a.Sig_p = Sig_p; a.mu_s = mu_s; a.Sig_s = Sig_s;
  a.A_list = A_list; a.B_list = B_list; a.C_list = C_list; a.info = info;
  HostParams hp{A, Bm, C, Q, R, mu0, S0};
#define X(n, p, m, k)                                                                                   \
  if (N == n && P == p && M == m && K == k) {                                                           \
    if (switching) { if (force_mem) run_fwd<Cfg<n, p, m, k, 1, true, true, true>>(a, hp, smooth);       \
                     else run_fwd<Cfg<n, p, m, k, 1, true, true, false>>(a, hp, smooth); }              \
    else { if (force_mem) run_fwd<Cfg<n, p, m, k, 1, false, false, true>>(a, hp, smooth);               \
           else run_fwd<Cfg<n, p, m, k, 1, false, false, false>>(a, hp, smooth); }                      \
    return 0;                                                                                           \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return -1;
}

template <class C> static void run_elbo(const Args& a, const HostParams& hp, float jitter, double* acc5) {
  std::vector<float> base(Base<C>::total);
  for (int i = 0; i < Base<C>::total; ++i) base_fill<C>(base.data(), i, hp.A, hp.Bm, hp.C, hp.Q, hp.R, hp.mu0, hp.S0);
  std::vector<float> tiles(FTiles<C>::warp_total + 4);
  Group<C::L, C::R> g{0, 1u};
  FTiles<C> tl{tiles.data(), 0};
  double acc[5] = {0, 0, 0, 0, 0};
  // cut every sequence into chunks of 3 steps to exercise the time-parallel chunk starts
  for (int b = 0; b < a.B; ++b)
    for (int t0 = 0; t0 < a.T; t0 += 3) elbo_sweep<C>(a, base.data(), tl, g, b, true, jitter, t0, t0 + 3 < a.T ? t0 + 3 : a.T, nullptr, acc);
  for (int i = 0; i < 5; ++i) acc5[i] = acc[i];
}

extern "C" int hostsim_elbo(int N, int P, int M, int K, int switching, int force_mem, int B, int T,
                            const float* Y, const float* U, const float* mask, const float* alpha, const float* eps,
                            const float* A, const float* Bm, const float* C, const float* Q, const float* R,
                            const float* mu0, const float* S0, float* mu_s, float* Sig_s, float jitter,
                            double* acc5, int* info) {
  Args a{};
  a.B = B; a.T = T; a.Y = Y; a.U = U; a.mask = mask; a.alpha = alpha; a.eps = eps;
  a.mu_s = mu_s; a.Sig_s = Sig_s; a.info = info;
  a.jitter_q = jitter; a.chol_diag = 0;
  HostParams hp{A, Bm, C, Q, R, mu0, S0};
#define X(n, p, m, k)                                                                                   \
  if (N == n && P == p && M == m && K == k) {                                                           \
    if (switching) { if (force_mem) run_elbo<Cfg<n, p, m, k, 1, true, true, true>>(a, hp, jitter, acc5);   \
                     else run_elbo<Cfg<n, p, m, k, 1, true, true, false>>(a, hp, jitter, acc5); }          \
    else { if (force_mem) run_elbo<Cfg<n, p, m, k, 1, false, false, true>>(a, hp, jitter, acc5);           \
           else run_elbo<Cfg<n, p, m, k, 1, false, false, false>>(a, hp, jitter, acc5); }                  \
    return 0;                                                                                           \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return -1;
}

template <class C> static void run_bwd(const Args& a, BwdArgs w, const HostParams& hp, double* gp, float** dbg, double* el5) {
  std::vector<float> base(Base<C>::total);
  for (int i = 0; i < Base<C>::total; ++i) base_fill<C>(base.data(), i, hp.A, hp.Bm, hp.C, hp.Q, hp.R, hp.mu0, hp.S0);
  std::vector<float> tiles(BTiles<C>::warp_total + 4);
  BTiles<C> tl{tiles.data(), 0};
  const size_t nn = (size_t)a.B * a.T * C::N * C::N, nv = (size_t)a.B * a.T * C::N;
  std::vector<float> wSf(nn), wSp(nn), wmf(nv), wmp(nv);
  w.w_Sig_f = wSf.data(); w.w_Sig_p = wSp.data(); w.w_mu_f = wmf.data(); w.w_mu_p = wmp.data();
  Group<C::L, C::R> g{0, 1u};
  for (int i = 0; i < GradAcc<C>::PSZ; ++i) gp[i] = 0.0;
  for (int b = 0; b < a.B; ++b) {
    GradAcc<C> acc;
    acc.zero();
    using GA = GradAcc<C>;
    const size_t BTs = (size_t)a.B * a.T;
    static std::vector<float> dA, dB, dQ, dCt;
    dA.assign(BTs * C::N * C::N, 0.f); dB.assign(BTs * C::N * C::M, 0.f); dQ.assign(BTs * C::N * C::N, 0.f); dCt.assign(BTs * C::N * C::P, 0.f);
    acc.dnA = dA.data(); acc.dnB = dB.data(); acc.dnQ = dQ.data(); acc.dnCt = dCt.data();
    acc.on = true;
    double el[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    bwd_sweep3<C>(a, w, base.data(), tl, g, b, true, acc, el);
    if (el5) for (int i = 0; i < 5; ++i) el5[i] += el[i];
    bwd_sweep4<C>(a, w, base.data(), tl, g, b, true, acc);
    acc.for_each(0, [&](int idx, float v) { gp[idx] += (double)v; });
    if (GA::DENSE) {   // host version of k_mode_contract for this sequence's steps
      for (int t = 0; t < a.T; ++t) {
        const size_t bt = (size_t)b * a.T + t;
        for (int k = 0; k < C::K; ++k) {
          const double al = a.alpha[bt * C::K + k];
          for (int e = 0; e < C::N * C::N; ++e) gp[GA::fA + k * C::N * C::N + e] += al * dA[bt * C::N * C::N + e];
          for (int e = 0; e < C::N * C::M; ++e) gp[GA::fB + k * C::N * C::M + e] += al * dB[bt * C::N * C::M + e];
          if (C::QPM) for (int e = 0; e < C::N * C::N; ++e) gp[GA::fQ + k * C::N * C::N + e] += al * dQ[bt * C::N * C::N + e];
          if (!C::CSH) for (int i = 0; i < C::N; ++i) for (int q = 0; q < C::P; ++q)
            gp[GA::fC + (k * C::P + q) * C::N + i] += al * dCt[bt * C::N * C::P + i * C::P + q];
        }
      }
    }
  }
  if (dbg) {
    memcpy(dbg[0], wSf.data(), nn * 4); memcpy(dbg[1], wSp.data(), nn * 4);
    memcpy(dbg[2], wmf.data(), nv * 4); memcpy(dbg[3], wmp.data(), nv * 4);
  }
}

extern "C" int hostsim_bwd(int N, int P, int M, int K, int switching, int force_mem, int B, int T,
                           const float* Y, const float* U, const float* mask, const float* alpha, const float* eps,
                           const float* A, const float* Bm, const float* C, const float* Q, const float* R,
                           const float* mu0, const float* S0,
                           float* mu_f, float* Sig_f, float* mu_p, float* Sig_p, float* mu_s, float* Sig_s,
                           float c_elbo, float jitter, const float** cot9,
                           float* dY, float* dU, float* dalpha, double* gparams, int* info, float** dbg, double* el5) {
  Args a{};
  a.B = B; a.T = T; a.Y = Y; a.U = U; a.mask = mask; a.alpha = alpha; a.eps = eps;
  a.mu_f = mu_f; a.Sig_f = Sig_f; a.mu_p = mu_p; a.Sig_p = Sig_p; a.mu_s = mu_s; a.Sig_s = Sig_s; a.info = info;
  a.jitter_q = jitter; a.chol_diag = 0;
  BwdArgs w{};
  w.c_mu_s = cot9[0]; w.c_Sig_s = cot9[1]; w.c_mu_f = cot9[2]; w.c_Sig_f = cot9[3]; w.c_mu_p = cot9[4]; w.c_Sig_p = cot9[5];
  w.c_A = cot9[6]; w.c_B = cot9[7]; w.c_C = cot9[8];
  w.dY = dY; w.dU = dU; w.dalpha = dalpha; w.c_elbo = c_elbo; w.jitter = jitter;
  w.with_elbo = el5 ? 1 : 0;
  if (el5) for (int i = 0; i < 5; ++i) el5[i] = 0.0;
  HostParams hp{A, Bm, C, Q, R, mu0, S0};
#define X(n, p, m, k)                                                                                   \
  if (N == n && P == p && M == m && K == k) {                                                           \
    if (switching) { if (force_mem) run_bwd<Cfg<n, p, m, k, 1, true, true, true>>(a, w, hp, gparams, dbg, el5);      \
                     else run_bwd<Cfg<n, p, m, k, 1, true, true, false>>(a, w, hp, gparams, dbg, el5); }             \
    else { if (force_mem) run_bwd<Cfg<n, p, m, k, 1, false, false, true>>(a, w, hp, gparams, dbg, el5);              \
           else run_bwd<Cfg<n, p, m, k, 1, false, false, false>>(a, w, hp, gparams, dbg, el5); }                     \
    return 0;                                                                                           \
  }
  KVAE_FOR_EACH_SHAPE(X)
#undef X
  return -1;
}
