// TEST TOOLING — host (g++) build of the device arithmetic headers with one lane per sequence.
// Lets the CPU-only build box check kvae_fwd/elbo/bwd.cuh against the oracle before any GPU
// time is spent.  Never loaded by the kalman_vae_b200 package (the product path is CUDA only).
#include <vector>
#include <cstring>
#include "../../kalman_vae_b200/csrc/kvae_configs.h"
#include "../../kalman_vae_b200/csrc/kvae_fwd.cuh"

using namespace kvae;

struct HostParams { const float *A, *Bm, *C, *Q, *R, *mu0, *S0; };

template <class C> static void run_fwd(const Args& a, const HostParams& hp, int smooth) {
  std::vector<float> base(Base<C>::total);
  for (int i = 0; i < Base<C>::total; ++i) base_fill<C>(base.data(), i, hp.A, hp.Bm, hp.C, hp.Q, hp.R, hp.mu0, hp.S0);
  std::vector<float> tiles(Tiles<C>::total + 4);
  Group<C::L, C::R> g{0};
  for (int b = 0; b < a.B; ++b) {
    float Sig[C::R][C::N], mu[C::N], mu_own[C::R];
    filter_sweep<C>(a, base.data(), tiles.data(), g, b, true, Sig, mu, mu_own);
    if (smooth) smoother_sweep<C>(a, base.data(), tiles.data(), g, b, true, Sig, mu_own);
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
