// kvae_ops.h — internal interface between the C ABI (kvae_capi.cu) and the per-shape
// translation units (kvae_shape.cu compiled once per KVAE_FOR_EACH_SHAPE entry).
#pragma once
#include <cuda_runtime.h>
#include "../../include/kvae_kalman.h"

namespace kvae {

// device-side view of a kvae_dp_comm (csrc/kvae_dp.cu): every rank's exchange buffer as mapped in this process.
// buffer layout (64-bit words): [0] step counter, [1] blocks-done counter, [16 ...] 2 slots x world x nf_pad LL words
struct DpView {
  unsigned long long* buf[16];
  int rank, world;
  unsigned long long nf_pad;
  int nparam;
};
constexpr int KV_DP_HDR_WORDS = 16;
bool kvae_dp_get_view(kvae_dp_comm* c, DpView* out);   // false if the communicator is not connected

struct BwdExtra {
  const float* eps; float jitter; float jitter_q; int chol_diag; const float* g_elbo; float* terms;
  const kvae_cotangents* cot; const kvae_grads* grads; void* workspace;
  const DpView* dp;   // non-null: kvae_kf_bwd_dp -- the final kernel also does the cross-rank exchange
};

template <int N, int P, int M, int K> struct ShapeOps {
  static bool lanes_ok(int lanes);
  static int fwd_grid(const kvae_dims& d);
  static int fwd_lstm(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, float* A_list, float* B_list,
                      float* C_list, const kvae_lstm& lw, float* alpha_out, int32_t* info, cudaStream_t s);
  static int fwd(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, float* A_list, float* B_list,
                 float* C_list, int32_t* info, cudaStream_t s);
  static size_t elbo_ws(const kvae_dims& d);
  static int elbo(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, const float* eps, float jitter,
                  float jitter_q, int chol_diag, float* terms, void* ws, int32_t* info, cudaStream_t s);
  static size_t bwd_ws(const kvae_dims& d);
  static int bwd(const kvae_dims& d, const kvae_inputs& in, const kvae_states& st, const BwdExtra& x, int32_t* info,
                 cudaStream_t s);
};

}  // namespace kvae
