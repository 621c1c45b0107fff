// kvae_regime.cu — SKVAE regime sampler: the T-step Gumbel-softmax Markov chain of
// SwitchingDynamicsParameter.compute_batch (kvae/kalman/switch_dyn_param.py:51-79) as ONE forward launch and ONE
// explicit-adjoint launch (the reference runs ~12 tiny ops per time step in a Python loop and lets autograd replay them).
//
//   t = 0 : l_0 = init_logits                    log p_0 = log(1/K)
//   t >= 1: l_t = y_{t-1}^T logits[:, t]         tp_t = y_{t-1}^T trans
//   y_t   = gumbel_softmax(l_t; g_t, tau, hard)  = softmax((l_t + g_t)/tau)            (soft)
//                                                = (onehot(argmax s) - s) + s           (hard, straight-through; the
//                                                  reference's fp32 expression y_hard - y_soft.detach() + y_soft)
//   log_q[t] = sum_k y_tk log_softmax(l_t)_k     log_p[t] = sum_k y_tk log(max(tp_tk, 1e-8))   (t=0: y_0k log(1/K))
//
// One thread owns one sequence (K <= 8 values in registers); the chain is sequential in t, sequences are independent.
// The Gumbel noise g [B,T,K] is an INPUT (drawn by the caller with torch, as the reference's gumbel_softmax does), so
// the launch is deterministic and testable against the reference.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../../include/kvae_kalman.h"

namespace {

// One thread per sequence reads its own K*K logits per step: every load instruction of a warp touches 32 different
// sectors, so the kernels are bound by L1 tag/wavefront throughput per SM, not by latency.  One-warp CTAs spread the
// warps over all SMs (B = 8192 -> 256 CTAs).
constexpr int RG_TPB = 32;

template <int K> struct Step {
  float s[K];      // softmax((l+g)/tau)
  float y[K];      // sample (soft: = s; hard: straight-through one-hot)
  float lsm[K];    // log_softmax(l)
};

template <int K>
__device__ __forceinline__ void sample_step(const float (&l)[K], const float (&g)[K], float tau, int hard, Step<K>& o) {
  float x[K];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) { x[k] = (l[k] + g[k]) / tau; mx = fmaxf(mx, x[k]); }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { o.s[k] = expf(x[k] - mx); sum += o.s[k]; }
  int arg = 0;
  float best = -1.f;
#pragma unroll
  for (int k = 0; k < K; ++k) { o.s[k] = o.s[k] / sum; if (o.s[k] > best) { best = o.s[k]; arg = k; } }   // first maximum, as torch.max
#pragma unroll
  for (int k = 0; k < K; ++k) o.y[k] = hard ? (((k == arg) ? 1.f : 0.f) - o.s[k]) + o.s[k] : o.s[k];
  float ml = -INFINITY;
#pragma unroll
  for (int k = 0; k < K; ++k) ml = fmaxf(ml, l[k]);
  float se = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) se += expf(l[k] - ml);
  const float lse = logf(se);
#pragma unroll
  for (int k = 0; k < K; ++k) o.lsm[k] = (l[k] - ml) - lse;
}

template <int K>
__global__ void __launch_bounds__(RG_TPB) k_regime_fwd(int B, int T, float tau, int hard, const float* __restrict__ logits,
                                                    const float* __restrict__ init_logits, const float* __restrict__ gumbel,
                                                    const float* __restrict__ trans, float* __restrict__ y_seq,
                                                    float* __restrict__ log_q, float* __restrict__ log_p) {
  __shared__ float tr[K * K];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) tr[i] = trans[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float yp[K];
  const float log_p0 = logf(1.0f / K);
  // software pipeline: the inputs of step t+1 do not depend on the chain, so they are fetched while step t computes
  float gn[K], Mn[K * K];
#pragma unroll
  for (int k = 0; k < K; ++k) gn[k] = gumbel[(size_t)b * T * K + k];
#pragma unroll
  for (int i = 0; i < K * K; ++i) Mn[i] = 0.f;
  for (int t = 0; t < T; ++t) {
    const size_t bt = (size_t)b * T + t;
    float l[K], g[K], M[K * K];
#pragma unroll
    for (int k = 0; k < K; ++k) g[k] = gn[k];
#pragma unroll
    for (int i = 0; i < K * K; ++i) M[i] = Mn[i];
    if (t + 1 < T) {
#pragma unroll
      for (int k = 0; k < K; ++k) gn[k] = gumbel[(bt + 1) * K + k];
#pragma unroll
      for (int i = 0; i < K * K; ++i) Mn[i] = logits[(bt + 1) * K * K + i];
    }
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) l[k] = init_logits[(size_t)b * K + k];
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) l[j] = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) l[j] = fmaf(yp[i], M[i * K + j], l[j]);
    }
    Step<K> st;
    sample_step<K>(l, g, tau, hard, st);
    float lq = 0.f, lp = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) lq = fmaf(st.y[k], st.lsm[k], lq);
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) lp = fmaf(st.y[k], log_p0, lp);
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float tp = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) tp = fmaf(yp[i], tr[i * K + j], tp);
        lp = fmaf(st.y[j], logf(fmaxf(tp, 1e-8f)), lp);
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) { y_seq[bt * K + k] = st.y[k]; yp[k] = st.y[k]; }
    log_q[bt] = lq;
    log_p[bt] = lp;
  }
}

// ---- vectorised variants.  One thread per sequence means every SCALAR load/store of a warp touches 32 different
// sectors (ncu: 31.6 sectors per request, long-scoreboard = 4 stall cycles per issued instruction even with a one-step
// prefetch).  CH consecutive steps of a sequence are contiguous in every tensor ([B,T,...] layout), so the kernels below
// move CH steps at a time with 128-bit accesses (CH*K*K and CH*K multiples of 4 floats; K = 3: CH = 4 -> 9 + 3 loads per
// 4 steps instead of 48) and prefetch a whole chunk ahead.  Requires T % CH == 0; otherwise the scalar kernels run.
template <int NF> __device__ __forceinline__ void ldv(const float* __restrict__ p, float (&o)[NF]) {
  static_assert(NF % 4 == 0, "vector width");
#pragma unroll
  for (int q = 0; q < NF / 4; ++q) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * q);
    o[4 * q] = v.x; o[4 * q + 1] = v.y; o[4 * q + 2] = v.z; o[4 * q + 3] = v.w;
  }
}
template <int NF> __device__ __forceinline__ void stv(float* __restrict__ p, const float (&o)[NF]) {
  static_assert(NF % 4 == 0, "vector width");
#pragma unroll
  for (int q = 0; q < NF / 4; ++q) *reinterpret_cast<float4*>(p + 4 * q) = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
}
template <int K> struct RgChunk { static constexpr int CH = (K == 3) ? 4 : (K == 2 ? 2 : 1); };   // K in {2,3,4,8}

template <int K>
__global__ void __launch_bounds__(RG_TPB) k_regime_fwd_v(int B, int T, float tau, int hard, const float* __restrict__ logits,
                                                      const float* __restrict__ init_logits, const float* __restrict__ gumbel,
                                                      const float* __restrict__ trans, float* __restrict__ y_seq,
                                                      float* __restrict__ log_q, float* __restrict__ log_p) {
  constexpr int CH = RgChunk<K>::CH, KK = K * K;
  __shared__ float tr[KK];
  for (int i = threadIdx.x; i < KK; i += blockDim.x) tr[i] = trans[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float yp[K];
#pragma unroll
  for (int k = 0; k < K; ++k) yp[k] = 0.f;
  const float log_p0 = logf(1.0f / K);
  float Mn[CH * KK], gn[CH * K];
  ldv<CH * KK>(logits + (size_t)b * T * KK, Mn);
  ldv<CH * K>(gumbel + (size_t)b * T * K, gn);
  for (int t0 = 0; t0 < T; t0 += CH) {
    const size_t bt0 = (size_t)b * T + t0;
    float Mc[CH * KK], gc[CH * K];
#pragma unroll
    for (int i = 0; i < CH * KK; ++i) Mc[i] = Mn[i];
#pragma unroll
    for (int i = 0; i < CH * K; ++i) gc[i] = gn[i];
    if (t0 + CH < T) {   // whole next chunk in flight while this one computes
      ldv<CH * KK>(logits + (bt0 + CH) * KK, Mn);
      ldv<CH * K>(gumbel + (bt0 + CH) * K, gn);
    }
    float yc[CH * K], lqc[CH], lpc[CH];
#pragma unroll
    for (int s = 0; s < CH; ++s) {
      float l[K], g[K];
#pragma unroll
      for (int k = 0; k < K; ++k) g[k] = gc[s * K + k];
      if (t0 + s == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) l[k] = init_logits[(size_t)b * K + k];
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) l[j] = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) l[j] = fmaf(yp[i], Mc[s * KK + i * K + j], l[j]);
      }
      Step<K> st;
      sample_step<K>(l, g, tau, hard, st);
      float lq = 0.f, lp = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) lq = fmaf(st.y[k], st.lsm[k], lq);
      if (t0 + s == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) lp = fmaf(st.y[k], log_p0, lp);
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float tp = 0.f;
#pragma unroll
          for (int i = 0; i < K; ++i) tp = fmaf(yp[i], tr[i * K + j], tp);
          lp = fmaf(st.y[j], logf(fmaxf(tp, 1e-8f)), lp);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) { yc[s * K + k] = st.y[k]; yp[k] = st.y[k]; }
      lqc[s] = lq;
      lpc[s] = lp;
    }
    stv<CH * K>(y_seq + bt0 * K, yc);
    if constexpr (CH % 4 == 0) {
      stv<CH>(log_q + bt0, lqc);
      stv<CH>(log_p + bt0, lpc);
    } else {
#pragma unroll
      for (int s = 0; s < CH; ++s) { log_q[bt0 + s] = lqc[s]; log_p[bt0 + s] = lpc[s]; }
    }
  }
}

// reverse-time adjoint: gradient of  sum <g_y, y_seq> + <g_logq, log_q> + <g_logp, log_p>  w.r.t. logits and init_logits
template <int K>
__global__ void __launch_bounds__(RG_TPB) k_regime_bwd(int B, int T, float tau, int hard, const float* __restrict__ logits,
                                                    const float* __restrict__ init_logits, const float* __restrict__ gumbel,
                                                    const float* __restrict__ trans, const float* __restrict__ y_seq,
                                                    const float* __restrict__ g_y, const float* __restrict__ g_logq,
                                                    const float* __restrict__ g_logp, float* __restrict__ d_logits,
                                                    float* __restrict__ d_init) {
  __shared__ float tr[K * K];
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) tr[i] = trans[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float log_p0 = logf(1.0f / K);
  float carry[K];   // d/dy_t coming from step t+1
#pragma unroll
  for (int k = 0; k < K; ++k) carry[k] = 0.f;
  // software pipeline: everything step t-1 reads (y_{t-2}, logits_{t-1}, gumbel_{t-1}, cotangents) is fetched during step t
  float ypn[K], gn[K], Mn[K * K], gyn[K], qbn, pbn;
  auto fetch = [&](int t) {
    const size_t bt = (size_t)b * T + t;
#pragma unroll
    for (int k = 0; k < K; ++k) { gn[k] = gumbel[bt * K + k]; gyn[k] = g_y ? g_y[bt * K + k] : 0.f; }
    qbn = g_logq ? g_logq[bt] : 0.f;
    pbn = g_logp ? g_logp[bt] : 0.f;
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) ypn[k] = y_seq[(bt - 1) * K + k];
#pragma unroll
      for (int i = 0; i < K * K; ++i) Mn[i] = logits[bt * K * K + i];
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) ypn[k] = 0.f;
#pragma unroll
      for (int i = 0; i < K * K; ++i) Mn[i] = 0.f;
    }
  };
  fetch(T - 1);
  for (int t = T - 1; t >= 0; --t) {
    const size_t bt = (size_t)b * T + t;
    float yp[K], l[K], g[K], M[K * K], gy[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { yp[k] = ypn[k]; g[k] = gn[k]; gy[k] = gyn[k]; }
#pragma unroll
    for (int i = 0; i < K * K; ++i) M[i] = Mn[i];
    const float qb = qbn, pb = pbn;
    if (t > 0) fetch(t - 1);
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) l[k] = init_logits[(size_t)b * K + k];
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) l[j] = 0.f;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) l[j] = fmaf(yp[i], M[i * K + j], l[j]);
    }
    Step<K> st;
    sample_step<K>(l, g, tau, hard, st);      // recomputed (bit-identical to the forward launch)
    float yb[K], lb[K], ypb[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { yb[k] = carry[k] + gy[k]; ypb[k] = 0.f; }
    // log_q = sum y lsm
    float sum_lsmb = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { yb[k] = fmaf(qb, st.lsm[k], yb[k]); sum_lsmb += qb * st.y[k]; }
#pragma unroll
    for (int k = 0; k < K; ++k) lb[k] = qb * st.y[k] - expf(st.lsm[k]) * sum_lsmb;   // log_softmax backward
    // log_p
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) yb[k] = fmaf(pb, log_p0, yb[k]);
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) {
        float tp = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) tp = fmaf(yp[i], tr[i * K + j], tp);
        yb[j] = fmaf(pb, logf(fmaxf(tp, 1e-8f)), yb[j]);
        const float tpb = (tp >= 1e-8f) ? pb * st.y[j] / tp : 0.f;                     // clamp_min backward
#pragma unroll
        for (int i = 0; i < K; ++i) ypb[i] = fmaf(tpb, tr[i * K + j], ypb[i]);
      }
    }
    // y = f(s), dy/ds = I (soft and straight-through hard);  s = softmax(x), x = (l + g)/tau
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) dot = fmaf(st.s[k], yb[k], dot);
#pragma unroll
    for (int k = 0; k < K; ++k) lb[k] += st.s[k] * (yb[k] - dot) / tau;
    if (t == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) d_init[(size_t)b * K + k] = lb[k];
#pragma unroll
      for (int i = 0; i < K * K; ++i) d_logits[bt * K * K + i] = 0.f;                  // logits[:, 0] is never read (:67)
    } else {
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          d_logits[bt * K * K + i * K + j] = yp[i] * lb[j];
          ypb[i] = fmaf(M[i * K + j], lb[j], ypb[i]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) carry[k] = ypb[k];
  }
}

// vectorised reverse-time adjoint (see k_regime_fwd_v): chunks of CH steps, processed last to first
template <int K>
__global__ void __launch_bounds__(RG_TPB) k_regime_bwd_v(int B, int T, float tau, int hard, const float* __restrict__ logits,
                                                      const float* __restrict__ init_logits, const float* __restrict__ gumbel,
                                                      const float* __restrict__ trans, const float* __restrict__ y_seq,
                                                      const float* __restrict__ g_y, const float* __restrict__ g_logq,
                                                      const float* __restrict__ g_logp, float* __restrict__ d_logits,
                                                      float* __restrict__ d_init) {
  constexpr int CH = RgChunk<K>::CH, KK = K * K;
  constexpr bool PF = (CH * KK <= 36);   // prefetch a whole chunk ahead when the registers allow it
  __shared__ float tr[KK];
  for (int i = threadIdx.x; i < KK; i += blockDim.x) tr[i] = trans[i];
  __syncthreads();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float log_p0 = logf(1.0f / K);
  struct Chunk { float M[CH * KK], g[CH * K], y[CH * K], gy[CH * K], q[CH], p[CH]; };
  auto fetch = [&](int t0, Chunk& c) {
    const size_t bt0 = (size_t)b * T + t0;
    ldv<CH * KK>(logits + bt0 * KK, c.M);
    ldv<CH * K>(gumbel + bt0 * K, c.g);
    ldv<CH * K>(y_seq + bt0 * K, c.y);
    if (g_y) ldv<CH * K>(g_y + bt0 * K, c.gy);
    else {
#pragma unroll
      for (int i = 0; i < CH * K; ++i) c.gy[i] = 0.f;
    }
#pragma unroll
    for (int s = 0; s < CH; ++s) { c.q[s] = g_logq ? g_logq[bt0 + s] : 0.f; c.p[s] = g_logp ? g_logp[bt0 + s] : 0.f; }
  };
  float carry[K];
#pragma unroll
  for (int k = 0; k < K; ++k) carry[k] = 0.f;
  Chunk nx;
  fetch(T - CH, nx);
  for (int t0 = T - CH; t0 >= 0; t0 -= CH) {
    const size_t bt0 = (size_t)b * T + t0;
    Chunk cu = nx;
    float yprev0[K];   // y_{t0-1}
    if (t0 > 0) {
      if constexpr (PF) {
        fetch(t0 - CH, nx);
#pragma unroll
        for (int k = 0; k < K; ++k) yprev0[k] = nx.y[(CH - 1) * K + k];
      } else {
#pragma unroll
        for (int k = 0; k < K; ++k) yprev0[k] = y_seq[(bt0 - 1) * K + k];
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k) yprev0[k] = 0.f;
    }
    float dl[CH * KK];
#pragma unroll
    for (int s = CH - 1; s >= 0; --s) {
      const int t = t0 + s;
      float yp[K], l[K], g[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { yp[k] = (s > 0) ? cu.y[(s - 1) * K + k] : yprev0[k]; g[k] = cu.g[s * K + k]; }
      if (t == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) l[k] = init_logits[(size_t)b * K + k];
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) l[j] = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i)
#pragma unroll
          for (int j = 0; j < K; ++j) l[j] = fmaf(yp[i], cu.M[s * KK + i * K + j], l[j]);
      }
      Step<K> st;
      sample_step<K>(l, g, tau, hard, st);
      const float qb = cu.q[s], pb = cu.p[s];
      float yb[K], lb[K], ypb[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { yb[k] = carry[k] + cu.gy[s * K + k]; ypb[k] = 0.f; }
      float sum_lsmb = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) { yb[k] = fmaf(qb, st.lsm[k], yb[k]); sum_lsmb += qb * st.y[k]; }
#pragma unroll
      for (int k = 0; k < K; ++k) lb[k] = qb * st.y[k] - expf(st.lsm[k]) * sum_lsmb;
      if (t == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) yb[k] = fmaf(pb, log_p0, yb[k]);
      } else {
#pragma unroll
        for (int j = 0; j < K; ++j) {
          float tp = 0.f;
#pragma unroll
          for (int i = 0; i < K; ++i) tp = fmaf(yp[i], tr[i * K + j], tp);
          yb[j] = fmaf(pb, logf(fmaxf(tp, 1e-8f)), yb[j]);
          const float tpb = (tp >= 1e-8f) ? pb * st.y[j] / tp : 0.f;
#pragma unroll
          for (int i = 0; i < K; ++i) ypb[i] = fmaf(tpb, tr[i * K + j], ypb[i]);
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) dot = fmaf(st.s[k], yb[k], dot);
#pragma unroll
      for (int k = 0; k < K; ++k) lb[k] += st.s[k] * (yb[k] - dot) / tau;
      if (t == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) d_init[(size_t)b * K + k] = lb[k];
#pragma unroll
        for (int i = 0; i < KK; ++i) dl[s * KK + i] = 0.f;
      } else {
#pragma unroll
        for (int i = 0; i < K; ++i) {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            dl[s * KK + i * K + j] = yp[i] * lb[j];
            ypb[i] = fmaf(cu.M[s * KK + i * K + j], lb[j], ypb[i]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) carry[k] = ypb[k];
    }
    stv<CH * KK>(d_logits + bt0 * KK, dl);
    if constexpr (!PF) { if (t0 > 0) fetch(t0 - CH, nx); }
  }
}

thread_local char g_rg_err[200] = "";
int rg_fail(int code, const char* msg) { snprintf(g_rg_err, sizeof(g_rg_err), "%s", msg); return code; }

struct DevGuard {
  int prev = -1; bool sw = false;
  explicit DevGuard(int dev) { if (dev >= 0) { cudaGetDevice(&prev); if (prev != dev) { cudaSetDevice(dev); sw = true; } } }
  ~DevGuard() { if (sw) cudaSetDevice(prev); }
};

}  // namespace

extern "C" {

const char* kvae_regime_last_error(void) { return g_rg_err; }

int kvae_regime_supported(int K) { return K >= 2 && K <= 8; }

int kvae_regime_sample_fwd(const kvae_regime_dims* d, const float* logits, const float* init_logits, const float* gumbel,
                           const float* trans, float* y_seq, float* log_q, float* log_p, int device, void* stream) {
  if (!d || !logits || !init_logits || !gumbel || !trans || !y_seq || !log_q || !log_p) return rg_fail(-1, "null argument");
  if (d->B <= 0 || d->T <= 0 || !(d->tau > 0.f)) return rg_fail(-1, "B, T, tau must be positive");
  if (!kvae_regime_supported(d->K)) return rg_fail(-2, "K must be in 2..8");
  DevGuard guard(device);
  const int grid = (d->B + RG_TPB - 1) / RG_TPB;
  cudaStream_t s = (cudaStream_t)stream;
  const float it = d->tau;
  (void)cudaGetLastError();
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(gumbel) | reinterpret_cast<uintptr_t>(y_seq) |
                        reinterpret_cast<uintptr_t>(log_q) | reinterpret_cast<uintptr_t>(log_p)) & 15) == 0;
  switch (d->K) {
#define KV_CASE(k) case k: k_regime_fwd<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, log_q, log_p); break;
#define KV_CASE_V(k) case k:                                                                                                          \
    if (vec_ok && d->T % RgChunk<k>::CH == 0)                                                                                          \
      k_regime_fwd_v<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, log_q, log_p);    \
    else                                                                                                                               \
      k_regime_fwd<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, log_q, log_p);      \
    break;
    KV_CASE_V(2) KV_CASE_V(3) KV_CASE_V(4) KV_CASE(5) KV_CASE(6) KV_CASE(7) KV_CASE_V(8)
#undef KV_CASE
#undef KV_CASE_V
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : rg_fail((int)e, cudaGetErrorString(e));
}

int kvae_regime_sample_bwd(const kvae_regime_dims* d, const float* logits, const float* init_logits, const float* gumbel,
                           const float* trans, const float* y_seq, const float* g_y, const float* g_logq, const float* g_logp,
                           float* d_logits, float* d_init, int device, void* stream) {
  if (!d || !logits || !init_logits || !gumbel || !trans || !y_seq || !d_logits || !d_init) return rg_fail(-1, "null argument");
  if (d->B <= 0 || d->T <= 0 || !(d->tau > 0.f)) return rg_fail(-1, "B, T, tau must be positive");
  if (!kvae_regime_supported(d->K)) return rg_fail(-2, "K must be in 2..8");
  DevGuard guard(device);
  const int grid = (d->B + RG_TPB - 1) / RG_TPB;
  cudaStream_t s = (cudaStream_t)stream;
  const float it = d->tau;
  (void)cudaGetLastError();
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(logits) | reinterpret_cast<uintptr_t>(gumbel) | reinterpret_cast<uintptr_t>(y_seq) |
                        reinterpret_cast<uintptr_t>(g_y) | reinterpret_cast<uintptr_t>(d_logits)) & 15) == 0;
  switch (d->K) {
#define KV_CASE(k) case k: k_regime_bwd<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, g_y, g_logq, g_logp, d_logits, d_init); break;
#define KV_CASE_V(k) case k:                                                                                                                        \
    if (vec_ok && d->T % RgChunk<k>::CH == 0)                                                                                                        \
      k_regime_bwd_v<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, g_y, g_logq, g_logp, d_logits, d_init); \
    else                                                                                                                                             \
      k_regime_bwd<k><<<grid, RG_TPB, 0, s>>>(d->B, d->T, it, d->hard, logits, init_logits, gumbel, trans, y_seq, g_y, g_logq, g_logp, d_logits, d_init);   \
    break;
    KV_CASE_V(2) KV_CASE_V(3) KV_CASE_V(4) KV_CASE(5) KV_CASE(6) KV_CASE(7) KV_CASE_V(8)
#undef KV_CASE
#undef KV_CASE_V
  }
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : rg_fail((int)e, cudaGetErrorString(e));
}

}  // extern "C"
