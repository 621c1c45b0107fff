// kvae_seq_bwd.cuh — explicit adjoint (sweeps 3 + 4) as a thread-per-sequence kernel with TMA-staged streams.
//
// Same arithmetic as csrc/kvae_bwd.cuh (the reference differentiates
// kvae/kalman/kalman_filter.py:31-401 with autograd, kvae/train/train.py:53) for the training case: gradient of
// g_elbo * elbo, no dense cotangents on the nine smooth outputs (those calls run on the lane-group kernels).
// One thread owns a sequence; see csrc/kvae_seq.cuh for the staging scheme.  Specific to this kernel:
//
//   * sweep 3 (t = 0..T-1) loads per step [Sigma_s, mu_s, Sigma_p, mu_p, eps](t+1) and Sigma_f(t), writes the scratch
//     (Sigma_f, mu_f)-bar(t), (Sigma_p, mu_p)-bar(t+1) and the partial dY / dalpha / dU of the step;
//     sweep 4 (t = T-1..0) loads the scratch, Sigma_p / mu_p (t), Sigma_f / mu_f (t-1) and the partial per-step
//     gradients (four steps per box, updated in place in shared memory and stored from there).
//   * parameter gradients dA_k = sum_{b,t} alpha_{t,k} Abar_t ... : a thread cannot afford K*(n*n + n*m + p*n) register
//     accumulators per sequence, and a warp-level reduction per step would cost more than the step.  Instead the four
//     threads of a quad exchange their per-step cotangents through a small shared-memory tile so that lane j of the
//     quad accumulates ROW j of Abar / Bbar / Ctbar (/ Qbar) of the quad's four sequences: K*(n + m + p) (+ K*n)
//     accumulators per lane, the same layout the lane-group kernel uses (GradAcc with L = 4), reduced once per CTA.
#pragma once
#include "kvae_seq.cuh"

namespace kvae {

struct alignas(64) SeqBwdMaps {
  CUtensorMap Y, U, alpha, mask;                          // inputs, four steps per box
  CUtensorMap dY, dalpha, dU;                             // per-step gradients, four steps per box
  CUtensorMap eps, mu_s, mu_p, mu_f, w_mu_f, w_mu_p;      // (4, T, B) rows
  CUtensorMap Sig_s, Sig_p, Sig_f, w_Sig_f, w_Sig_p;      // (16, T, B) rows, SWIZZLE_64B
};

template <class C> struct SeqBwdPlan {
  static constexpr int P = C::P, K = C::K, N = C::N, M = C::M;
  using TV = RowTile<1>;
  using TM = RowTile<4>;
  static constexpr int TMB = TM::bytes, TVB = TV::bytes;   // 2048, 512
  // input chunk (single buffer, re-armed one step ahead of its first use)
  static constexpr int in_Y = 0, in_U = in_Y + 32 * 4 * P * 4, in_al = in_U + 32 * 4 * M * 4, in_m = in_al + 32 * 4 * K * 4;
  static constexpr int in_bytes = kv_align_up(in_m + 32 * 4 * 4, 512);
  static __host__ __device__ constexpr uint32_t in_tx(bool has_u, bool has_m) {
    return 32u * 4 * P * 4 + (has_u ? 32u * 4 * M * 4 : 0u) + 32u * 4 * K * 4 + (has_m ? 32u * 4 * 4 : 0u);
  }
  // per-lane row of the gradient-exchange tile: [Abar 16 | Bbar 16 | Ctbar 4P | (Qbar 16) | alpha_ab K | alpha_c K | pad]; the row
  // stride is an odd multiple of 16 bytes so that the 128-bit row accesses of eight consecutive lanes (and of the two
  // quads x four rows of the transposed read) fall into distinct bank groups
  static constexpr int ax_A = 0, ax_B = 16, ax_C = 32, ax_Q = 32 + 4 * P, ax_al = ax_Q + (C::QPM ? 16 : 0);
  static constexpr int ax_row0 = kv_align_up(ax_al + 2 * K, 4);
  static constexpr int ax_row = (ax_row0 / 4) % 2 == 1 ? ax_row0 : ax_row0 + 4;
  static constexpr int ax_bytes = kv_align_up(32 * ax_row * 4, 512);
  // per-step gradient chunk [dY 4P | dalpha 4K | dU 4M] tiles (rows of 4W floats, no swizzle)
  static constexpr int gc_Y = 0, gc_al = gc_Y + 32 * 4 * P * 4, gc_U = gc_al + 32 * 4 * K * 4;
  static __host__ __device__ constexpr int gc_bytes(bool has_du) { return kv_align_up(gc_U + (has_du ? 32 * 4 * M * 4 : 0), 512); }
  static __host__ __device__ constexpr uint32_t gc_tx(bool has_du) { return 32u * 4 * P * 4 + 32u * 4 * K * 4 + (has_du ? 32u * 4 * M * 4 : 0u); }
  // sweep 3 state buffer: Sigma_s(t+1) | Sigma_p(t+1) | Sigma_f(t) | mu_s(t+1) | mu_p(t+1) | eps(t+1)
  static constexpr int s3_Ss = 0, s3_Sp = TMB, s3_Sf = 2 * TMB, s3_ms = 3 * TMB, s3_mp = s3_ms + TVB, s3_ep = s3_mp + TVB;
  static constexpr int s3_bytes = 3 * TMB + 3 * TVB;
  static constexpr uint32_t s3_tx = 3u * TMB + 3u * TVB;
  // sweep 3 output tiles: (Sigma_f)-bar(t) | (Sigma_p)-bar(t+1) | (mu_f)-bar(t) | (mu_p)-bar(t+1)
  static constexpr int o3_Sf = 0, o3_Sp = TMB, o3_mf = 2 * TMB, o3_mp = 2 * TMB + TVB;
  static constexpr int o3_bytes = 2 * TMB + 2 * TVB;
  // sweep 4 state buffer: Sigma_f-bar | Sigma_p-bar | Sigma_p | Sigma_f(t-1) | mu_f-bar | mu_p-bar | mu_p | mu_f(t-1)
  static constexpr int s4_Sfb = 0, s4_Spb = TMB, s4_Sp = 2 * TMB, s4_Sv = 3 * TMB, s4_mfb = 4 * TMB, s4_mpb = s4_mfb + TVB,
                       s4_mp = s4_mpb + TVB, s4_mv = s4_mp + TVB;
  static constexpr int s4_bytes = 4 * TMB + 4 * TVB;
  // layout of one warp: [input chunk | exchange tile | sweep region]
  static constexpr int oIn = 0, oAx = in_bytes, oSw = oAx + ax_bytes;
  // sweep 3: two state buffers, output tiles, gradient chunk; sweep 4: two state buffers, two gradient chunks
  static __host__ __device__ constexpr int warp_bytes(bool has_du) {
    const int s3 = 2 * s3_bytes + o3_bytes + gc_bytes(has_du);
    const int s4 = 2 * s4_bytes + 2 * gc_bytes(has_du);
    return kv_align_up(oSw + (s3 > s4 ? s3 : s4), 512);
  }
};

template <class C> size_t seq_bwd_smem(int warps, bool has_du) {
  const size_t red = sizeof(float) * (size_t)GradAcc<C>::PSZ;   // CTA reduction of the parameter gradients reuses the warps' region
  size_t body = (size_t)warps * SeqBwdPlan<C>::warp_bytes(has_du);
  if (body < red) body = red;
  return 512 + (size_t)kv_align_up((int)sizeof(float) * Base<C>::total, 512) + body;
}
// warps per CTA of the backward kernel: ~36.5 KB of staging per warp; three two-warp CTAs (77 KB each incl. the
// per-CTA reserve) fit an SM -> 6 warps per SM
inline int seq_bwd_warps_per_cta(int) {
  static const int forced = seq_env_int("KVAE_SEQ_BWD_WARPS", 0);   // development knob
  return (forced == 1 || forced == 2) ? forced : 2;
}
inline int seq_bwd_grid(int B) { const int per = 32 * seq_bwd_warps_per_cta(B); return (B + per - 1) / per; }

// one step's inputs out of the staged chunk holding step t (s = t & 3)
template <class C>
__device__ __forceinline__ void seqb_read_step(const unsigned char* in, int lane, int s, bool has_u, bool has_m, StepIn<C>& cur) {
  using PL = SeqBwdPlan<C>;
  constexpr int P = C::P, M = C::M, K = C::K;
  const float* y = reinterpret_cast<const float*>(in + PL::in_Y) + lane * 4 * P + s * P;
#pragma unroll
  for (int j = 0; j < P; ++j) cur.y[j] = y[j];
  if (has_u) load_row<M>(reinterpret_cast<const float*>(in + PL::in_U) + lane * 4 * M + s * M, cur.u);
  else {
#pragma unroll
    for (int j = 0; j < M; ++j) cur.u[j] = 0.f;
  }
  const float* al = reinterpret_cast<const float*>(in + PL::in_al) + lane * 4 * K + s * K;
#pragma unroll
  for (int k = 0; k < K; ++k) cur.al[k] = al[k];
  cur.m = has_m ? (reinterpret_cast<const float*>(in + PL::in_m))[lane * 4 + s] : 1.0f;
}

// ---------------------------------------------------------------------------------------
// Parameter-gradient accumulation through the quad exchange (see the file header).  acc has the L = 4 layout:
// lane j of a quad owns row j.  (A,B,Q)-bar are weighted with al_ab, C^T-bar with al_c (sweep 3 pairs the A/B/Q
// cotangents of step t+1 with the C cotangent of step t).  dal_ab / dal_c receive this SEQUENCE's contractions
// <Xbar, X_k> with the base matrices (the dalpha parts).  Inactive lanes (tail of the batch) publish zeros.
// ---------------------------------------------------------------------------------------
template <class C, class C4>
__device__ __forceinline__ void seq_acc_exchange(float* ax, int lane, bool active, const float* base, const float (&al_ab)[C::K],
                                                 const float (&al_c)[C::K], const float (&Ab)[4][4], const float (&Bb)[4][4],
                                                 const float (&Ctb)[4][C::P], const float (&Qb)[4][4], bool with_ab,
                                                 GradAcc<C4>& acc, float (&dal_ab)[C::K], float (&dal_c)[C::K]) {
  using PL = SeqBwdPlan<C>;
  using GA = GradAcc<C4>;
  constexpr int K = C::K, P = C::P, N = C::N, M = C::M;
  // own contractions with the base matrices (dalpha)
#pragma unroll
  for (int k = 0; k < K; ++k) {
    float s = 0.f, sc = 0.f;
#pragma unroll
    for (int r = 0; r < N; ++r) {
      if (with_ab) {
        float row[N];
        load_row<N>(base + Base<C>::oA + (k * N + r) * Base<C>::ldA, row);
        s = kv_dot<N>(Ab[r], row, s);
        float rowb[M];
        load_row<M>(base + Base<C>::oB + (k * N + r) * Base<C>::ldB, rowb);
        s = kv_dot<M>(Bb[r], rowb, s);
        if constexpr (C::QPM) {
          float rowq[N];
          load_row<N>(base + Base<C>::oQ + (k * N + r) * Base<C>::ldQ, rowq);
          s = kv_dot<N>(Qb[r], rowq, s);
        }
      }
      if constexpr (!C::CSH) {
        float rowc[P];
        load_row<P>(base + Base<C>::oCt + (k * N + r) * Base<C>::ldCt, rowc);
#pragma unroll
        for (int q = 0; q < P; ++q) sc = fmaf(Ctb[r][q], rowc[q], sc);
      }
    }
    dal_ab[k] = s;
    dal_c[k] = sc;
  }
  // publish the step's cotangents, then accumulate row (lane & 3) of the quad's four sequences
  float* mine = ax + lane * PL::ax_row;
  __syncwarp();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float t4[4];
    if (with_ab) {
#pragma unroll
      for (int j = 0; j < 4; ++j) t4[j] = active ? Ab[r][j] : 0.f;
      store_row<4>(mine + PL::ax_A + 4 * r, t4);
#pragma unroll
      for (int j = 0; j < 4; ++j) t4[j] = active ? Bb[r][j] : 0.f;
      store_row<4>(mine + PL::ax_B + 4 * r, t4);
      if constexpr (C::QPM) {
#pragma unroll
        for (int j = 0; j < 4; ++j) t4[j] = active ? Qb[r][j] : 0.f;
        store_row<4>(mine + PL::ax_Q + 4 * r, t4);
      }
    }
#pragma unroll
    for (int q = 0; q < P; ++q) mine[PL::ax_C + r * P + q] = active ? Ctb[r][q] : 0.f;   // C^T-bar rows, row-major [4][P]
  }
#pragma unroll
  for (int k = 0; k < K; ++k) { mine[PL::ax_al + k] = active ? al_ab[k] : 0.f; mine[PL::ax_al + K + k] = active ? al_c[k] : 0.f; }
  __syncwarp();
  const int j = lane & 3, q0 = lane & ~3;
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const float* other = ax + (q0 + s) * PL::ax_row;
    float cr[P], a_ab[K], a_c[K];
#pragma unroll
    for (int q = 0; q < P; ++q) cr[q] = other[PL::ax_C + j * P + q];
#pragma unroll
    for (int k = 0; k < K; ++k) { a_ab[k] = other[PL::ax_al + k]; a_c[k] = other[PL::ax_al + K + k]; }
    if constexpr (C::CSH) {
#pragma unroll
      for (int q = 0; q < P; ++q) acc.v[GA::oCt + q] += cr[q];
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < P; ++q) acc.v[GA::oCt + k * P + q] = fmaf(a_c[k], cr[q], acc.v[GA::oCt + k * P + q]);
    }
    if (with_ab) {
      float ar[4], br[4];
      load_row<4>(other + PL::ax_A + 4 * j, ar);
      load_row<4>(other + PL::ax_B + 4 * j, br);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        kv_fma2(acc.v[GA::oA + k * N + 0], acc.v[GA::oA + k * N + 1], a_ab[k], a_ab[k], ar[0], ar[1]);
        kv_fma2(acc.v[GA::oA + k * N + 2], acc.v[GA::oA + k * N + 3], a_ab[k], a_ab[k], ar[2], ar[3]);
        kv_fma2(acc.v[GA::oB + k * M + 0], acc.v[GA::oB + k * M + 1], a_ab[k], a_ab[k], br[0], br[1]);
        kv_fma2(acc.v[GA::oB + k * M + 2], acc.v[GA::oB + k * M + 3], a_ab[k], a_ab[k], br[2], br[3]);
      }
      if constexpr (C::QPM) {
        float qr[4];
        load_row<4>(other + PL::ax_Q + 4 * j, qr);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          kv_fma2(acc.v[GA::oQ + k * N + 0], acc.v[GA::oQ + k * N + 1], a_ab[k], a_ab[k], qr[0], qr[1]);
          kv_fma2(acc.v[GA::oQ + k * N + 2], acc.v[GA::oQ + k * N + 3], a_ab[k], a_ab[k], qr[2], qr[3]);
        }
      }
    }
  }
}

template <class C>
__global__ void __launch_bounds__(64) k_seq_bwd(Args a, BwdArgs w, BasePtrs bp, const __grid_constant__ SeqBwdMaps mp,
                                                const float* __restrict__ g_elbo, const float* __restrict__ terms,
                                                float* __restrict__ partials, double* __restrict__ elbo_partials) {
  static_assert(C::L == 1 && C::N == 4 && C::M == 4, "thread-per-sequence kernels: z_dim = u_dim = 4");
  constexpr int N = C::N, P = C::P, M = C::M, K = C::K, R = C::R;
  using C4 = Cfg<C::N, C::P, C::M, C::K, 4, C::QPM, C::CSH>;
  using GA = GradAcc<C4>;
  using PL = SeqBwdPlan<C>;
  using TV = typename PL::TV;
  using TM = typename PL::TM;
  constexpr int MAXW = 2;
  extern __shared__ unsigned char seq_smem_raw[];
  __shared__ __align__(8) unsigned long long bars[MAXW][6];
  __shared__ double nred[MAXW];
  __shared__ double ered[MAXW][5];
  unsigned char* sm = seq_smem_raw + ((512u - (tma::s32(seq_smem_raw) & 511u)) & 511u);
  float* base = reinterpret_cast<float*>(sm);
  stage_base<C>(base, bp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const bool has_du = w.dU != nullptr;
  unsigned char* body = sm + kv_align_up((int)sizeof(float) * Base<C>::total, 512);
  unsigned char* wsm = body + warp * PL::warp_bytes(has_du);
  const int b0 = blockIdx.x * blockDim.x + warp * 32;
  const int b = b0 + lane;
  const bool active = b < a.B;
  const bool warp_on = b0 < a.B;
  const int T = a.T, nchunk = T >> 2;
  const bool has_u = a.U != nullptr, has_m = a.mask != nullptr;
  const Group<1, R> g{0, 0xffffffffu};
  const TileRef NT{base, 0};   // L = 1: tiles are never dereferenced
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) tma::bar_init(tma::s32(&bars[warp][i]), 1);
    tma::bar_init_fence();
  }
  __syncwarp();
  SeqBar bar_in{tma::s32(&bars[warp][0]), 0u}, bar_s0{tma::s32(&bars[warp][1]), 0u}, bar_s1{tma::s32(&bars[warp][2]), 0u};
  SeqBar bar_g0{tma::s32(&bars[warp][3]), 0u}, bar_g1{tma::s32(&bars[warp][4]), 0u};

  // normaliser (see k_bwd): the forward kernel's per-CTA mask sums, same fixed-order fp64 sum in every CTA
  float inv_norm = 1.0f;
  if (w.mask_part) {
    double v = 0.0;
    for (int i = threadIdx.x; i < w.n_mask_part; i += blockDim.x) v += (double)w.mask_part[i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) nred[warp] = v;
    __syncthreads();
    double tot = 0.0;
    for (int wq = 0; wq < nwarps; ++wq) tot += nred[wq];
    inv_norm = (float)(1.0 / (tot < 1.0 ? 1.0 : tot));
  }
  const float c = (*g_elbo) * (w.with_elbo ? inv_norm : terms[6]);
  const CholOpt co = chol_opt(a, w.jitter);
  const uint64_t pol_ef = tma::policy_evict_first(), pol_el = tma::policy_evict_last();

  GA acc;
  acc.zero();
  acc.on = true;
  double el[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  bool ok = true, ok_s = true, ok_q = true;
  int bad0 = 0;
  unsigned char* ib = wsm + PL::oIn;
  float* ax = reinterpret_cast<float*>(wsm + PL::oAx);
  unsigned char* sw = wsm + PL::oSw;

  auto issue_in = [&](int cidx) {   // elected lane: the input streams of chunk cidx
    tma::bar_expect(bar_in.addr, PL::in_tx(has_u, has_m));
    tma::load2d(tma::s32(ib + PL::in_Y), &mp.Y, 4 * cidx * P, b0, bar_in.addr, pol_el);
    if (has_u) tma::load2d(tma::s32(ib + PL::in_U), &mp.U, 4 * cidx * M, b0, bar_in.addr, pol_el);
    tma::load2d(tma::s32(ib + PL::in_al), &mp.alpha, 4 * cidx * K, b0, bar_in.addr, pol_el);
    if (has_m) tma::load2d(tma::s32(ib + PL::in_m), &mp.mask, 4 * cidx, b0, bar_in.addr, pol_el);
  };

  if (warp_on) {
    // ================================================================== sweep 3: ELBO + smoother adjoint, t = 0..T-1
    unsigned char* st0 = sw;
    unsigned char* st1 = sw + PL::s3_bytes;
    unsigned char* o3 = sw + 2 * PL::s3_bytes;
    unsigned char* gc = o3 + PL::o3_bytes;
    auto issue_s3 = [&](int t) {   // states step t needs: (t+1) rows and Sigma_f(t); buffer t & 1
      unsigned char* buf = (t & 1) ? st1 : st0;
      const uint32_t bar = (t & 1) ? bar_s1.addr : bar_s0.addr;
      tma::bar_expect(bar, PL::s3_tx);
      tma::load3d(tma::s32(buf + PL::s3_Ss), &mp.Sig_s, 0, t + 1, b0, bar, pol_ef);
      tma::load3d(tma::s32(buf + PL::s3_Sp), &mp.Sig_p, 0, t + 1, b0, bar, pol_el);
      tma::load3d(tma::s32(buf + PL::s3_Sf), &mp.Sig_f, 0, t, b0, bar, pol_el);
      tma::load3d(tma::s32(buf + PL::s3_ms), &mp.mu_s, 0, t + 1, b0, bar, pol_ef);
      tma::load3d(tma::s32(buf + PL::s3_mp), &mp.mu_p, 0, t + 1, b0, bar, pol_el);
      tma::load3d(tma::s32(buf + PL::s3_ep), &mp.eps, 0, t + 1, b0, bar, pol_ef);
    };
    if (tma::elect_one()) {
      issue_in(0);
      // prologue: Sigma_s(0), mu_s(0), eps(0) into the slots of buffer 1 (first completion of its barrier)
      tma::bar_expect(bar_s1.addr, (uint32_t)(PL::TMB + 2 * PL::TVB));
      tma::load3d(tma::s32(st1 + PL::s3_Ss), &mp.Sig_s, 0, 0, b0, bar_s1.addr, pol_ef);
      tma::load3d(tma::s32(st1 + PL::s3_ms), &mp.mu_s, 0, 0, b0, bar_s1.addr, pol_ef);
      tma::load3d(tma::s32(st1 + PL::s3_ep), &mp.eps, 0, 0, b0, bar_s1.addr, pol_ef);
      if (T > 1) issue_s3(0);
    }
    ElboConst<C> ec;
    bad0 = elbo_const<C>(g, base, NT, co, ec);
    RegView<N, N> LQc_v{ec.LQ};

    ElboStep<C> es;
    float eps_cur[N];
    {
      bar_s1.wait();
      float Ss0[R][N], ms0[R];
      TM::ld_mat(reinterpret_cast<const float*>(st1 + PL::s3_Ss), lane, Ss0);
      TV::ld(reinterpret_cast<const float*>(st1 + PL::s3_ms), lane, 0, ms0);
      TV::ld(reinterpret_cast<const float*>(st1 + PL::s3_ep), lane, 0, eps_cur);
      ok_s = elbo_sample_rows<C>(g, NT, NT, Ss0, ms0, co, eps_cur, es) && ok_s;
    }
    bar_in.wait();
    StepIn<C> in;
    seqb_read_step<C>(ib, lane, 0, has_u, has_m, in);
    float xbar[N], Ssb[R][N], msb[R], dal_carry[K], du_carry[M];
    KV_UNROLL for (int r = 0; r < R; ++r) { xbar[r] = 0.f; msb[r] = 0.f; KV_UNROLL for (int j = 0; j < N; ++j) Ssb[r][j] = 0.f; }
    KV_UNROLL for (int k = 0; k < K; ++k) dal_carry[k] = 0.f;
    KV_UNROLL for (int j = 0; j < M; ++j) du_carry[j] = 0.f;
    __syncwarp();   // prologue reads of buffer 1 are done before step 1's loads are issued into it

#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      const bool has_next = (t + 1 < T);
      const int s = t & 3;
      // inputs of step t+1 (they become `in` of the next iteration).  The chunk buffer holds the chunk of t+1: it was
      // re-armed after the reads of step t-1 when t+1 starts a new chunk.
      StepIn<C> in1;
      if (has_next) {
        if (((t + 1) & 3) == 0) bar_in.wait();
        seqb_read_step<C>(ib, lane, (t + 1) & 3, has_u, has_m, in1);
      } else {
        in1 = in;
      }
      // prefetch: states of step t+1 into the other buffer; the next input chunk once step t+1 is its last reader
      __syncwarp();
      if (tma::elect_one()) {
        if (t + 2 < T) issue_s3(t + 1);
        if (((t + 1) & 3) == 3 && t + 2 < T) issue_in((t + 2) >> 2);
      }
      float A1[R][N];
      mix_A<C>(base, in1.al, 0, A1);
      float Ab[R][N], Bb[R][M], Qb[R][N], Ctb[R][P];
      KV_UNROLL for (int r = 0; r < R; ++r) {
        KV_UNROLL for (int j = 0; j < N; ++j) { Ab[r][j] = 0.f; Qb[r][j] = 0.f; }
        KV_UNROLL for (int j = 0; j < M; ++j) Bb[r][j] = 0.f;
        KV_UNROLL for (int q = 0; q < P; ++q) Ctb[r][q] = 0.f;
      }
      float dy[P], du1[M], xbar_next[N];
      KV_UNROLL for (int q = 0; q < P; ++q) dy[q] = 0.f;
      KV_UNROLL for (int j = 0; j < M; ++j) du1[j] = 0.f;
      KV_UNROLL for (int r = 0; r < R; ++r) xbar_next[r] = 0.f;
      ElboStep<C> es1;
      float Ss1[R][N], ms1[R], Sp1[R][N], mp1[R], Sf[R][N], eps1[N];
      KV_UNROLL for (int j = 0; j < N; ++j) eps1[j] = 0.f;
      if (has_next) {
        if (t & 1) bar_s1.wait(); else bar_s0.wait();
        const unsigned char* sb = (t & 1) ? st1 : st0;
        TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s3_Ss), lane, Ss1);
        TV::ld(reinterpret_cast<const float*>(sb + PL::s3_ms), lane, 0, ms1);
        TV::ld(reinterpret_cast<const float*>(sb + PL::s3_ep), lane, 0, eps1);
      }
      // ---------------------------------------------------------------- A.3 ELBO adjoint at t
      float zbar[N];
      KV_UNROLL for (int r = 0; r < R; ++r) zbar[r] = xbar[r];
      float v_tr = 0.f, v_em = 0.f, v_in = 0.f, v_en = 0.f;
      if (has_next) {
        ok_s = elbo_sample_rows<C>(g, NT, NT, Ss1, ms1, co, eps1, es1) && ok_s;
        float B1[R][M];
        mix_B<C>(base, in1.al, 0, B1);
        float x[N];
        KV_UNROLL for (int r = 0; r < R; ++r) {
          float s1 = 0.f, s2 = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) s1 = fmaf(A1[r][j], es.z[j], s1);
          KV_UNROLL for (int j = 0; j < M; ++j) s2 = fmaf(B1[r][j], in1.u[j], s2);
          x[r] = es1.z_own[r] - (s1 + s2);
        }
        if constexpr (C::QPM) {
          float Q1[R][N], Qs[R][N], LQ[R][N], invdQ[N], dgQ[R];
          mix_Q<C>(base, in1.al, 0, Q1);
          sym_jitter_rows<C>(g, Q1, NT, co.diag_q ? 0.f : co.jq, Qs);
          unsigned clq;
          ok_q = chol_dist_opt<1, R>(g, Qs, LQ, invdQ, dgQ, co.diag_q, clq) && ok_q;
          RegView<N, N> LQ_v{LQ};
          solve_vec_l<N>(x, LQ_v, invdQ);
          if (w.with_elbo) {
            float q2 = 0.f, ld = 0.f;
            KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(x[j], x[j], q2);
            KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(dgQ[r]);
            v_tr = -0.5f * (N * KV_LOG2PI + q2) - ld;
          }
          solve_vec_lt<N>(x, LQ_v, invdQ);        // x := q = Qj^-1 x
          float Qi[R][N];
          KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Qi[r][j] = (r == j) ? 1.f : 0.f;
          solve_rows_llt<R, N>(Qi, LQ_v, invdQ);
          KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) {
            const bool live = !co.diag_q || (r == j && !((clq >> j) & 1u));   // diagonal fallback: see kvae_bwd.cuh
            Qb[r][j] += live ? 0.5f * c * (x[r] * x[j] - Qi[r][j]) : 0.f;
          }
        } else {
          solve_vec_l<N>(x, LQc_v, ec.invdQ);
          if (w.with_elbo) {
            float q2 = 0.f;
            KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(x[j], x[j], q2);
            v_tr = -0.5f * (N * KV_LOG2PI + q2) - ec.logdetQ;
          }
          solve_vec_lt<N>(x, LQc_v, ec.invdQ);
        }
        KV_UNROLL for (int j = 0; j < N; ++j) xbar_next[j] = -c * x[j];
        {
          RegView<N, N> A1_v{A1};
          float av[R];
          matTvec_own<C>(A1_v, 0, xbar_next, av);
          KV_UNROLL for (int r = 0; r < R; ++r) zbar[r] -= av[r];
        }
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) Ab[r][j] = fmaf(-xbar_next[r], es.z[j], Ab[r][j]);
          KV_UNROLL for (int j = 0; j < M; ++j) Bb[r][j] = fmaf(-xbar_next[r], in1.u[j], Bb[r][j]);
          KV_UNROLL for (int j = 0; j < M; ++j) du1[j] = fmaf(-B1[r][j], xbar_next[r], du1[j]);
        }
      }
      {   // emission at t
        float Ct[R][P];
        mix_Ct<C>(base, in.al, 0, Ct);
        float e[P];
        KV_UNROLL for (int q = 0; q < P; ++q) {
          float sacc = 0.f;
          KV_UNROLL for (int r = 0; r < R; ++r) sacc = fmaf(Ct[r][q], es.z_own[r], sacc);
          e[q] = in.y[q] - sacc;
        }
        RegView<P, P> LR_v{ec.LR};
        solve_vec_l<P>(e, LR_v, ec.invdR);
        if (w.with_elbo) {
          float q2 = 0.f;
          KV_UNROLL for (int q = 0; q < P; ++q) q2 = fmaf(e[q], e[q], q2);
          v_em = (-0.5f * (P * KV_LOG2PI + q2) - ec.logdetR) * in.m;
        }
        solve_vec_lt<P>(e, LR_v, ec.invdR);
        KV_UNROLL for (int q = 0; q < P; ++q) { e[q] = -c * in.m * e[q]; dy[q] += e[q]; }
        KV_UNROLL for (int r = 0; r < R; ++r) {
          float sacc = 0.f;
          KV_UNROLL for (int q = 0; q < P; ++q) {
            Ctb[r][q] = fmaf(-es.z_own[r], e[q], Ctb[r][q]);
            sacc = fmaf(Ct[r][q], e[q], sacc);
          }
          zbar[r] -= sacc;
        }
      }
      if (t == 0) {   // zbar_0 += -c Sigma0^-1 (z_0 - mu0)
        float S0[R][N], L0[R][N], invd0[N], dg0[R];
        copy_rows<C, N>(base + Base<C>::oS0, 0, S0);
        ok = chol_dist<1, R>(g, S0, L0, invd0, dg0) && ok;
        RegView<N, N> L0_v{L0};
        float wv[N];
        KV_UNROLL for (int j = 0; j < N; ++j) wv[j] = es.z[j] - base[Base<C>::oMu0 + j];
        solve_vec_l<N>(wv, L0_v, invd0);
        if (w.with_elbo) {
          float q2 = 0.f, ld = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) q2 = fmaf(wv[j], wv[j], q2);
          KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(dg0[r]);
          v_in = -0.5f * (N * KV_LOG2PI + q2) - ld;
        }
        solve_vec_lt<N>(wv, L0_v, invd0);
        KV_UNROLL for (int r = 0; r < R; ++r) zbar[r] = fmaf(-c, wv[r], zbar[r]);
      }
      if (w.with_elbo) {
        float ld = 0.f, e2 = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) ld += logf(es.dg[r]);
        KV_UNROLL for (int j = 0; j < N; ++j) e2 = fmaf(eps_cur[j], eps_cur[j], e2);
        v_en = 0.5f * e2 + 0.5f * N * KV_LOG2PI + ld;
        if (active) {
          el[0] += (double)v_tr; el[1] += (double)v_em; el[2] += (double)v_in; el[3] += (double)v_en;
          el[4] += (double)in.m;
        }
      }
      // mu_s-bar += zbar ; Sigma_s-bar += sym(Ls^-T Phi Ls^-1)
      KV_UNROLL for (int r = 0; r < R; ++r) msb[r] += zbar[r];
      {
        RegView<N, N> Ls_v{es.Ls};
        float v_own[R];
        matTvec_own<C>(Ls_v, 0, zbar, v_own);
        float Phi[R][N];
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) {
          const float ve = v_own[r] * eps_cur[j];
          const bool live_d = !co.diag_s || !((es.clamped >> r) & 1u);
          Phi[r][j] = (j < r) ? (co.diag_s ? 0.f : ve) : ((j == r && live_d) ? 0.5f * (ve + c) : 0.f);
        }
        solve_rows_l<R, N>(Phi, Ls_v, es.invd);
        float Zt[R][N];
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Zt[r][j] = Phi[j][r];
        solve_rows_l<R, N>(Zt, Ls_v, es.invd);
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Ssb[r][j] += 0.5f * (Zt[r][j] + Zt[j][r]);
      }
      // ---------------------------------------------------------------- A.4 smoother adjoint at t
      float Sfb[R][N], mfb[R], Ssb1[R][N], msb1[R], Spb1[R][N], mpb1[R];
      if (has_next) {
        const unsigned char* sb = (t & 1) ? st1 : st0;
        TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s3_Sp), lane, Sp1);
        TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s3_Sf), lane, Sf);
        TV::ld(reinterpret_cast<const float*>(sb + PL::s3_mp), lane, 0, mp1);
        float J[R][N], LU[R][N], invu[N];
        ok = smoother_gain<C>(g, NT, NT, Sf, A1, Sp1, J, LU, invu) && ok;
        RegView<N, N> A1_v{A1}, LU_v{LU}, J_v{J};
        float D[R][N], d[N];
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) D[r][j] = Ss1[r][j] - Sp1[r][j];
          d[r] = ms1[r] - mp1[r];
        }
        float Gs[R][N];
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Gs[r][j] = 0.5f * (Ssb[r][j] + Ssb[j][r]);
        KV_UNROLL for (int r = 0; r < R; ++r) { mfb[r] = msb[r]; KV_UNROLL for (int j = 0; j < N; ++j) Sfb[r][j] = Gs[r][j]; }
        float GJ[R][N];
        mm_RS<false>(Gs, J_v, GJ);
        float Jb[R][N];
        {
          RegView<N, N> D_v{D};
          mm_RSt<false>(GJ, D_v, Jb);
          mm_RS<true>(GJ, D_v, Jb);
        }
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Jb[r][j] = fmaf(msb[r], d[j], Jb[r][j]);
        float Db[R][N], db[R];
        {
          RegView<N, N> GJ_v{GJ};
          mm_StS<false, R, N, N>(J_v, 0, GJ_v, Db);
          matTvec_own<C>(J_v, 0, msb, db);
        }
        KV_UNROLL for (int r = 0; r < R; ++r) { msb1[r] = db[r]; KV_UNROLL for (int j = 0; j < N; ++j) Ssb1[r][j] = Db[r][j]; }
        solve_rows_lut<R, N>(Jb, LU_v, invu);                       // W-bar = J-bar Sp1^-T
        {
          RegView<N, N> Wb_v{Jb};
          float JW[R][N];
          mm_StS<false, R, N, N>(J_v, 0, Wb_v, JW);
          KV_UNROLL for (int r = 0; r < R; ++r) {
            KV_UNROLL for (int j = 0; j < N; ++j) Spb1[r][j] = -(Db[r][j] + JW[r][j]);
            mpb1[r] = -db[r];
          }
          mm_RS<true>(Jb, A1_v, Sfb);
          RegView<N, N> Sf_v{Sf};
          mm_StS<true, R, N, N>(Wb_v, 0, Sf_v, Ab);
        }
      } else {
        KV_UNROLL for (int r = 0; r < R; ++r) {
          mfb[r] = msb[r]; msb1[r] = 0.f; mpb1[r] = 0.f;
          KV_UNROLL for (int j = 0; j < N; ++j) { Sfb[r][j] = Ssb[r][j]; Ssb1[r][j] = 0.f; Spb1[r][j] = 0.f; }
        }
      }
      // ---------------------------------------------------------------- A.0 adjoint of this sweep's parts
      // (A,B,Q)-bar belong to step t+1 (weights alpha_{t+1}), C^T-bar to step t (alpha_t): one exchange carries both
      float dal_next[K], dal_c[K];
      seq_acc_exchange<C, C4>(ax, lane, active, base, in1.al, in.al, Ab, Bb, Ctb, Qb, has_next, acc, dal_next, dal_c);
      // ---------------------------------------------------------------- stores of step t
      tma::wait_read0();
      __syncwarp();
      TM::st_mat(reinterpret_cast<float*>(o3 + PL::o3_Sf), lane, Sfb);
      TV::st(reinterpret_cast<float*>(o3 + PL::o3_mf), lane, 0, mfb);
      if (has_next) {
        TM::st_mat(reinterpret_cast<float*>(o3 + PL::o3_Sp), lane, Spb1);
        TV::st(reinterpret_cast<float*>(o3 + PL::o3_mp), lane, 0, mpb1);
      }
      {
        float* gy = reinterpret_cast<float*>(gc + PL::gc_Y) + lane * 4 * P + s * P;
        KV_UNROLL for (int q = 0; q < P; ++q) gy[q] = dy[q];
        float* ga = reinterpret_cast<float*>(gc + PL::gc_al) + lane * 4 * K + s * K;
        KV_UNROLL for (int k = 0; k < K; ++k) ga[k] = dal_carry[k] + dal_c[k];
        if (has_du) store_row<M>(reinterpret_cast<float*>(gc + PL::gc_U) + lane * 4 * M + s * M, du_carry);
      }
      tma::fence_async();
      __syncwarp();
      if (tma::elect_one()) {
        tma::store3d(&mp.w_Sig_f, 0, t, b0, tma::s32(o3 + PL::o3_Sf), pol_el);
        tma::store3d(&mp.w_mu_f, 0, t, b0, tma::s32(o3 + PL::o3_mf), pol_el);
        if (has_next) {
          tma::store3d(&mp.w_Sig_p, 0, t + 1, b0, tma::s32(o3 + PL::o3_Sp), pol_el);
          tma::store3d(&mp.w_mu_p, 0, t + 1, b0, tma::s32(o3 + PL::o3_mp), pol_el);
        }
        if (s == 3) {
          const int cidx = t >> 2;
          tma::store2d(&mp.dY, 4 * cidx * P, b0, tma::s32(gc + PL::gc_Y), pol_el);
          tma::store2d(&mp.dalpha, 4 * cidx * K, b0, tma::s32(gc + PL::gc_al), pol_el);
          if (has_du) tma::store2d(&mp.dU, 4 * cidx * M, b0, tma::s32(gc + PL::gc_U), pol_el);
        }
        tma::commit();
      }
      // carry
      KV_UNROLL for (int k = 0; k < K; ++k) dal_carry[k] = dal_next[k];
      KV_UNROLL for (int j = 0; j < M; ++j) du_carry[j] = du1[j];
      KV_UNROLL for (int r = 0; r < R; ++r) {
        xbar[r] = xbar_next[r];
        msb[r] = msb1[r];
        KV_UNROLL for (int j = 0; j < N; ++j) Ssb[r][j] = Ssb1[r][j];
      }
      if (has_next) es = es1;
      KV_UNROLL for (int j = 0; j < N; ++j) eps_cur[j] = eps1[j];
      in = in1;
    }

    // ================================================================== sweep 4: filter + mixing adjoint, t = T-1..0
    // sweep 3's scratch and partial gradients must have LANDED before they are read back
    tma::wait_all0();
    __syncwarp();
    const int gcb = PL::gc_bytes(has_du);
    unsigned char* q0 = sw;
    unsigned char* q1 = sw + PL::s4_bytes;
    unsigned char* g0 = sw + 2 * PL::s4_bytes;
    unsigned char* g1 = g0 + gcb;
    auto issue_s4 = [&](int i) {   // iteration i handles t = T-1-i; buffer i & 1
      const int t = T - 1 - i;
      unsigned char* buf = (i & 1) ? q1 : q0;
      const uint32_t bar = (i & 1) ? bar_s1.addr : bar_s0.addr;
      const bool first = (t == 0);   // no scratch for (Sigma_p, mu_p)-bar(0), no filtered belief before step 0
      tma::bar_expect(bar, first ? (uint32_t)(2 * PL::TMB + 2 * PL::TVB) : (uint32_t)(4 * PL::TMB + 4 * PL::TVB));
      tma::load3d(tma::s32(buf + PL::s4_Sfb), &mp.w_Sig_f, 0, t, b0, bar, pol_ef);
      tma::load3d(tma::s32(buf + PL::s4_Sp), &mp.Sig_p, 0, t, b0, bar, pol_ef);
      tma::load3d(tma::s32(buf + PL::s4_mfb), &mp.w_mu_f, 0, t, b0, bar, pol_ef);
      tma::load3d(tma::s32(buf + PL::s4_mp), &mp.mu_p, 0, t, b0, bar, pol_ef);
      if (!first) {
        tma::load3d(tma::s32(buf + PL::s4_Spb), &mp.w_Sig_p, 0, t, b0, bar, pol_ef);
        tma::load3d(tma::s32(buf + PL::s4_Sv), &mp.Sig_f, 0, t - 1, b0, bar, pol_ef);
        tma::load3d(tma::s32(buf + PL::s4_mpb), &mp.w_mu_p, 0, t, b0, bar, pol_ef);
        tma::load3d(tma::s32(buf + PL::s4_mv), &mp.mu_f, 0, t - 1, b0, bar, pol_ef);
      }
    };
    auto issue_gc = [&](int j) {   // gradient chunk nchunk-1-j; buffer j & 1
      const int cidx = nchunk - 1 - j;
      unsigned char* buf = (j & 1) ? g1 : g0;
      const uint32_t bar = (j & 1) ? bar_g1.addr : bar_g0.addr;
      tma::bar_expect(bar, PL::gc_tx(has_du));
      tma::load2d(tma::s32(buf + PL::gc_Y), &mp.dY, 4 * cidx * P, b0, bar, pol_ef);
      tma::load2d(tma::s32(buf + PL::gc_al), &mp.dalpha, 4 * cidx * K, b0, bar, pol_ef);
      if (has_du) tma::load2d(tma::s32(buf + PL::gc_U), &mp.dU, 4 * cidx * M, b0, bar, pol_ef);
    };
    // the input chunk buffer still holds the LAST chunk (sweep 3 ended there); the state / gradient barriers restart
    if (tma::elect_one()) {
      issue_s4(0);
      issue_gc(0);
    }
    const float* Rm = base + Base<C>::oR;
    float Sf_carry[R][N], mf_carry[R];
    KV_UNROLL for (int r = 0; r < R; ++r) { mf_carry[r] = 0.f; KV_UNROLL for (int j = 0; j < N; ++j) Sf_carry[r][j] = 0.f; }
    int jg = 0;
#pragma unroll 1
    for (int i = 0; i < T; ++i) {
      const int t = T - 1 - i;
      const int s = t & 3;
      if (s == 3 && i > 0) bar_in.wait();                 // chunk of t was re-armed during step t+1
      seqb_read_step<C>(ib, lane, s, has_u, has_m, in);
      if (s == 3) tma::wait_read0();   // the gradient tile re-armed below was the source of a store one step ago
      __syncwarp();
      if (tma::elect_one()) {
        if (t > 0) issue_s4(i + 1);
        if (s == 0 && t > 0) issue_in((t - 1) >> 2);      // step t was the chunk's last reader
        if (s == 3 && t >= 4) issue_gc(jg + 1);           // first step of this gradient chunk: fetch the next one
      }
      float A[R][N], Bm[R][M], Ct[R][P];
      mix_A<C>(base, in.al, 0, A);
      mix_B<C>(base, in.al, 0, Bm);
      mix_Ct<C>(base, in.al, 0, Ct);
      if (i & 1) bar_s1.wait(); else bar_s0.wait();
      const unsigned char* sb = (i & 1) ? q1 : q0;
      float Sp[R][N], mup[R];
      TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s4_Sp), lane, Sp);
      TV::ld(reinterpret_cast<const float*>(sb + PL::s4_mp), lane, 0, mup);
      // recompute the gain
      RegView<N, P> Ct_v{Ct};
      GainOut<C> go;
      ok = gain<C>(g, base, Sp, mup, Ct, Ct_v, in.y, in.m, go) && ok;
      float G[R][N];
      mm_RSt<false>(go.Kg, Ct_v, G);
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) G[r][j] = ((r == j) ? 1.0f : 0.0f) - G[r][j];
      float Gf[R][N], mfb[R];
      {
        float Sfb[R][N];
        TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s4_Sfb), lane, Sfb);
        TV::ld(reinterpret_cast<const float*>(sb + PL::s4_mfb), lane, 0, mfb);
        KV_UNROLL for (int r = 0; r < R; ++r) {
          mfb[r] += mf_carry[r];
          KV_UNROLL for (int j = 0; j < N; ++j) Sfb[r][j] += Sf_carry[r][j];
        }
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Gf[r][j] = 0.5f * (Sfb[r][j] + Sfb[j][r]);
      }
      float Spb[R][N], mpb[R];
      if (t > 0) {
        TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s4_Spb), lane, Spb);
        TV::ld(reinterpret_cast<const float*>(sb + PL::s4_mpb), lane, 0, mpb);
      } else {
        KV_UNROLL for (int r = 0; r < R; ++r) { mpb[r] = 0.f; KV_UNROLL for (int j = 0; j < N; ++j) Spb[r][j] = 0.f; }
      }
      RegView<N, N> G_v{G}, Sp_v{Sp};
      float GfG[R][N];
      mm_RS<false>(Gf, G_v, GfG);
      float Gb[R][N];
      mm_RSt<false>(GfG, Sp_v, Gb);
      mm_RS<true>(GfG, Sp_v, Gb);
      {
        RegView<N, N> GfG_v{GfG};
        mm_StS<true, R, N, N>(G_v, 0, GfG_v, Spb);
      }
      RegView<N, P> K_v{go.Kg};
      float GfK[R][P];
      mm_RS<false>(Gf, K_v, GfK);
      float Kb[R][P];
      mm_RS<false>(Gb, Ct_v, Kb);
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
        float sacc = 0.f;
        KV_UNROLL for (int q2 = 0; q2 < P; ++q2) sacc = fmaf(GfK[r][q2], Rm[q * P + q2] + Rm[q2 * P + q], sacc);
        Kb[r][q] = sacc - Kb[r][q] + mfb[r] * go.r[q];
      }
      float Ctb[R][P];
      {
        RegView<N, N> Gb_v{Gb};
        mm_StS<false, R, N, P>(Gb_v, 0, K_v, Ctb);
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Ctb[r][q] = -Ctb[r][q];
      }
      float K0b[R][P];
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) K0b[r][q] = in.m * Kb[r][q];
      float rb[P], M1t[P][P];
      KV_UNROLL for (int q = 0; q < P; ++q) {
        float sacc = 0.f;
        KV_UNROLL for (int r = 0; r < R; ++r) sacc = fmaf(go.Kg[r][q], mfb[r], sacc);
        rb[q] = sacc;
        KV_UNROLL for (int q2 = 0; q2 < P; ++q2) {
          float s2 = 0.f;
          KV_UNROLL for (int r = 0; r < R; ++r) s2 = fmaf(K0b[r][q], go.K0[r][q2], s2);
          M1t[q2][q] = s2;                                           // M1[q][q2] = (K0-bar^T K0)[q][q2]
        }
      }
      float Sb[P][P];
      {
        RegView<P, P> Lc_v{go.Lc};
        solve_rows_llt<P, P>(M1t, Lc_v, go.invd);
        KV_UNROLL for (int q = 0; q < P; ++q) KV_UNROLL for (int q2 = 0; q2 < P; ++q2) Sb[q][q2] = -0.5f * (M1t[q2][q] + M1t[q][q2]);
      }
      float Pb[R][P];
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) Pb[r][q] = K0b[r][q];
      {
        RegView<P, P> Lc_v{go.Lc};
        solve_rows_llt<R, P>(Pb, Lc_v, go.invd);
      }
      {
        float P2[R][P];
        mm_StS<false, R, N, P>(Sp_v, 0, Ct_v, P2);
        RegView<N, P> Pb_v{Pb};
        mm_StS<true, R, N, P>(Sp_v, 0, Pb_v, Ctb);
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
          float sacc = Ctb[r][q];
          KV_UNROLL for (int q2 = 0; q2 < P; ++q2) {
            sacc = fmaf(go.Pm[r][q2], Sb[q][q2], sacc);
            sacc = fmaf(P2[r][q2], Sb[q2][q], sacc);
          }
          Ctb[r][q] = sacc - mup[r] * rb[q];
        }
      }
      {
        float T3m[R][P];
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
          float sacc = Pb[r][q];
          KV_UNROLL for (int q2 = 0; q2 < P; ++q2) sacc = fmaf(Ct[r][q2], Sb[q2][q], sacc);
          T3m[r][q] = sacc;
        }
        mm_RSt<true>(T3m, Ct_v, Spb);
      }
      KV_UNROLL for (int r = 0; r < R; ++r) {
        float sacc = 0.f;
        KV_UNROLL for (int q = 0; q < P; ++q) sacc = fmaf(Ct[r][q], rb[q], sacc);
        mpb[r] = mpb[r] + mfb[r] - sacc;
      }
      // A-bar = Sp-bar' (A Sprev^T) + Sp-bar'^T (A Sprev) + mu_p-bar' mu_prev^T
      float Ab[R][N];
      {
        float Sprev[R][N], muprev[N];
        if (t > 0) {
          TM::ld_mat(reinterpret_cast<const float*>(sb + PL::s4_Sv), lane, Sprev);
          TV::ld(reinterpret_cast<const float*>(sb + PL::s4_mv), lane, 0, muprev);
        } else {
          copy_rows<C, N>(base + Base<C>::oS0, 0, Sprev);
          load_row<N>(base + Base<C>::oMu0, muprev);
          if (a.Sig_init && active) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_init + ((long)b * N + r) * N, Sprev[r]); }
          if (a.mu_init && active) load_row<N>(a.mu_init + (long)b * N, muprev);
        }
        RegView<N, N> Sv{Sprev};
        float M1[R][N], M2[R][N];
        mm_RSt<false>(A, Sv, M1);
        mm_RS<false>(A, Sv, M2);
        RegView<N, N> M1_v{M1}, M2_v{M2}, Spb_v{Spb};
        mm_RS<false>(Spb, M1_v, Ab);
        mm_StS<true, R, N, N>(Spb_v, 0, M2_v, Ab);
        KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Ab[r][j] = fmaf(mpb[r], muprev[j], Ab[r][j]);
      }
      float Bb[R][M], du[M];
      KV_UNROLL for (int j = 0; j < M; ++j) du[j] = 0.f;
      KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < M; ++j) {
        Bb[r][j] = mpb[r] * in.u[j];
        du[j] = fmaf(Bm[r][j], mpb[r], du[j]);
      }
      // carry to t-1
      {
        RegView<N, N> A_v{A};
        float T4m[R][N];
        mm_RS<false>(Spb, A_v, T4m);
        RegView<N, N> T4_v{T4m};
        mm_StS<false, R, N, N>(A_v, 0, T4_v, Sf_carry);
        matTvec_own<C>(A_v, 0, mpb, mf_carry);
      }
      // A.0
      float dal[K], dal_c4[K];
      seq_acc_exchange<C, C4>(ax, lane, active, base, in.al, in.al, Ab, Bb, Ctb, Spb, true, acc, dal, dal_c4);
      KV_UNROLL for (int k = 0; k < K; ++k) dal[k] += dal_c4[k];
      // partial gradients of step t (left by sweep 3): updated in place in the staged chunk
      if (s == 3) { if (jg & 1) bar_g1.wait(); else bar_g0.wait(); }
      unsigned char* gb = (jg & 1) ? g1 : g0;
      {
        float* gy = reinterpret_cast<float*>(gb + PL::gc_Y) + lane * 4 * P + s * P;
        KV_UNROLL for (int q = 0; q < P; ++q) gy[q] += rb[q];
        float* ga = reinterpret_cast<float*>(gb + PL::gc_al) + lane * 4 * K + s * K;
        KV_UNROLL for (int k = 0; k < K; ++k) ga[k] += dal[k];
        if (has_du) {
          float* gu = reinterpret_cast<float*>(gb + PL::gc_U) + lane * 4 * M + s * M;
          KV_UNROLL for (int j = 0; j < M; ++j) gu[j] += du[j];
        }
      }
      if (s == 0) {   // chunk complete: store it
        tma::fence_async();
        __syncwarp();
        if (tma::elect_one()) {
          const int cidx = t >> 2;
          tma::store2d(&mp.dY, 4 * cidx * P, b0, tma::s32(gb + PL::gc_Y), pol_ef);
          tma::store2d(&mp.dalpha, 4 * cidx * K, b0, tma::s32(gb + PL::gc_al), pol_ef);
          if (has_du) tma::store2d(&mp.dU, 4 * cidx * M, b0, tma::s32(gb + PL::gc_U), pol_ef);
          tma::commit();
        }
        ++jg;
      }
    }
    tma::wait_all0();
  }
  if (active) kv_info_or(a.info, bad0 | (ok ? 0 : KV_INFO_PIVOT) | (ok_s ? 0 : KV_INFO_CHOL_S) | (ok_q ? 0 : KV_INFO_CHOL_Q));

  // ---------------------------------------------------------------- per-CTA partial sums (fixed order -> deterministic)
  if (w.with_elbo) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double v = el[i];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) ered[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      double v = 0.0;
      for (int wq = 0; wq < nwarps; ++wq) v += ered[wq][threadIdx.x];
      elbo_partials[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
  }
  {
    // parameter gradients: quads of a warp by xor-shuffles, warps through shared memory (the staging tiles are dead)
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < GA::count; ++i) acc.v[i] += __shfl_xor_sync(0xffffffffu, acc.v[i], off);
    }
    __syncthreads();
    float* red = reinterpret_cast<float*>(body);
    for (int i = threadIdx.x; i < GA::PSZ; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    for (int wq = 0; wq < nwarps; ++wq) {
      if (warp == wq && lane < 4) acc.for_each(lane, [&](int idx, float v) { red[idx] += v; });
      __syncthreads();
    }
    for (int i = threadIdx.x; i < GA::PSZ; i += blockDim.x) partials[(size_t)blockIdx.x * GA::PSZ + i] = red[i];
  }
}

// workspace of the thread-per-sequence backward: scratch [B,T,..] x 4 | ELBO partials | parameter-gradient rows
template <class C> size_t seq_bwd_ws_bytes(int B, int T) {
  using GA = GradAcc<Cfg<C::N, C::P, C::M, C::K, 4, C::QPM, C::CSH>>;
  const size_t BT = (size_t)B * T;
  const size_t nn = align256(sizeof(float) * BT * C::N * C::N);
  const size_t nv = align256(sizeof(float) * BT * C::N);
  const size_t rows = (size_t)seq_bwd_grid(B);
  return 2 * nn + 2 * nv + align256(sizeof(double) * 5 * rows) + align256(sizeof(float) * rows * GA::PSZ);
}

// can the adjoint of this call run on the thread-per-sequence kernel?
inline bool seq_bwd_eligible(const Args& a, const BwdArgs& w, const float* g_elbo) {
  const bool any_cot = w.c_mu_s || w.c_Sig_s || w.c_mu_f || w.c_Sig_f || w.c_mu_p || w.c_Sig_p || w.c_A || w.c_B || w.c_C;
  return seq_eligible(a) && g_elbo != nullptr && !any_cot && !w.elbo_only && a.eps != nullptr;
}

template <class C>
int launch_seq_bwd(const Args& a, BwdArgs w, const BasePtrs& bp, const float* g_elbo, const float* terms, void* ws,
                   GradPtrs gp, cudaStream_t s, const DpView* dp = nullptr) {
  using C4 = Cfg<C::N, C::P, C::M, C::K, 4, C::QPM, C::CSH>;
  using GA = GradAcc<C4>;
  (void)cudaGetLastError();
  const int B = a.B, T = a.T;
  const size_t BT = (size_t)B * T;
  const size_t nn = align256(sizeof(float) * BT * C::N * C::N);
  const size_t nv = align256(sizeof(float) * BT * C::N);
  char* p = reinterpret_cast<char*>(ws);
  w.w_Sig_f = reinterpret_cast<float*>(p); p += nn;
  w.w_Sig_p = reinterpret_cast<float*>(p); p += nn;
  w.w_mu_f = reinterpret_cast<float*>(p); p += nv;
  w.w_mu_p = reinterpret_cast<float*>(p); p += nv;
  const int grid = seq_bwd_grid(B);
  double* elbo_partials = reinterpret_cast<double*>(p); p += align256(sizeof(double) * 5 * (size_t)grid);
  float* partials = reinterpret_cast<float*>(p);

  SeqBwdMaps mp;
  memset(&mp, 0, sizeof(mp));
  bool okm = make_chunk_map(&mp.Y, a.Y, B, T, C::P) && make_chunk_map(&mp.alpha, a.alpha, B, T, C::K) &&
             make_chunk_map(&mp.dY, w.dY, B, T, C::P) && make_chunk_map(&mp.dalpha, w.dalpha, B, T, C::K);
  if (a.U) okm = okm && make_chunk_map(&mp.U, a.U, B, T, C::M);
  if (a.mask) okm = okm && make_chunk_map(&mp.mask, a.mask, B, T, 1);
  if (w.dU) okm = okm && make_chunk_map(&mp.dU, w.dU, B, T, C::M);
  okm = okm && make_row_map(&mp.eps, a.eps, B, T, C::N) && make_row_map(&mp.mu_s, a.mu_s, B, T, C::N) &&
        make_row_map(&mp.mu_p, a.mu_p, B, T, C::N) && make_row_map(&mp.mu_f, a.mu_f, B, T, C::N) &&
        make_row_map(&mp.w_mu_f, w.w_mu_f, B, T, C::N) && make_row_map(&mp.w_mu_p, w.w_mu_p, B, T, C::N) &&
        make_row_map(&mp.Sig_s, a.Sig_s, B, T, C::N * C::N) && make_row_map(&mp.Sig_p, a.Sig_p, B, T, C::N * C::N) &&
        make_row_map(&mp.Sig_f, a.Sig_f, B, T, C::N * C::N) && make_row_map(&mp.w_Sig_f, w.w_Sig_f, B, T, C::N * C::N) &&
        make_row_map(&mp.w_Sig_p, w.w_Sig_p, B, T, C::N * C::N);
  if (!okm) return -6;
  const int warps = seq_bwd_warps_per_cta(B);
  const bool has_du = w.dU != nullptr;
  size_t smb = seq_bwd_smem<C>(warps, has_du);
  if (smb < seq_smem_floor()) smb = seq_smem_floor();
  cudaError_t e = cudaFuncSetAttribute(k_seq_bwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
  if (e != cudaSuccess) return (int)e;
  k_seq_bwd<C><<<grid, 32 * warps, smb, s>>>(a, w, bp, mp, g_elbo, terms, partials, elbo_partials);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  // final kernel: identical to the lane-group path
  constexpr int psz = GA::PSZ;
  const int nparam_blocks = (psz + 3) / 4;
  ScaleJob sj{{nullptr, nullptr, nullptr}, {0, 0, 0}};
  int scale_blocks = 0;
  const int no_scale = (w.raw_sums || w.mask_part != nullptr) ? 1 : 0;
  auto fill_sj = [&] {
    sj.p[0] = w.dY; sj.n[0] = (long)BT * C::P;
    sj.p[1] = w.dalpha; sj.n[1] = (long)BT * C::K;
    sj.p[2] = w.dU; sj.n[2] = w.dU ? (long)BT * C::M : 0;
    const long total4 = (sj.n[0] + sj.n[1] + sj.n[2]) / 4;
    int sb = (int)((total4 + 128 * 8 - 1) / (128 * 8));
    if (sb < 1) sb = 1;
    if (sb > 148 * 8) sb = 148 * 8;
    return sb;
  };
  if (w.with_elbo && !no_scale) scale_blocks = fill_sj();
  if (dp) {
    if (dp->nparam != psz) return -5;
    const int sb = fill_sj();
    k_bwd_final_dp<<<nparam_blocks + sb, 128, 0, s>>>(partials, grid, psz, C::K * C::N * C::N, C::K * C::N * C::M,
                                                      C::K * C::P * C::N, gp, elbo_partials, grid, w.terms_out, nparam_blocks, sj,
                                                      *dp, a.info);
    return (int)cudaGetLastError();
  }
  k_bwd_final<<<nparam_blocks + scale_blocks, 128, 0, s>>>(partials, grid, psz, C::K * C::N * C::N, C::K * C::N * C::M,
                                                           C::K * C::P * C::N, gp, w.with_elbo ? elbo_partials : nullptr, grid,
                                                           w.terms_out, no_scale, nparam_blocks, sj);
  return (int)cudaGetLastError();
}

}  // namespace kvae
