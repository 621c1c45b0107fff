// kvae_fwd.cuh — forward recursion of the Kalman hot path, one sequence per lane group.
//
//   sweep 1 (t = 0..T-1)  : A.0 mixing + A.1 filter step   (kalman_filter.py:31-104, :107-201;
//                           dyn_param.py:58-60; switch_dyn_param.py:82-86)
//   sweep 2 (t = T-2..0)  : A.2 RTS smoother step          (kalman_filter.py:204-237, :240-279)
//
// Equation labels (A.x) follow SURVEY.md Appendix A.  Tensor layouts are the reference's
// batch-major [B,T,...] contiguous fp32 layouts (kalman_filter.py:193-201).
#pragma once
#ifndef KV_SMOOTH_PF_L2
#define KV_SMOOTH_PF_L2 4   // lane-group smoother: L2 prefetch distance beyond the one-ahead register loads (0 = off)
#endif

#include "kvae_prims.cuh"

namespace kvae {

// Arguments shared by all sweeps (plain pointers; nullable where noted).
struct Args {
  int B, T;
  // inputs
  const float* Y;      // [B,T,P]
  const float* U;      // [B,T,M]   nullable -> zeros
  const float* mask;   // [B,T]     nullable -> ones
  const float* alpha;  // [B,T,K]
  const float* eps;    // [B,T,N]   (ELBO / backward)
  // state tensors (outputs of the forward pass, inputs of ELBO / backward)
  float* mu_f;   float* Sig_f;   // [B,T,N], [B,T,N,N]
  float* mu_p;   float* Sig_p;
  float* mu_s;   float* Sig_s;   // nullable in the forward pass -> filter only
  float* A_list; float* B_list; float* C_list;  // nullable -> not materialised
  // optional initial belief per sequence (single-step / chunked filtering); nullable -> mu0/Sigma0
  const float* mu_init;   // [B,N]
  const float* Sig_init;  // [B,N,N]
  // optional explicit per-step matrices (forward kernels only): when dA is given the kernels read
  // A_t,B_t,C_t(,Q_t) from these instead of mixing the base matrices (filter_step / smooth_step forms,
  // kalman_filter.py:31,204).  Same layouts as A_list/B_list/C_list.
  const float* dA;   // [B,T,N,N]
  const float* dB;   // [B,T,N,M]
  const float* dC;   // [B,T,P,N]
  const float* dQ;   // [B,T,N,N]  nullable -> base Q
  int smooth_only;   // 1: skip the filter sweep, smooth from the filtered states already in mu_f / Sig_f
  float* mask_part;  // optional out (forward kernel): per-CTA sum of the mask, see kvae_states.mask_partials
  // optional outs (forward kernels): the projections KVAE.impute forms afterwards (model.py:280-281, 287-288),
  // a_filt[b,t] = C_t mu_{t|t}, a_smooth[b,t] = C_t mu_{t|T}  ([B,T,P]); nullable
  float* a_filt;
  float* a_smooth;
  // ELBO factorisations (kalman_filter.py:282-302 _safe_cholesky): jitter added to sym(Q_t) (the one added to sym(Sigma_s)
  // travels with the ELBO / adjoint calls) and the final rung of the reference's ladder, L = diag(sqrt(clamp(diag, 1e-6))),
  // selected per matrix family: bit 0 = Sigma_smooth, bit 1 = Q
  float jitter_q;
  int chol_diag;
  int* info;  // device status word, OR of KV_INFO_* bits
};
// status bits: a pivot of the filter / smoother / R / Sigma_0 factorisations was not positive (the reference raises
// LinAlgError); the Cholesky of sym(Sigma_smooth)+jitter / sym(Q_t)+jitter failed (the reference retries with 10x jitter);
// 2 is the data-parallel exchange time-out (kvae_kernels.cuh)
#define KV_INFO_PIVOT 1
#define KV_INFO_CHOL_S 4
#define KV_INFO_CHOL_Q 8
KV_FN void kv_info_or(int* info, int code) {
#if defined(__CUDA_ARCH__)
  if (code) atomicOr(info, code);
#else
  if (code) *info |= code;
#endif
}

constexpr int pad4(int x) { return (x + 3) & ~3; }

// Base parameters staged once per CTA in shared memory (C^T is stored transposed per mode so that
// a lane's rows of C_t^T are contiguous).
// Row stride of a staged base matrix.  Each lane reads ITS OWN rows of A_k/B_k/Q_k/C_k^T when it mixes, i.e. the
// lanes of a warp read rows that are `cols` floats apart: for cols = 8 or 16 that is a 2-/4-way bank conflict on every
// 128-bit load (measured for n = 16: 16 wavefronts per LDS.128 instead of 4, half of all shared-memory wavefronts of
// the forward kernel).  Four floats of padding per row make the eight lanes of a quarter-warp hit disjoint banks.
template <class C> constexpr int bld(int cols) { return (C::L > 1 && cols % 8 == 0) ? cols + 4 : cols; }

template <class C> struct Base {
  static constexpr int ldA = bld<C>(C::N), ldB = bld<C>(C::M), ldCt = bld<C>(C::P), ldQ = bld<C>(C::N);
  static constexpr int oA = 0;
  static constexpr int oB = oA + pad4(C::K * C::N * ldA);
  static constexpr int oCt = oB + pad4(C::K * C::N * ldB);
  static constexpr int oQ = oCt + pad4(C::KC * C::N * ldCt);
  static constexpr int oR = oQ + pad4(C::KQ * C::N * ldQ);
  static constexpr int oMu0 = oR + pad4(C::P * C::P);
  static constexpr int oS0 = oMu0 + pad4(C::N);
  static constexpr int total = oS0 + pad4(C::N * C::N);
};

// element-wise fill of the base block; `i` strides over [0,total) (thread-strided on device)
template <class C>
KV_FN void base_fill(float* base, int i, const float* A, const float* Bm, const float* Cm, const float* Q,
                     const float* Rm, const float* mu0, const float* S0) {
  using BL = Base<C>;
  constexpr int N = C::N, P = C::P, M = C::M, K = C::K;
  float v = 0.f;
  if (i < BL::oB) { int row = i / BL::ldA, col = i % BL::ldA; if (row < K * N && col < N) v = A[row * N + col]; }
  else if (i < BL::oCt) { int j = i - BL::oB, row = j / BL::ldB, col = j % BL::ldB; if (row < K * N && col < M) v = Bm[row * M + col]; }
  else if (i < BL::oQ) {
    int j = i - BL::oCt, row = j / BL::ldCt, a = j % BL::ldCt;
    if (row < C::KC * N && a < P) { int k = row / N, rn = row % N; v = Cm[(k * P + a) * N + rn]; }
  }
  else if (i < BL::oR) { int j = i - BL::oQ, row = j / BL::ldQ, col = j % BL::ldQ; if (row < C::KQ * N && col < N) v = Q[row * N + col]; }
  else if (i < BL::oMu0) { int j = i - BL::oR; if (j < P * P) v = Rm[j]; }
  else if (i < BL::oS0) { int j = i - BL::oMu0; if (j < N) v = mu0[j]; }
  else { int j = i - BL::oS0; if (j < N * N) v = S0[j]; }
  base[i] = v;
}

// Per-warp scratch tiles of the forward / ELBO kernels: 4 [n x n] + 2 [n x p] + 2 vector slots.
template <class C> using FTiles = TileSet<C::L, C::R, C::P, 4, 2, 2, C::MEM>;

// ---------------------------------------------------------------------------------------
// per-step inputs
// ---------------------------------------------------------------------------------------
template <class C> struct StepIn {
  float y[C::P];
  float u[C::M];
  float al[C::K];
  float m;
};
template <class C> KV_FN void load_step(const Args& a, long bt, StepIn<C>& s) {
  load_row<C::P>(a.Y + bt * C::P, s.y);
  if (a.U) load_row<C::M>(a.U + bt * C::M, s.u);
  else { KV_UNROLL for (int j = 0; j < C::M; ++j) s.u[j] = 0.f; }
  if (a.alpha) load_row<C::K>(a.alpha + bt * C::K, s.al);
  else { KV_UNROLL for (int k = 0; k < C::K; ++k) s.al[k] = 0.f; }
  s.m = a.mask ? a.mask[bt] : 1.0f;
}

// ---------------------------------------------------------------------------------------
// Input staging: the per-step inputs of a sequence (y_t, u_t, alpha_t, mask_t: 4*(P+M+K+1) bytes) are
// tiny, so fetching them step by step costs one L1 wavefront per stream per group per step.  Instead a
// group fetches FOUR time steps at once with 128-bit loads (one 16-byte piece per lane and load), keeps
// the chunk pending in registers while it works on the previous one, then drops it into a per-group
// shared-memory slot from which every lane reads its step with broadcast loads.
// Needs T % 4 == 0 (all pieces 16-byte aligned); otherwise the callers use load_step().
// ---------------------------------------------------------------------------------------
template <class C, bool WITH_YUM> struct InStage {
  static constexpr int P = C::P, M = C::M, K = C::K, L = C::L;
  static constexpr int oY = 0, oU = 4 * P, oA = oU + 4 * M, oM = oA + 4 * K;
  static constexpr int chunk_floats = oM + 4;
  static constexpr int NPIECE = WITH_YUM ? (P + M + K + 1) : K;   // 16-byte pieces per 4-step chunk
  static constexpr int PER_LANE = (NPIECE + L - 1) / L;
  // per-group slot, padded so that the groups of a warp start 4 banks apart
  static constexpr int group_floats = pad_res(chunk_floats, L == 1 ? 4 : (L < 32 ? 4 : L));
  f4 pend[PER_LANE];
  float* slot;

  // piece q of the chunk starting at step index bt0 (= b*T + t0, t0 % 4 == 0)
  KV_FN f4 fetch(const Args& a, long bt0, int q) const {
    const float* src = nullptr;
    if (WITH_YUM) {
      if (q < P) src = a.Y + bt0 * P + 4 * q;
      else if (q < P + M) src = a.U ? a.U + bt0 * M + 4 * (q - P) : nullptr;
      else if (q < P + M + K) { if (a.alpha == nullptr) return f4{0.f, 0.f, 0.f, 0.f}; src = a.alpha + bt0 * K + 4 * (q - P - M); }
      else src = a.mask ? a.mask + bt0 + 0 : nullptr;
      if (src == nullptr) { const float v = (q < P + M) ? 0.f : 1.f; return f4{v, v, v, v}; }
    } else {
      if (a.alpha == nullptr) return f4{0.f, 0.f, 0.f, 0.f};
      src = a.alpha + bt0 * K + 4 * q;
    }
    return *reinterpret_cast<const f4*>(src);
  }
  KV_FN int slot_off(int q) const {
    if (WITH_YUM) return 4 * q;            // Y pieces, U pieces, alpha pieces, mask piece are laid out back to back
    return oA + 4 * q;
  }
  KV_FN void prefetch(const Args& a, const Group<C::L, C::R>& g, long bt0) {
    KV_UNROLL for (int i = 0; i < PER_LANE; ++i) {
      const int q = g.lane + i * L;
      if (q < NPIECE) pend[i] = fetch(a, bt0, q);
    }
  }
  KV_FN void commit(const Group<C::L, C::R>& g) {
    g.sync();
    KV_UNROLL for (int i = 0; i < PER_LANE; ++i) {
      const int q = g.lane + i * L;
      if (q < NPIECE) *reinterpret_cast<f4*>(slot + slot_off(q)) = pend[i];
    }
    g.sync();
  }
  KV_FN void read(int s, StepIn<C>& in) const {   // s = step within the chunk (0..3)
    if (WITH_YUM) {
      KV_UNROLL for (int j = 0; j < P; ++j) in.y[j] = slot[oY + s * P + j];
      KV_UNROLL for (int j = 0; j < M; ++j) in.u[j] = slot[oU + s * M + j];
      in.m = slot[oM + s];
    }
    KV_UNROLL for (int k = 0; k < K; ++k) in.al[k] = slot[oA + s * K + k];
  }
};

// ---------------------------------------------------------------------------------------
// A.0  mixing: own rows of A_t, B_t, C_t^T, Q_t
// ---------------------------------------------------------------------------------------
template <class C, int COLS, int LD = COLS>
KV_FN void mix_one(const float* __restrict__ basek, int modes, const float (&al)[C::K], int row0, float (&out)[C::R][COLS]) {
  KV_UNROLL for (int r = 0; r < C::R; ++r) KV_UNROLL for (int j = 0; j < COLS; ++j) out[r][j] = 0.f;
  KV_UNROLL for (int k = 0; k < C::K; ++k) {
    if (k < modes) {
      KV_UNROLL for (int r = 0; r < C::R; ++r) {
        float row[COLS];
        load_row<COLS>(basek + (k * C::N + row0 + r) * LD, row);
        kv_axpy<COLS>(al[k], row, out[r]);
      }
    }
  }
}
template <class C, int COLS, int LD = COLS>
KV_FN void copy_rows(const float* __restrict__ src, int row0, float (&out)[C::R][COLS]) {
  KV_UNROLL for (int r = 0; r < C::R; ++r) load_row<COLS>(src + (row0 + r) * LD, out[r]);
}
template <class C> KV_FN void mix_A(const float* base, const float (&al)[C::K], int row0, float (&A)[C::R][C::N]) {
  mix_one<C, C::N, Base<C>::ldA>(base + Base<C>::oA, C::K, al, row0, A);
}
template <class C> KV_FN void mix_B(const float* base, const float (&al)[C::K], int row0, float (&Bm)[C::R][C::M]) {
  mix_one<C, C::M, Base<C>::ldB>(base + Base<C>::oB, C::K, al, row0, Bm);
}
template <class C> KV_FN void mix_Ct(const float* base, const float (&al)[C::K], int row0, float (&Ct)[C::R][C::P]) {
  if constexpr (C::CSH) copy_rows<C, C::P, Base<C>::ldCt>(base + Base<C>::oCt, row0, Ct);
  else mix_one<C, C::P, Base<C>::ldCt>(base + Base<C>::oCt, C::K, al, row0, Ct);
}
template <class C> KV_FN void mix_Q(const float* base, const float (&al)[C::K], int row0, float (&Q)[C::R][C::N]) {
  if constexpr (C::QPM) mix_one<C, C::N, Base<C>::ldQ>(base + Base<C>::oQ, C::K, al, row0, Q);
  else copy_rows<C, C::N, Base<C>::ldQ>(base + Base<C>::oQ, row0, Q);
}

// out[q] = sum_j C_t[q][j] v[j] from the lane's rows of C_t^T and its entries of v (all lanes of the group take part)
template <class C>
KV_FN void project_obs(const Group<C::L, C::R>& g, const float (&Ct)[C::R][C::P], const float (&v_own)[C::R], float (&out)[C::P]) {
  KV_UNROLL for (int q = 0; q < C::P; ++q) {
    float s = 0.f;
    KV_UNROLL for (int r = 0; r < C::R; ++r) s = fmaf(Ct[r][q], v_own[r], s);
    out[q] = s;
  }
  g.allreduce(out);
}

// explicit-matrix variants (forward kernels): read the step's rows from dense per-step tensors when given
template <class C> KV_FN void get_A(const Args& a, const float* base, const float (&al)[C::K], int row0, long bt, float (&A)[C::R][C::N]) {
  if (a.dA) { KV_UNROLL for (int r = 0; r < C::R; ++r) load_row<C::N>(a.dA + (bt * C::N + row0 + r) * C::N, A[r]); }
  else mix_A<C>(base, al, row0, A);
}
template <class C> KV_FN void get_B(const Args& a, const float* base, const float (&al)[C::K], int row0, long bt, float (&Bm)[C::R][C::M]) {
  if (a.dA) { KV_UNROLL for (int r = 0; r < C::R; ++r) load_row<C::M>(a.dB + (bt * C::N + row0 + r) * C::M, Bm[r]); }
  else mix_B<C>(base, al, row0, Bm);
}
template <class C> KV_FN void get_Ct(const Args& a, const float* base, const float (&al)[C::K], int row0, long bt, float (&Ct)[C::R][C::P]) {
  if (a.dA) { KV_UNROLL for (int r = 0; r < C::R; ++r) KV_UNROLL for (int q = 0; q < C::P; ++q) Ct[r][q] = a.dC[(bt * C::P + q) * C::N + row0 + r]; }
  else mix_Ct<C>(base, al, row0, Ct);
}
template <class C> KV_FN void get_Q(const Args& a, const float* base, const float (&al)[C::K], int row0, long bt, float (&Q)[C::R][C::N]) {
  if (a.dA && a.dQ) { KV_UNROLL for (int r = 0; r < C::R; ++r) load_row<C::N>(a.dQ + (bt * C::N + row0 + r) * C::N, Q[r]); }
  else mix_Q<C>(base, al, row0, Q);
}

// Everything a filter step recomputes that the adjoint also needs.
template <class C> struct GainOut {
  float Pm[C::R][C::P];   // own rows of P = Sigma_p C^T
  float K0[C::R][C::P];   // unmasked gain rows
  float Kg[C::R][C::P];   // masked gain rows
  float r[C::P];          // innovation (replicated)
  float yp[C::P];         // predicted observation C mu_p (replicated)
  float Lc[C::P][C::P];   // chol(S) (replicated)
  float invd[C::P];
};

// gain part of A.1:  S = sym(C Sp C^T + R), K0 = P S^-1, K = m K0, r = y - C mu_p.
// Ct: own rows of C^T; Ctf: full view of C^T; Sp: own rows of Sigma_p; mup_own: own entries of mu_p.
template <class C, class VC>
KV_FN bool gain(const Group<C::L, C::R>& g, const float* base, const float (&Sp)[C::R][C::N], const float (&mup_own)[C::R],
                const float (&Ct)[C::R][C::P], const VC& Ctf, const float (&y)[C::P], float m, GainOut<C>& o) {
  constexpr int P = C::P, R = C::R;
  mm_RS<false>(Sp, Ctf, o.Pm);  // P[r][a] = sum_j Sp[r][j] Ct[j][a]     (kalman_filter.py:82)
  // partial S = C P and y_pred = C mu_p over own rows, then all-reduce
  float red[P * P + P];
  KV_UNROLL for (int a = 0; a < P; ++a) {
    KV_UNROLL for (int b = 0; b < P; ++b) {
      float s = 0.f;
      KV_UNROLL for (int r = 0; r < R; ++r) s = fmaf(Ct[r][a], o.Pm[r][b], s);
      red[a * P + b] = s;
    }
    float s = 0.f;
    KV_UNROLL for (int r = 0; r < R; ++r) s = fmaf(Ct[r][a], mup_own[r], s);
    red[P * P + a] = s;
  }
  g.allreduce(red);
  const float* Rm = base + Base<C>::oR;
  float S[P][P];
  KV_UNROLL for (int a = 0; a < P; ++a) KV_UNROLL for (int b = 0; b < P; ++b) S[a][b] = red[a * P + b] + Rm[a * P + b];  // :78
  float Ss[P][P];
  KV_UNROLL for (int a = 0; a < P; ++a) KV_UNROLL for (int b = 0; b < P; ++b) Ss[a][b] = 0.5f * (S[a][b] + S[b][a]);     // :79
  const bool ok = chol_small<P>(Ss, o.Lc, o.invd);
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int a = 0; a < P; ++a) o.K0[r][a] = o.Pm[r][a];
  RegView<P, P> Lv{o.Lc};
  solve_rows_llt<R, P>(o.K0, Lv, o.invd);                                                    // :89
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int a = 0; a < P; ++a) o.Kg[r][a] = m * o.K0[r][a];  // :92
  KV_UNROLL for (int a = 0; a < P; ++a) { o.yp[a] = red[P * P + a]; o.r[a] = y[a] - red[P * P + a]; }   // :73-75
  return ok;
}

// ---------------------------------------------------------------------------------------
// One filter step (A.1) given the step's mixed matrices: shared by the lane-group sweep below and by the
// thread-per-sequence kernel with staged streams (csrc/kvae_seq.cuh).
// ---------------------------------------------------------------------------------------
struct NoHook {
  static constexpr bool ON = false;
};
template <class C> struct FilterStepOut {
  float mup[C::R], muf[C::R];
  float Sp[C::R][C::N], Sf[C::R][C::N];
};
template <class C, class Hook = NoHook>
KV_FN bool filter_step_math(const Group<C::L, C::R>& g, const float* base, const FTiles<C>& tl, const StepIn<C>& cur,
                            const float (&A)[C::R][C::N], const float (&Bm)[C::R][C::M], const float (&Ct)[C::R][C::P],
                            const float (&Q)[C::R][C::N], const float (&Sig)[C::R][C::N], const float (&mu)[C::N],
                            FilterStepOut<C>& o, Hook* hook = nullptr) {
  constexpr int N = C::N, P = C::P, M = C::M, R = C::R, L = C::L;
  constexpr bool MEM = C::MEM;
  const TileRef X0 = tl.nn(0), X1 = tl.nn(1), X2 = tl.nn(2), CB = tl.np(0), KB = tl.np(1);
  const int row0 = g.row0();
  float (&mup)[R] = o.mup;
  float (&muf)[R] = o.muf;
  float (&Sp)[R][N] = o.Sp;
  float (&Sf)[R][N] = o.Sf;
  // predict (A.1): mu_p = A mu + B u ; Sigma_p = (A Sigma) A^T + Q        (kalman_filter.py:65-67)
  KV_UNROLL for (int r = 0; r < R; ++r) {
    float s1 = 0.f, s2 = 0.f;
    KV_UNROLL for (int j = 0; j < N; ++j) s1 = fmaf(A[r][j], mu[j], s1);
    KV_UNROLL for (int j = 0; j < M; ++j) s2 = fmaf(Bm[r][j], cur.u[j], s2);
    mup[r] = s1 + s2;
  }
  {
    auto Sf_v = publish<MEM, L, R, N>(g, Sig, X0);
    float Mx[R][N];
    mm_RS<false>(A, Sf_v, Mx);
    auto A_v = publish<MEM, L, R, N>(g, A, X1);
    mm_RSt<false>(Mx, A_v, Sp);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Sp[r][j] += Q[r][j];
  }
  // update
  auto Ct_v = publish<MEM, L, R, P>(g, Ct, CB);
  GainOut<C> go;
  const bool ok = gain<C>(g, base, Sp, mup, Ct, Ct_v, cur.y, cur.m, go);
  if constexpr (Hook::ON) hook->after_gain(cur, go);
  KV_UNROLL for (int r = 0; r < R; ++r) {
    float s = 0.f;
    KV_UNROLL for (int q = 0; q < P; ++q) s = fmaf(go.Kg[r][q], go.r[q], s);
    muf[r] = mup[r] + s;                                                     // :96
  }
  // Joseph form: Sigma_f = sym((G Sp) G^T + (K R) K^T), G = I - K C          (:99-101)
  float G[R][N];
  mm_RSt<false>(go.Kg, Ct_v, G);
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) G[r][j] = ((row0 + r == j) ? 1.0f : 0.0f) - G[r][j];
  float Xm[R][N];
  {
    auto Sp_v = publish<MEM, L, R, N>(g, Sp, X2);
    float T1[R][N];
    mm_RS<false>(G, Sp_v, T1);
    auto G_v = publish<MEM, L, R, N>(g, G, X1);
    mm_RSt<false>(T1, G_v, Xm);
    float KR[R][P];
    const float* Rm = base + Base<C>::oR;
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int q = 0; q < P; ++q) {
      float s = 0.f;
      KV_UNROLL for (int q2 = 0; q2 < P; ++q2) s = fmaf(go.Kg[r][q2], Rm[q2 * P + q], s);
      KR[r][q] = s;
    }
    auto K_v = publish<MEM, L, R, P>(g, go.Kg, KB);
    float KRK[R][N];
    mm_RSt<false>(KR, K_v, KRK);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Xm[r][j] += KRK[r][j];
  }
  {
    auto X_v = publish<MEM, L, R, N>(g, Xm, X0);
    float Xt[R][N];
    tr_rows<R, N>(X_v, row0, Xt);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Sf[r][j] = 0.5f * (Xm[r][j] + Xt[r][j]);
  }

  return ok;
}

// ---------------------------------------------------------------------------------------
// sweep 1: filter.  On return Sig (own rows) and mu (replicated) hold the last filtered belief.
// ---------------------------------------------------------------------------------------
// Hook: compile-time extension points of the filter loop (csrc/kvae_lstm.cuh runs the LSTM dynamics network there);
// NoHook compiles to nothing.
template <class C, class Hook = NoHook>
KV_FN void filter_sweep(const Args& a, const float* base, const FTiles<C>& tl, const Group<C::L, C::R>& g, int b, bool active,
                        float* stage_slot, float (&Sig)[C::R][C::N], float (&mu)[C::N], float (&mu_own)[C::R],
                        float* msum_out = nullptr, Hook* hook = nullptr) {
  constexpr int N = C::N, P = C::P, M = C::M, R = C::R, L = C::L;
  constexpr bool MEM = C::MEM;
  const TileRef X0 = tl.nn(0), X1 = tl.nn(1), X2 = tl.nn(2), CB = tl.np(0), KB = tl.np(1), VB = tl.vec(0);
  float msum = 0.f;   // sum_t mask_t of this sequence (the ELBO normaliser's numerator, kalman_filter.py:392)
  const int row0 = g.row0();
  const int T = a.T;
  bool ok = true;

  // initial belief
  if (a.Sig_init) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_init + ((long)b * N + row0 + r) * N, Sig[r]); }
  else copy_rows<C, N>(base + Base<C>::oS0, row0, Sig);
  if (a.mu_init) load_row<N>(a.mu_init + (long)b * N, mu);
  else load_row<N>(base + Base<C>::oMu0, mu);

  StepIn<C> cur, nxt;
  const bool staged = (T % 4 == 0);
  InStage<C, true> ins;
  ins.slot = stage_slot;
  if (staged) {
    ins.prefetch(a, g, (long)b * T);
    ins.commit(g);
  } else {
    load_step<C>(a, (long)b * T, cur);
  }
  for (int t = 0; t < T; ++t) {
    const long bt = (long)b * T + t;
    if (staged) ins.read(t & 3, cur);
    else if (t + 1 < T) load_step<C>(a, bt + 1, nxt);  // software prefetch of the next step's inputs

    if constexpr (Hook::ON) hook->before_step(g, bt, cur);   // alpha_t from the in-kernel dynamics network
    float A[R][N], Bm[R][M], Ct[R][P], Q[R][N];
    get_A<C>(a, base, cur.al, row0, bt, A);
    get_B<C>(a, base, cur.al, row0, bt, Bm);
    get_Ct<C>(a, base, cur.al, row0, bt, Ct);
    get_Q<C>(a, base, cur.al, row0, bt, Q);
    // first step of a chunk: start fetching the next chunk.  Issued AFTER the slot / mixing loads: global loads placed
    // directly in front of shared-memory loads end up on the same scoreboard and the first FMA of the mixing then
    // waits for the whole L2 round trip (measured: 40 % of the stall samples of this kernel before the reordering)
    if (staged && (t & 3) == 0 && t + 4 < T) ins.prefetch(a, g, bt + 4);

    FilterStepOut<C> fo;
    msum += cur.m;
    ok = filter_step_math<C, Hook>(g, base, tl, cur, A, Bm, Ct, Q, Sig, mu, fo, hook) && ok;
    const float (&mup)[R] = fo.mup;
    const float (&muf)[R] = fo.muf;
    const float (&Sp)[R][N] = fo.Sp;
    const float (&Sf)[R][N] = fo.Sf;

    if (a.a_filt) {   // C_t mu_{t|t} (model.py:287-288): C_t^T rows are in registers
      float af[P];
      project_obs<C>(g, Ct, muf, af);
      if (active && g.lane == 0) store_row<P>(a.a_filt + bt * P, af);
    }
#ifdef KV_DIAG_NO_STORES   // diagnostic build only (tools/README.md): how long is the bare dependent chain without its outputs?
    if (active && a.T < 0) {
#else
    if (active) {
#endif
      KV_UNROLL for (int r = 0; r < R; ++r) {
        store_row<N>(a.Sig_p + (bt * N + row0 + r) * N, Sp[r]);
        store_row<N>(a.Sig_f + (bt * N + row0 + r) * N, Sf[r]);
      }
      store_row<R>(a.mu_p + bt * N + row0, mup);
      store_row<R>(a.mu_f + bt * N + row0, muf);
      if (a.A_list) { KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(a.A_list + (bt * N + row0 + r) * N, A[r]); }
      if (a.B_list) { KV_UNROLL for (int r = 0; r < R; ++r) store_row<M>(a.B_list + (bt * N + row0 + r) * M, Bm[r]); }
      if (a.C_list) {
        if constexpr (L == 1) {
          KV_UNROLL for (int q = 0; q < P; ++q) {
            float crow[N];
            KV_UNROLL for (int j = 0; j < N; ++j) crow[j] = Ct[j][q];
            store_row<N>(a.C_list + (bt * P + q) * N, crow);
          }
        } else {
          KV_UNROLL for (int q = 0; q < P; ++q) KV_UNROLL for (int r = 0; r < R; ++r) a.C_list[(bt * P + q) * N + row0 + r] = Ct[r][q];
        }
      }
    }
    // carry
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Sig[r][j] = Sf[r][j];
    allgather<MEM, L, R>(g, muf, VB, mu);
    KV_UNROLL for (int r = 0; r < R; ++r) mu_own[r] = muf[r];
    if (staged) {
      if ((t & 3) == 3 && t + 1 < T) ins.commit(g);   // next chunk: publish the pending one
    } else {
      cur = nxt;
    }
  }
  if (msum_out) *msum_out = msum;
  if (!ok && active) kv_info_or(a.info, KV_INFO_PIVOT);
}

// ---------------------------------------------------------------------------------------
// smoother gain (A.2), shared with the adjoint:  J = (Sf A1^T) Sp1^-1 with a general (unsymmetric)
// LU of Sp1, as the reference does (kalman_filter.py:229).  On return LU (own rows) / invu hold the
// factor and XL holds its published copy.
// ---------------------------------------------------------------------------------------
template <class C>
KV_FN bool smoother_gain(const Group<C::L, C::R>& g, TileRef XA, TileRef XL, const float (&Sf)[C::R][C::N],
                         const float (&A1)[C::R][C::N], const float (&Sp1)[C::R][C::N], float (&J)[C::R][C::N],
                         float (&LU)[C::R][C::N], float (&invu)[C::N]) {
  constexpr int N = C::N, R = C::R, L = C::L;
  auto A_v = publish<C::MEM, L, R, N>(g, A1, XA);
  mm_RSt<false>(Sf, A_v, J);                                 // W = Sf A1^T
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) LU[r][j] = Sp1[r][j];
  const bool ok = lu_dist<L, R>(g, LU, invu);
  auto L_v = publish<C::MEM, L, R, N>(g, LU, XL);
  solve_rows_lu<R, N>(J, L_v, invu);                         // J = W Sp1^-1
  return ok;
}

// states a smoother step reads: Sigma_f/mu_f at t and Sigma_p/mu_p at t+1 (own rows / entries)
template <class C> struct SmoothIn {
  float Sf[C::R][C::N], Sp1[C::R][C::N], muf[C::R], mup1[C::R];
};
template <class C> KV_FN void load_smooth_in(const Args& a, long bt, int row0, SmoothIn<C>& s) {
  KV_UNROLL for (int r = 0; r < C::R; ++r) {
    load_row<C::N>(a.Sig_f + (bt * C::N + row0 + r) * C::N, s.Sf[r]);
    load_row<C::N>(a.Sig_p + ((bt + 1) * C::N + row0 + r) * C::N, s.Sp1[r]);
  }
  load_row<C::R>(a.mu_f + bt * C::N + row0, s.muf);
  load_row<C::R>(a.mu_p + (bt + 1) * C::N + row0, s.mup1);
}

// ---------------------------------------------------------------------------------------
// One RTS smoother step (A.2): Sig / mus enter as the smoothed belief at t+1 and leave as the one at t.  Shared by the
// lane-group sweep below and the thread-per-sequence kernel (csrc/kvae_seq.cuh).
// ---------------------------------------------------------------------------------------
template <class C>
KV_FN bool smoother_step_math(const Group<C::L, C::R>& g, const FTiles<C>& tl, const float (&Sf)[C::R][C::N],
                              const float (&Sp1)[C::R][C::N], const float (&muf)[C::R], const float (&mup1)[C::R],
                              const float (&A1)[C::R][C::N], float (&Sig)[C::R][C::N], float (&mus)[C::R]) {
  constexpr int N = C::N, R = C::R, L = C::L;
  constexpr bool MEM = C::MEM;
  const TileRef X0 = tl.nn(0), X1 = tl.nn(1), X2 = tl.nn(2), VB = tl.vec(0);
  const int row0 = g.row0();
  float J[R][N], LU[R][N], invu[N];
  const bool ok = smoother_gain<C>(g, X0, X1, Sf, A1, Sp1, J, LU, invu);
  // mu_s = mu_f + J (mu_s1 - mu_p1)                                            (:232)
  float d_own[R], d[N];
  KV_UNROLL for (int r = 0; r < R; ++r) d_own[r] = mus[r] - mup1[r];
  allgather<MEM, L, R>(g, d_own, VB, d);
  KV_UNROLL for (int r = 0; r < R; ++r) {
    float s = 0.f;
    KV_UNROLL for (int j = 0; j < N; ++j) s = fmaf(J[r][j], d[j], s);
    mus[r] = muf[r] + s;
  }
  // Sigma_s = sym(Sf + (J D) J^T), D = Sigma_s1 - Sp1                          (:234-235)
  float D[R][N];
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) D[r][j] = Sig[r][j] - Sp1[r][j];
  float Xm[R][N];
  {
    auto D_v = publish<MEM, L, R, N>(g, D, X2);
    float T1[R][N];
    mm_RS<false>(J, D_v, T1);
    auto J_v = publish<MEM, L, R, N>(g, J, X0);
    mm_RSt<false>(T1, J_v, Xm);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Xm[r][j] = Sf[r][j] + Xm[r][j];
  }
  {
    auto X_v = publish<MEM, L, R, N>(g, Xm, X1);
    float Xt[R][N];
    tr_rows<R, N>(X_v, row0, Xt);
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) Sig[r][j] = 0.5f * (Xm[r][j] + Xt[r][j]);
  }
  return ok;
}

// ---------------------------------------------------------------------------------------
// sweep 2: RTS smoother, t = T-2..0.  Sig / mus (own rows / entries) enter as the last filtered
// belief (= smoothed belief at T-1).
// ---------------------------------------------------------------------------------------
template <class C>
KV_FN void smoother_sweep(const Args& a, const float* base, const FTiles<C>& tl, const Group<C::L, C::R>& g, int b, bool active,
                          float* stage_slot, float (&Sig)[C::R][C::N], float (&mus)[C::R]) {
  constexpr int N = C::N, R = C::R, L = C::L;
  constexpr bool MEM = C::MEM;
  const TileRef X0 = tl.nn(0), X1 = tl.nn(1), X2 = tl.nn(2), VB = tl.vec(0);
  const int row0 = g.row0();
  const int T = a.T;
  bool ok = true;
  {  // t = T-1: copied, not symmetrised (kalman_filter.py:251-256)
    const long bt = (long)b * T + (T - 1);
    if (active) {
      KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(a.Sig_s + (bt * N + row0 + r) * N, Sig[r]);
      store_row<R>(a.mu_s + bt * N + row0, mus);
    }
  }
  const bool staged = (T % 4 == 0) && T >= 4;
  InStage<C, false> ins;
  ins.slot = stage_slot;
  if (staged) {   // chunk holding alpha_{T-1}, then keep one chunk ahead (towards t = 0)
    ins.prefetch(a, g, (long)b * T + (T - 4));
    ins.commit(g);
    if (T >= 8) ins.prefetch(a, g, (long)b * T + (T - 8));
  }
  const bool pf_l2 = T >= 64;   // (see k_seq_fwd: short sequences gain nothing)
  (void)pf_l2;
  SmoothIn<C> pf;   // software prefetch: the states of the step after the current one
  if (T >= 2) load_smooth_in<C>(a, (long)b * T + (T - 2), row0, pf);
  for (int t = T - 2; t >= 0; --t) {
    const long bt = (long)b * T + t;
    float al1[C::K];
    if (staged) {
      StepIn<C> tmp;
      ins.read((t + 1) & 3, tmp);
      KV_UNROLL for (int k = 0; k < C::K; ++k) al1[k] = tmp.al[k];
      // alpha_{t+1} was the first step of its chunk: switch to the previous chunk (its prefetch follows the mixing)
      if (((t + 1) & 3) == 0 && t + 1 >= 4) ins.commit(g);
    } else {
      if (a.alpha) load_row<C::K>(a.alpha + (bt + 1) * C::K, al1);
      else { KV_UNROLL for (int k = 0; k < C::K; ++k) al1[k] = 0.f; }
    }
    // this step's states were fetched one iteration ago; start fetching the next step's now
    float Sf[R][N], Sp1[R][N], muf[R], mup1[R];
    KV_UNROLL for (int r = 0; r < R; ++r) {
      KV_UNROLL for (int j = 0; j < N; ++j) { Sf[r][j] = pf.Sf[r][j]; Sp1[r][j] = pf.Sp1[r][j]; }
      muf[r] = pf.muf[r]; mup1[r] = pf.mup1[r];
    }
    float A1[R][N];
    get_A<C>(a, base, al1, row0, bt + 1, A1);
    if (a.a_smooth) {   // C_{t+1} mu_{t+1|T} (model.py:280-281): alpha_{t+1} and the smoothed mean at t+1 are in hand here
      float Ct1[R][C::P], as[C::P];
      get_Ct<C>(a, base, al1, row0, bt + 1, Ct1);
      project_obs<C>(g, Ct1, mus, as);
      if (active && g.lane == 0) store_row<C::P>(a.a_smooth + (bt + 1) * C::P, as);
    }
    // (issued AFTER the mixing loads: directly in front of them the prefetch shared a scoreboard with the LDS of the
    //  mixing, so the first mixing FMA waited for the whole L2 round trip -- 40 % of this kernel's stall samples)
    if (staged && ((t + 1) & 3) == 0 && t + 1 >= 8) ins.prefetch(a, g, (long)b * T + (t + 1) - 8);
    if (t > 0) load_smooth_in<C>(a, bt - 1, row0, pf);
#if defined(__CUDA_ARCH__) && KV_SMOOTH_PF_L2 > 0
    if (pf_l2 && t > KV_SMOOTH_PF_L2) {   // long sequences: pull the rows of a later iteration into L2 (no registers held)
      const long bq = bt - 1 - KV_SMOOTH_PF_L2;
      KV_UNROLL for (int r = 0; r < R; ++r) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Sig_f + (bq * N + row0 + r) * N));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.Sig_p + ((bq + 1) * N + row0 + r) * N));
      }
      if (g.lane == 0) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mu_f + bq * N));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mu_p + (bq + 1) * N));
      }
    }
#endif
    ok = smoother_step_math<C>(g, tl, Sf, Sp1, muf, mup1, A1, Sig, mus) && ok;
    if (active) {
      KV_UNROLL for (int r = 0; r < R; ++r) store_row<N>(a.Sig_s + (bt * N + row0 + r) * N, Sig[r]);
      store_row<R>(a.mu_s + bt * N + row0, mus);
    }
  }
  if (a.a_smooth) {   // t = 0 (alpha_0 was never needed by the smoother itself)
    const long bt0 = (long)b * T;
    float al0[C::K], Ct0[R][C::P], as[C::P];
    if (a.alpha) load_row<C::K>(a.alpha + bt0 * C::K, al0);
    else { KV_UNROLL for (int k = 0; k < C::K; ++k) al0[k] = 0.f; }
    get_Ct<C>(a, base, al0, row0, bt0, Ct0);
    project_obs<C>(g, Ct0, mus, as);
    if (active && g.lane == 0) store_row<C::P>(a.a_smooth + bt0 * C::P, as);
  }
  if (!ok && active) kv_info_or(a.info, KV_INFO_PIVOT);
}

}  // namespace kvae
