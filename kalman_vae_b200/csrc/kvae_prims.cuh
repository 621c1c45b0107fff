// kvae_prims.cuh — row-distributed small-matrix primitives for the Kalman hot path.
//
// A sequence is owned by a GROUP of L lanes of one warp.  Every n-row matrix of the recursion
// (Sigma, A_t, B_t, C_t^T, K, J, ...) is distributed by rows: lane l of the group holds rows
// [l*R, l*R+R), R = n/L, in registers.  A product needs one operand "fully visible"; for L>1
// that operand is PUBLISHED into a per-group shared-memory tile (row stride padded so that
// 128-bit row writes and broadcast row reads are bank-conflict free) and read back with
// 128-bit broadcast loads; for L==1 the "view" simply aliases the registers, so the whole
// recursion runs out of the register file with no shared-memory traffic at all.
//
// The same header compiles for the host (L==1 only): tests/hostsim builds it with g++ to check
// the arithmetic against the oracle on the CPU-only build box.  That build is test tooling and
// is never loaded by the package.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KV_FN __host__ __device__ __forceinline__
#else
#define KV_FN inline __attribute__((always_inline))
#endif
#if defined(__CUDA_ARCH__)
#define KV_UNROLL _Pragma("unroll")
#else
#define KV_UNROLL
#endif

namespace kvae {

// reciprocal square root to ~1 ulp: hardware approximation (MUFU.RSQ, 2 ulp) + one Newton step; one
// special-function op on the critical path of a Cholesky pivot instead of sqrt followed by a reciprocal
// reciprocal to ~1 ulp: MUFU.RCP (1 ulp) + one Newton step, no slow-path branch; an LU pivot is on the critical path
// of every smoother step (four sequential pivots for n = 4), the IEEE division routine costs ~3x the latency
KV_FN float kv_rcp(float x) {
#if defined(__CUDA_ARCH__)
  float y0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(x));
  return fmaf(y0, fmaf(-x, y0, 1.0f), y0);
#else
  return 1.0f / x;
#endif
}
KV_FN float kv_rsqrt(float s) {
#if defined(__CUDA_ARCH__)
  const float y0 = rsqrtf(s);
  const float e = fmaf(-s * y0, y0, 1.0f);
  return fmaf(0.5f * y0, e, y0);
#else
  return 1.0f / sqrtf(s);
#endif
}

struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };

// ---------------------------------------------------------------------------------------
// Packed fp32 arithmetic (sm_100: fma.rn.f32x2 -> FFMA2, two IEEE fused multiply-adds per issue slot).
// The kernels are instruction-issue bound, not FMA-pipe bound, so halving the FMA issue count pays directly.
// ptxas folds the mov.b64 packs into register-pair operands (and a repeated scalar into a broadcast operand):
// the SASS has no extra moves (profiles/r02_sass_opcodes.txt).  Host build (tests/hostsim): plain fmaf.
//   (c0,c1) = (a0,a1) * (b0,b1) + (c0,c1)
// ---------------------------------------------------------------------------------------
KV_FN void kv_fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
#if defined(__CUDA_ARCH__)
  unsigned long long ra, rb, rc;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c0), "f"(c1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(rc));
#else
  c0 = fmaf(a0, b0, c0);
  c1 = fmaf(a1, b1, c1);
#endif
}
// c[j] += a * y[j], j = 0..NC-1 (pairs packed; identical rounding to the scalar loop)
template <int NC> KV_FN void kv_axpy(float a, const float (&y)[NC], float (&c)[NC]) {
  KV_UNROLL for (int j = 0; j + 1 < NC; j += 2) kv_fma2(c[j], c[j + 1], a, a, y[j], y[j + 1]);
  if constexpr (NC % 2 == 1) c[NC - 1] = fmaf(a, y[NC - 1], c[NC - 1]);
}
// s + sum_k x[k] y[k]: even and odd k accumulate in the two halves of one packed register (s rides on the even half),
// i.e. (s + x0 y0 + x2 y2 + ..) + (x1 y1 + x3 y3 + ..)
template <int KD> KV_FN float kv_dot(const float (&x)[KD], const float (&y)[KD], float s) {
  if constexpr (KD >= 4 && KD % 2 == 0) {
    float s1 = 0.f;
    KV_UNROLL for (int k = 0; k < KD; k += 2) kv_fma2(s, s1, x[k], x[k + 1], y[k], y[k + 1]);
    return s + s1;
  } else {
    KV_UNROLL for (int k = 0; k < KD; ++k) s = fmaf(x[k], y[k], s);
    return s;
  }
}

// ---------------------------------------------------------------------------------------
// Static problem description.
//   N,P,M,K : z_dim, a_dim, u_dim, number of mixture modes
//   L       : lanes per sequence (divides N); R rows per lane
//   QPM     : Q_t = sum_k alpha_k Q_k (switching) instead of one fixed Q (lstm)
//   CSH     : C_t = C_0 (switching) instead of sum_k alpha_k C_k
//   FM      : force the shared-memory ("published") code path even for L==1 (host tests)
// ---------------------------------------------------------------------------------------
template <int N_, int P_, int M_, int K_, int L_, bool QPM_, bool CSH_, bool FM_ = false>
struct Cfg {
  static constexpr int N = N_, P = P_, M = M_, K = K_, L = L_, R = N_ / L_;
  static constexpr bool QPM = QPM_, CSH = CSH_;
  static constexpr bool MEM = (L_ > 1) || FM_;
  static constexpr int KC = CSH_ ? 1 : K_;
  static constexpr int KQ = QPM_ ? K_ : 1;
  static_assert(N_ % L_ == 0, "lanes per sequence must divide z_dim");
  static_assert(L_ == 1 || L_ == 2 || L_ == 4 || L_ == 8 || L_ == 16 || L_ == 32, "L must be a power of two <= 32");
};

// padded leading dimension of a published [rows x COLS] tile
template <int COLS> struct ld_of { static constexpr int v = (COLS % 8 == 0) ? COLS + 4 : COLS; };

// ---------------------------------------------------------------------------------------
// vector row access (16-byte when the row length allows it)
// ---------------------------------------------------------------------------------------
template <int COLS> KV_FN void load_row(const float* __restrict__ p, float (&o)[COLS]) {
  if constexpr (COLS % 4 == 0) {
    KV_UNROLL for (int q = 0; q < COLS / 4; ++q) {
      f4 v = *reinterpret_cast<const f4*>(p + 4 * q);
      o[4 * q] = v.x; o[4 * q + 1] = v.y; o[4 * q + 2] = v.z; o[4 * q + 3] = v.w;
    }
  } else if constexpr (COLS % 2 == 0) {
    KV_UNROLL for (int q = 0; q < COLS / 2; ++q) {
      f2 v = *reinterpret_cast<const f2*>(p + 2 * q);
      o[2 * q] = v.x; o[2 * q + 1] = v.y;
    }
  } else {
    KV_UNROLL for (int q = 0; q < COLS; ++q) o[q] = p[q];
  }
}
template <int COLS> KV_FN void store_row(float* __restrict__ p, const float (&o)[COLS]) {
  if constexpr (COLS % 4 == 0) {
    KV_UNROLL for (int q = 0; q < COLS / 4; ++q) {
      f4 v; v.x = o[4 * q]; v.y = o[4 * q + 1]; v.z = o[4 * q + 2]; v.w = o[4 * q + 3];
      *reinterpret_cast<f4*>(p + 4 * q) = v;
    }
  } else if constexpr (COLS % 2 == 0) {
    KV_UNROLL for (int q = 0; q < COLS / 2; ++q) {
      f2 v; v.x = o[2 * q]; v.y = o[2 * q + 1];
      *reinterpret_cast<f2*>(p + 2 * q) = v;
    }
  } else {
    KV_UNROLL for (int q = 0; q < COLS; ++q) p[q] = o[q];
  }
}

// ---------------------------------------------------------------------------------------
// Group of L lanes
// ---------------------------------------------------------------------------------------
template <int L, int R> struct Group {
  int lane;       // lane within the group
  unsigned mask;  // the warp lanes of this group: all group collectives name exactly these lanes, so groups
                  // of one warp may diverge from each other (e.g. different time chunks) without deadlock
  KV_FN int row0() const { return L == 1 ? 0 : lane * R; }
  KV_FN void sync() const {
#if defined(__CUDA_ARCH__)
    if constexpr (L > 1) __syncwarp(mask);
#endif
  }
  // value held by lane `src` of this group
  KV_FN float bcast(float v, int src) const {
#if defined(__CUDA_ARCH__)
    if constexpr (L > 1) return __shfl_sync(mask, v, src, L);
#endif
    (void)src;
    return v;
  }
  // butterfly all-reduce: every lane ends with the bit-identical sum
  template <int CNT> KV_FN void allreduce(float (&v)[CNT]) const {
#if defined(__CUDA_ARCH__)
    if constexpr (L > 1) {
      KV_UNROLL for (int off = L / 2; off >= 1; off >>= 1) {
        KV_UNROLL for (int i = 0; i < CNT; ++i) v[i] += __shfl_xor_sync(mask, v[i], off);
      }
    }
#endif
  }
  KV_FN float allreduce1(float v) const {
    float a[1] = {v};
    allreduce(a);
    return a[0];
  }
};

// ---------------------------------------------------------------------------------------
// Shared-memory tile geometry.  Tiles are allocated per WARP (G = 32/L groups each).
//   * n = 4, L = 4 (eight sequences per warp, rows of <= 4 floats): rows of the eight groups are
//     interleaved and rotated, element (r,c) of group g at  r*(G*COLS) + ((g + 2r) mod G)*COLS + c.
//     Row publishes (128-bit, one row per lane), broadcast row reads and the scalar transposed reads
//     at(k, lane) are then all bank-conflict free.
//   * otherwise: group-major, padded row stride, group stride padded to == L (mod 32) floats so that
//     the groups of a warp start in different banks.
// ---------------------------------------------------------------------------------------
struct TileRef { float* p; int g; };

// A/B knobs of the n = 4, L = 4 tile layout (see TileGeom): KV_TILE_INTER = 0 selects the group-major layout there too;
// KV_TILE_RES4 is the residue (mod 32 floats) of its group stride.
#ifndef KV_TILE_INTER
#define KV_TILE_INTER 1
#endif
#ifndef KV_TILE_RES4
#define KV_TILE_RES4 4
#endif
constexpr int pad_res(int x, int L) {  // smallest y >= x, y % 4 == 0, y % 32 == max(L,4) % 32 (L > 1)
  int y = (x + 3) & ~3;
  if (L > 1) { const int want = (L == 4 ? KV_TILE_RES4 : (L < 4 ? 4 : L)) % 32; while (y % 32 != want) y += 4; }
  return y;
}
template <int L, int R, int COLS> struct TileGeom {
  static constexpr int ROWS = L * R;
  static constexpr int G = (L == 1) ? 1 : 32 / L;
  static constexpr bool INTER = KV_TILE_INTER && (L == 4) && (R == 1) && (COLS <= 4);
  static constexpr int LD = INTER ? G * COLS : ld_of<COLS>::v;
  static constexpr int group_floats = INTER ? ROWS * COLS : pad_res(ROWS * ld_of<COLS>::v, L);
  static constexpr int warp_floats = group_floats * G;
  // offset of row r of group g inside the tile whose TileRef.p is used (p already includes the
  // group offset for the group-major layout)
  static KV_FN int row_off(int g, int r) {
    if constexpr (INTER) return r * LD + ((g + 2 * r) & (G - 1)) * COLS;
    else { (void)g; return r * LD; }
  }
};
template <int L, int R> struct VecGeom {  // replicated n-vector slot per group
  static constexpr int G = (L == 1) ? 1 : 32 / L;
  static constexpr int group_floats = (L * R + 3) & ~3;
  static constexpr int warp_floats = group_floats * G;
};

// ---------------------------------------------------------------------------------------
// Fully visible matrix views
// ---------------------------------------------------------------------------------------
template <int L, int R, int COLS> struct MemView {  // tile in shared (or host) memory
  using Geo = TileGeom<L, R, COLS>;
  const float* p;
  int g;
  KV_FN float at(int r, int c) const { return p[Geo::row_off(g, r) + c]; }
  KV_FN void row(int r, float (&o)[COLS]) const { load_row<COLS>(p + Geo::row_off(g, r), o); }
};
template <int ROWS, int COLS> struct RegView {  // alias of a register array (L == 1)
  const float (*v)[COLS];
  KV_FN float at(int r, int c) const { return v[r][c]; }
  KV_FN void row(int r, float (&o)[COLS]) const {
    KV_UNROLL for (int c = 0; c < COLS; ++c) o[c] = v[r][c];
  }
};
template <bool MEM, int L, int R, int COLS> struct view_of { using type = RegView<L * R, COLS>; };
template <int L, int R, int COLS> struct view_of<true, L, R, COLS> { using type = MemView<L, R, COLS>; };

// publish the lane's R rows of an [L*R x COLS] matrix; returns the full view.
// For RegView the view ALIASES x: x must stay unmodified while the view is in use.
template <bool MEM, int L, int R, int COLS>
KV_FN typename view_of<MEM, L, R, COLS>::type publish(const Group<L, R>& g, const float (&x)[R][COLS], TileRef buf) {
  if constexpr (MEM) {
    using Geo = TileGeom<L, R, COLS>;
    g.sync();  // everyone finished reading the previous contents of buf
    KV_UNROLL for (int r = 0; r < R; ++r) store_row<COLS>(buf.p + Geo::row_off(buf.g, g.row0() + r), x[r]);
    g.sync();
    return MemView<L, R, COLS>{buf.p, buf.g};
  } else {
    (void)g; (void)buf;
    return RegView<L * R, COLS>{x};
  }
}

// all-gather of a distributed n-vector (lane holds R entries) into a replicated one
template <bool MEM, int L, int R>
KV_FN void allgather(const Group<L, R>& g, const float (&x)[R], TileRef vbuf, float (&full)[L * R]) {
  if constexpr (MEM) {
    g.sync();
    KV_UNROLL for (int r = 0; r < R; ++r) vbuf.p[g.row0() + r] = x[r];
    g.sync();
    load_row<L * R>(vbuf.p, full);
  } else {
    (void)g; (void)vbuf;
    KV_UNROLL for (int r = 0; r < R; ++r) full[r] = x[r];
  }
}

// The tiles of one warp: NNN [n x n] tiles, NNP [n x p] tiles, NV vector slots.
template <int L, int R, int P, int NNN, int NNP, int NV, bool MEM> struct TileSet {
  using GNN = TileGeom<L, R, L * R>;
  using GNP = TileGeom<L, R, P>;
  using GV = VecGeom<L, R>;
  static constexpr int oNP = NNN * GNN::warp_floats;
  static constexpr int oV = oNP + NNP * GNP::warp_floats;
  static constexpr int warp_total = MEM ? (oV + NV * GV::warp_floats) : 0;
  static constexpr int groups_per_warp = GNN::G;
  float* base;  // this warp's region
  int g;        // group index inside the warp
  KV_FN TileRef nn(int i) const { return TileRef{base + i * GNN::warp_floats + (GNN::INTER ? 0 : g * GNN::group_floats), g}; }
  KV_FN TileRef np(int i) const { return TileRef{base + oNP + i * GNP::warp_floats + (GNP::INTER ? 0 : g * GNP::group_floats), g}; }
  KV_FN TileRef vec(int i) const { return TileRef{base + oV + i * GV::warp_floats + g * GV::group_floats, g}; }
};

// ---------------------------------------------------------------------------------------
// products.  X: local rows; Y / Xv: fully visible views.
// ---------------------------------------------------------------------------------------
// C[r][j] (+)= sum_k X[r][k] * Y(k,j)
template <bool ACC, int R, int KD, int NC, class V>
KV_FN void mm_RS(const float (&X)[R][KD], const V& Y, float (&C)[R][NC]) {
  if constexpr (!ACC) {
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < NC; ++j) C[r][j] = 0.f;
  }
  KV_UNROLL for (int k = 0; k < KD; ++k) {
    float yr[NC];
    Y.row(k, yr);
    KV_UNROLL for (int r = 0; r < R; ++r) kv_axpy<NC>(X[r][k], yr, C[r]);
  }
}
// C[r][j] (+)= sum_k X[r][k] * Y(j,k)
template <bool ACC, int R, int KD, int NC, class V>
KV_FN void mm_RSt(const float (&X)[R][KD], const V& Y, float (&C)[R][NC]) {
  KV_UNROLL for (int j = 0; j < NC; ++j) {
    float yr[KD];
    Y.row(j, yr);
    KV_UNROLL for (int r = 0; r < R; ++r) C[r][j] = kv_dot<KD>(X[r], yr, ACC ? C[r][j] : 0.f);
  }
}
// C[r][j] (+)= sum_k Xv(k, row0+r) * Y(k,j)      (i.e. own rows of Xv^T * Y)
template <bool ACC, int R, int KD, int NC, class VX, class VY>
KV_FN void mm_StS(const VX& Xv, int row0, const VY& Y, float (&C)[R][NC]) {
  if constexpr (!ACC) {
    KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < NC; ++j) C[r][j] = 0.f;
  }
  KV_UNROLL for (int k = 0; k < KD; ++k) {
    float yr[NC];
    Y.row(k, yr);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      kv_axpy<NC>(Xv.at(k, row0 + r), yr, C[r]);
    }
  }
}
// own rows of the transpose: O[r][j] = Xv(j, row0+r)
template <int R, int NC, class VX> KV_FN void tr_rows(const VX& Xv, int row0, float (&O)[R][NC]) {
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < NC; ++j) O[r][j] = Xv.at(j, row0 + r);
}

// ---------------------------------------------------------------------------------------
// Cholesky (lower), replicated small matrix.  Only the lower triangle of a is read.
// Returns false if a pivot was not positive.
// ---------------------------------------------------------------------------------------
template <int D> KV_FN bool chol_small(const float (&a)[D][D], float (&l)[D][D], float (&invd)[D]) {
  bool ok = true;
  KV_UNROLL for (int j = 0; j < D; ++j) {
    float s = a[j][j];
    KV_UNROLL for (int q = 0; q < j; ++q) s = fmaf(-l[j][q], l[j][q], s);
    ok = ok && (s > 0.f);
    invd[j] = kv_rsqrt(s);
    const float d = s * invd[j];
    l[j][j] = d;
    KV_UNROLL for (int i = j + 1; i < D; ++i) {
      float v = a[i][j];
      KV_UNROLL for (int q = 0; q < j; ++q) v = fmaf(-l[i][q], l[j][q], v);
      l[i][j] = v * invd[j];
    }
    KV_UNROLL for (int i = 0; i < j; ++i) l[i][j] = 0.f;
  }
  return ok;
}

// Row-distributed Cholesky of an [N x N] matrix (N = L*R): lane owns rows row0..row0+R-1 of
// a (lower triangle read) and of l.  invd (1/diag) is replicated.  Column j is finished by
// broadcasting the pivot row from its owner lane with warp shuffles.
template <int L, int R>
KV_FN bool chol_dist(const Group<L, R>& g, const float (&a)[R][L * R], float (&l)[R][L * R], float (&invd)[L * R],
                     float (&dg_own)[R]) {
  constexpr int N = L * R;
  bool ok = true;
  KV_UNROLL for (int r = 0; r < R; ++r) dg_own[r] = 1.0f;
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) l[r][j] = 0.f;
  KV_UNROLL for (int j = 0; j < N; ++j) {
    const int owner = j / R, jr = j % R;
    float lj[N];  // row j of L, entries q<j (replicated)
    KV_UNROLL for (int q = 0; q < j; ++q) lj[q] = g.bcast(l[jr][q], owner);
    float s = g.bcast(a[jr][j], owner);
    KV_UNROLL for (int q = 0; q < j; ++q) s = fmaf(-lj[q], lj[q], s);
    ok = ok && (s > 0.f);
    invd[j] = kv_rsqrt(s);
    const float d = s * invd[j];
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const int i = g.row0() + r;
      float v = a[r][j];
      KV_UNROLL for (int q = 0; q < j; ++q) v = fmaf(-l[r][q], lj[q], v);
      v *= invd[j];
      l[r][j] = (i > j) ? v : ((i == j) ? d : 0.f);
      if (i == j) dg_own[r] = d;
    }
  }
  return ok;
}

// chol_dist, or -- diag != 0 -- the last rung of the reference's _safe_cholesky ladder (kalman_filter.py:298-302):
// L = diag(sqrt(clamp(diag(a), 1e-6))) (a: symmetrised, no jitter).  Never fails in that mode.
template <int L, int R>
KV_FN bool chol_dist_opt(const Group<L, R>& g, const float (&a)[R][L * R], float (&l)[R][L * R], float (&invd)[L * R],
                         float (&dg_own)[R], int diag, unsigned& clamped) {
  clamped = 0u;   // bit j: diagonal entry j was clamped (no gradient flows through it); replicated over the group
  if (!diag) return chol_dist<L, R>(g, a, l, invd, dg_own);
  constexpr int N = L * R;
  KV_UNROLL for (int r = 0; r < R; ++r) KV_UNROLL for (int j = 0; j < N; ++j) l[r][j] = 0.f;
  KV_UNROLL for (int j = 0; j < N; ++j) {
    const int owner = j / R, jr = j % R;
    float d = g.bcast(a[jr][j], owner);
    if (!(d > 1e-6f)) clamped |= 1u << j;
    d = sqrtf(d > 1e-6f ? d : 1e-6f);
    invd[j] = 1.0f / d;
    KV_UNROLL for (int r = 0; r < R; ++r) {
      if (g.row0() + r == j) { l[r][j] = d; dg_own[r] = d; }
    }
  }
  return true;
}

// x := x (Lc Lc^T)^-1 for each local row (Lc fully visible, invd = 1/diag(Lc) replicated).
template <int R, int D, class V>
KV_FN void solve_rows_llt(float (&x)[R][D], const V& Lc, const float (&invd)[D]) {
  // y Lc^T = b  (forward, dot form with row j)
  KV_UNROLL for (int j = 0; j < D; ++j) {
    float lj[D];
    Lc.row(j, lj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      float s = x[r][j];
      KV_UNROLL for (int q = 0; q < j; ++q) s = fmaf(-x[r][q], lj[q], s);
      x[r][j] = s * invd[j];
    }
  }
  // z Lc = y    (backward, axpy form with row j)
  KV_UNROLL for (int j = D - 1; j >= 0; --j) {
    float lj[D];
    Lc.row(j, lj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const float xj = x[r][j] * invd[j];
      x[r][j] = xj;
      KV_UNROLL for (int q = 0; q < j; ++q) x[r][q] = fmaf(-xj, lj[q], x[r][q]);
    }
  }
}
// x := x Lc^-1 only (solve z Lc = x), rows local
template <int R, int D, class V>
KV_FN void solve_rows_l(float (&x)[R][D], const V& Lc, const float (&invd)[D]) {
  KV_UNROLL for (int j = D - 1; j >= 0; --j) {
    float lj[D];
    Lc.row(j, lj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const float xj = x[r][j] * invd[j];
      x[r][j] = xj;
      KV_UNROLL for (int q = 0; q < j; ++q) x[r][q] = fmaf(-xj, lj[q], x[r][q]);
    }
  }
}
// w := Lc^-1 v for a replicated vector (forward substitution)
template <int D, class V> KV_FN void solve_vec_l(float (&v)[D], const V& Lc, const float (&invd)[D]) {
  KV_UNROLL for (int j = 0; j < D; ++j) {
    float lj[D];
    Lc.row(j, lj);
    float s = v[j];
    KV_UNROLL for (int q = 0; q < j; ++q) s = fmaf(-v[q], lj[q], s);
    v[j] = s * invd[j];
  }
}
// w := Lc^-T v for a replicated vector (backward substitution, axpy form)
template <int D, class V> KV_FN void solve_vec_lt(float (&v)[D], const V& Lc, const float (&invd)[D]) {
  KV_UNROLL for (int j = D - 1; j >= 0; --j) {
    float lj[D];
    Lc.row(j, lj);
    const float xj = v[j] * invd[j];
    v[j] = xj;
    KV_UNROLL for (int q = 0; q < j; ++q) v[q] = fmaf(-xj, lj[q], v[q]);
  }
}

// ---------------------------------------------------------------------------------------
// Row-distributed LU without pivoting (in place: unit-lower L below the diagonal, U on/above).
// Used for the smoother gain, whose matrix Sigma_{t+1|t} = A Sigma A^T + Q_t is NOT symmetric when
// the learnable Q_k are not (switch_dyn_param.py:18-23; the reference solves it with a general LU,
// kalman_filter.py:229).  Its symmetric part is positive definite, so elimination without pivoting
// is stable.  invu = 1/diag(U), replicated.  Pivot row k is broadcast from its owner lane.
// ---------------------------------------------------------------------------------------
template <int L, int R>
KV_FN bool lu_dist(const Group<L, R>& g, float (&a)[R][L * R], float (&invu)[L * R]) {
  constexpr int N = L * R;
  bool ok = true;
  KV_UNROLL for (int k = 0; k < N; ++k) {
    const int owner = k / R, kr = k % R;
    float uk[N];
    KV_UNROLL for (int j = k; j < N; ++j) uk[j] = g.bcast(a[kr][j], owner);
    ok = ok && (uk[k] > 0.f);
    invu[k] = kv_rcp(uk[k]);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const bool below = (g.row0() + r) > k;
      const float f = below ? a[r][k] * invu[k] : 0.f;
      a[r][k] = below ? f : a[r][k];
      KV_UNROLL for (int j = k + 1; j < N; ++j) a[r][j] = fmaf(-f, uk[j], a[r][j]);
    }
  }
  return ok;
}
// x := x (LU)^-1 for each local row  (x A = b  ->  y U = b, x L = y; both in axpy form with row j)
template <int R, int D, class V>
KV_FN void solve_rows_lu(float (&x)[R][D], const V& LU, const float (&invu)[D]) {
  KV_UNROLL for (int j = 0; j < D; ++j) {
    float uj[D];
    LU.row(j, uj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const float yj = x[r][j] * invu[j];
      x[r][j] = yj;
      KV_UNROLL for (int q = j + 1; q < D; ++q) x[r][q] = fmaf(-yj, uj[q], x[r][q]);
    }
  }
  KV_UNROLL for (int j = D - 1; j >= 1; --j) {
    float lj[D];
    LU.row(j, lj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      const float xj = x[r][j];
      KV_UNROLL for (int q = 0; q < j; ++q) x[r][q] = fmaf(-xj, lj[q], x[r][q]);
    }
  }
}
// x := x (LU)^-T for each local row  (x U^T L^T = b  ->  y L^T = b, then x U^T = y; dot form with row j)
template <int R, int D, class V>
KV_FN void solve_rows_lut(float (&x)[R][D], const V& LU, const float (&invu)[D]) {
  KV_UNROLL for (int j = 1; j < D; ++j) {
    float lj[D];
    LU.row(j, lj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      float s = x[r][j];
      KV_UNROLL for (int q = 0; q < j; ++q) s = fmaf(-x[r][q], lj[q], s);
      x[r][j] = s;
    }
  }
  KV_UNROLL for (int j = D - 1; j >= 0; --j) {
    float uj[D];
    LU.row(j, uj);
    KV_UNROLL for (int r = 0; r < R; ++r) {
      float s = x[r][j];
      KV_UNROLL for (int q = j + 1; q < D; ++q) s = fmaf(-x[r][q], uj[q], s);
      x[r][j] = s * invu[j];
    }
  }
}

}  // namespace kvae
