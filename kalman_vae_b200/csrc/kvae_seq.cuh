// kvae_seq.cuh — thread-per-sequence kernels with TMA-staged streams (z_dim = u_dim = 4).
//
// One THREAD owns one sequence for a whole sweep: every matrix of the recursion lives in its registers, products are
// plain register arithmetic in packed fp32 (FFMA2), there is no cross-lane exchange at all.  That costs ~5x fewer
// instructions per sequence-step than the lane-group kernels (csrc/kvae_kernels.cuh), which are instruction-issue
// bound.  What a thread per sequence cannot do is talk to global memory directly: the 32 sequences of a warp are
// T*W*4 bytes apart in every [B,T,W] tensor, so a direct 128-bit access touches 32 lines per instruction (measured
// on B200, tools/microbench/ub_stream.cu: 1.5 TB/s for stores against 6.6 TB/s for the lane-group pattern).  All
// streams therefore go through shared memory and the TMA engine (async proxy, no LSU wavefronts):
//
//   * per-step STATE streams (rows of W = 4, 8 or 16 floats: mu_*, Sigma_*, A/B/C lists, scratch): the tensor is
//     described as (W, T, B) with box (W, 1, 32): ONE cp.async.bulk.tensor per warp, stream and time step moves the
//     rows of the warp's 32 sequences between global memory and a [32][W] shared-memory tile.  Tiles with 32- / 64-byte
//     rows use the TMA 32B / 64B swizzle so that the thread-per-row 128-bit accesses are bank-conflict free.
//   * the tiny per-step INPUT streams (y_t, u_t, alpha_t, mask_t, and dY / dalpha / dU going out): rows of 8 / 12 / 4
//     bytes cannot be TMA boxes, so FOUR time steps travel together: tensor (T*W, B), box (4W, 32) -> [32][4W].
//     Needs T % 4 == 0; other lengths run on the lane-group kernels.
//   * chunk loads complete on an mbarrier (complete_tx) and are re-armed one step before their first use; stores leave
//     from a single staging tile per warp and stream (bulk_group; the next step waits for the tile to be READ, not
//     for the write to land).
//   * the TMA engine moves ~14-18 bytes per cycle and SM for such short rows (a 16-byte row costs ~2 cycles, a 64-byte
//     row ~4.5; tools/microbench/ub_stream.cu), i.e. 4-5 TB/s for the whole chip: a kernel that pushes ALL its traffic
//     through it is TMA-bound below the HBM roofline (measured: forward 2.3 TB/s algorithmic at 75 % TMA occupancy).
//     So the two engines share the work: the 64-/32-byte rows and the input chunks use TMA, the 16-byte mean rows are
//     stored with plain 128-bit stores, and the states a sweep READS BACK (smoother: Sigma_f(t), Sigma_p(t+1), ...)
//     are fetched with ordinary vector loads one step ahead (each thread reads whole sectors of its own rows, which
//     the LSU path sustains at 4.1 TB/s).
//
// Reference arithmetic: kvae/kalman/kalman_filter.py:31-279 (sweeps 1-2 here), :305-401 + autograd (sweeps 3-4,
// csrc/kvae_seq_bwd.cuh).  The step arithmetic is the same code as the lane-group kernels (filter_step_math,
// smoother_step_math, ... instantiated with L = 1: "publish" aliases registers).
#pragma once
#ifndef KV_SEQ_PF_L2
#define KV_SEQ_PF_L2 6   // smoother: L2 prefetch distance (iterations beyond the two-ahead register loads); 0 = off
#endif

#include <cuda.h>
#include "kvae_kernels.cuh"

namespace kvae {
namespace tma {

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane of the (converged) warp: the form ptxas recognises, so the uniform-datapath TMA instructions in the
// branch are issued once instead of inside a per-lane serialisation loop (the plain `lane == 0` test compiles to an
// ELECT / BRA.U.ANY loop around EVERY UTMALDG / UTMASTG: ~100 issue cycles each)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P; }" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void load2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void load3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void store2d(const CUtensorMap* m, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
               ::"l"(m), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void store3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
// L2 eviction-priority hints: streams that nobody reads again soon (A/B/C lists, smoothed states, final gradients) are
// written evict_first so that they do not push the re-read streams (filtered / predicted states, adjoint scratch) out
// of L2 between the sweep that writes them and the sweep that reads them back; last-use loads are evict_first too.
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// (measured on B200, B = 262 144, T = 20: the hints cost 5 % in the forward kernel and change nothing in the adjoint:
//  compiled in only with -DKV_SEQ_L2_HINTS=1)
#ifndef KV_SEQ_L2_HINTS
#define KV_SEQ_L2_HINTS 0
#endif
#if !KV_SEQ_L2_HINTS
__device__ __forceinline__ void load2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar, uint64_t) { load2d(dst, m, c0, c1, bar); }
__device__ __forceinline__ void load3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar, uint64_t) { load3d(dst, m, c0, c1, c2, bar); }
__device__ __forceinline__ void store2d(const CUtensorMap* m, int c0, int c1, uint32_t src, uint64_t) { store2d(m, c0, c1, src); }
__device__ __forceinline__ void store3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src, uint64_t) { store3d(m, c0, c1, c2, src); }
#else
__device__ __forceinline__ void load2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void load3d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, uint32_t bar, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
               ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void store2d(const CUtensorMap* m, int c0, int c1, uint32_t src, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
               ::"l"(m), "r"(c0), "r"(c1), "r"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void store3d(const CUtensorMap* m, int c0, int c1, int c2, uint32_t src, uint64_t pol) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
               ::"l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(src), "l"(pol) : "memory");
}
#endif
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }   // all but the newest group
__device__ __forceinline__ void wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes of this thread to shared memory become visible to the async proxy (TMA store source)
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma

// ---------------------------------------------------------------------------------------
// [32][W] tile of one warp, W = 4 * PIECES floats per row, laid out as the TMA box with the swizzle of its row size:
// 16-byte piece q of row r lives at  r*W*4 + 16*(q ^ sw(r))  bytes, sw(r) = (r>>2)&1 for 32-byte rows (SWIZZLE_32B:
// address bit 4 ^= bit 7), (r>>1)&3 for 64-byte rows (SWIZZLE_64B: bits 4-5 ^= bits 7-8), 0 otherwise.  The tile base
// is 1024-byte aligned, so tile-relative and absolute address bits agree.  With it the 128-bit accesses of eight
// consecutive lanes (one shared-memory phase) cover all 32 banks exactly once.
// ---------------------------------------------------------------------------------------
template <int PIECES> struct RowTile {
  static constexpr int W = 4 * PIECES;
  static constexpr int floats = 32 * W;
  static constexpr uint32_t bytes = 32u * W * 4u;
  static __device__ __forceinline__ int off(int r, int q) {
    if constexpr (PIECES == 2) return r * W + 4 * (q ^ ((r >> 2) & 1));
    else if constexpr (PIECES == 4) return r * W + 4 * (q ^ ((r >> 1) & 3));
    else return r * W + 4 * q;
  }
  static __device__ __forceinline__ void ld(const float* tile, int r, int q, float (&o)[4]) {
    const f4 v = *reinterpret_cast<const f4*>(tile + off(r, q));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  static __device__ __forceinline__ void st(float* tile, int r, int q, const float (&o)[4]) {
    f4 v; v.x = o[0]; v.y = o[1]; v.z = o[2]; v.w = o[3];
    *reinterpret_cast<f4*>(tile + off(r, q)) = v;
  }
  // a [PIECES][4] row-major matrix of the lane's sequence (row r of the tile)
  static __device__ __forceinline__ void ld_mat(const float* tile, int r, float (&m)[PIECES][4]) {
#pragma unroll
    for (int q = 0; q < PIECES; ++q) ld(tile, r, q, m[q]);
  }
  static __device__ __forceinline__ void st_mat(float* tile, int r, const float (&m)[PIECES][4]) {
#pragma unroll
    for (int q = 0; q < PIECES; ++q) st(tile, r, q, m[q]);
  }
};

// tensor maps of the forward kernel (host: make_seq_fwd_maps)
struct alignas(64) SeqFwdMaps {
  CUtensorMap Y, U, alpha, mask;                     // (T*W, B), box (4W, 32)
  CUtensorMap Sig_p, Sig_f, Sig_s, A, B;             // (16, T, B), box (16, 1, 32), SWIZZLE_64B
  CUtensorMap C;                                     // (4P, T, B), box (4P, 1, 32), swizzle of a 16P-byte row
};

__host__ __device__ constexpr int kv_align_up(int x, int a) { return (x + a - 1) / a * a; }

// Shared-memory plan of one warp of k_seq_fwd (bytes).  Swizzled tiles need their base aligned to the swizzle period
// (512 bytes for SWIZZLE_64B, 256 for SWIZZLE_32B); every tile is 512-byte aligned.  Sweep 2 reuses sweep 1's region.
// sets of output staging tiles per warp: with two, step t+1 fills its tiles while the TMA engine still reads step t's
#ifndef KV_SEQ_OUT_BUFS
#define KV_SEQ_OUT_BUFS 2
#endif
template <class C> struct SeqFwdPlan {
  static constexpr int P = C::P, K = C::K, M = C::M;
  using TM = RowTile<4>;    // Sigma / A / B rows
  using TC = RowTile<P>;    // C rows ([P][4])
  // output tiles of one step (sweep 2 reuses oSp for Sigma_s), then two input-chunk buffers (the next chunk is requested
  // three steps before its first use; the 78 KB floor per CTA leaves room for them)
  static constexpr int oSp = 0, oSf = oSp + TM::bytes, oA = oSf + TM::bytes, oB = oA + TM::bytes;
  static constexpr int oC = oB + TM::bytes;
  static constexpr int in_Y = 0, in_U = in_Y + 32 * 4 * P * 4, in_al = in_U + 32 * 4 * M * 4, in_m = in_al + 32 * 4 * K * 4;
  static constexpr int in_bytes = kv_align_up(in_m + 32 * 4 * 4, 512);
  static __host__ __device__ constexpr uint32_t in_tx(bool has_u, bool has_m) {
    return 32u * 4 * P * 4 + (has_u ? 32u * 4 * M * 4 : 0u) + 32u * 4 * K * 4 + (has_m ? 32u * 4 * 4 : 0u);
  }
  static constexpr int out_bytes = kv_align_up(oC + TC::bytes, 512);
  static constexpr int oIn = KV_SEQ_OUT_BUFS * out_bytes;
  static constexpr int warp_bytes = kv_align_up(oIn + 2 * in_bytes, 512);   // two input-chunk buffers
};

#ifndef KV_SEQ_MAXWARPS
#define KV_SEQ_MAXWARPS 4
#endif
template <class C> constexpr size_t seq_fwd_smem(int warps) {
  return 512 + (size_t)kv_align_up((int)sizeof(float) * Base<C>::total, 512) + (size_t)warps * SeqFwdPlan<C>::warp_bytes;
}
// warps per CTA.  A warp stages 19 KB; four-warp CTAs, two per SM (see seq_smem_floor); small batches use two-warp CTAs so
// that the few warps spread over all SMs and land on distinct schedulers.
inline int seq_env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }
inline int seq_warps_per_cta(int B) {
  static const int forced = seq_env_int("KVAE_SEQ_FWD_WARPS", 0);   // development knob
  if (forced >= 1 && forced <= KV_SEQ_MAXWARPS) return forced;
  return ((B + 31) / 32 > 148 * 8) ? KV_SEQ_MAXWARPS : 2;
}
// Resident CTAs per SM are capped at TWO by requesting at least 78 KB of shared memory per CTA (3 x 79 KB > 227 KB) --
// and not more, because the unified L1/shared array is carved per SM: what the CTAs do not take stays L1, which the
// smoother's read-back loads and the adjoint's register spills live in.  Measured on B200 (B = 262 144, T = 20;
// profiles/r02_seq_occupancy_sweep.log): forward 0.92 ms at three CTAs per SM, 0.78 ms at two with ~70 KB of L1, 1.05 ms
// at two with 23 KB of L1 (100 KB requested per CTA); adjoint 2.31 / 1.93 / 2.6 ms.  KVAE_SEQ_SMEM_FLOOR overrides
// the floor (development knob).
inline size_t seq_smem_floor() { static const int v = seq_env_int("KVAE_SEQ_SMEM_FLOOR", 78 * 1024); return (size_t)v; }
inline int seq_grid(int B) { const int per = 32 * seq_warps_per_cta(B); return (B + per - 1) / per; }

struct SeqBar {   // mbarrier + the parity of its next completion
  uint32_t addr, phase;
  __device__ __forceinline__ void wait() { tma::bar_wait(addr, phase & 1u); ++phase; }
};

// step inputs from a staged 4-step chunk ([32][4W] rows, no swizzle)
template <class C>
__device__ __forceinline__ void seq_read_step(const unsigned char* in, int lane, int s, bool has_u, bool has_m, StepIn<C>& cur) {
  using PL = SeqFwdPlan<C>;
  constexpr int P = C::P, M = C::M, K = C::K;
  const float* y = reinterpret_cast<const float*>(in + PL::in_Y) + lane * 4 * P + s * P;
#pragma unroll
  for (int j = 0; j < P; ++j) cur.y[j] = y[j];
  if (has_u) {
    const float* u = reinterpret_cast<const float*>(in + PL::in_U) + lane * 4 * M + s * M;
    load_row<M>(u, cur.u);
  } else {
#pragma unroll
    for (int j = 0; j < M; ++j) cur.u[j] = 0.f;
  }
  const float* al = reinterpret_cast<const float*>(in + PL::in_al) + lane * 4 * K + s * K;
#pragma unroll
  for (int k = 0; k < K; ++k) cur.al[k] = al[k];
  cur.m = has_m ? (reinterpret_cast<const float*>(in + PL::in_m))[lane * 4 + s] : 1.0f;
}

template <class C>
__global__ void __launch_bounds__(32 * KV_SEQ_MAXWARPS, 2) k_seq_fwd(Args a, BasePtrs bp, const __grid_constant__ SeqFwdMaps mp,
                                                                  int smooth) {
  static_assert(C::L == 1 && C::N == 4 && C::M == 4, "thread-per-sequence kernels: z_dim = u_dim = 4");
  constexpr int N = C::N, P = C::P, M = C::M, K = C::K, R = C::R;
  using PL = SeqFwdPlan<C>;
  using TM = typename PL::TM;
  using TC = typename PL::TC;
  extern __shared__ unsigned char seq_smem_raw[];
  __shared__ __align__(8) unsigned long long bars[KV_SEQ_MAXWARPS][6];
  __shared__ float mred[KV_SEQ_MAXWARPS];
  unsigned char* sm = seq_smem_raw + ((512u - (tma::s32(seq_smem_raw) & 511u)) & 511u);
  float* base = reinterpret_cast<float*>(sm);
  stage_base<C>(base, bp);   // __syncthreads inside
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = sm + kv_align_up((int)sizeof(float) * Base<C>::total, 512) + warp * PL::warp_bytes;
  const int b0 = blockIdx.x * blockDim.x + warp * 32;   // first sequence of this warp
  const int b = b0 + lane;
  const bool active = b < a.B;
  const bool warp_on = b0 < a.B;
  const int T = a.T, nchunk = T >> 2;
  const bool has_u = a.U != nullptr, has_m = a.mask != nullptr;
  const Group<1, R> g{0, 0xffffffffu};
  const FTiles<C> tl{base, 0};   // L = 1: never dereferenced (views alias registers)
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) tma::bar_init(tma::s32(&bars[warp][i]), 1);
    tma::bar_init_fence();
  }
  __syncwarp();

  float Sig[R][N], mu[N], mus[R];
  float msum = 0.f;
  bool ok = true;
  const uint64_t pol_ef = tma::policy_evict_first(), pol_el = tma::policy_evict_last();
  if (warp_on) {
    // ------------------------------------------------------------------ sweep 1: filter
    copy_rows<C, N>(base + Base<C>::oS0, 0, Sig);
    load_row<N>(base + Base<C>::oMu0, mu);
    if (a.Sig_init && active) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_init + ((long)b * N + r) * N, Sig[r]); }
    if (a.mu_init && active) load_row<N>(a.mu_init + (long)b * N, mu);
    SeqBar bar_in[2] = {{tma::s32(&bars[warp][0]), 0u}, {tma::s32(&bars[warp][1]), 0u}};
    unsigned char* ib0 = wsm + PL::oIn;
    auto issue_in = [&](int c) {   // elected lane: the four input streams of chunk c into buffer c & 1
      unsigned char* ib = ib0 + (c & 1) * PL::in_bytes;
      const uint32_t bar = bar_in[c & 1].addr;
      tma::bar_expect(bar, PL::in_tx(has_u, has_m));
      tma::load2d(tma::s32(ib + PL::in_Y), &mp.Y, 4 * c * P, b0, bar, pol_ef);
      if (has_u) tma::load2d(tma::s32(ib + PL::in_U), &mp.U, 4 * c * M, b0, bar, pol_ef);
      tma::load2d(tma::s32(ib + PL::in_al), &mp.alpha, 4 * c * K, b0, bar, smooth ? pol_el : pol_ef);
      if (has_m) tma::load2d(tma::s32(ib + PL::in_m), &mp.mask, 4 * c, b0, bar, pol_ef);
    };
    if (tma::elect_one()) issue_in(0);
    bar_in[0].wait();
    StepIn<C> cur;
    seq_read_step<C>(ib0, lane, 0, has_u, has_m, cur);
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      {
        // inputs of step t+1 (they become `cur` of the next iteration).  Chunk c+1 is requested at the first step of
        // chunk c (its buffer held chunk c-1, whose last step was read two steps ago): three steps before its first use
        StepIn<C> nxt = cur;
        if (t + 1 < T) {
          const int c1 = (t + 1) >> 2;
          if (((t + 1) & 3) == 0) bar_in[c1 & 1].wait();
          seq_read_step<C>(ib0 + (c1 & 1) * PL::in_bytes, lane, (t + 1) & 3, has_u, has_m, nxt);
        }
        __syncwarp();
        if ((t & 3) == 0 && t + 4 < T && tma::elect_one()) issue_in((t >> 2) + 1);
        float A[R][N], Bm[R][M], Ct[R][P], Q[R][N];
        mix_A<C>(base, cur.al, 0, A);
        mix_B<C>(base, cur.al, 0, Bm);
        mix_Ct<C>(base, cur.al, 0, Ct);
        mix_Q<C>(base, cur.al, 0, Q);
        // this step's set of staging tiles is free once the stores that last used it have READ it
        unsigned char* ot = wsm + (KV_SEQ_OUT_BUFS == 2 ? (t & 1) * PL::out_bytes : 0);
        float* tSp = reinterpret_cast<float*>(ot + PL::oSp);
        float* tSf = reinterpret_cast<float*>(ot + PL::oSf);
        float* tA = reinterpret_cast<float*>(ot + PL::oA);
        float* tB = reinterpret_cast<float*>(ot + PL::oB);
        float* tC = reinterpret_cast<float*>(ot + PL::oC);
        if (KV_SEQ_OUT_BUFS == 2) tma::wait_read1(); else tma::wait_read0();   // all lanes: only the electing lane has groups pending
        __syncwarp();
        if (a.A_list) TM::st_mat(tA, lane, A);
        if (a.B_list) TM::st_mat(tB, lane, Bm);
        if (a.C_list) {
          float Cm[P][4];
#pragma unroll
          for (int q = 0; q < P; ++q)
#pragma unroll
            for (int j = 0; j < N; ++j) Cm[q][j] = Ct[j][q];
          TC::st_mat(tC, lane, Cm);
        }
        FilterStepOut<C> fo;
        msum += active ? cur.m : 0.f;
        ok = filter_step_math<C>(g, base, tl, cur, A, Bm, Ct, Q, Sig, mu, fo) && ok;
        TM::st_mat(tSp, lane, fo.Sp);
        TM::st_mat(tSf, lane, fo.Sf);
        if (active) {   // 16-byte rows: plain stores (see the file header)
          store_row<N>(a.mu_p + ((long)b * T + t) * N, fo.mup);
          store_row<N>(a.mu_f + ((long)b * T + t) * N, fo.muf);
          if (a.a_filt) {   // C_t mu_{t|t} (model.py:287-288)
            float af[P];
            project_obs<C>(g, Ct, fo.muf, af);
            store_row<P>(a.a_filt + ((long)b * T + t) * P, af);
          }
        }
        tma::fence_async();
        __syncwarp();
        if (tma::elect_one()) {
          tma::store3d(&mp.Sig_p, 0, t, b0, tma::s32(tSp), pol_el);   // read back by the smoother sweep
          tma::store3d(&mp.Sig_f, 0, t, b0, tma::s32(tSf), pol_el);
          if (a.A_list) tma::store3d(&mp.A, 0, t, b0, tma::s32(tA), pol_ef);
          if (a.B_list) tma::store3d(&mp.B, 0, t, b0, tma::s32(tB), pol_ef);
          if (a.C_list) tma::store3d(&mp.C, 0, t, b0, tma::s32(tC), pol_ef);
          tma::commit();
        }
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) Sig[r][j] = fo.Sf[r][j];
          mu[r] = fo.muf[r];
        }
        cur = nxt;
      }
    }
    KV_UNROLL for (int r = 0; r < R; ++r) mus[r] = mu[r];
  }
  if (a.mask_part) {   // per-CTA sum of the mask, fixed order (see k_filter_smooth)
    float v = msum;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) mred[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int wq = 0; wq < (int)(blockDim.x >> 5); ++wq) tot += mred[wq];
      a.mask_part[blockIdx.x] = tot;
    }
  }
  if (warp_on && smooth) {
    // ------------------------------------------------------------------ sweep 2: RTS smoother, t = T-2 .. 0
    // sweep 1's TMA stores must have LANDED: the smoother reads Sigma_f / Sigma_p back with ordinary loads
    tma::wait_all0();
    __syncwarp();
    float* tSs0 = reinterpret_cast<float*>(wsm + PL::oSp);
    float* tSs1 = reinterpret_cast<float*>(wsm + (KV_SEQ_OUT_BUFS == 2 ? PL::oSf : PL::oSp));
    float* tSs = tSs1;
    const int bl = active ? b : a.B - 1;   // tail lanes re-read the last sequence and store nothing
    // t = T-1: copied, not symmetrised (kalman_filter.py:251-256)
    TM::st_mat(tSs, lane, Sig);
    if (active) store_row<N>(a.mu_s + ((long)b * T + (T - 1)) * N, mus);
    tma::fence_async();
    __syncwarp();
    if (tma::elect_one()) {
      tma::store3d(&mp.Sig_s, 0, T - 1, b0, tma::s32(tSs), pol_ef);
      tma::commit();
    }
    // What an iteration needs -- Sigma_f(t), Sigma_p(t+1), mu_f(t), mu_p(t+1), alpha_{t+1} -- is requested TWO iterations
    // ahead (with the machine full one iteration does not cover the HBM latency: 21 % of this kernel's stall samples sat on
    // the first use of a one-step prefetch).  The two engines share the work: the 64-byte covariance rows come through the
    // TMA engine into a ring of three [Sigma_f | Sigma_p] tile pairs (mbarrier per slot; the ring reuses sweep 1's
    // staging area), the 16-byte mean rows and alpha through plain vector loads into two alternating register sets.
    // (All of it through vector loads is bound by L1 wavefronts -- 32 lines per instruction: 0.91 ms of smoother time at
    //  B = 65 536, T = 200; all of it through TMA is bound by the TMA engine, see the file header.)
    struct VecPf { float muf[R], mup1[R], al[K]; };
    VecPf pfA, pfB;
    auto load_vec = [&](int t, VecPf& pf) {   // means of iteration t and alpha_{t+1}
      const long bt = (long)bl * T + t;
      load_row<R>(a.mu_f + bt * N, pf.muf);
      load_row<R>(a.mu_p + (bt + 1) * N, pf.mup1);
      load_row<K>(a.alpha + (bt + 1) * K, pf.al);
    };
    unsigned char* ring = wsm + 2 * (int)TM::bytes;   // behind the two Sigma_s tiles; 3 x 4 KB, inside sweep 1's staging area
    static_assert(2 * TM::bytes + 3 * 2 * TM::bytes <= (unsigned)PL::warp_bytes, "smoother ring exceeds the warp's staging area");
    SeqBar bar_st[3] = {{tma::s32(&bars[warp][2]), 0u}, {tma::s32(&bars[warp][3]), 0u}, {tma::s32(&bars[warp][4]), 0u}};
    auto issue_st = [&](int t, int slot) {   // elected lane
      unsigned char* buf = ring + slot * (2 * (int)TM::bytes);
      tma::bar_expect(bar_st[slot].addr, 2u * TM::bytes);
      tma::load3d(tma::s32(buf), &mp.Sig_f, 0, t, b0, bar_st[slot].addr, pol_ef);
      tma::load3d(tma::s32(buf + TM::bytes), &mp.Sig_p, 0, t + 1, b0, bar_st[slot].addr, pol_ef);
    };
    // ... and their lines are pulled into L2 KV_SEQ_PF_L2 iterations ahead (prefetch.global.L2, no registers): with the
    // machine full an HBM round trip is longer than two smoother iterations, and the first use of a two-ahead register
    // load was still 21 % of this kernel's stall samples (profiles/r02_ncu_hot_instructions_B262144_thread_per_sequence.txt)
    const bool pf_on = T >= 64;   // (short sequences: the prologue costs more than the prefetch saves; T = 20: +2 %, T = 200: -4 %)
    auto pf_l2 = [&](int t) {
#if KV_SEQ_PF_L2 > 0
      if (!pf_on) return;
      const long bt = (long)bl * T + t;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mu_f + bt * N));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.mu_p + (bt + 1) * N));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.alpha + (bt + 1) * K));
#else
      (void)t;
#endif
    };
#if KV_SEQ_PF_L2 > 0
    for (int t = T - 4; t > T - 4 - KV_SEQ_PF_L2 && t >= 0; --t) pf_l2(t);
#endif
    if (T >= 2) load_vec(T - 2, pfA);
    if (T >= 3) load_vec(T - 3, pfB);
    if (tma::elect_one()) {
      if (T >= 2) issue_st(T - 2, 0);
      if (T >= 3) issue_st(T - 3, 1);
    }
    int slot = 0;
    auto smooth_iter = [&](int t, VecPf& pf) {
      float muf[R], mup1[R], al1[K];
      KV_UNROLL for (int r = 0; r < R; ++r) { muf[r] = pf.muf[r]; mup1[r] = pf.mup1[r]; }
      KV_UNROLL for (int k = 0; k < K; ++k) al1[k] = pf.al[k];
      float A1[R][N];
      mix_A<C>(base, al1, 0, A1);
      if (a.a_smooth && active) {   // C_{t+1} mu_{t+1|T} (model.py:280-281): both are in hand at the top of step t
        float Ct1[R][P], as[P];
        mix_Ct<C>(base, al1, 0, Ct1);
        project_obs<C>(g, Ct1, mus, as);
        store_row<P>(a.a_smooth + ((long)b * T + t + 1) * P, as);
      }
      // requests for iteration t-2: the register set just consumed, and the ring slot iteration t+1 finished with
      // (every lane passed the __syncwarp of that iteration after its last read of the slot)
      if (t >= 2) {
        load_vec(t - 2, pf);
        if (tma::elect_one()) issue_st(t - 2, slot == 0 ? 2 : slot - 1);
      }
#if KV_SEQ_PF_L2 > 0
      if (t >= 2 + KV_SEQ_PF_L2) pf_l2(t - 2 - KV_SEQ_PF_L2);
#endif
      bar_st[slot].wait();
      const float* sb = reinterpret_cast<const float*>(ring + slot * (2 * (int)TM::bytes));
      float Sf[R][N], Sp1[R][N];
      TM::ld_mat(sb, lane, Sf);
      TM::ld_mat(sb + TM::floats, lane, Sp1);
      ok = smoother_step_math<C>(g, tl, Sf, Sp1, muf, mup1, A1, Sig, mus) && ok;
      float* tSs = (t & 1) ? tSs1 : tSs0;      // (T is even: step T-1 used tSs1, step T-2 uses tSs0, ...)
      if (KV_SEQ_OUT_BUFS == 2) tma::wait_read1(); else tma::wait_read0();
      __syncwarp();
      TM::st_mat(tSs, lane, Sig);
      if (active) store_row<N>(a.mu_s + ((long)b * T + t) * N, mus);
      tma::fence_async();
      __syncwarp();
      if (tma::elect_one()) {
        tma::store3d(&mp.Sig_s, 0, t, b0, tma::s32(tSs), pol_ef);
        tma::commit();
      }
      slot = (slot == 2) ? 0 : slot + 1;
    };
#pragma unroll 1
    for (int t = T - 2; t >= 0; t -= 2) {
      smooth_iter(t, pfA);
      if (t >= 1) smooth_iter(t - 1, pfB);
    }
    if (a.a_smooth && active) {   // t = 0
      float al0[K], Ct0[R][P], as[P];
      load_row<K>(a.alpha + (long)b * T * K, al0);
      mix_Ct<C>(base, al0, 0, Ct0);
      project_obs<C>(g, Ct0, mus, as);
      store_row<P>(a.a_smooth + (long)b * T * P, as);
    }
  }
  if (warp_on) tma::wait_all0();   // the staging tiles must outlive their stores
  if (!ok && active) kv_info_or(a.info, KV_INFO_PIVOT);
}


// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// state stream [B,T,W]: (W, T, B), box (W, 1, 32), swizzle of a 4W-byte row (see RowTile)
inline bool make_row_map(CUtensorMap* m, const float* p, int B, int T, int W) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc || !p) return false;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * 4 * (cuuint64_t)T};
  cuuint32_t box[3] = {(cuuint32_t)W, 1, 32};
  cuuint32_t es[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = (W == 8) ? CU_TENSOR_MAP_SWIZZLE_32B : (W == 16) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// input stream [B,T,W] moved four steps at a time: (T*W, B), box (4W, 32)
inline bool make_chunk_map(CUtensorMap* m, const float* p, int B, int T, int W) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc || !p) return false;
  cuuint64_t dims[2] = {(cuuint64_t)T * W, (cuuint64_t)B};
  cuuint64_t strides[1] = {(cuuint64_t)T * W * 4};
  cuuint32_t box[2] = {(cuuint32_t)(4 * W), 32};
  cuuint32_t es[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// can this call run on the thread-per-sequence kernels?  (explicit per-step matrices and smooth-only calls are the
// per-step forms filter_step / smooth_step: T = 1 or 2, lane-group kernels)
inline bool seq_eligible(const Args& a) {
  return (a.T % 4 == 0) && a.dA == nullptr && !a.smooth_only && a.alpha != nullptr;
}

template <class C> int launch_seq_fwd(const Args& a, const BasePtrs& bp, int smooth, cudaStream_t s) {
  (void)cudaGetLastError();
  SeqFwdMaps mp;
  memset(&mp, 0, sizeof(mp));
  const int B = a.B, T = a.T;
  bool okm = make_chunk_map(&mp.Y, a.Y, B, T, C::P) && make_chunk_map(&mp.alpha, a.alpha, B, T, C::K);
  if (a.U) okm = okm && make_chunk_map(&mp.U, a.U, B, T, C::M);
  if (a.mask) okm = okm && make_chunk_map(&mp.mask, a.mask, B, T, 1);
  okm = okm && make_row_map(&mp.Sig_p, a.Sig_p, B, T, C::N * C::N) && make_row_map(&mp.Sig_f, a.Sig_f, B, T, C::N * C::N);
  if (smooth) okm = okm && make_row_map(&mp.Sig_s, a.Sig_s, B, T, C::N * C::N);
  if (a.A_list) okm = okm && make_row_map(&mp.A, a.A_list, B, T, C::N * C::N);
  if (a.B_list) okm = okm && make_row_map(&mp.B, a.B_list, B, T, C::N * C::M);
  if (a.C_list) okm = okm && make_row_map(&mp.C, a.C_list, B, T, C::P * C::N);
  if (!okm) return -6;   // tensor-map encoding failed (driver entry point missing or misaligned pointer)
  const int warps = seq_warps_per_cta(B);
  size_t sm = seq_fwd_smem<C>(warps);
  if (sm < seq_smem_floor()) sm = seq_smem_floor();
  cudaError_t e = cudaFuncSetAttribute(k_seq_fwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  k_seq_fwd<C><<<seq_grid(B), 32 * warps, sm, s>>>(a, bp, mp, smooth);
  return (int)cudaGetLastError();
}

}  // namespace kvae
