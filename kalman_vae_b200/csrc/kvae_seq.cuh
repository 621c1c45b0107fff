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
//   * loads are double buffered (mbarrier complete_tx, the next step / chunk in flight while this one is computed);
//     stores leave from a single staging tile per warp (bulk_group; the next step waits for the tile to be READ, not
//     for the write to land).
//
// Reference arithmetic: kvae/kalman/kalman_filter.py:31-279 (sweeps 1-2 here), :305-401 + autograd (sweeps 3-4,
// csrc/kvae_seq_bwd.cuh).  The step arithmetic is the same code as the lane-group kernels (filter_step_math,
// smoother_step_math, ... instantiated with L = 1: "publish" aliases registers).
#pragma once
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
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
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
  CUtensorMap mu_p, mu_f, mu_s;                      // (4, T, B), box (4, 1, 32)
  CUtensorMap Sig_p, Sig_f, Sig_s, A, B;             // (16, T, B), box (16, 1, 32), SWIZZLE_64B
  CUtensorMap C;                                     // (4P, T, B), box (4P, 1, 32), swizzle of a 16P-byte row
};

__host__ __device__ constexpr int kv_align_up(int x, int a) { return (x + a - 1) / a * a; }

// Shared-memory plan of one warp of k_seq_fwd (bytes; every tile 1024-byte aligned).  Sweep 2 reuses sweep 1's region.
template <class C> struct SeqFwdPlan {
  static constexpr int P = C::P, K = C::K;
  using TV = RowTile<1>;    // mu rows
  using TM = RowTile<4>;    // Sigma / A / B rows
  using TC = RowTile<P>;    // C rows ([P][4])
  // sweep 1: output tiles of one step, then the two input-chunk buffers
  static constexpr int oSp = 0, oSf = oSp + TM::bytes, oA = oSf + TM::bytes, oB = oA + TM::bytes;
  static constexpr int oC = oB + TM::bytes;
  static constexpr int oMp = kv_align_up(oC + TC::bytes, 1024), oMf = oMp + 1024;
  static constexpr int in_Y = 0, in_U = in_Y + 32 * 4 * P * 4, in_al = in_U + 32 * 16 * 4, in_m = in_al + 32 * 4 * K * 4;
  static constexpr int in_bytes = kv_align_up(in_m + 32 * 4 * 4, 1024);
  static __host__ __device__ constexpr uint32_t in_tx(bool has_u, bool has_m) {
    return 32u * 4 * P * 4 + (has_u ? 32u * 16 * 4 : 0u) + 32u * 4 * K * 4 + (has_m ? 32u * 4 * 4 : 0u);
  }
  static constexpr int oIn0 = oMf + 1024, oIn1 = oIn0 + in_bytes;
  static constexpr int sweep1 = oIn1 + in_bytes;
  // sweep 2: two state buffers [Sigma_f(t) | Sigma_p(t+1) | mu_f(t) | mu_p(t+1)], the output tiles, two alpha chunks
  static constexpr int st_Sf = 0, st_Sp = TM::bytes, st_mf = 2 * TM::bytes, st_mp = 2 * TM::bytes + 512;
  static constexpr int st_bytes = 2 * TM::bytes + 1024;
  static constexpr uint32_t st_tx = 2u * TM::bytes + 2u * TV::bytes;
  static constexpr int oSt0 = 0, oSt1 = st_bytes, oSs = 2 * st_bytes, oMs = oSs + TM::bytes;
  static constexpr int al_bytes = kv_align_up(32 * 4 * K * 4, 128);
  static constexpr int oAl0 = oMs + 1024, oAl1 = oAl0 + al_bytes;
  static constexpr int sweep2 = kv_align_up(oAl1 + al_bytes, 1024);
  static constexpr int warp_bytes = sweep1 > sweep2 ? sweep1 : sweep2;
};

template <class C> constexpr size_t seq_fwd_smem(int warps) {
  return 1024 + (size_t)kv_align_up((int)sizeof(float) * Base<C>::total, 1024) + (size_t)warps * SeqFwdPlan<C>::warp_bytes;
}
// warps per CTA: small batches are spread over as many SMs as possible with two warps per CTA (distinct schedulers)
inline int seq_warps_per_cta(int B) { return (B <= 148 * 64) ? 2 : 4; }
inline int seq_grid(int B) { const int per = 32 * seq_warps_per_cta(B); return (B + per - 1) / per; }

#ifndef KV_SEQ_MAXWARPS
#define KV_SEQ_MAXWARPS 4
#endif

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
__global__ void __launch_bounds__(32 * KV_SEQ_MAXWARPS) k_seq_fwd(Args a, BasePtrs bp, const __grid_constant__ SeqFwdMaps mp,
                                                                  int smooth) {
  static_assert(C::L == 1 && C::N == 4 && C::M == 4, "thread-per-sequence kernels: z_dim = u_dim = 4");
  constexpr int N = C::N, P = C::P, M = C::M, K = C::K, R = C::R;
  using PL = SeqFwdPlan<C>;
  using TV = typename PL::TV;
  using TM = typename PL::TM;
  using TC = typename PL::TC;
  extern __shared__ unsigned char seq_smem_raw[];
  __shared__ __align__(8) unsigned long long bars[KV_SEQ_MAXWARPS][6];
  __shared__ float mred[KV_SEQ_MAXWARPS];
  unsigned char* sm = seq_smem_raw + ((1024u - (tma::s32(seq_smem_raw) & 1023u)) & 1023u);
  float* base = reinterpret_cast<float*>(sm);
  stage_base<C>(base, bp);   // __syncthreads inside
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* wsm = sm + kv_align_up((int)sizeof(float) * Base<C>::total, 1024) + warp * PL::warp_bytes;
  const int b0 = blockIdx.x * blockDim.x + warp * 32;   // first sequence of this warp
  const int b = b0 + lane;
  const bool active = b < a.B;
  const bool warp_on = b0 < a.B;
  const int T = a.T, nchunk = T >> 2;
  const bool has_u = a.U != nullptr, has_m = a.mask != nullptr;
  const Group<1, R> g{0, 0xffffffffu};
  const FTiles<C> tl{base, 0};   // L = 1: never dereferenced (views alias registers)
  const uint32_t bar_in0 = tma::s32(&bars[warp][0]), bar_in1 = tma::s32(&bars[warp][1]);
  const uint32_t bar_st0 = tma::s32(&bars[warp][2]), bar_st1 = tma::s32(&bars[warp][3]);
  const uint32_t bar_al0 = tma::s32(&bars[warp][4]), bar_al1 = tma::s32(&bars[warp][5]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) tma::bar_init(tma::s32(&bars[warp][i]), 1);
    tma::bar_init_fence();
  }
  __syncwarp();

  float Sig[R][N], mu[N], mus[R];
  float msum = 0.f;
  bool ok = true;
  if (warp_on) {
    // ------------------------------------------------------------------ sweep 1: filter
    copy_rows<C, N>(base + Base<C>::oS0, 0, Sig);
    load_row<N>(base + Base<C>::oMu0, mu);
    if (a.Sig_init && active) { KV_UNROLL for (int r = 0; r < R; ++r) load_row<N>(a.Sig_init + ((long)b * N + r) * N, Sig[r]); }
    if (a.mu_init && active) load_row<N>(a.mu_init + (long)b * N, mu);
    auto issue_in = [&](int c) {   // lane 0: the four input streams of chunk c
      unsigned char* buf = wsm + ((c & 1) ? PL::oIn1 : PL::oIn0);
      const uint32_t bar = (c & 1) ? bar_in1 : bar_in0;
      tma::bar_expect(bar, PL::in_tx(has_u, has_m));
      tma::load2d(tma::s32(buf + PL::in_Y), &mp.Y, 4 * c * P, b0, bar);
      if (has_u) tma::load2d(tma::s32(buf + PL::in_U), &mp.U, 4 * c * M, b0, bar);
      tma::load2d(tma::s32(buf + PL::in_al), &mp.alpha, 4 * c * K, b0, bar);
      if (has_m) tma::load2d(tma::s32(buf + PL::in_m), &mp.mask, 4 * c, b0, bar);
    };
    if (tma::elect_one()) issue_in(0);
    float* tSp = reinterpret_cast<float*>(wsm + PL::oSp);
    float* tSf = reinterpret_cast<float*>(wsm + PL::oSf);
    float* tA = reinterpret_cast<float*>(wsm + PL::oA);
    float* tB = reinterpret_cast<float*>(wsm + PL::oB);
    float* tC = reinterpret_cast<float*>(wsm + PL::oC);
    float* tMp = reinterpret_cast<float*>(wsm + PL::oMp);
    float* tMf = reinterpret_cast<float*>(wsm + PL::oMf);
    for (int c = 0; c < nchunk; ++c) {
      if (c + 1 < nchunk && tma::elect_one()) issue_in(c + 1);   // its buffer was released by the __syncwarp ending chunk c-1
      tma::bar_wait((c & 1) ? bar_in1 : bar_in0, (uint32_t)((c >> 1) & 1));
      const unsigned char* ib = wsm + ((c & 1) ? PL::oIn1 : PL::oIn0);
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const int t = 4 * c + s;
        StepIn<C> cur;
        seq_read_step<C>(ib, lane, s, has_u, has_m, cur);
        float A[R][N], Bm[R][M], Ct[R][P], Q[R][N];
        mix_A<C>(base, cur.al, 0, A);
        mix_B<C>(base, cur.al, 0, Bm);
        mix_Ct<C>(base, cur.al, 0, Ct);
        mix_Q<C>(base, cur.al, 0, Q);
        // the staging tiles are free once the previous step's stores have READ them
        tma::wait_read0();   // all lanes: only the electing lane has groups pending
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
        TV::st(tMp, lane, 0, fo.mup);
        TV::st(tMf, lane, 0, fo.muf);
        tma::fence_async();
        __syncwarp();
        if (tma::elect_one()) {
          tma::store3d(&mp.Sig_p, 0, t, b0, tma::s32(tSp));
          tma::store3d(&mp.Sig_f, 0, t, b0, tma::s32(tSf));
          tma::store3d(&mp.mu_p, 0, t, b0, tma::s32(tMp));
          tma::store3d(&mp.mu_f, 0, t, b0, tma::s32(tMf));
          if (a.A_list) tma::store3d(&mp.A, 0, t, b0, tma::s32(tA));
          if (a.B_list) tma::store3d(&mp.B, 0, t, b0, tma::s32(tB));
          if (a.C_list) tma::store3d(&mp.C, 0, t, b0, tma::s32(tC));
          tma::commit();
        }
        KV_UNROLL for (int r = 0; r < R; ++r) {
          KV_UNROLL for (int j = 0; j < N; ++j) Sig[r][j] = fo.Sf[r][j];
          mu[r] = fo.muf[r];
        }
      }
      __syncwarp();   // every lane is done with input buffer c & 1
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
    // sweep 1's stores must have landed (the first step re-reads Sigma_f(T-2), Sigma_p(T-1)) and left the tiles
    tma::wait_all0();
    __syncwarp();
    float* tSs = reinterpret_cast<float*>(wsm + PL::oSs);
    float* tMs = reinterpret_cast<float*>(wsm + PL::oMs);
    auto issue_st = [&](int i) {   // lane 0: states of iteration i (t = T-2-i)
      const int t = T - 2 - i;
      unsigned char* buf = wsm + ((i & 1) ? PL::oSt1 : PL::oSt0);
      const uint32_t bar = (i & 1) ? bar_st1 : bar_st0;
      tma::bar_expect(bar, PL::st_tx);
      tma::load3d(tma::s32(buf + PL::st_Sf), &mp.Sig_f, 0, t, b0, bar);
      tma::load3d(tma::s32(buf + PL::st_Sp), &mp.Sig_p, 0, t + 1, b0, bar);
      tma::load3d(tma::s32(buf + PL::st_mf), &mp.mu_f, 0, t, b0, bar);
      tma::load3d(tma::s32(buf + PL::st_mp), &mp.mu_p, 0, t + 1, b0, bar);
    };
    auto issue_al = [&](int j) {   // lane 0: alpha chunk nchunk-1-j
      const int c = nchunk - 1 - j;
      const uint32_t bar = (j & 1) ? bar_al1 : bar_al0;
      tma::bar_expect(bar, 32u * 4 * K * 4);
      tma::load2d(tma::s32(wsm + ((j & 1) ? PL::oAl1 : PL::oAl0)), &mp.alpha, 4 * c * K, b0, bar);
    };
    if (tma::elect_one()) {
      issue_al(0);
      if (T >= 2) issue_st(0);
    }
    // t = T-1: copied, not symmetrised (kalman_filter.py:251-256)
    TM::st_mat(tSs, lane, Sig);
    TV::st(tMs, lane, 0, mus);
    tma::fence_async();
    __syncwarp();
    if (tma::elect_one()) {
      tma::store3d(&mp.Sig_s, 0, T - 1, b0, tma::s32(tSs));
      tma::store3d(&mp.mu_s, 0, T - 1, b0, tma::s32(tMs));
      tma::commit();
    }
    int jal = 0;   // alpha chunk counter (0 = last chunk)
    tma::bar_wait(bar_al0, 0u);
    for (int i = 0; i + 2 <= T; ++i) {
      const int t = T - 2 - i;
      if (t > 0 && tma::elect_one()) issue_st(i + 1);          // buffer released by the __syncwarp ending iteration i-1
      // alpha_{t+1}
      const int c1 = (t + 1) >> 2;
      if (nchunk - 1 - c1 != jal) {                     // (t+1) moved into the previous chunk
        ++jal;
        tma::bar_wait((jal & 1) ? bar_al1 : bar_al0, (uint32_t)((jal >> 1) & 1));
      }
      if (((t + 1) & 3) == 3 && c1 > 0 && tma::elect_one()) issue_al(jal + 1);   // first use of chunk c1: fetch chunk c1-1
      float al1[K];
      {
        const float* al = reinterpret_cast<const float*>(wsm + ((jal & 1) ? PL::oAl1 : PL::oAl0)) + lane * 4 * K + ((t + 1) & 3) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) al1[k] = al[k];
      }
      float A1[R][N];
      mix_A<C>(base, al1, 0, A1);
      tma::bar_wait((i & 1) ? bar_st1 : bar_st0, (uint32_t)((i >> 1) & 1));
      const unsigned char* sb = wsm + ((i & 1) ? PL::oSt1 : PL::oSt0);
      float Sf[R][N], Sp1[R][N], muf[R], mup1[R];
      TM::ld_mat(reinterpret_cast<const float*>(sb + PL::st_Sf), lane, Sf);
      TM::ld_mat(reinterpret_cast<const float*>(sb + PL::st_Sp), lane, Sp1);
      TV::ld(reinterpret_cast<const float*>(sb + PL::st_mf), lane, 0, muf);
      TV::ld(reinterpret_cast<const float*>(sb + PL::st_mp), lane, 0, mup1);
      ok = smoother_step_math<C>(g, tl, Sf, Sp1, muf, mup1, A1, Sig, mus) && ok;
      tma::wait_read0();   // all lanes: only the electing lane has groups pending
      __syncwarp();
      TM::st_mat(tSs, lane, Sig);
      TV::st(tMs, lane, 0, mus);
      tma::fence_async();
      __syncwarp();   // also: every lane is done with state buffer i & 1 and (at a chunk change) the old alpha chunk
      if (tma::elect_one()) {
        tma::store3d(&mp.Sig_s, 0, t, b0, tma::s32(tSs));
        tma::store3d(&mp.mu_s, 0, t, b0, tma::s32(tMs));
        tma::commit();
      }
    }
  }
  if (warp_on) tma::wait_all0();   // the staging tiles must outlive their stores
  if (!ok && active) *a.info = 1;
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
  okm = okm && make_row_map(&mp.mu_p, a.mu_p, B, T, C::N) && make_row_map(&mp.mu_f, a.mu_f, B, T, C::N) &&
        make_row_map(&mp.Sig_p, a.Sig_p, B, T, C::N * C::N) && make_row_map(&mp.Sig_f, a.Sig_f, B, T, C::N * C::N);
  if (smooth) okm = okm && make_row_map(&mp.mu_s, a.mu_s, B, T, C::N) && make_row_map(&mp.Sig_s, a.Sig_s, B, T, C::N * C::N);
  if (a.A_list) okm = okm && make_row_map(&mp.A, a.A_list, B, T, C::N * C::N);
  if (a.B_list) okm = okm && make_row_map(&mp.B, a.B_list, B, T, C::N * C::M);
  if (a.C_list) okm = okm && make_row_map(&mp.C, a.C_list, B, T, C::P * C::N);
  if (!okm) return -6;   // tensor-map encoding failed (driver entry point missing or misaligned pointer)
  const int warps = seq_warps_per_cta(B);
  const size_t sm = seq_fwd_smem<C>(warps);
  cudaError_t e = cudaFuncSetAttribute(k_seq_fwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_fwd_smem<C>(KV_SEQ_MAXWARPS));
  if (e != cudaSuccess) return (int)e;
  k_seq_fwd<C><<<seq_grid(B), 32 * warps, sm, s>>>(a, bp, mp, smooth);
  return (int)cudaGetLastError();
}

}  // namespace kvae
