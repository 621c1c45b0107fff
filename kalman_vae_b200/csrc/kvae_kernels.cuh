// kvae_kernels.cuh — __global__ entry points: one CTA of 128 threads = 128/L lane groups, each
// group owning one sequence for the whole sweep; base matrices staged once per CTA in shared
// memory; per-group publish tiles behind them.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "kvae_bwd.cuh"
#include "kvae_lstm.cuh"

namespace kvae {

// threads per CTA: 4 warps (the base matrices are staged once per CTA: for n = 16 they are 37 KB, so 4 warps per CTA
// measured 19 % faster in the forward kernel than 2)
#ifndef KV_TPB_SMALL
#define KV_TPB_SMALL 128
#endif
#ifndef KV_TPB_LARGE
#define KV_TPB_LARGE 128
#endif
// (n = 16 with 8 lanes per sequence has 4 groups per warp: its tiles only fit with 2 warps per CTA)
template <class C> constexpr int TPB = (C::N >= 16) ? (C::L >= 16 ? KV_TPB_LARGE : 64) : KV_TPB_SMALL;
// backward kernel: for n = 16, L = 16 eight warps per CTA (7 [n x n] tiles per warp + the 33 KB of staged base matrices
// = 225 KB, 255 registers x 256 threads = the whole register file): ONE CTA per SM either way, so twice the warps
#ifndef KV_TPB_BWD_LARGE
#define KV_TPB_BWD_LARGE 256
#endif
template <class C> constexpr int TPBB = (C::N >= 16 && C::L >= 16) ? KV_TPB_BWD_LARGE : TPB<C>;

struct BasePtrs { const float *A, *Bm, *C, *Q, *R, *mu0, *S0; };

template <class C> __device__ __forceinline__ float* stage_base(float* smem, const BasePtrs& bp) {
  for (int i = threadIdx.x; i < Base<C>::total; i += blockDim.x)
    base_fill<C>(smem, i, bp.A, bp.Bm, bp.C, bp.Q, bp.R, bp.mu0, bp.S0);
  __syncthreads();
  return smem + Base<C>::total;
}

template <class C> __device__ __forceinline__ Group<C::L, C::R> this_group() {
  // every kernel keeps the control flow of all groups of a warp identical (CTA-uniform time chunks, tail
  // groups mirror a valid sequence), so the group collectives can name the whole warp: sub-warp masks are
  // legal too but let the hardware split the warp, which costs ~15-40 % here (measured)
  return Group<C::L, C::R>{(int)(threadIdx.x % C::L), 0xffffffffu};
}
template <class C> constexpr size_t smem_bytes() {
  return sizeof(float) * (size_t)(Base<C>::total + (TPB<C> / 32) * FTiles<C>::warp_total +
                                  (TPB<C> / C::L) * InStage<C, true>::group_floats);
}
// this thread's tile set: per-warp region + group index inside the warp
template <class TS> __device__ __forceinline__ TS warp_tiles(float* tiles_all, int L) {
  return TS{tiles_all + (threadIdx.x >> 5) * TS::warp_total, (int)((threadIdx.x & 31) / L)};
}

template <class C>
__global__ void __launch_bounds__(TPB<C>) k_filter_smooth(Args a, BasePtrs bp, int smooth) {
  extern __shared__ f4 smem_raw[];
  float* base = reinterpret_cast<float*>(smem_raw);
  float* tiles_all = stage_base<C>(base, bp);
  constexpr int GPB = TPB<C> / C::L;
  const int gi = threadIdx.x / C::L;
  const Group<C::L, C::R> g = this_group<C>();
  int b = blockIdx.x * GPB + gi;
  const bool active = b < a.B;
  if (!active) b = a.B - 1;  // tail groups recompute the last sequence and store nothing
  const FTiles<C> tl = warp_tiles<FTiles<C>>(tiles_all, C::L);
  float* stage_slot = tiles_all + (TPB<C> / 32) * FTiles<C>::warp_total + gi * InStage<C, true>::group_floats;
  float Sig[C::R][C::N], mu[C::N], mu_own[C::R];
  if (a.smooth_only) {   // filtered states given: the belief at T-1 is the smoother's starting point
    const long btl = (long)b * a.T + (a.T - 1);
    for (int r = 0; r < C::R; ++r) load_row<C::N>(a.Sig_f + (btl * C::N + g.row0() + r) * C::N, Sig[r]);
    load_row<C::R>(a.mu_f + btl * C::N + g.row0(), mu_own);
  } else {
    float msum = 0.f;
    filter_sweep<C>(a, base, tl, g, b, active, stage_slot, Sig, mu, mu_own, &msum);
    if (a.mask_part) {   // per-CTA sum of the mask (fixed order): lets the adjoint apply 1/max(sum mask,1) inside its sweep
      __shared__ float mred[TPB<C> / 32];
      float v = (active && g.lane == 0) ? msum : 0.f;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31) == 0) mred[threadIdx.x >> 5] = v;
      __syncthreads();
      if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int wq = 0; wq < TPB<C> / 32; ++wq) tot += mred[wq];
        a.mask_part[blockIdx.x] = tot;
      }
    }
  }
  if (smooth) smoother_sweep<C>(a, base, tl, g, b, active, stage_slot, Sig, mu_own);
}

// ------------------------------------------------------------------------------------------------
// filter sweep with the LSTM dynamics network in the loop (csrc/kvae_lstm.cuh): lstm variant only
// ------------------------------------------------------------------------------------------------
template <class C> constexpr size_t smem_bytes_lstm() {
  return smem_bytes<C>() + sizeof(float) * (size_t)(LstmGeo<C>::total + (TPB<C> / C::L) * LstmGeo<C>::HP);
}
template <class C>
__global__ void __launch_bounds__(TPB<C>) k_filter_lstm(Args a, BasePtrs bp, LstmPtrs lw) {
  extern __shared__ f4 smem_raw[];
  float* base = reinterpret_cast<float*>(smem_raw);
  float* tiles_all = stage_base<C>(base, bp);
  constexpr int GPB = TPB<C> / C::L;
  const int gi = threadIdx.x / C::L;
  const Group<C::L, C::R> g = this_group<C>();
  int b = blockIdx.x * GPB + gi;
  const bool active = b < a.B;
  if (!active) b = a.B - 1;
  const FTiles<C> tl = warp_tiles<FTiles<C>>(tiles_all, C::L);
  float* stage_slot = tiles_all + (TPB<C> / 32) * FTiles<C>::warp_total + gi * InStage<C, true>::group_floats;
  float* lstm_w = tiles_all + (TPB<C> / 32) * FTiles<C>::warp_total + GPB * InStage<C, true>::group_floats;
  for (int i = threadIdx.x; i < LstmGeo<C>::total; i += TPB<C>) lstm_w[i] = lstm_weight_at<C>(lw, i);
  __syncthreads();
  LstmHook<C> hook;
  hook.W = lstm_w;
  hook.hbuf = lstm_w + LstmGeo<C>::total + gi * LstmGeo<C>::HP;
  hook.alpha_out = lw.alpha_out;
  hook.on = active;
  hook.init(g, lw, b);
  float Sig[C::R][C::N], mu[C::N], mu_own[C::R];
  filter_sweep<C, LstmHook<C>>(a, base, tl, g, b, active, stage_slot, Sig, mu, mu_own, nullptr, &hook);
  hook.finish(g, lw, b);
}
template <class C> int launch_fwd_lstm(const Args& a, const BasePtrs& bp, const LstmPtrs& lw, cudaStream_t s) {
  (void)cudaGetLastError();
  constexpr int GPB = TPB<C> / C::L;
  const size_t sm = smem_bytes_lstm<C>();
  if (sm > 48 * 1024) {   // per DEVICE attribute: set on every launch (cheap, idempotent) so that a second GPU of the process works too
    cudaError_t e = cudaFuncSetAttribute(k_filter_lstm<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (a.B + GPB - 1) / GPB;
  constexpr int tpb = TPB<C>;
  k_filter_lstm<C><<<grid, tpb, sm, s>>>(a, bp, lw);
  return (int)cudaGetLastError();
}

template <class C> int fwd_grid_of(int B) { return (B + TPB<C> / C::L - 1) / (TPB<C> / C::L); }

template <class C> int launch_fwd(const Args& a, const BasePtrs& bp, int smooth, cudaStream_t s) {
  (void)cudaGetLastError();  // do not inherit a stale (non-sticky) error from an earlier call
  constexpr int GPB = TPB<C> / C::L;
  const size_t sm = smem_bytes<C>();
  if (sm > 48 * 1024) {   // per DEVICE attribute: set on every launch (cheap, idempotent) so that a second GPU of the process works too
    cudaError_t e = cudaFuncSetAttribute(k_filter_smooth<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
  }
  const int grid = (a.B + GPB - 1) / GPB;
  constexpr int tpb = TPB<C>;
  k_filter_smooth<C><<<grid, tpb, sm, s>>>(a, bp, smooth);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && getenv("KVAE_DEBUG")) fprintf(stderr, "[kvae] k_filter_smooth launch failed: smem=%zu grid=%d L=%d: %s\n", sm, grid, C::L, cudaGetErrorString(e));
  return (int)e;
}


// ------------------------------------------------------------------------------------------------
// ELBO: per-CTA partial sums (double) -> one-block final reduction
// ------------------------------------------------------------------------------------------------
template <class C> constexpr size_t smem_bytes_elbo() { return smem_bytes<C>(); }

// Time-parallel kernels cut every sequence into `chunks` pieces of TC = ceil(T/chunks) steps so that small
// batches still give each SM sub-partition several warps (>= ~6); long chunks keep the per-chunk overhead
// (one or two extra Cholesky factorisations at the chunk start) small.
inline int pick_chunks(int B, int T, int L) {
  const long want_groups = 148L * 4 * 6 * (32 / L);
  long ch = (want_groups + B - 1) / B;
  const long max_ch = T >= 8 ? T / 4 : 1;
  if (ch > max_ch) ch = max_ch;
  if (ch < 1) ch = 1;
  return (int)ch;
}
struct ChunkMap { int chunks, tc; };
inline ChunkMap make_chunks(int B, int T, int L) {
  ChunkMap cm;
  cm.chunks = pick_chunks(B, T, L);
  cm.tc = (T + cm.chunks - 1) / cm.chunks;
  cm.chunks = (T + cm.tc - 1) / cm.tc;   // no empty chunk
  return cm;
}
// CTA -> (chunk, block of sequences): every group of a CTA works on the SAME time chunk of different
// sequences, so control flow stays warp-uniform; inactive tail groups mirror the last sequence, store nothing
template <class C>
__device__ __forceinline__ bool chunk_of(const Args& a, ChunkMap cm, int& b, int& t0, int& t1) {
  constexpr int GPB = TPB<C> / C::L;
  const int c = blockIdx.x % cm.chunks;
  b = (blockIdx.x / cm.chunks) * GPB + threadIdx.x / C::L;
  const bool active = b < a.B;
  if (!active) b = a.B - 1;
  t0 = c * cm.tc;
  t1 = t0 + cm.tc < a.T ? t0 + cm.tc : a.T;
  return active && t0 < a.T;
}
template <class C>
__global__ void __launch_bounds__(TPB<C>) k_elbo(Args a, BasePtrs bp, float jitter, ChunkMap cm, double* __restrict__ partials) {
  extern __shared__ f4 smem_raw[];
  float* base = reinterpret_cast<float*>(smem_raw);
  float* tiles_all = stage_base<C>(base, bp);
  const Group<C::L, C::R> g = this_group<C>();
  int b, t0, t1;
  const bool active = chunk_of<C>(a, cm, b, t0, t1);
  const FTiles<C> tl = warp_tiles<FTiles<C>>(tiles_all, C::L);
  double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  elbo_sweep<C>(a, base, tl, g, b, active, jitter, t0, t1, nullptr, acc);
  // block reduction (fixed order -> deterministic)
  __shared__ double red[TPB<C> / 32][5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    double v = acc[i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
#pragma unroll
    for (int wq = 0; wq < TPB<C> / 32; ++wq) v += red[wq][threadIdx.x];
    partials[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
  }
}

// terms: [0] trans [1] emiss [2] init [3] entropy [4] sum(mask) [5] elbo [6] 1/max(sum mask,1) [7] 0
// one warp per sum, lanes stride over the per-CTA partials, fixed shuffle tree -> deterministic
static __global__ void k_elbo_final(const double* __restrict__ partials, int nblocks, float* __restrict__ terms) {
  __shared__ double tot[5];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < 5) {
    double v = 0.0;
    for (int i = lane; i < nblocks; i += 32) v += partials[(size_t)i * 5 + w];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) tot[w] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double n = tot[4] < 1.0 ? 1.0 : tot[4];
    for (int j = 0; j < 5; ++j) terms[j] = (float)tot[j];
    terms[5] = (float)((tot[0] + tot[1] + tot[2] + tot[3]) / n);
    terms[6] = (float)(1.0 / n);
    terms[7] = 0.f;
  }
}

template <class C> int chunk_grid(int B, int chunks) {
  constexpr int GPB = TPB<C> / C::L;
  return ((B + GPB - 1) / GPB) * chunks;
}
template <class C> size_t elbo_ws_bytes(int B, int T) {
  return sizeof(double) * 5 * (size_t)chunk_grid<C>(B, make_chunks(B, T, C::L).chunks);
}

template <class C> int launch_elbo(const Args& a, const BasePtrs& bp, float jitter, float* terms, void* ws, cudaStream_t s) {
  constexpr int GPB = TPB<C> / C::L;
  const size_t sm = smem_bytes_elbo<C>();
  if (sm > 48 * 1024) {   // per DEVICE attribute: set on every launch (cheap, idempotent) so that a second GPU of the process works too
    cudaError_t e = cudaFuncSetAttribute(k_elbo<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
  }
  const ChunkMap cm = make_chunks(a.B, a.T, C::L);
  const int grid = chunk_grid<C>(a.B, cm.chunks);
  constexpr int tpb = TPB<C>;
  k_elbo<C><<<grid, tpb, sm, s>>>(a, bp, jitter, cm, reinterpret_cast<double*>(ws));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  k_elbo_final<<<1, 160, 0, s>>>(reinterpret_cast<const double*>(ws), grid, terms);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// backward: sweeps 3 + 4 per group, parameter gradients reduced per CTA (fixed order), then a
// second kernel sums the per-CTA partials.
// ------------------------------------------------------------------------------------------------
struct GradPtrs { float *dA, *dB, *dC, *dQ; };
struct DensePtrs { float *A, *B, *Q, *Ct; };

template <class C> constexpr size_t smem_floats_bwd() {
  constexpr size_t tiles = (size_t)Base<C>::total + (size_t)(TPBB<C> / 32) * BTiles<C>::warp_total;
  constexpr size_t red = (size_t)Base<C>::total + (size_t)GradAcc<C>::PSZ;
  return tiles > red ? tiles : red;
}

// CTA reduction of the per-lane parameter-gradient accumulators in a fixed order (deterministic):
// groups of a warp by xor-shuffles, warps through shared memory, one partial row per CTA.
template <class C>
__device__ __forceinline__ void cta_reduce_acc(GradAcc<C>& acc, const Group<C::L, C::R>& g, bool active, float* red,
                                               float* __restrict__ partial_row) {
  if (!active) acc.zero();
#pragma unroll
  for (int off = C::L; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < GradAcc<C>::nreduce; ++i) acc.v[i] += __shfl_xor_sync(0xffffffffu, acc.v[i], off);
  }
  __syncthreads();   // tiles are dead: reuse them
  for (int i = threadIdx.x; i < GradAcc<C>::PSZ; i += TPBB<C>) red[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int wq = 0; wq < TPBB<C> / 32; ++wq) {
    if (warp == wq && lane < C::L) acc.for_each(g.row0(), [&](int idx, float v) { red[idx] += v; });
    __syncthreads();
  }
  for (int i = threadIdx.x; i < GradAcc<C>::PSZ; i += TPBB<C>) partial_row[i] = red[i];
}

// DENSE gradient mode: out[k][e] = sum_{bt in chunk} alpha[bt][k] * X[bt][e]  (one thread per column e, K register
// accumulators, X streamed once with coalesced loads); partial rows are reduced by k_bwd_final.
template <int K>
static __global__ void __launch_bounds__(256) k_mode_contract(const float* __restrict__ alpha, const float* __restrict__ X, long BT,
                                                              int E, int chunk, float* __restrict__ rows, int psz, int foff,
                                                              int transposeP, int Ncols) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= E) return;
  const long bt0 = (long)blockIdx.y * chunk;
  const long bt1 = bt0 + chunk < BT ? bt0 + chunk : BT;
  float acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.f;
  for (long bt = bt0; bt < bt1; ++bt) {
    const float x = X[bt * E + e];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = fmaf(alpha[bt * K + k], x, acc[k]);
  }
  // flat index: row-major [k][e], or for the transposed C^T scratch ([N][P] per step -> dC [k][P][N]) swap (i,a)
  int fe = e;
  if (transposeP) { const int i = e / transposeP, a_ = e % transposeP; fe = a_ * Ncols + i; }
#pragma unroll
  for (int k = 0; k < K; ++k) rows[(size_t)blockIdx.y * psz + foff + k * E + fe] = acc[k];
}

template <class C>
__global__ void __launch_bounds__(TPBB<C>) k_bwd(Args a, BwdArgs w, BasePtrs bp, const float* __restrict__ g_elbo,
                                                  const float* __restrict__ terms, float* __restrict__ partials, DensePtrs dn,
                                                  double* __restrict__ elbo_partials) {
  extern __shared__ f4 smem_raw[];
  float* base = reinterpret_cast<float*>(smem_raw);
  float* tiles_all = stage_base<C>(base, bp);
  constexpr int GPB = TPBB<C> / C::L;
  const int gi = threadIdx.x / C::L;
  const Group<C::L, C::R> g = this_group<C>();
  int b = blockIdx.x * GPB + gi;
  const bool active = b < a.B;
  if (!active) b = a.B - 1;
  const BTiles<C> tl = warp_tiles<BTiles<C>>(tiles_all, C::L);
  float inv_norm = 1.0f;
  if (w.mask_part) {   // normaliser from the forward kernel's per-CTA mask sums: same fixed-order fp64 sum in every CTA
    __shared__ double nred[TPBB<C> / 32];
    double v = 0.0;
    for (int i = threadIdx.x; i < w.n_mask_part; i += TPBB<C>) v += (double)w.mask_part[i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) nred[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int wq = 0; wq < TPBB<C> / 32; ++wq) tot += nred[wq];
    inv_norm = (float)(1.0 / (tot < 1.0 ? 1.0 : tot));
  }
  w.c_elbo = g_elbo ? (*g_elbo) * (w.with_elbo ? inv_norm : terms[6]) : 0.f;
  GradAcc<C> acc;
  acc.zero();
  acc.on = active;
  acc.dnA = dn.A; acc.dnB = dn.B; acc.dnQ = dn.Q; acc.dnCt = dn.Ct;
  double el[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  bwd_sweep3<C>(a, w, base, tl, g, b, active, acc, el);
  if (!w.elbo_only) bwd_sweep4<C>(a, w, base, tl, g, b, active, acc);
  if (w.with_elbo) {   // per-CTA partial sums of the ELBO value terms (fixed order -> deterministic)
    __shared__ double red[TPBB<C> / 32][5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      double v = el[i];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
      double v = 0.0;
#pragma unroll
      for (int wq = 0; wq < TPBB<C> / 32; ++wq) v += red[wq][threadIdx.x];
      elbo_partials[(size_t)blockIdx.x * 5 + threadIdx.x] = v;
    }
  }
  cta_reduce_acc<C>(acc, g, active, tiles_all, partials + (size_t)blockIdx.x * GradAcc<C>::PSZ);
}

// Final kernel of the backward pass.
//  * blocks [0, nparam_blocks): sum the per-CTA parameter-gradient partials and scatter into dA | dB | dC | dQ: one warp
//    per parameter element, lanes stride over the CTAs, fp64 accumulation, fixed shuffle tree -> deterministic.
//  * fused-ELBO mode (elbo_partials != nullptr): every block first re-derives the normaliser 1/max(sum mask,1) from
//    the per-CTA ELBO partials (same fixed order in every block); block 0 also writes terms[8].  Unless `raw`, the
//    parameter gradients are multiplied by the normaliser and the remaining blocks scale dY / dalpha / dU in place
//    (the adjoint ran with c = g_elbo because the mask sum was not known yet; it is linear in c).
struct ScaleJob { float* p[3]; long n[3]; };
static __global__ void __launch_bounds__(128) k_bwd_final(const float* __restrict__ partials, int nblocks, int psz, int nA, int nB,
                                                          int nC, GradPtrs gp, const double* __restrict__ elbo_partials,
                                                          int n_elbo_partials, float* __restrict__ terms, int raw,
                                                          int nparam_blocks, ScaleJob sj) {
  __shared__ double tot[5];
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float scale = 1.f;
  if (elbo_partials) {
    const bool all5 = (blockIdx.x == 0);
    for (int q = wp; q < 5; q += 4) {
      if (q == 4 || all5) {
        double v = 0.0;
        for (int i = lane; i < n_elbo_partials; i += 32) v += elbo_partials[(size_t)i * 5 + q];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) tot[q] = v;
      }
    }
    __syncthreads();
    const double nrm = tot[4] < 1.0 ? 1.0 : tot[4];
    if (all5 && threadIdx.x == 0) {
      for (int j = 0; j < 5; ++j) terms[j] = (float)tot[j];
      terms[5] = (float)((tot[0] + tot[1] + tot[2] + tot[3]) / nrm);
      terms[6] = (float)(1.0 / nrm);
      terms[7] = 0.f;
    }
    if (!raw) scale = (float)(1.0 / nrm);
  }
  if ((int)blockIdx.x >= nparam_blocks) {   // in-place scaling of the per-step gradients
    if (scale == 1.f) return;
    const long nthreads = (long)(gridDim.x - nparam_blocks) * blockDim.x;
    const long tid = (long)(blockIdx.x - nparam_blocks) * blockDim.x + threadIdx.x;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float* p = sj.p[j];
      if (!p) continue;
      const long n4 = sj.n[j] >> 2;
      f4* p4 = reinterpret_cast<f4*>(p);
      for (long i = tid; i < n4; i += nthreads) { f4 v = p4[i]; v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; p4[i] = v; }
      for (long i = (n4 << 2) + tid; i < sj.n[j]; i += nthreads) p[i] *= scale;
    }
    return;
  }
  const int i = blockIdx.x * (blockDim.x >> 5) + wp;
  if (i >= psz) return;
  double v = 0.0;
  for (int blk = lane; blk < nblocks; blk += 32) v += (double)partials[(size_t)blk * psz + i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if (lane != 0) return;
  const float f = (float)v * scale;
  if (i < nA) gp.dA[i] = f;
  else if (i < nA + nB) gp.dB[i - nA] = f;
  else if (i < nA + nB + nC) gp.dC[i - nA - nB] = f;
  else if (gp.dQ) gp.dQ[i - nA - nB - nC] = f;
}

// Data-parallel variant of k_bwd_final (kvae_kf_bwd_dp): the local reduction, the cross-rank exchange over NVLink peer
// memory (8-byte {value, step} words pushed into every rank's buffer, see csrc/kvae_dp.cu), the GLOBAL normalisation and
// the scaling of the rank's dY/dalpha/dU in ONE launch:
//   A  block 0 reduces the per-CTA ELBO partials and pushes the five sums; every parameter warp reduces its element over
//      the local per-CTA rows and pushes it to all ranks (fire and forget);
//   B  every block polls (local memory) the five sums of every rank -> global normaliser;
//   C  parameter warps poll their element from every rank, add in rank order (fp64), scale, store; the other blocks scale
//      dY / dalpha / dU.  No block waits for another block of the same launch (a warp only needs what it pushed itself
//      and block 0's sums; block 0 is scheduled first), so partial residency of the grid cannot dead-lock.
//   The last block to finish advances the rank's step counter (every block has read it by then).
__device__ __forceinline__ void dp_st_ll(unsigned long long* p, float v, unsigned step) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(step) : "memory");
}
__device__ __forceinline__ bool dp_ld_ll(const unsigned long long* p, unsigned step, float& v) {
  unsigned lo, hi;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p) : "memory");
    if (hi == step) break;
    if (clock64() - t0 > 40000000000LL) { v = 0.f; return false; }   // ~20 s: a peer died; do not hang the GPU
  }
  v = __uint_as_float(lo);
  return true;
}
static __global__ void __launch_bounds__(128) k_bwd_final_dp(const float* __restrict__ partials, int nblocks, int psz, int nA,
                                                             int nB, int nC, GradPtrs gp,
                                                             const double* __restrict__ elbo_partials, int n_elbo_partials,
                                                             float* __restrict__ terms, int nparam_blocks, ScaleJob sj,
                                                             DpView dp, int* __restrict__ info) {
  __shared__ float part[16][5];
  __shared__ double tot[5];
  __shared__ int bad;
  unsigned long long* mine = dp.buf[dp.rank];
  const unsigned step = (unsigned)mine[0] + 1u;
  const int s = (int)(step & 1u);
  const int wp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int world = dp.world;
  auto word = [&](int dst, int src, int i) {
    return dp.buf[dst] + KV_DP_HDR_WORDS + ((size_t)(s * world + src)) * dp.nf_pad + i;
  };
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  // ---- A: local reductions, pushed to every rank
  const bool is_param = (int)blockIdx.x < nparam_blocks;
  const int i = blockIdx.x * (blockDim.x >> 5) + wp;
  if (blockIdx.x == 0) {
    for (int q = wp; q < 5; q += 4) {
      double v = 0.0;
      for (int k = lane; k < n_elbo_partials; k += 32) v += elbo_partials[(size_t)k * 5 + q];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane < world) dp_st_ll(word(lane, dp.rank, psz + q), (float)v, step);
    }
  }
  if (is_param && i < psz) {
    double v = 0.0;
    for (int blk = lane; blk < nblocks; blk += 32) v += (double)partials[(size_t)blk * psz + i];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane < world) dp_st_ll(word(lane, dp.rank, i), (float)v, step);
  }
  // ---- B: poll (LOCAL memory).  Lane r of a parameter warp waits for rank r's copy of the warp's element while the
  //         first 5*world threads of the block wait for the ELBO sums: all round trips of a block are in flight together
  float xr = 0.f;
  bool okp = true;
  if (is_param && i < psz && lane < world) okp = dp_ld_ll(word(dp.rank, lane, i), step, xr);
  if ((int)threadIdx.x < 5 * world) {
    const int q = threadIdx.x / 5, j = threadIdx.x % 5;
    float v;
    if (!dp_ld_ll(word(dp.rank, q, psz + j), step, v)) bad = 1;
    part[q][j] = v;
  }
  if (!okp) bad = 1;
  __syncthreads();
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int r = 0; r < world; ++r) v += (double)part[r][threadIdx.x];   // rank order, fp64
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  const double nrm = tot[4] < 1.0 ? 1.0 : tot[4];
  const float scale = (float)(1.0 / nrm);
  if (bad) {
    if (threadIdx.x == 0) atomicOr(info, 2);
  } else {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      for (int j = 0; j < 5; ++j) terms[j] = (float)tot[j];
      terms[5] = (float)((tot[0] + tot[1] + tot[2] + tot[3]) / nrm);
      terms[6] = scale;
      terms[7] = 0.f;
    }
    // ---- C
    if (is_param) {
      if (i < psz) {   // warp-uniform
        double v = 0.0;
        for (int r = 0; r < world; ++r) v += (double)__shfl_sync(0xffffffffu, xr, r);   // rank order, fp64: same bits on every rank
        if (lane == 0) {
          const float f = (float)v * scale;
          if (i < nA) gp.dA[i] = f;
          else if (i < nA + nB) gp.dB[i - nA] = f;
          else if (i < nA + nB + nC) gp.dC[i - nA - nB] = f;
          else if (gp.dQ) gp.dQ[i - nA - nB - nC] = f;
        }
      }
    } else {
      const long nthreads = (long)(gridDim.x - nparam_blocks) * blockDim.x;
      const long tid = (long)(blockIdx.x - nparam_blocks) * blockDim.x + threadIdx.x;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        float* p = sj.p[j];
        if (!p) continue;
        const long n4 = sj.n[j] >> 2;
        f4* p4 = reinterpret_cast<f4*>(p);
        for (long k = tid; k < n4; k += nthreads) { f4 v = p4[k]; v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; p4[k] = v; }
        for (long k = (n4 << 2) + tid; k < sj.n[j]; k += nthreads) p[k] *= scale;
      }
    }
  }
  // ---- the last block to finish advances the step counter
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned* done = reinterpret_cast<unsigned*>(mine + 1);
    const unsigned prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) { *done = 0u; mine[0] = (unsigned long long)step; }
  }
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

template <class C> inline int contract_chunks(long BT) {
  long n = BT / 256;
  if (n < 1) n = 1;
  if (n > 592) n = 592;
  return (int)n;
}
template <class C> size_t bwd_ws_bytes(int B, int T) {
  constexpr int GPB = TPBB<C> / C::L;
  using GA = GradAcc<C>;
  const size_t BT = (size_t)B * T;
  const size_t nn = align256(sizeof(float) * BT * C::N * C::N);
  const size_t nv = align256(sizeof(float) * BT * C::N);
  size_t rows = (size_t)((B + GPB - 1) / GPB);
  size_t dense = 0;
  if (GA::DENSE) {
    rows += (size_t)contract_chunks<C>((long)BT);
    dense = nn + align256(sizeof(float) * BT * C::N * C::M) + (C::QPM ? nn : 0) +
            (C::CSH ? 0 : align256(sizeof(float) * BT * C::N * C::P));
  }
  const size_t ep = align256(sizeof(double) * 5 * (size_t)((B + GPB - 1) / GPB));
  return 2 * nn + 2 * nv + dense + ep + align256(sizeof(float) * rows * GA::PSZ);
}

template <class C>
int launch_bwd(const Args& a, BwdArgs w, const BasePtrs& bp, const float* g_elbo, const float* terms, void* ws,
               GradPtrs gp, cudaStream_t s, const DpView* dp = nullptr) {
  constexpr int GPB = TPBB<C> / C::L;
  using GA = GradAcc<C>;
  const size_t sm = sizeof(float) * smem_floats_bwd<C>();
  (void)cudaGetLastError();
  if (sm > 48 * 1024) {   // per DEVICE attribute: set on every launch (cheap, idempotent) so that a second GPU of the process works too
    cudaError_t e = cudaFuncSetAttribute(k_bwd<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
  }
  const size_t BT = (size_t)a.B * a.T;
  const size_t nn = align256(sizeof(float) * BT * C::N * C::N);
  const size_t nv = align256(sizeof(float) * BT * C::N);
  char* p = reinterpret_cast<char*>(ws);
  w.w_Sig_f = reinterpret_cast<float*>(p); p += nn;
  w.w_Sig_p = reinterpret_cast<float*>(p); p += nn;
  w.w_mu_f = reinterpret_cast<float*>(p); p += nv;
  w.w_mu_p = reinterpret_cast<float*>(p); p += nv;
  if (w.elbo_only) { w.w_Sig_f = w.e_dSig; w.w_mu_f = w.e_dmu; }   // the scratch IS the requested output
  DensePtrs dn{nullptr, nullptr, nullptr, nullptr};
  if (GA::DENSE) {
    dn.A = reinterpret_cast<float*>(p); p += nn;
    dn.B = reinterpret_cast<float*>(p); p += align256(sizeof(float) * BT * C::N * C::M);
    if (C::QPM) { dn.Q = reinterpret_cast<float*>(p); p += nn; }
    if (!C::CSH) { dn.Ct = reinterpret_cast<float*>(p); p += align256(sizeof(float) * BT * C::N * C::P); }
  }
  const int grid = (a.B + GPB - 1) / GPB;
  double* elbo_partials = reinterpret_cast<double*>(p); p += align256(sizeof(double) * 5 * (size_t)grid);
  float* partials = reinterpret_cast<float*>(p);
  constexpr int psz = GA::PSZ;
  constexpr int tpb = TPBB<C>;
  k_bwd<C><<<grid, tpb, sm, s>>>(a, w, bp, g_elbo, terms, partials, dn, elbo_partials);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  int rows = grid;
  if (GA::DENSE) {
    const int nch = contract_chunks<C>((long)BT);
    const int chunk = (int)((BT + nch - 1) / nch);
    float* crow = partials + (size_t)grid * psz;
    cudaMemsetAsync(crow, 0, sizeof(float) * (size_t)nch * psz, s);   // C (shared) columns of these rows stay zero
    constexpr int K = C::K;
    k_mode_contract<K><<<dim3((C::N * C::N + 255) / 256, nch), 256, 0, s>>>(a.alpha, dn.A, (long)BT, C::N * C::N, chunk, crow, psz, GA::fA, 0, 0);
    k_mode_contract<K><<<dim3((C::N * C::M + 255) / 256, nch), 256, 0, s>>>(a.alpha, dn.B, (long)BT, C::N * C::M, chunk, crow, psz, GA::fB, 0, 0);
    if (C::QPM)
      k_mode_contract<K><<<dim3((C::N * C::N + 255) / 256, nch), 256, 0, s>>>(a.alpha, dn.Q, (long)BT, C::N * C::N, chunk, crow, psz, GA::fQ, 0, 0);
    if (!C::CSH)
      k_mode_contract<K><<<dim3((C::N * C::P + 255) / 256, nch), 256, 0, s>>>(a.alpha, dn.Ct, (long)BT, C::N * C::P, chunk, crow, psz, GA::fC, C::P, C::N);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    rows += nch;
  }
  const int nparam_blocks = (psz + 3) / 4;
  ScaleJob sj{{nullptr, nullptr, nullptr}, {0, 0, 0}};
  int scale_blocks = 0;
  const int no_scale = (w.raw_sums || w.mask_part != nullptr) ? 1 : 0;   // normaliser already inside the sweep (or left out)
  if (w.with_elbo && !no_scale) {
    sj.p[0] = w.dY; sj.n[0] = (long)BT * C::P;
    sj.p[1] = w.dalpha; sj.n[1] = (long)BT * C::K;
    sj.p[2] = w.dU; sj.n[2] = w.dU ? (long)BT * C::M : 0;
    const long total4 = (sj.n[0] + sj.n[1] + sj.n[2]) / 4;
    scale_blocks = (int)((total4 + 128 * 8 - 1) / (128 * 8));
    if (scale_blocks < 1) scale_blocks = 1;
    if (scale_blocks > 148 * 8) scale_blocks = 148 * 8;
  }
  if (dp) {   // data parallel: reduction + exchange + global normalisation + scaling in one launch
    if (dp->nparam != psz) return -5;
    sj.p[0] = w.dY; sj.n[0] = (long)BT * C::P;
    sj.p[1] = w.dalpha; sj.n[1] = (long)BT * C::K;
    sj.p[2] = w.dU; sj.n[2] = w.dU ? (long)BT * C::M : 0;
    const long total4 = (sj.n[0] + sj.n[1] + sj.n[2]) / 4;
    int sb = (int)((total4 + 128 * 8 - 1) / (128 * 8));
    if (sb < 1) sb = 1;
    if (sb > 148 * 8) sb = 148 * 8;
    k_bwd_final_dp<<<nparam_blocks + sb, 128, 0, s>>>(partials, rows, psz, C::K * C::N * C::N, C::K * C::N * C::M,
                                                      C::K * C::P * C::N, gp, elbo_partials, grid, w.terms_out, nparam_blocks, sj,
                                                      *dp, a.info);
    return (int)cudaGetLastError();
  }
  k_bwd_final<<<nparam_blocks + scale_blocks, 128, 0, s>>>(partials, rows, psz, C::K * C::N * C::N, C::K * C::N * C::M,
                                                           C::K * C::P * C::N, gp, w.with_elbo ? elbo_partials : nullptr, grid,
                                                           w.terms_out, no_scale, nparam_blocks, sj);
  return (int)cudaGetLastError();
}

}  // namespace kvae
