// kvae_kernels.cuh — __global__ entry points: one CTA of 128 threads = 128/L lane groups, each
// group owning one sequence for the whole sweep; base matrices staged once per CTA in shared
// memory; per-group publish tiles behind them.
#pragma once
#include <cuda_runtime.h>
#include "kvae_fwd.cuh"

namespace kvae {

constexpr int kThreads = 128;

struct BasePtrs { const float *A, *Bm, *C, *Q, *R, *mu0, *S0; };

template <class C> __device__ __forceinline__ float* stage_base(float* smem, const BasePtrs& bp) {
  for (int i = threadIdx.x; i < Base<C>::total; i += blockDim.x)
    base_fill<C>(smem, i, bp.A, bp.Bm, bp.C, bp.Q, bp.R, bp.mu0, bp.S0);
  __syncthreads();
  return smem + Base<C>::total;
}

template <class C> constexpr size_t smem_bytes() {
  return sizeof(float) * (size_t)(Base<C>::total + (kThreads / C::L) * Tiles<C>::total);
}

template <class C>
__global__ void __launch_bounds__(kThreads) k_filter_smooth(Args a, BasePtrs bp, int smooth) {
  extern __shared__ f4 smem_raw[];
  float* base = reinterpret_cast<float*>(smem_raw);
  float* tiles_all = stage_base<C>(base, bp);
  constexpr int GPB = kThreads / C::L;
  const int gi = threadIdx.x / C::L;
  Group<C::L, C::R> g{(int)(threadIdx.x % C::L)};
  int b = blockIdx.x * GPB + gi;
  const bool active = b < a.B;
  if (!active) b = a.B - 1;  // tail groups recompute the last sequence and store nothing
  float* tiles = tiles_all + gi * Tiles<C>::total;
  float Sig[C::R][C::N], mu[C::N], mu_own[C::R];
  filter_sweep<C>(a, base, tiles, g, b, active, Sig, mu, mu_own);
  if (smooth) smoother_sweep<C>(a, base, tiles, g, b, active, Sig, mu_own);
}

template <class C> int launch_fwd(const Args& a, const BasePtrs& bp, int smooth, cudaStream_t s) {
  constexpr int GPB = kThreads / C::L;
  const size_t sm = smem_bytes<C>();
  static bool attr_set = false;  // benign race: idempotent
  if (sm > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_filter_smooth<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  const int grid = (a.B + GPB - 1) / GPB;
  k_filter_smooth<C><<<grid, kThreads, sm, s>>>(a, bp, smooth);
  return (int)cudaGetLastError();
}

}  // namespace kvae
