// ub_stream.cu — micro-benchmark of the access patterns the Kalman sweeps can use for their per-sequence
// streams ([B,T,W] fp32, W floats per sequence-step, one warp = 32 sequences when a thread owns a sequence):
//   st_l1     thread-per-sequence, direct 128-bit stores (32 lines per instruction)
//   st_l4     four lanes per sequence, direct 128-bit stores (8 lines per instruction)
//   st_tma    thread-per-sequence, rows staged in shared memory, ONE cp.async.bulk.tensor store per warp and TAU steps
//   st_blk    thread-per-sequence, rows staged in shared memory, one 1-D cp.async.bulk per THREAD and TAU steps
//   ld_l1 / ld_tma  the same for loads
// plus FFMA vs FFMA2 (fma.rn.f32x2) issue throughput.  Development tooling (numbers -> DESIGN.md); not part of the package.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_stream ub_stream.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (EncodeFn)fn;
}
// [B][T][W] fp32 viewed as 3-D (W, T, B), box (W, tau, 32)
static CUtensorMap make_map(EncodeFn enc, float* p, int B, int T, int W, int tau) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * 4 * T};
  cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)tau, 32};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (W=%d tau=%d)\n", (int)r, W, tau); exit(1); }
  return m;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int W> __global__ void __launch_bounds__(128) st_l1(float* out, int B, int T) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  float4 v = make_float4(b, 1.f, 2.f, 3.f);
  for (int t = 0; t < T; ++t) {
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)b * T + t) * W);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) { v.y += 1.f; dst[q] = v; }
  }
}
template <int W> __global__ void __launch_bounds__(128) st_l4(float* out, int B, int T) {   // W == 16: one row per lane
  const int b = (blockIdx.x * 128 + threadIdx.x) / 4, l = threadIdx.x & 3;
  if (b >= B) return;
  float4 v = make_float4(b, 1.f, 2.f, 3.f);
  for (int t = 0; t < T; ++t) {
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)b * T + t) * W);
    v.y += 1.f;
    dst[l] = v;
  }
}
// staging layout per warp: [32 seq][TAU][W] floats (= the TMA box, dense)
template <int W, int TAU> __global__ void __launch_bounds__(128) st_tma(const __grid_constant__ CUtensorMap map, int B, int T) {
  extern __shared__ __align__(128) float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wbuf = sm + warp * (32 * TAU * W);
  const int b0 = blockIdx.x * 128 + warp * 32;
  if (b0 >= B) return;
  float4 v = make_float4(b0 + lane, 1.f, 2.f, 3.f);
  for (int t0 = 0; t0 < T; t0 += TAU) {
    // previous store must have finished READING the buffer
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int s = 0; s < TAU; ++s) {
      float4* dst = reinterpret_cast<float4*>(wbuf + (lane * TAU + s) * W);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) { v.y += 1.f; dst[q] = v; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                   :: "l"(&map), "r"(0), "r"(t0), "r"(b0), "r"(smem_u32(wbuf)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <int W, int TAU> __global__ void __launch_bounds__(128) st_blk(float* out, int B, int T) {
  extern __shared__ __align__(128) float sm[];
  const int b = blockIdx.x * 128 + threadIdx.x;
  float* tbuf = sm + threadIdx.x * (TAU * W);
  if (b >= B) return;
  float4 v = make_float4(b, 1.f, 2.f, 3.f);
  for (int t0 = 0; t0 < T; t0 += TAU) {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
#pragma unroll
    for (int s = 0; s < TAU; ++s) {
      float4* dst = reinterpret_cast<float4*>(tbuf + s * W);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) { v.y += 1.f; dst[q] = v; }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(out + ((size_t)b * T + t0) * W), "r"(smem_u32(tbuf)), "r"(TAU * W * 4) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int W> __global__ void __launch_bounds__(128) ld_l1(const float* in, float* sink, int B, int T) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    const float4* src = reinterpret_cast<const float4*>(in + ((size_t)b * T + t) * W);
#pragma unroll
    for (int q = 0; q < W / 4; ++q) { float4 v = src[q]; acc += v.x + v.y + v.z + v.w; }
  }
  if (acc == 12345.678f) sink[b] = acc;
}
// double-buffered TMA loads: box (W, TAU, 32) per warp, mbarrier per buffer
template <int W, int TAU> __global__ void __launch_bounds__(128) ld_tma(const __grid_constant__ CUtensorMap map, float* sink, int B, int T) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) unsigned long long bars[4][2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wbuf = sm + warp * (2 * 32 * TAU * W);
  const int b0 = blockIdx.x * 128 + warp * 32;
  if (b0 >= B) return;
  const uint32_t bar0 = smem_u32(&bars[warp][0]), bar1 = smem_u32(&bars[warp][1]);
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  constexpr uint32_t BYTES = 32 * TAU * W * 4;
  auto issue = [&](int chunk) {
    const uint32_t bar = (chunk & 1) ? bar1 : bar0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 :: "r"(smem_u32(wbuf + (chunk & 1) * 32 * TAU * W)), "l"(&map), "r"(0), "r"(chunk * TAU), "r"(b0), "r"(bar) : "memory");
  };
  const int nchunk = T / TAU;
  if (lane == 0) issue(0);
  float acc = 0.f;
  for (int c = 0; c < nchunk; ++c) {
    if (lane == 0 && c + 1 < nchunk) issue(c + 1);
    const uint32_t bar = (c & 1) ? bar1 : bar0;
    const uint32_t parity = (c >> 1) & 1;
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
    const float* buf = wbuf + (c & 1) * 32 * TAU * W;
#pragma unroll
    for (int s = 0; s < TAU; ++s) {
      const float4* src = reinterpret_cast<const float4*>(buf + (lane * TAU + s) * W);
#pragma unroll
      for (int q = 0; q < W / 4; ++q) { float4 v = src[q]; acc += v.x + v.y + v.z + v.w; }
    }
    __syncwarp();   // everyone finished reading before the buffer is refilled (issue(c+2) happens next iteration)
  }
  if (acc == 12345.678f) sink[b0 + lane] = acc;
}

// ---- FMA issue throughput: 8 independent chains per thread
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters) {
  float a[8], x = threadIdx.x * 1e-3f, y = 0.999f;
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], y, x);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters) {
  unsigned long long a[8], xy, yy;
  float2 t;
#pragma unroll
  for (int i = 0; i < 8; ++i) { t = make_float2(i, i + 0.5f); a[i] = *reinterpret_cast<unsigned long long*>(&t); }
  t = make_float2(threadIdx.x * 1e-3f, threadIdx.x * 2e-3f); xy = *reinterpret_cast<unsigned long long*>(&t);
  t = make_float2(0.999f, 0.998f); yy = *reinterpret_cast<unsigned long long*>(&t);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(yy), "l"(xy));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { t = *reinterpret_cast<float2*>(&a[i]); s += t.x + t.y; }
  out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <class F> static float time_ms(F&& f, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

template <int W> static void run_width(EncodeFn enc, int B, int T) {
  const size_t n = (size_t)B * T * W;
  float *buf, *sink;
  CK(cudaMalloc(&buf, n * 4)); CK(cudaMalloc(&sink, (size_t)B * 4));
  CK(cudaMemset(buf, 0, n * 4));
  const int grid = (B + 127) / 128;
  const double gb = n * 4 / 1e9;
  auto rep = [&](const char* name, float ms) { printf("W=%2d B=%d T=%d  %-22s %8.3f ms  %7.1f GB/s\n", W, B, T, name, ms, gb / (ms * 1e-3)); fflush(stdout); };
  rep("st_l1 (direct)", time_ms([&] { st_l1<W><<<grid, 128>>>(buf, B, T); }));
  if (W == 16) rep("st_l4 (direct)", time_ms([&] { st_l4<W><<<(B * 4 + 127) / 128, 128>>>(buf, B, T); }));
  {
    CUtensorMap m1 = make_map(enc, buf, B, T, W, 1), m4 = make_map(enc, buf, B, T, W, 4);
    CK(cudaFuncSetAttribute(st_tma<W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(st_blk<W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(ld_tma<W, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    rep("st_tma tau=1", time_ms([&] { st_tma<W, 1><<<grid, 128, 4 * 32 * 1 * W * 4>>>(m1, B, T); }));
    rep("st_tma tau=4", time_ms([&] { st_tma<W, 4><<<grid, 128, 4 * 32 * 4 * W * 4>>>(m4, B, T); }));
    rep("st_blk tau=4 (1-D/thread)", time_ms([&] { st_blk<W, 4><<<grid, 128, 128 * 4 * W * 4>>>(buf, B, T); }));
    rep("ld_l1 (direct)", time_ms([&] { ld_l1<W><<<grid, 128>>>(buf, sink, B, T); }));
    rep("ld_tma tau=1", time_ms([&] { ld_tma<W, 1><<<grid, 128, 4 * 2 * 32 * 1 * W * 4>>>(m1, sink, B, T); }));
    rep("ld_tma tau=4", time_ms([&] { ld_tma<W, 4><<<grid, 128, 4 * 2 * 32 * 4 * W * 4>>>(m4, sink, B, T); }));
  }
  CK(cudaFree(buf)); CK(cudaFree(sink));
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 262144, T = argc > 2 ? atoi(argv[2]) : 20;
  EncodeFn enc = get_encode();
  run_width<16>(enc, B, T);
  run_width<4>(enc, B, T);
  run_width<8>(enc, B, T);
  {
    float* out; CK(cudaMalloc(&out, 148 * 8 * 256 * 4));
    const int iters = 20000;
    float t1 = time_ms([&] { k_ffma<<<148 * 8, 256>>>(out, iters); });
    float t2 = time_ms([&] { k_ffma2<<<148 * 8, 256>>>(out, iters); });
    const double fma = 148.0 * 8 * 256 * 8 * iters;
    printf("FFMA : %.3f ms  %.1f TFMA/s (scalar FMAs)\n", t1, fma / (t1 * 1e-3) / 1e12);
    printf("FFMA2: %.3f ms  %.1f TFMA/s (scalar FMAs, 2 per instruction)\n", t2, 2 * fma / (t2 * 1e-3) / 1e12);
  }
  return 0;
}
