// Register-resident FMA-chain microbenchmarks.  SURVEY.md section 6 / BASELINE.md section 3: the
// FP32 and FP64 CUDA-core peaks are not in MEASURED_PEAKS.json, so the consensus kernel's roofline
// denominator is measured here on the same device, same clocks, right next to the timed run.
//   kind 0: scalar fp32 FFMA, 16 independent chains per thread
//   kind 1: packed fp32x2 FFMA2 (fma.rn.f32x2, sm_100+), 16 independent chains (32 lane-FMAs)
//   kind 2: fp64 DFMA, 8 independent chains
#include "engine.h"

namespace lsqr {

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

template <int KIND>
__global__ void __launch_bounds__(256) fma_bench_kernel(int iters, float* __restrict__ sink) {
  const float seed = 1.0f + 1e-7f * (float)threadIdx.x;
  if (KIND == 0) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = seed + i;
    const float m = 0.999999f, c = 1e-6f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += a[i];
    if (s == 123.456f) sink[0] = s;
  } else if (KIND == 1) {
    unsigned long long a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { const float v = seed + i; a[i] = ((unsigned long long)__float_as_uint(v) << 32) | __float_as_uint(v + 0.5f); }
    const unsigned long long m = ((unsigned long long)__float_as_uint(0.999999f) << 32) | __float_as_uint(0.999998f);
    const unsigned long long c = ((unsigned long long)__float_as_uint(1e-6f) << 32) | __float_as_uint(2e-6f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fma2(a[i], m, c);
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= a[i];
    if (s == 0x123456789abcdefull) sink[0] = 1.0f;
  } else {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = (double)seed + i;
    const double m = 0.999999, c = 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < 8; u++)
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 123.456) sink[0] = (float)s;
  }
}

void launch_fma_bench(int kind, int iters, int blocks, int threads, float* sink, cudaStream_t s) {
  if (kind == 0) fma_bench_kernel<0><<<blocks, threads, 0, s>>>(iters, sink);
  else if (kind == 1) fma_bench_kernel<1><<<blocks, threads, 0, s>>>(iters, sink);
  else fma_bench_kernel<2><<<blocks, threads, 0, s>>>(iters, sink);
}

}  // namespace lsqr
