// R&D harness (not part of the product): how do the FMA pipe (FFMA / FFMA2) and the ALU pipe
// (FSETP, IADD3, LEA.HI) share issue slots on sm_100a?  Independent register chains, no memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_lab.cu -o tools/bin/pipe_lab
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// per loop iteration: NF2 FFMA2 (distinct 3-register-pair operands), NF1 scalar FFMA (3 distinct regs),
// NLEA shift-adds (acc += x >> 31), NSP (FSETP + predicated add) pairs.
template <int NF2, int NF1, int NLEA, int NSP, bool BCAST = false>
__global__ void __launch_bounds__(256) pipe_kernel(int iters, float* sink, float seedf) {
  u64 a2[16], b2[8], c2[16];
  float a1[16], b1[8], c1[16];
  uint32_t acc[8], cnt[8];
  const float s = seedf + threadIdx.x * 1e-7f;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    a2[i] = ((u64)__float_as_uint(s + i) << 32) | __float_as_uint(s - i);
    c2[i] = ((u64)__float_as_uint(1e-6f * i) << 32) | __float_as_uint(2e-6f * i);
    a1[i] = s + 0.5f * i; c1[i] = 1e-6f * i;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) { b2[i] = ((u64)__float_as_uint(0.99999f - 1e-6f * i) << 32) | __float_as_uint(0.99998f + 1e-6f * i); b1[i] = 0.99999f - 1e-6f * i; acc[i] = i; cnt[i] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < NF2; i++) {
        if (BCAST) { u64 bb; asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b1[i % 8])); a2[i % 16] = ffma2(bb, a2[i % 16], c2[(i + u) % 16]); }
        else a2[i % 16] = ffma2(a2[i % 16], b2[i % 8], c2[(i + u) % 16]);
      }
#pragma unroll
      for (int i = 0; i < NF1; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1[i % 16]) : "f"(b1[i % 8]), "f"(c1[(i + u) % 16]));
#pragma unroll
      for (int i = 0; i < NLEA; i++) asm volatile("{ .reg .u32 t; shr.u32 t, %1, 31; add.u32 %0, %0, t; }" : "+r"(acc[i % 8]) : "r"(__float_as_uint(a1[(i + u) % 16])));
#pragma unroll
      for (int i = 0; i < NSP; i++) asm volatile("{ .reg .pred p; setp.lt.f32 p, %1, %2; @p add.u32 %0, %0, 1; }" : "+r"(cnt[i % 8]) : "f"(a1[(i + u) % 16]), "f"(c1[i % 16]));
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) r += a1[i] + __uint_as_float((uint32_t)a2[i]) + __uint_as_float((uint32_t)(a2[i] >> 32));
#pragma unroll
  for (int i = 0; i < 8; i++) r += (float)(acc[i] + cnt[i]);
  if (r == 123.456f) sink[0] = r;
}

template <int NF2, int NF1, int NLEA, int NSP, bool BCAST = false>
void run(const char* name, float* sink) {
  const int iters = 2048, blocks = 148 * 8;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CK(cudaEventRecord(e0));
    pipe_kernel<NF2, NF1, NLEA, NSP, BCAST><<<blocks, 256>>>(iters, sink, 1.0f);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  // cycles per SMSP per loop body (one 'u' slice) at 1965 MHz: warps per SMSP = blocks*8/(148*4) = 16
  const double warp_iters = (double)blocks * 8 * iters * 4 / (148.0 * 4);
  const double cyc = best * 1e-3 * 1.965e9 / warp_iters;
  printf("%-44s FFMA2=%2d FFMA=%2d LEA=%2d SETP+ADD=%2d : %7.3f ms  %6.2f cycles per body (sum of naive issue = %d)\n", name, NF2, NF1, NLEA, NSP, best, cyc,
         NF2 + NF1 + NLEA + 2 * NSP);
}

int main() {
  float* sink; CK(cudaMalloc(&sink, 64));
  run<12, 0, 0, 0>("FFMA2 only", sink);
  run<12, 0, 0, 0, true>("FFMA2 with one .F32 broadcast operand", sink);
  run<0, 12, 0, 0>("FFMA (3 distinct regs) only", sink);
  run<0, 0, 8, 0>("LEA/shift-add only", sink);
  run<0, 0, 0, 8>("FSETP+@P ADD only", sink);
  run<8, 0, 4, 0>("v6 ratio   FFMA2:LEA = 2:1", sink);
  run<12, 0, 0, 8>("v5 ratio   FFMA2:(SETP+ADD) = 3:2", sink);
  run<12, 0, 8, 0>("FFMA2:LEA = 3:2", sink);
  run<12, 0, 4, 0>("FFMA2:LEA = 3:1", sink);
  run<0, 12, 0, 4>("scalar v0 ratio FFMA:(SETP+ADD) = 3:1", sink);
  run<6, 0, 6, 0>("FFMA2:LEA = 1:1", sink);
  run<6, 0, 0, 3>("FFMA2:(SETP+ADD) = 2:1", sink);
  return 0;
}
