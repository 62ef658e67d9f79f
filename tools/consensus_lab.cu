// R&D harness (not part of the product): inner-loop variants of the plane consensus kernel,
// timed on the GPU box to choose the production formulation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo tools/consensus_lab.cu -o gpurun_out/consensus_lab
// Each variant counts |n.p + md| < delta for H hypotheses x N points and must produce identical counts.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ u64 pack2_opaque(float lo, float hi) { u64 d; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint32_t set_lt(float a, float b) { uint32_t m; asm("set.lt.u32.f32 %0, %1, %2;" : "=r"(m) : "f"(a), "f"(b)); return m; }

constexpr int TILE = 512;

// VARIANT 0: scalar fmaf, cnt += cond                     (what nvcc makes of the obvious code)
// VARIANT 1: scalar fmaf, set.lt mask, cnt -= m0 + m1     (one IADD3 per two evals)
// VARIANT 2: fma.rn.f32x2 over point pairs, set.lt masks, IADD3
// VARIANT 3: fma.rn.f32x2, squared residual scaled so that inlier <=> bit 30 clear, cnt += t >> 30 (counts OUTLIERS)
// VARIANT 4: like 2 but the predicate form: setp + selp-free "cnt += p" written with asm predicates (IADD3.X style)
template <int VARIANT, int R, int THREADS>
__global__ void __launch_bounds__(THREADS) lab_kernel(const float* __restrict__ px, const float* __restrict__ py, const float* __restrict__ pz,
                                                        uint32_t n_tiles, uint32_t tiles_per_chunk, const float4* __restrict__ hyp, uint32_t H,
                                                        float delta, uint32_t* __restrict__ counts) {
  __shared__ __align__(16) float sx[TILE], sy[TILE], sz[TILE];
  const int tid = threadIdx.x;
  const uint32_t hbase = blockIdx.x * (THREADS * R);
  float4 h[R];
  uint32_t cnt[R];
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t idx = hbase + r * THREADS + tid;
    h[r] = idx < H ? hyp[idx] : make_float4(0, 0, 0, 1e30f);
    cnt[r] = 0;
  }
  // variant 3 scales the hypothesis so that the test becomes t = (s*c)^2 < 2, c = sqrt(2)/delta
  u64 hx2[R], hy2[R], hz2[R], hd2[R];
  if (VARIANT >= 2) {
    const float c = (VARIANT == 3) ? 1.41421356237f / delta : 1.0f;
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (VARIANT == 12) {
        hx2[r] = pack2_opaque(h[r].x, h[r].x); hy2[r] = pack2_opaque(h[r].y, h[r].y);
        hz2[r] = pack2_opaque(h[r].z, h[r].z); hd2[r] = pack2_opaque(h[r].w, h[r].w);
      } else {
      hx2[r] = pack2(h[r].x * c, h[r].x * c); hy2[r] = pack2(h[r].y * c, h[r].y * c);
      hz2[r] = pack2(h[r].z * c, h[r].z * c); hd2[r] = pack2(h[r].w * c, h[r].w * c);
      }
    }
  }
  if (VARIANT == 11) {
#pragma unroll
    for (int r = 0; r < R; r++) { h[r].x /= delta; h[r].y /= delta; h[r].z /= delta; h[r].w /= delta; }
  }
  uint32_t total[R];
#pragma unroll
  for (int r = 0; r < R; r++) total[r] = 0;
  uint32_t opaque_one, opaque_msb;
  asm volatile("mov.u32 %0, 1;" : "=r"(opaque_one));
  asm volatile("mov.u32 %0, 0x80000000;" : "=r"(opaque_msb));
  const uint32_t t0 = blockIdx.y * tiles_per_chunk, t1 = min(t0 + tiles_per_chunk, n_tiles);
  for (uint32_t t = t0; t < t1; t++) {
    __syncthreads();
    for (int i = tid; i < TILE; i += THREADS) { sx[i] = px[(size_t)t * TILE + i]; sy[i] = py[(size_t)t * TILE + i]; sz[i] = pz[(size_t)t * TILE + i]; }
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < TILE; i += 4) {
      if (VARIANT <= 1 || (VARIANT >= 8 && VARIANT <= 11)) {
        const float4 x = *reinterpret_cast<const float4*>(sx + i), y = *reinterpret_cast<const float4*>(sy + i), z = *reinterpret_cast<const float4*>(sz + i);
        const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, zs[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int r = 0; r < R; r++) {
          if (VARIANT == 8 || (VARIANT == 10 && (r % 3) != 0)) {          // scalar FFMA, FSETP + predicated add
            float a[4];
#pragma unroll
            for (int u = 0; u < 4; u++) a[u] = fabsf(fmaf(h[r].x, xs[u], fmaf(h[r].y, ys[u], fmaf(h[r].z, zs[u], h[r].w))));
            uint32_t c0 = cnt[r];
            asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                "setp.lt.f32 p0, %1, %5;\n\tsetp.lt.f32 p1, %2, %5;\n\tsetp.lt.f32 p2, %3, %5;\n\tsetp.lt.f32 p3, %4, %5;\n\t"
                "@p0 add.u32 %0, %0, 1;\n\t@p1 add.u32 %0, %0, 1;\n\t@p2 add.u32 %0, %0, 1;\n\t@p3 add.u32 %0, %0, 1;\n\t}"
                : "+r"(c0) : "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(delta));
            cnt[r] = c0;
          } else if (VARIANT == 9 || VARIANT == 10) {                      // scalar FFMA, sign bit of s^2 - delta^2
            const float md2 = -delta * delta;
#pragma unroll
            for (int u = 0; u < 4; u++) { const float sv = fmaf(h[r].x, xs[u], fmaf(h[r].y, ys[u], fmaf(h[r].z, zs[u], h[r].w))); cnt[r] += __float_as_uint(fmaf(sv, sv, md2)) >> 31; }
          } else if (VARIANT == 11) {                                       // scalar FFMA on q = s/delta, packed f16 compare + accumulate
            float q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) q[u] = fmaf(h[r].x, xs[u], fmaf(h[r].y, ys[u], fmaf(h[r].z, zs[u], h[r].w)));
            uint32_t acc = cnt[r];
            asm("{\n\t.reg .b32 h0, h1, m0, m1, one;\n\t"
                "mov.b32 one, 0x3c003c00;\n\t"
                "cvt.rz.f16x2.f32 h0, %2, %1;\n\tcvt.rz.f16x2.f32 h1, %4, %3;\n\t"
                "abs.f16x2 h0, h0;\n\tabs.f16x2 h1, h1;\n\t"
                "set.lt.f16x2.f16x2 m0, h0, one;\n\tset.lt.f16x2.f16x2 m1, h1, one;\n\t"
                "add.f16x2 %0, %0, m0;\n\tadd.f16x2 %0, %0, m1;\n\t}"
                : "+r"(acc) : "f"(q[0]), "f"(q[1]), "f"(q[2]), "f"(q[3]));
            cnt[r] = acc;
          } else if (VARIANT == 0) {
#pragma unroll
            for (int u = 0; u < 4; u++) { const float s = fmaf(h[r].x, xs[u], fmaf(h[r].y, ys[u], fmaf(h[r].z, zs[u], h[r].w))); cnt[r] += fabsf(s) < delta ? 1u : 0u; }
          } else {
            uint32_t m[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const float s = fmaf(h[r].x, xs[u], fmaf(h[r].y, ys[u], fmaf(h[r].z, zs[u], h[r].w))); m[u] = set_lt(fabsf(s), delta); }
            cnt[r] = cnt[r] - m[0] - m[1];
            cnt[r] = cnt[r] - m[2] - m[3];
          }
        }
      } else {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(sx + i), y = *reinterpret_cast<const ulonglong2*>(sy + i), z = *reinterpret_cast<const ulonglong2*>(sz + i);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const u64 s01 = ffma2(hx2[r], x.x, ffma2(hy2[r], y.x, ffma2(hz2[r], z.x, hd2[r])));
          const u64 s23 = ffma2(hx2[r], x.y, ffma2(hy2[r], y.y, ffma2(hz2[r], z.y, hd2[r])));
          if (VARIANT == 2 || VARIANT == 4) {  // mask / select forms
            float a, b, c, d;
            unpack2(s01, a, b); unpack2(s23, c, d);
            if (VARIANT == 2) {
              const uint32_t m0 = set_lt(fabsf(a), delta), m1 = set_lt(fabsf(b), delta), m2 = set_lt(fabsf(c), delta), m3 = set_lt(fabsf(d), delta);
              cnt[r] = cnt[r] - m0 - m1;
              cnt[r] = cnt[r] - m2 - m3;
            } else {
              // carry-in form: cnt = cnt + 0 + p0 + p1 via addc chains
              uint32_t c0 = cnt[r];
              asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .u32 t0, t1, t2, t3;\n\t"
                  "setp.lt.f32 p0, %1, %5;\n\tsetp.lt.f32 p1, %2, %5;\n\tsetp.lt.f32 p2, %3, %5;\n\tsetp.lt.f32 p3, %4, %5;\n\t"
                  "selp.u32 t0, 1, 0, p0;\n\tselp.u32 t1, 1, 0, p1;\n\tselp.u32 t2, 1, 0, p2;\n\tselp.u32 t3, 1, 0, p3;\n\t"
                  "add.u32 t0, t0, t1;\n\tadd.u32 t2, t2, t3;\n\tadd.u32 t0, t0, t2;\n\tadd.u32 %0, %0, t0;\n\t}"
                  : "+r"(c0) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)), "f"(fabsf(d)), "f"(delta));
              cnt[r] = c0;
            }
          } else if ((VARIANT == 13 || VARIANT == 14) && (r % 3) != 0) {
            float a, b, c, d;
            unpack2(s01, a, b); unpack2(s23, c, d);
            uint32_t c0 = cnt[r];
            if (VARIANT == 13) {
              asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                  "setp.lt.f32 p0, %1, %5;\n\tsetp.lt.f32 p1, %2, %5;\n\tsetp.lt.f32 p2, %3, %5;\n\tsetp.lt.f32 p3, %4, %5;\n\t"
                  "@p0 add.u32 %0, %0, %6;\n\t@p1 add.u32 %0, %0, %6;\n\t@p2 add.u32 %0, %0, %6;\n\t@p3 add.u32 %0, %0, %6;\n\t}"
                  : "+r"(c0) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)), "f"(fabsf(d)), "f"(delta), "r"(opaque_one));
            } else {
              asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t.reg .u32 t;\n\t"
                  "setp.lt.f32 p0, %1, %5;\n\tsetp.lt.f32 p1, %2, %5;\n\tsetp.lt.f32 p2, %3, %5;\n\tsetp.lt.f32 p3, %4, %5;\n\t"
                  "@p0 shf.l.wrap.b32 t, %6, %0, 1;\n\t@p0 mov.u32 %0, t;\n\t}"
                  : "+r"(c0) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)), "f"(fabsf(d)), "f"(delta), "r"(opaque_msb));
            }
            cnt[r] = c0;
          } else if (VARIANT == 5 || ((VARIANT == 7 || VARIANT == 12) && (r % 3) != 0)) {
            // FSETP + predicated IADD, forced through PTX so that ptxas cannot turn it into add+select
            float a, b, c, d;
            unpack2(s01, a, b); unpack2(s23, c, d);
            uint32_t c0 = cnt[r];
            asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                "setp.lt.f32 p0, %1, %5;\n\tsetp.lt.f32 p1, %2, %5;\n\tsetp.lt.f32 p2, %3, %5;\n\tsetp.lt.f32 p3, %4, %5;\n\t"
                "@p0 add.u32 %0, %0, 1;\n\t@p1 add.u32 %0, %0, 1;\n\t@p2 add.u32 %0, %0, 1;\n\t@p3 add.u32 %0, %0, 1;\n\t}"
                : "+r"(c0) : "f"(fabsf(a)), "f"(fabsf(b)), "f"(fabsf(c)), "f"(fabsf(d)), "f"(delta));
            cnt[r] = c0;
          } else if (VARIANT == 6 || VARIANT == 7 || VARIANT == 12 || VARIANT == 13 || VARIANT == 14) {
            // w = s^2 - delta^2 on the FMA pipe; the sign bit of w is the inlier flag: cnt += bits(w) >> 31
            const u64 md2 = pack2(-delta * delta, -delta * delta);
            const u64 w01 = ffma2(s01, s01, md2), w23 = ffma2(s23, s23, md2);
            float a, b, c, d;
            unpack2(w01, a, b); unpack2(w23, c, d);
            cnt[r] += (__float_as_uint(a) >> 31);
            cnt[r] += (__float_as_uint(b) >> 31);
            cnt[r] += (__float_as_uint(c) >> 31);
            cnt[r] += (__float_as_uint(d) >> 31);
          } else {
            const u64 q01 = fmul2(s01, s01), q23 = fmul2(s23, s23);
            cnt[r] += (uint32_t)(q01 >> 62) + (uint32_t)((q01 >> 30) & 3u);   // bit 30 of each half (sign bits are 0)
            cnt[r] += (uint32_t)(q23 >> 62) + (uint32_t)((q23 >> 30) & 3u);
          }
        }
      }
    }
    if (VARIANT == 11) {   // flush the packed f16 counters (exact up to 2048 per half; a tile adds at most 256)
#pragma unroll
      for (int r = 0; r < R; r++) {
        const __half2 hv = *reinterpret_cast<__half2*>(&cnt[r]);
        total[r] += (uint32_t)__half2float(__low2half(hv)) + (uint32_t)__half2float(__high2half(hv));
        cnt[r] = 0;
      }
    }
  }
  if (VARIANT == 11) {
#pragma unroll
    for (int r = 0; r < R; r++) cnt[r] = total[r];
  }
#pragma unroll
  for (int r = 0; r < R; r++) {
    const uint32_t idx = hbase + r * THREADS + tid;
    if (idx < H && cnt[r]) atomicAdd(&counts[idx], cnt[r]);
  }
}

template <int VARIANT, int R, int THREADS>
double run(const char* name, const float* px, const float* py, const float* pz, uint32_t N, const float4* hyp, uint32_t H, float delta, uint32_t* counts,
           std::vector<uint32_t>& out, bool outliers) {
  const uint32_t n_tiles = N / TILE;
  const uint32_t hb = (H + THREADS * R - 1) / (THREADS * R);
  uint32_t chunks = (148 * 16 + hb - 1) / hb;
  if (chunks > n_tiles) chunks = n_tiles;
  const uint32_t tpc = (n_tiles + chunks - 1) / chunks;
  chunks = (n_tiles + tpc - 1) / tpc;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CK(cudaMemset(counts, 0, sizeof(uint32_t) * H));
    CK(cudaEventRecord(e0));
    lab_kernel<VARIANT, R, THREADS><<<dim3(hb, chunks), THREADS>>>(px, py, pz, n_tiles, tpc, hyp, H, delta, counts);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  out.resize(H);
  CK(cudaMemcpy(out.data(), counts, sizeof(uint32_t) * H, cudaMemcpyDeviceToHost));
  if (outliers) for (auto& c : out) c = N - c;
  const double evals = (double)N * H;
  const double rate = evals / (best * 1e-3);
  printf("%-34s R=%2d T=%3d  %8.3f ms  %7.3f T evals/s  cycles/eval/SMSP@1965MHz = %.2f\n", name, R, THREADS, best, rate / 1e12, 148.0 * 4 * 32 * 1.965e9 / rate);
  return rate;
}

int main(int argc, char** argv) {
  const uint32_t N = 1u << 20, H = 1u << 18;
  std::vector<float> hx(N), hy(N), hz(N);
  std::vector<float4> hh(H);
  srand(1);
  auto u = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
  for (uint32_t i = 0; i < N; i++) { hx[i] = 1000 * u(); hy[i] = 1000 * u(); hz[i] = (i % 5 < 3) ? 0.4f * u() : 1000 * u(); }
  for (uint32_t i = 0; i < H; i++) {
    float a = 0.02f * u(), b = 0.02f * u(), c = 1.f, n = sqrtf(a * a + b * b + c * c);
    hh[i] = make_float4(a / n, b / n, c / n, 0.3f * u());
  }
  float *px, *py, *pz; float4* hyp; uint32_t* counts;
  CK(cudaMalloc(&px, 4 * N)); CK(cudaMalloc(&py, 4 * N)); CK(cudaMalloc(&pz, 4 * N)); CK(cudaMalloc(&hyp, 16 * H)); CK(cudaMalloc(&counts, 4 * H));
  CK(cudaMemcpy(px, hx.data(), 4 * N, cudaMemcpyHostToDevice)); CK(cudaMemcpy(py, hy.data(), 4 * N, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pz, hz.data(), 4 * N, cudaMemcpyHostToDevice)); CK(cudaMemcpy(hyp, hh.data(), 16 * H, cudaMemcpyHostToDevice));
  const float delta = 0.5f;
  std::vector<uint32_t> ref, got;
  auto check = [&](const char* nm) { size_t bad = 0; long long dsum = 0; for (uint32_t i = 0; i < H; i++) { if (ref[i] != got[i]) bad++; dsum += llabs((long long)ref[i] - got[i]); } printf("    check %-28s mismatching hyps %zu, total |diff| %lld (of %.3g inliers)\n", nm, bad, dsum, (double)N * 0.6 * H * 0.5); };
  run<0, 8, 256>("v0 scalar, cnt += cond", px, py, pz, N, hyp, H, delta, counts, ref, false);
  run<1, 8, 256>("v1 scalar, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false); check("v1");
  run<1, 4, 256>("v1 scalar, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<1, 16, 128>("v1 scalar, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<2, 8, 256>("v2 f32x2, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false); check("v2");
  run<2, 4, 256>("v2 f32x2, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<2, 8, 128>("v2 f32x2, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<2, 12, 128>("v2 f32x2, set.lt + IADD3", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<3, 8, 256>("v3 f32x2, square + bit30", px, py, pz, N, hyp, H, delta, counts, got, true); check("v3");
  run<3, 4, 256>("v3 f32x2, square + bit30", px, py, pz, N, hyp, H, delta, counts, got, true);
  run<4, 8, 256>("v4 f32x2, setp/selp adds", px, py, pz, N, hyp, H, delta, counts, got, false); check("v4");
  run<5, 8, 256>("v5 f32x2, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false); check("v5");
  run<5, 4, 256>("v5 f32x2, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<5, 8, 128>("v5 f32x2, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<5, 6, 256>("v5 f32x2, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<6, 8, 256>("v6 f32x2, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false); check("v6");
  run<6, 4, 256>("v6 f32x2, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<6, 8, 128>("v6 f32x2, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<6, 6, 256>("v6 f32x2, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<7, 9, 256>("v7 hybrid 1/3 sign, 2/3 setp", px, py, pz, N, hyp, H, delta, counts, got, false); check("v7");
  run<7, 6, 256>("v7 hybrid 1/3 sign, 2/3 setp", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<7, 9, 128>("v7 hybrid 1/3 sign, 2/3 setp", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<13, 9, 256>("v13 hybrid, @p add opaque reg", px, py, pz, N, hyp, H, delta, counts, got, false); check("v13");
  run<13, 9, 128>("v13 hybrid, @p add opaque reg", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<13, 12, 128>("v13 hybrid, @p add opaque reg", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<12, 9, 256>("v12 hybrid, packed (non-bcast) hyps", px, py, pz, N, hyp, H, delta, counts, got, false); check("v12");
  run<12, 6, 256>("v12 hybrid, packed (non-bcast) hyps", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<12, 9, 128>("v12 hybrid, packed (non-bcast) hyps", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<8, 8, 256>("v8 scalar, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false); check("v8");
  run<8, 16, 128>("v8 scalar, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<8, 16, 256>("v8 scalar, setp + @p add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<9, 8, 256>("v9 scalar, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false); check("v9");
  run<9, 16, 128>("v9 scalar, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<9, 16, 256>("v9 scalar, s^2-d^2 sign bit", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<10, 9, 256>("v10 scalar hybrid", px, py, pz, N, hyp, H, delta, counts, got, false); check("v10");
  run<10, 15, 128>("v10 scalar hybrid", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<10, 15, 256>("v10 scalar hybrid", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<10, 12, 256>("v10 scalar hybrid", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<11, 8, 256>("v11 scalar, f16x2 set+add", px, py, pz, N, hyp, H, delta, counts, got, false); check("v11");
  run<11, 16, 128>("v11 scalar, f16x2 set+add", px, py, pz, N, hyp, H, delta, counts, got, false);
  run<11, 16, 256>("v11 scalar, f16x2 set+add", px, py, pz, N, hyp, H, delta, counts, got, false);
  return 0;
}
