// Issue-rate microbenchmark of the instructions the softmax warps are made of (sm_100a).  One CTA of 256 threads per SM
// (2 warps per scheduler, as in the attention kernels); each thread runs independent chains so latency is hidden; the
// result is cycles per warp-instruction per scheduler.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
enum { K_EX2, K_FFMA, K_FFMA2, K_FADD2, K_F2FP, K_FMNMX, K_LEA, K_EX2_FFMA2, K_EX2_F2FP, K_N };
template <int KIND>
__global__ void bench(float* out, long long* cyc, int iters, float s) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = s * (threadIdx.x + i) * 1e-3f;
  u64 y[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] = pk(x[2 * i], x[2 * i + 1]);
  uint32_t z[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] = threadIdx.x + i;
  const u64 c2 = pk(s, s);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (KIND == K_EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (KIND == K_FFMA) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(s));
      if (KIND == K_FMNMX) asm volatile("max.ftz.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(s));
      if (KIND == K_LEA) asm volatile("{ .reg .u32 t; shl.b32 t, %0, 23; add.u32 %0, t, %1; }" : "+r"(z[i]) : "r"(z[(i + 1) & 15]));
      if (KIND == K_F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(z[i]) : "f"(x[i]), "f"(x[(i + 1) & 15]));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == K_FFMA2) { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(y[i]) : "l"(c2)); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(y[i]) : "l"(c2)); }
      if (KIND == K_FADD2) { asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(c2)); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[i]) : "l"(c2)); }
    }
    if (KIND == K_EX2_FFMA2) {   // 8 MUFU + 8 FFMA2 interleaved: do the two pipes overlap?
#pragma unroll
      for (int i = 0; i < 8; ++i) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(y[i]) : "l"(c2)); }
    }
    if (KIND == K_EX2_F2FP) {    // 8 MUFU + 8 bf16x2 packs
#pragma unroll
      for (int i = 0; i < 8; ++i) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i])); asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(z[i]) : "f"(x[8 + i]), "f"(x[(9 + i) & 15])); }
    }
  }
  long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc += x[i] + __uint_as_float(z[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; upk(y[i], a, b); acc += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int KIND>
void run(const char* name, int per_iter) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 256 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  bench<KIND><<<148, 256>>>(out, cyc, 10, 0.5f);
  bench<KIND><<<148, 256>>>(out, cyc, iters, 0.5f);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 2 warps per scheduler each issue per_iter instructions per iteration
  printf("%-12s %6.2f cycles per warp-instruction per scheduler   (%s)\n", name, (double)h / iters / (2.0 * per_iter), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<K_EX2>("MUFU.EX2", 16); run<K_FFMA>("FFMA", 16); run<K_FFMA2>("FFMA2", 16); run<K_FADD2>("FADD2", 16);
  run<K_F2FP>("F2FP.BF16", 16); run<K_FMNMX>("FMNMX", 16); run<K_LEA>("SHL+IADD", 16);
  run<K_EX2_FFMA2>("EX2+FFMA2", 16); run<K_EX2_F2FP>("EX2+F2FP", 16);
  return 0;
}
