// tcgen05.mma issue/throughput microbenchmark (one CTA per SM, one issuing thread): measures cycles per MMA for
//   SS vs TS (A in TMEM), K-major vs MN-major B, N in {64,128,256}, 1/2/4 independent accumulator chains,
// with and without concurrent tcgen05.ld traffic from 4 other warps.   Operands are whatever is in smem: timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -Ifms_fsdp_b200/csrc scripts/mma_bench.cu -o /tmp/mma_bench && /tmp/mma_bench
#include "common.cuh"
#include <cstdio>
using namespace b200;

template <int N, bool TS, bool B_MN, int CHAINS>
__global__ void __launch_bounds__(192, 1) mma_bench(long long* out, int nrep, int ldtm) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 65536);
  constexpr uint32_t idesc = make_idesc_bf16(128, N, false, B_MN);
  if (warp == 0) {
    if (lane == 0) {
      // D chains at columns 0, N, 2N..(<=256 cols used by D) ; A (TS) at columns 384..
      const long long t0 = clock64();
      const uint64_t bd0 = B_MN ? make_smem_desc(b_s, 8192, 1024) : make_smem_desc(b_s, 0, 1024);
      const uint64_t ad0 = make_smem_desc(a_s, 0, 1024);
      for (int r = 0; r < nrep; r += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {             // walk the 4 K-steps of a 64-wide swizzle atom (constant offsets)
          const int c = k % CHAINS;
          const uint64_t bd = bd0 + (B_MN ? (k * 2048 >> 4) : (k * 32 >> 4));
          if constexpr (TS) umma_bf16_ts(tmem + c * N, tmem + 384 + k * 8, bd, idesc, 1);
          else umma_bf16_ss(tmem + c * N, ad0 + (k * 32 >> 4), bd, idesc, 1);
        }
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      stop = 1;
    }
    __syncwarp();
  } else if (warp >= 2 && ldtm) {
    // background TMEM reads (like softmax warps): 32x32b.x32 loads from columns 256..383
    const uint32_t la = static_cast<uint32_t>((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    while (!stop) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + la + 256, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[i];
    }
    if (acc == 0x12345678) out[3] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// The attention-backward MMA streams with their exact addressing: 'S-type' = 2 x 8 SS MMAs (M128 N64 K16, both K-major,
// operands walking two 64-wide swizzle chunks), 'A-type' = 2 x 4 TS MMAs (M128 N128 K16, B MN-major).  fill: 0 = smem as
// found (zeros), 1 = random bf16 data.
__global__ void __launch_bounds__(320, 1) mma_bench_attn(long long* out, int nrep, int fill, int which, int spin) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  uint32_t* w = reinterpret_cast<uint32_t*>(smem);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (i >> 7);
    // two bf16 in roughly [-2, 2): sign | exponent 0x3f/0x40 | random mantissa
    w[i] = fill ? ((h & 0x807f807fu) | 0x3f803f80u) : 0u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (fill && warp >= 2) {   // random A operands in TMEM columns 0..255 (P / dS positions)
    uint32_t v[16];
    for (int c = 0; c < 256; c += 16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (((c + i) * 40503u + threadIdx.x * 2654435761u) & 0x807f807fu) | 0x3f003f00u;
      tmem_st_32x32b_x16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t x1 = smem_u32(smem), x2 = x1 + 32768, y1 = x1 + 65536, y2 = y1 + 16384;
  constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, false, false);
  constexpr uint32_t idesc_a = make_idesc_bf16(128, 128, false, true);
  if (warp == 0 && lane == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < nrep; ++r) {
      if (which & 1) {
        const uint32_t t1 = tmem + (r & 1) * 128, t2 = t1 + 64;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32, yo = (kk >> 2) * 8192 + (kk & 3) * 32;
          umma_bf16_ss(t1, make_smem_desc(x1 + xo, 0, 1024), make_smem_desc(y1 + yo, 0, 1024), idesc_t, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32, yo = (kk >> 2) * 8192 + (kk & 3) * 32;
          umma_bf16_ss(t2, make_smem_desc(x2 + xo, 0, 1024), make_smem_desc(y2 + yo, 0, 1024), idesc_t, kk != 0);
        }
      }
      if (which & 2) {
        const uint32_t tP = tmem + (r & 1) * 128, tdS = tP + 64;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t co = 32 * (t >> 1) + 8 * (t & 1);
          umma_bf16_ts(tmem + 256, tP + co, make_smem_desc(y2 + t * 2048, 8192, 1024), idesc_a, 1);
          umma_bf16_ts(tmem + 384, tdS + co, make_smem_desc(y1 + t * 2048, 8192, 1024), idesc_a, 1);
        }
      }
    }
    const long long t1c = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2c = clock64();
    if (blockIdx.x == 0) { out[0] = t1c - t0; out[1] = t2c - t0; }
    mbar_arrive(&bar2);
  } else if (warp >= 2 && spin) {
    // like the softmax warps waiting for their tile: spin == 1 every lane polls, spin == 2 one lane polls
    if (spin == 1 || lane == 0) mbar_wait(&bar2, 0);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int N, bool TS, bool B_MN>
void run(const char* name, long long* d) {
  const int nrep = 512;
  for (int ldtm = 0; ldtm < 2; ++ldtm)
    for (int chains = 1; chains <= (N == 256 ? 1 : 2); chains *= 2) {
      auto k = chains == 1 ? mma_bench<N, TS, B_MN, 1> : mma_bench<N, TS, B_MN, 2>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      long long h[2] = {0, 0};
      for (int w = 0; w < 2; ++w) {
        k<<<148, 192, 200 * 1024>>>(d, nrep, ldtm);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
      }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("%-28s N=%3d chains=%d ldtm=%d : issue %.1f cyc/mma, complete %.1f cyc/mma (floor %d)\n", name, N, chains, ldtm,
             (double)h[0] / nrep, (double)h[1] / nrep, 128 * N / 256);
    }
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(mma_bench_attn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int which = 1; which <= 3; ++which)
    for (int fill = 0; fill < 3; ++fill) {
      long long h[2];
      const int spin = fill;   // 0: idle waiters, 1: 8 warps x 32 lanes poll an mbarrier, 2: 8 warps x 1 lane poll
      for (int w = 0; w < 2; ++w) { mma_bench_attn<<<148, 320, 200 * 1024>>>(d, 64, 1, which, spin); cudaDeviceSynchronize(); }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("attn-bwd stream %s pollers=%s : %.0f cycles per iteration (floor: S 16x32=512 [measured N=64 rate 16x48=768], A 8x64=512)\n",
             which == 1 ? "S-type (16 SS N64)" : which == 2 ? "A-type (8 TS N128)" : "S+A", spin == 0 ? "none" : spin == 1 ? "256 threads" : "8 threads", (double)h[1] / 64);
    }
  run<64, false, false>("SS  A:K  B:K", d);
  run<128, false, false>("SS  A:K  B:K", d);
  run<256, false, false>("SS  A:K  B:K", d);
  run<64, false, true>("SS  A:K  B:MN", d);
  run<128, false, true>("SS  A:K  B:MN", d);
  run<64, true, false>("TS  A:tmem B:K", d);
  run<128, true, false>("TS  A:tmem B:K", d);
  run<64, true, true>("TS  A:tmem B:MN", d);
  run<128, true, true>("TS  A:tmem B:MN", d);
  return 0;
}
