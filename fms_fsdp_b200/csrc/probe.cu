// Hardware-layout probes (run once on the box by scripts/gpu_diag.py probe; not on any hot path).
// ts_mma_probe: D[128,64] = A[128,64] * B[64,64]^T with A staged in TENSOR MEMORY (tcgen05.st, row per lane,
// two bf16 per 32-bit column) and consumed by tcgen05.mma's TMEM-A form -- validates the operand layout that the
// attention kernels rely on to keep P = softmax(S) out of shared memory.
#include "common.cuh"
#include "tensormap.h"

namespace b200 {

__global__ void __launch_bounds__(128, 1)
ts_mma_probe_kernel(const __grid_constant__ CUtensorMap tmB, const __nv_bfloat16* __restrict__ A, float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_b = reinterpret_cast<uint64_t*>(smem + 8192);
  uint64_t* bar_d = bar_b + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar_d + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar_b, 1);
    mbar_init(bar_d, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const int row = warp * 32 + lane;
  // A row -> 32 packed columns at TMEM columns [64, 96)
  uint32_t w0[16], w1[16];
  const uint32_t* arow = reinterpret_cast<const uint32_t*>(A + row * 64);
#pragma unroll
  for (int i = 0; i < 16; ++i) { w0[i] = arow[i]; w1[i] = arow[16 + i]; }
  tmem_st_32x32b_x16(tmem + lane_addr + 64, w0);
  tmem_st_32x32b_x16(tmem + lane_addr + 80, w1);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_b, 64 * 64 * 2);
    tma_load_2d(smem, &tmB, bar_b, 0, 0);
    mbar_wait(bar_b, 0);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, false, false);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16_ts(tmem, tmem + 64 + k * 8, make_smem_desc(smem_u32(smem) + k * 32, 0, 1024), idesc, k != 0);
    umma_commit(bar_d);
  }
  mbar_wait(bar_d, 0);
  tc_fence_after();
  uint32_t v0[32], v1[32];
  tmem_ld_32x32b_x32(tmem + lane_addr, v0);
  tmem_ld_32x32b_x32(tmem + lane_addr + 32, v1);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    D[row * 64 + i] = __uint_as_float(v0[i]);
    D[row * 64 + 32 + i] = __uint_as_float(v1[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

}  // namespace b200

extern "C" int b200_ts_mma_probe(const void* A, const void* B, float* D, cudaStream_t st) {
  using namespace b200;
  CUtensorMap tm;
  if (make_tmap_2d_bf16(&tm, B, 64, 64, 64, 64, 64)) return -3;
  const int smem = 8192 + 1024 + 64;
  ts_mma_probe_kernel<<<1, 128, smem, st>>>(tm, (const __nv_bfloat16*)A, D);
  return (int)cudaGetLastError();
}
