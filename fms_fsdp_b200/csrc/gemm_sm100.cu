// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] (+)= A * B  with fp32 accumulation in TMEM.
//
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> 4-stage smem ring -> tcgen05.mma (UMMA 128x256x16,
//   issued by one thread) -> 2 x 256-column fp32 accumulators in TMEM -> tcgen05.ld epilogue that
//   overlaps the next tile's main loop.
//
// One kernel serves the three GEMMs of a linear layer (SURVEY.md K1/K4/K6/K8 and their backward):
//   forward  y  = x  W^T : A K-major [M,K],  B K-major [N,K]            ("nt")
//   dgrad    dx = dy W   : A K-major [M,Nr], B MN-major (W is [Nr,K])   ("nn")
//   wgrad    dW = dy^T x : A MN-major (dy is [Mr,N]), B MN-major        ("tn")
// The MN-major cases use the tensor core's transposed-operand smem layout, so no transposes are
// ever materialised.  Epilogues: plain store, +residual (attention/MLP output projections),
// accumulate into C (gradient accumulation), bf16 or fp32 output.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).
#include "common.cuh"
#include "tensormap.h"

namespace b200 {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int TMEM_COLS = 512;  // two 256-column accumulators
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_ACCUM = 2 };


// L2-friendly rasterisation (see gemm2_sm100.cu): bands of GROUP_M m-tiles, n fastest across a band.
constexpr int GROUP_M = 16;
B200_DEVINL void tile_coords(int t, int m_tiles, int n_tiles, int& mt, int& nt) {
  const int per_band = GROUP_M * n_tiles;
  const int band = t / per_band;
  const int first_m = band * GROUP_M;
  const int band_m = min(GROUP_M, m_tiles - first_m);
  const int r = t - band * per_band;
  mt = first_m + r % band_m;
  nt = r / band_m;
}

struct GemmParams {
  int M, N, K;          // C is [M,N], reduction length K
  int ldc, ldr;         // row strides (elements) of C and residual
  void* C;
  const void* R;        // residual (bf16) or nullptr
  int m_tiles, n_tiles;
  // batched mode (BATCH = true): nb0 x nb1 independent problems; operands come through rank-4 tensor maps
  // {inner, outer, b0, b1}; C (and R) of problem (b0, b1) start at b0 * sc0 + b1 * sc1 elements
  int nb0, nb1;
  long long sc0, sc1;
};

// BATCH: many small independent GEMMs in one persistent launch (Mamba2 SSD chunk products: C_c B_c^T, masked-score x X,
// chunk states ...).  The tile loop runs over (problem, m-tile, n-tile); TMA coordinates get the two batch indices.
template <bool A_MN, bool B_MN, int EPI, typename OutT, bool BATCH = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms must sit on 1024 B boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  const int tiles_per_problem = p.m_tiles * p.n_tiles;
  const int num_tiles = BATCH ? tiles_per_problem * p.nb0 * p.nb1 : tiles_per_problem;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mt, nt, b0 = 0, b1 = 0;
        if constexpr (BATCH) {
          const int prob = t / tiles_per_problem;
          b0 = prob % p.nb0; b1 = prob / p.nb0;
          tile_coords(t - prob * tiles_per_problem, p.m_tiles, p.n_tiles, mt, nt);
        } else {
          tile_coords(t, p.m_tiles, p.n_tiles, mt, nt);
        }
        const int m0 = mt * BM, n0 = nt * BN;
        auto load = [&](void* dst, const CUtensorMap* tm, int c0, int c1) {
          if constexpr (BATCH) tma_load_4d(dst, tm, &full_bar[stage], c0, c1, b0, b1);
          else tma_load_2d(dst, tm, &full_bar[stage], c0, c1);
        };
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          const int k0 = kb * BK;
          if constexpr (!A_MN) {
            load(sa, &tmA, k0, m0);                                    // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)                          // boxes {64 m, 64 k-rows}
              load(sa + j * (64 * BK * 2), &tmA, m0 + 64 * j, k0);
          }
          if constexpr (!B_MN) {
            load(sb, &tmB, k0, n0);                                    // box {64 k, 256 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              load(sb + j * (64 * BK * 2), &tmB, n0 + 64 * j, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {   // single elected lane: back-to-back UTCHMMA, descriptors = per-stage base + constant
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint64_t a0 = make_smem_desc(sa, A_MN ? 64 * BK * 2 : 0, 1024);
          const uint64_t b0 = make_smem_desc(sb, B_MN ? 64 * BK * 2 : 0, 1024);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = a0 + ((A_MN ? k * (UMMA_K * 128) : k * (UMMA_K * 2)) >> 4);
            const uint64_t bdesc = b0 + ((B_MN ? k * (UMMA_K * 128) : k * (UMMA_K * 2)) >> 4);
            umma_bf16_ss(tmem_d, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);                 // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ============================== epilogue ==============================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mt, nt;
      size_t boff = 0;
      if constexpr (BATCH) {
        const int prob = t / tiles_per_problem;
        boff = (size_t)(prob % p.nb0) * p.sc0 + (size_t)(prob / p.nb0) * p.sc1;
        tile_coords(t - prob * tiles_per_problem, p.m_tiles, p.n_tiles, mt, nt);
      } else {
        tile_coords(t, p.m_tiles, p.n_tiles, mt, nt);
      }
      const int m0 = mt * BM, n0 = nt * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      OutT* crow = reinterpret_cast<OutT*>(p.C) + boff + static_cast<size_t>(row) * p.ldc;
      const __nv_bfloat16* rrow = reinterpret_cast<const __nv_bfloat16*>(p.R) + boff + static_cast<size_t>(row) * p.ldr;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 64) {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32b_x32(taddr + c, v0);
        tmem_ld_32x32b_x32(taddr + c + 32, v1);
        tmem_ld_wait();
        const int col = n0 + c;
        if (row_ok && col < p.N) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t* v = h ? v1 : v0;
            const int cb = col + 32 * h;
            if (cb >= p.N) break;
            if constexpr (sizeof(OutT) == 2) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {  // 8 columns (16 B) per store
                if (cb + g * 8 >= p.N) break;
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[g * 8 + i]);
                if constexpr (EPI == EPI_RESIDUAL) {
                  uint4 r = *reinterpret_cast<const uint4*>(rrow + cb + g * 8);
                  float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c2 = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
                  f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y; f[4] += c2.x; f[5] += c2.y; f[6] += d.x; f[7] += d.y;
                } else if constexpr (EPI == EPI_ACCUM) {
                  uint4 r = *reinterpret_cast<const uint4*>(crow + cb + g * 8);
                  float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y), c2 = unpack_bf16x2(r.z), d = unpack_bf16x2(r.w);
                  f[0] += a.x; f[1] += a.y; f[2] += b.x; f[3] += b.y; f[4] += c2.x; f[5] += c2.y; f[6] += d.x; f[7] += d.y;
                }
                uint4 o;
                o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                *reinterpret_cast<uint4*>(crow + cb + g * 8) = o;
              }
            } else {
#pragma unroll
              for (int g = 0; g < 8; ++g) {  // 4 fp32 columns (16 B) per store
                if (cb + g * 4 >= p.N) break;
                float4 o;
                o.x = __uint_as_float(v[g * 4 + 0]); o.y = __uint_as_float(v[g * 4 + 1]);
                o.z = __uint_as_float(v[g * 4 + 2]); o.w = __uint_as_float(v[g * 4 + 3]);
                if constexpr (EPI == EPI_RESIDUAL) {
                  uint2 r = *reinterpret_cast<const uint2*>(rrow + cb + g * 4);
                  float2 a = unpack_bf16x2(r.x), b = unpack_bf16x2(r.y);
                  o.x += a.x; o.y += a.y; o.z += b.x; o.w += b.y;
                } else if constexpr (EPI == EPI_ACCUM) {
                  float4 r = *reinterpret_cast<const float4*>(crow + cb + g * 4);
                  o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                }
                *reinterpret_cast<float4*>(crow + cb + g * 4) = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <bool A_MN, bool B_MN, int EPI, typename OutT, bool BATCH = false>
static int launch_one(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05<A_MN, B_MN, EPI, OutT, BATCH>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  long long tiles = (long long)p.m_tiles * p.n_tiles * (BATCH ? (long long)p.nb0 * p.nb1 : 1);
  int grid = tiles < sm_count() ? (int)tiles : sm_count();
  kern<<<grid, GEMM_THREADS, SMEM_BYTES, stream>>>(tmA, tmB, p);
  return (int)cudaGetLastError();
}

template <bool A_MN, bool B_MN>
static int dispatch_epi(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int epi, int out_fp32,
                        cudaStream_t s) {
  if (out_fp32) {
    if (epi == EPI_STORE) return launch_one<A_MN, B_MN, EPI_STORE, float>(a, b, p, s);
    if (epi == EPI_RESIDUAL) return launch_one<A_MN, B_MN, EPI_RESIDUAL, float>(a, b, p, s);
    return launch_one<A_MN, B_MN, EPI_ACCUM, float>(a, b, p, s);
  }
  if (epi == EPI_STORE) return launch_one<A_MN, B_MN, EPI_STORE, __nv_bfloat16>(a, b, p, s);
  if (epi == EPI_RESIDUAL) return launch_one<A_MN, B_MN, EPI_RESIDUAL, __nv_bfloat16>(a, b, p, s);
  return launch_one<A_MN, B_MN, EPI_ACCUM, __nv_bfloat16>(a, b, p, s);
}

// batched dispatch: store / accumulate epilogues, bf16 or fp32 output
template <bool A_MN, bool B_MN>
static int dispatch_batched(const CUtensorMap& a, const CUtensorMap& b, const GemmParams& p, int epi, int out_fp32,
                            cudaStream_t s) {
  if (out_fp32) {
    if (epi == EPI_ACCUM) return launch_one<A_MN, B_MN, EPI_ACCUM, float, true>(a, b, p, s);
    return launch_one<A_MN, B_MN, EPI_STORE, float, true>(a, b, p, s);
  }
  if (epi == EPI_ACCUM) return launch_one<A_MN, B_MN, EPI_ACCUM, __nv_bfloat16, true>(a, b, p, s);
  return launch_one<A_MN, B_MN, EPI_STORE, __nv_bfloat16, true>(a, b, p, s);
}

inline int make_tmap_4d_bf16(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t nb0, uint64_t nb1,
                             uint64_t ld, uint64_t s0, uint64_t s1, uint32_t box_inner, uint32_t box_outer) {
  uint64_t dims[4] = {inner, outer, nb0, nb1};
  uint64_t strides[3] = {ld * 2, s0 * 2, s1 * 2};
  uint32_t box[4] = {box_inner, box_outer, 1, 1};
  return make_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace b200

// Batched C[b0,b1] = op(A[b0,b1]) op(B[b0,b1]) (+C): nb0 x nb1 problems of identical shape; sa*/sb*/sc* are the element
// strides of the two batch indices (each a multiple of 8 elements; use any valid stride when the count is 1).
extern "C" int b200_bgemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int lda, int ldb, int ldc,
                               int nb0, int nb1, long long sa0, long long sa1, long long sb0, long long sb1,
                               long long sc0, long long sc1, int a_mn, int b_mn, int epi, int out_fp32,
                               cudaStream_t stream) {
  using namespace b200;
  if (epi == EPI_RESIDUAL) return -1;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_4d_bf16(&tmA, A, K, M, nb0, nb1, lda, sa0, sa1, BK, BM);
  else       rc = make_tmap_4d_bf16(&tmA, A, M, K, nb0, nb1, lda, sa0, sa1, 64, BK);
  if (rc) return 1000 - rc;
  if (!b_mn) rc = make_tmap_4d_bf16(&tmB, B, K, N, nb0, nb1, ldb, sb0, sb1, BK, BN);
  else       rc = make_tmap_4d_bf16(&tmB, B, N, K, nb0, nb1, ldb, sb0, sb1, 64, BK);
  if (rc) return 2000 - rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldr = 0; p.C = C; p.R = nullptr;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.nb0 = nb0; p.nb1 = nb1; p.sc0 = sc0; p.sc1 = sc1;
  if (a_mn) {
    if (b_mn) return dispatch_batched<true, true>(tmA, tmB, p, epi, out_fp32, stream);
    return dispatch_batched<true, false>(tmA, tmB, p, epi, out_fp32, stream);
  }
  if (b_mn) return dispatch_batched<false, true>(tmA, tmB, p, epi, out_fp32, stream);
  return dispatch_batched<false, false>(tmA, tmB, p, epi, out_fp32, stream);
}

// C[M,N] = op(A) op(B) (+R | +C).  a_mn: A is stored [K,M] (M contiguous) else [M,K];  b_mn: B is stored [K,N]
// (N contiguous) else [N,K].  lda/ldb/ldc/ldr are row strides in elements.  Returns 0 or a CUDA error code.
extern "C" int b200_gemm_bf16(const void* A, const void* B, void* C, const void* R, int M, int N, int K, int lda,
                              int ldb, int ldc, int ldr, int a_mn, int b_mn, int epi, int out_fp32,
                              cudaStream_t stream) {
  using namespace b200;
  CUtensorMap tmA, tmB;
  int rc;
  if (!a_mn) rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM);
  else       rc = make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK);
  if (rc) return 1000 - rc;
  if (!b_mn) rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN);
  else       rc = make_tmap_2d_bf16(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK);
  if (rc) return 2000 - rc;
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.ldr = ldr; p.C = C; p.R = R;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.nb0 = p.nb1 = 1; p.sc0 = p.sc1 = 0;
  if (a_mn) {
    if (b_mn) return dispatch_epi<true, true>(tmA, tmB, p, epi, out_fp32, stream);
    return dispatch_epi<true, false>(tmA, tmB, p, epi, out_fp32, stream);
  }
  if (b_mn) return dispatch_epi<false, true>(tmA, tmB, p, epi, out_fp32, stream);
  return dispatch_epi<false, false>(tmA, tmB, p, epi, out_fp32, stream);
}
