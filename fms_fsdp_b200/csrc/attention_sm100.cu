// Causal GQA flash attention for sm_100a on tcgen05 tensor cores.
//
// Input is the fused, already-roped projection qkv [B*S, (H+2*KVH)*HD] (bf16); one 2-D TMA map with a
// {64 x 128|64}-element box serves Q, K and V.  All GEMMs run as tcgen05.mma with fp32 accumulators in
// TMEM; softmax / dS math runs in "row-owner" threads (thread t <-> TMEM lane t, so row reductions need
// no shuffles) that read S with tcgen05.ld and hand P / dS back to the tensor core through
// 128B-swizzled shared memory.
//
//  forward   CTA = (q tile of 128, head, batch).  S_j = Q K_j^T (double-buffered in TMEM so QK_{j+1}
//            overlaps softmax_j), online softmax with lazy rescale (O is only rescaled when the row max
//            grows by > 2^8), O += P_j V_j with V consumed as an MN-major operand.
//  backward  two kernels with one skeleton (templated):  MODE_DKDV: CTA = (kv tile 128, kv head, batch),
//            streams (Q_i, dO_i) tiles of 64 rows, computes S^T = K Q^T and dP^T = V dO^T, accumulates
//            dV += P^T dO and dK += dS^T Q in TMEM (GQA group summed in the accumulator, no atomics);
//            MODE_DQ: CTA = (q tile 128, head, batch), streams (K_j, V_j) tiles of 64, dQ += dS K.
//            The score GEMMs are recomputed in both kernels (7 GEMMs instead of 5) in exchange for a
//            deterministic, atomic-free dQ.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer / TMEM owner, warps 2-5 row owners.
#include "common.cuh"
#include "tensormap.h"

namespace b200 {

constexpr int ATT_THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

B200_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Write 8 consecutive bf16 (16 B) of row `row`, 16-byte chunk index `chunk16` (0..7) into a K-major
// SWIZZLE_128B tile whose rows are 128 B.
B200_DEVINL void st_swz128(uint8_t* tile, int row, int chunk16, uint4 v) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk16 ^ (row & 7)) << 4)) = v;
}

// ============================================================================================ forward
template <int HD>
struct FwdCfg {
  static constexpr int NCH = HD / 64;              // 64-column chunks of the head dim
  static constexpr int TILE_BYTES = 128 * HD * 2;  // a 128-row Q/K/V tile
  static constexpr int KV_STAGES = 2;
  static constexpr int P_BYTES = 128 * 128 * 2;
  static constexpr int SMEM = TILE_BYTES * (1 + 2 * KV_STAGES) + P_BYTES + 1024 + 256;
};

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                int S, int H, int KVH, float scale_log2, int n_qt) {
  using C = FwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::TILE_BYTES;
  uint8_t* sV = sK + C::KV_STAGES * C::TILE_BYTES;
  uint8_t* sP = sV + C::KV_STAGES * C::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::P_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2]
  uint64_t* p_full = bars + 11;
  uint64_t* pv_done = bars + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy (late) q tiles first
  const int qt = n_qt - 1 - blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (H / KVH);
  const int n_kv = qt + 1;  // causal: kv tiles 0..qt
  const int row0 = b * S + qt * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_S[2] = {tmem, tmem + 128};
  const uint32_t tmem_O = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, C::TILE_BYTES);
      for (int c = 0; c < C::NCH; ++c) tma_load_2d(sQ + c * 16384, &tm, q_full, h * HD + 64 * c, row0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int krow = b * S + j * 128;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], C::TILE_BYTES);
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(sK + st * C::TILE_BYTES + c * 16384, &tm, &k_full[st], (H + kvh) * HD + 64 * c, krow);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], C::TILE_BYTES);
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(sV + st * C::TILE_BYTES + c * 16384, &tm, &v_full[st], (H + KVH + kvh) * HD + 64 * c, krow);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
    auto issue_qk = [&](int j) {
      const int st = j & 1;
      mbar_wait(&k_full[st], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t qa = smem_u32(sQ), ka = smem_u32(sK + st * C::TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          umma_bf16_ss(tmem_S[j & 1], make_smem_desc(qa + off, 0, 1024), make_smem_desc(ka + off, 0, 1024), idesc_qk,
                       kk != 0);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j & 1;
      mbar_wait(p_full, j & 1);
      mbar_wait(&v_full[st], (j >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + st * C::TILE_BYTES);
#pragma unroll
        for (int t = 0; t < 8; ++t) {  // 16 kv rows per step
          const uint32_t poff = (t >> 2) * 16384 + (t & 3) * 32;
          umma_bf16_ss(tmem_O, make_smem_desc(pa + poff, 0, 1024), make_smem_desc(va + t * 2048, 16384, 1024),
                       idesc_pv, (j | t) != 0);
        }
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ row owners: online softmax, O rescale, epilogue
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;         // row within the q tile == TMEM lane
    const int q_idx = qt * 128 + r;       // position in the sequence
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tmem_S[j & 1] + lane_addr;
      const bool diag = (j == qt);
      const int kv0 = j * 128;
      // pass 1: row max (scaled log2 domain)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ts + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(v[i]) * scale_log2;
          const int kv = kv0 + c + i;
          if ((diag && kv > q_idx) || kv >= S) x = -INFINITY;
          mx = fmaxf(mx, x);
        }
      }
      const float m_new = fmaxf(m_ref, mx);
      // P buffer and O are free only once PV_{j-1} has retired
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
      tc_fence_after();
      const bool grow = (m_new - m_ref) > 8.f;  // lazy rescale threshold (values stay < 2^8 above the reference)
      if (j == 0) {
        m_ref = m_new;
      } else if (__any_sync(0xffffffffu, grow)) {
        const float f = exp2f(m_ref - m_new);
        l_sum *= f;
        m_ref = m_new;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_O + lane_addr + c, v);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w0[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            w1[i] = __float_as_uint(__uint_as_float(v[16 + i]) * f);
          }
          tmem_st_32x32b_x16(tmem_O + lane_addr + c, w0);
          tmem_st_32x32b_x16(tmem_O + lane_addr + c + 16, w1);
        }
        tmem_st_wait();
      }
      // pass 2: p = 2^(x - m_ref) -> bf16 -> swizzled smem (A operand of the PV GEMM)
#pragma unroll 1
      for (int c = 0; c < 128; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(ts + c, v);
        tmem_ld_wait();
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(v[i]) * scale_log2;
          const int kv = kv0 + c + i;
          if ((diag && kv > q_idx) || kv >= S) x = -INFINITY;
          p[i] = exp2f(x - m_ref);
          l_sum += p[i];
        }
        uint8_t* tile = sP + (c >> 6) * 16384;
        const int cb = (c & 63) >> 3;  // first 16-byte chunk of this group within the 128 B row
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(p[g * 8 + 0], p[g * 8 + 1]); u.y = pack_bf16x2(p[g * 8 + 2], p[g * 8 + 3]);
          u.z = pack_bf16x2(p[g * 8 + 4], p[g * 8 + 5]); u.w = pack_bf16x2(p[g * 8 + 6], p[g * 8 + 7]);
          st_swz128(tile, r, cb + g, u);
        }
      }
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16 ; lse
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.f / l_sum;
    if (q_idx < S) {
      __nv_bfloat16* orow = o + (static_cast<size_t>(b) * S + q_idx) * (H * HD) + h * HD;
#pragma unroll 1
      for (int c = 0; c < HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_O + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]) * inv_l, __uint_as_float(v[g * 8 + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]) * inv_l, __uint_as_float(v[g * 8 + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]) * inv_l, __uint_as_float(v[g * 8 + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]) * inv_l, __uint_as_float(v[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
        }
      }
      lse[(static_cast<size_t>(b) * H + h) * S + q_idx] = (m_ref + log2f(l_sum)) * LN2;
    } else {
      // still drain TMEM reads are not required; nothing to store
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ===================================================================================== forward v2
// Two 128-row Q tiles per CTA ("ping-pong"): the tiles share every K/V load, each has its own softmax
// warpgroup, and P never touches shared memory -- it is written back into the TMEM columns of S (two bf16 per
// 32-bit column, row per lane) and consumed by the PV GEMM through tcgen05.mma's TMEM-A operand form.  While one
// tile's softmax runs, the tensor core works on the other tile's QK^T / PV.
//   TMEM: S0/P0 [0,128)  S1/P1 [128,256)  O0 [256,256+HD)  O1 [384,384+HD)
//   smem: Q 2 x tile, K 2 stages, V 2 stages  (192 KiB at HD=128)
//   warps: 0 TMA, 1 MMA, 2-3 idle, 4-7 softmax tile 0, 8-11 softmax tile 1   (384 threads)
constexpr int ATT2_THREADS = 384;

template <int HD>
struct Fwd2Cfg {
  static constexpr int NCH = HD / 64;
  static constexpr int TILE_BYTES = 128 * HD * 2;
  static constexpr int SMEM = TILE_BYTES * (2 + 2 + 2) + 1024 + 256;
};

template <int HD>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tm, __nv_bfloat16* __restrict__ o, float* __restrict__ lse,
                 int S, int H, int KVH, float scale_log2, int n_pt) {
  using C = Fwd2Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                              // [2 tiles]
  uint8_t* sK = sQ + 2 * C::TILE_BYTES;            // [2 stages]
  uint8_t* sV = sK + 2 * C::TILE_BYTES;            // [2 stages]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * C::TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* k_empty = bars + 3;    // [2]
  uint64_t* v_full = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;    // [2]
  uint64_t* s_full = bars + 9;     // [2 tiles]
  uint64_t* p_full = bars + 11;    // [2 tiles]
  uint64_t* pv_done = bars + 13;   // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = n_pt - 1 - blockIdx.x;  // pair-tile index, heavy first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (H / KVH);
  const int row0 = b * S + pt * 256;
  // kv tiles (128 rows): tile t of the pair attends kv tiles 0 .. 2*pt + t (clipped to the sequence)
  const int n_kv_total = (S + 127) / 128;
  const int n_kv[2] = {min(2 * pt + 1, n_kv_total), (pt * 256 + 128 < S) ? min(2 * pt + 2, n_kv_total) : 0};
  const int n_it = max(n_kv[0], n_kv[1]);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 2 * C::TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(sQ + t * C::TILE_BYTES + c * 16384, &tm, q_full, h * HD + 64 * c, row0 + 128 * t);
      for (int j = 0; j < n_it; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int krow = b * S + j * 128;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], C::TILE_BYTES);
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(sK + st * C::TILE_BYTES + c * 16384, &tm, &k_full[st], (H + kvh) * HD + 64 * c, krow);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], C::TILE_BYTES);
        for (int c = 0; c < C::NCH; ++c)
          tma_load_2d(sV + st * C::TILE_BYTES + c * 16384, &tm, &v_full[st], (H + KVH + kvh) * HD + 64 * c, krow);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
    mbar_wait(q_full, 0);
    // Tensor-pipe issue order (the pipe executes in issue order, which is what makes the S/P aliasing safe):
    //   QK0_0 QK1_0 | PV0_0 QK0_1 PV1_0 QK1_1 | PV0_1 QK0_2 PV1_1 QK1_2 | ...
    // QK_t(j+1) overwrites S_t only after PV_t(j) -- which reads P_t from the same TMEM columns -- was issued.
    // every MMA batch is issued by ONE elected lane of the converged warp (elect.sync): back-to-back UTCHMMA with
    // descriptors = base + constant, no per-MMA ELECT loop / descriptor rebuild
    const uint64_t qd0 = make_smem_desc(smem_u32(sQ), 0, 1024), kd0 = make_smem_desc(smem_u32(sK), 0, 1024);
    const uint64_t vd0 = make_smem_desc(smem_u32(sV), 16384, 1024);
    constexpr uint64_t TILE16 = C::TILE_BYTES >> 4;   // second tile / stage: start-address field + TILE_BYTES/16
    auto qk = [&](int t, int j) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
        umma_bf16_ss(tmem + t * 128, qd0 + t * TILE16 + off, kd0 + (j & 1) * TILE16 + off, idesc_qk, kk != 0);
      }
      umma_commit(&s_full[t]);
    };
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      if (0 < n_kv[0]) qk(0, 0);
      if (0 < n_kv[1]) qk(1, 0);
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_it; ++j) {
      const int st = j & 1;
      const bool next = j + 1 < n_it;
      mbar_wait(&v_full[st], (j >> 1) & 1);
      if (next) mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
      for (int t = 0; t < 2; ++t) {
        if (j < n_kv[t]) {
          mbar_wait(&p_full[t], j & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 8; ++k)  // 16 kv rows per step; P: 8 packed TMEM columns per step
              umma_bf16_ts(tmem + 256 + t * 128, tmem + t * 128 + k * 8, vd0 + st * TILE16 + ((k * 2048) >> 4), idesc_pv,
                           (j | k) != 0);
            umma_commit(&pv_done[t]);
            if (next && j + 1 < n_kv[t]) qk(t, j + 1);
          }
          __syncwarp();
        } else if (next && j + 1 < n_kv[t]) {
          if (elect_one()) qk(t, j + 1);
          __syncwarp();
        }
      }
      if (elect_one()) {
        umma_commit(&v_empty[st]);
        if (next) umma_commit(&k_empty[(j + 1) & 1]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------ softmax warpgroup of tile t (row owner = TMEM lane)
    const int t = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int q_idx = pt * 256 + t * 128 + r;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const uint32_t tS = tmem + t * 128 + lane_addr;
    const uint32_t tO = tmem + 256 + t * 128 + lane_addr;
    const int nk = n_kv[t];
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < nk; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      const bool diag = (j == 2 * pt + t);
      const int kv0 = j * 128;
      uint32_t v[128];
      tmem_ld_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_32x32b_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_32x32b_x32(tS + 64, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld_32x32b_x32(tS + 96, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_ld_wait();
      // row max on the RAW scores (scale > 0 commutes with max); masking only on the diagonal / ragged last tile
      if (diag || kv0 + 128 > S) {
#pragma unroll
        for (int i = 0; i < 128; ++i) {
          const int kv = kv0 + i;
          if (kv > q_idx || kv >= S) v[i] = 0xff800000u;  // -inf
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(v[i]));
        mx1 = fmaxf(mx1, __uint_as_float(v[i + 1]));
      }
      const float mx = fmaxf(mx0, mx1) * scale_log2;
      const float m_new = fmaxf(m_ref, mx);
      const bool grow = (m_new - m_ref) > 8.f;
      if (j == 0) {
        m_ref = m_new;
      } else if (__any_sync(0xffffffffu, grow)) {
        mbar_wait(&pv_done[t], (j - 1) & 1);   // O_t is stable only once PV_t of the previous tile retired
        tc_fence_after();
        const float f = exp2f(m_ref - m_new);
        l_sum *= f;
        m_ref = m_new;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tO + c, ov);
          tmem_ld_wait();
          uint32_t w0[16], w1[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            w0[i] = __float_as_uint(__uint_as_float(ov[i]) * f);
            w1[i] = __float_as_uint(__uint_as_float(ov[16 + i]) * f);
          }
          tmem_st_32x32b_x16(tO + c, w0);
          tmem_st_32x32b_x16(tO + c + 16, w1);
        }
      }
      // P = 2^(x - m_ref) packed two bf16 per column, written over S's own columns (in order => no hazard)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = exp2f(fmaf(__uint_as_float(v[c * 32 + 2 * i]), scale_log2, -m_ref));
          const float p1 = exp2f(fmaf(__uint_as_float(v[c * 32 + 2 * i + 1]), scale_log2, -m_ref));
          l_sum += p0 + p1;
          w[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x32b_x16(tS + c * 16, w);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[t]);
    }
    if (nk > 0) {
      mbar_wait(&pv_done[t], (nk - 1) & 1);
      tc_fence_after();
      if (q_idx < S) {
        const float inv_l = 1.f / l_sum;
        __nv_bfloat16* orow = o + (static_cast<size_t>(b) * S + q_idx) * (H * HD) + h * HD;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t ov[32];
          tmem_ld_32x32b_x32(tO + c, ov);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(ov[g * 8 + 0]) * inv_l, __uint_as_float(ov[g * 8 + 1]) * inv_l);
            u.y = pack_bf16x2(__uint_as_float(ov[g * 8 + 2]) * inv_l, __uint_as_float(ov[g * 8 + 3]) * inv_l);
            u.z = pack_bf16x2(__uint_as_float(ov[g * 8 + 4]) * inv_l, __uint_as_float(ov[g * 8 + 5]) * inv_l);
            u.w = pack_bf16x2(__uint_as_float(ov[g * 8 + 6]) * inv_l, __uint_as_float(ov[g * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow + c + g * 8) = u;
          }
        }
        lse[(static_cast<size_t>(b) * H + h) * S + q_idx] = (m_ref + log2f(l_sum)) * LN2;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ===================================================================================== backward prep
// delta[b,h,s] = sum_d dO[b,s,h,d] * O[b,s,h,d]      (one warp per (row, head))
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ o,
                                  float* __restrict__ delta, const float* __restrict__ lse, float* __restrict__ lse2,
                                  int B, int S, int H, int HD, int ld) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= B * S * H) return;
  const int h = gw % H;
  const size_t bs = gw / H;
  const __nv_bfloat16* a = dout + (bs * H + h) * HD;
  const __nv_bfloat16* c = o + (bs * H + h) * HD;
  float s = 0.f;
  for (int d = lane * 2; d < HD; d += 64) {
    float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(a + d));
    float2 y = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(c + d));
    s += x.x * y.x + x.y * y.y;
  }
  s = warp_sum(s);
  if (lane == 0) {
    const int b = (int)(bs / S), sq = (int)(bs % S);
    delta[((size_t)b * H + h) * ld + sq] = s;
    if (lse2) lse2[((size_t)b * H + h) * ld + sq] = lse[((size_t)b * H + h) * S + sq] * LOG2E;  // log2-domain copy
  }
}

// ============================================================================================ backward
// Skeleton shared by both kernels:
//   resident pair (X1, X2): 128 rows x HD        streamed pair (Y1, Y2): 64 rows x HD, 2 stages
//   T1 = X1 Y1^T, T2 = X2 Y2^T   (128 x 64 fp32 in TMEM, double buffered)
//   row owners turn (T1, T2) into bf16 tiles W1 (and W2) of shape [128 x 64] in smem
//   DKDV: X=(K_j, V_j), Y=(Q_i, dO_i): W1 = P^T, W2 = dS^T; acc1(dV) += W1 Y2, acc2(dK) += W2 Y1
//   DQ  : X=(Q_i, dO_i), Y=(K_j, V_j): W2 = dS;             acc2(dQ) += W2 Y1
enum { MODE_DKDV = 0, MODE_DQ = 1 };

template <int HD>
struct BwdCfg {
  static constexpr int NCH = HD / 64;
  static constexpr int X_BYTES = 128 * HD * 2;   // one resident operand
  static constexpr int Y_BYTES = 64 * HD * 2;    // one streamed operand
  static constexpr int Y_CHUNK = 64 * 128;       // bytes of one 64-row x 64-col chunk
  static constexpr int W_BYTES = 128 * 64 * 2;
  static constexpr int STAGES = 2;
  static constexpr int SMEM = 2 * X_BYTES + STAGES * 2 * Y_BYTES + 2 * W_BYTES + 64 * 2 * 4 * 2 + 1024 + 256;
};

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                const __grid_constant__ CUtensorMap tm_do128, const __grid_constant__ CUtensorMap tm_do64,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                int S, int H, int KVH, float scale, int n_t128) {
  using C = BwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem;
  uint8_t* sX2 = sX1 + C::X_BYTES;
  uint8_t* sY = sX2 + C::X_BYTES;                       // [stage][Y1 | Y2]
  uint8_t* sW1 = sY + C::STAGES * 2 * C::Y_BYTES;
  uint8_t* sW2 = sW1 + C::W_BYTES;
  float* sStat = reinterpret_cast<float*>(sW2 + C::W_BYTES);  // [2 buffers][lse2 64 | delta 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 64 * 2 * 4 * 2);
  uint64_t* x_full = bars;
  uint64_t* y_full = bars + 1;     // [2]
  uint64_t* y_empty = bars + 3;    // [2]
  uint64_t* t_full = bars + 5;     // [2]
  uint64_t* w_full = bars + 7;
  uint64_t* acc_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = H / KVH;
  const int b = blockIdx.z;
  const float scale_log2 = scale * LOG2E;

  // iteration space of the streamed (64-row) tiles
  int t128, head_lo, head_n, kvh;
  if constexpr (MODE == MODE_DKDV) {
    t128 = blockIdx.x;                 // kv tile; small index = most work, scheduled first
    kvh = blockIdx.y;
    head_lo = kvh * G;
    head_n = G;
  } else {
    t128 = n_t128 - 1 - blockIdx.x;    // q tile; large index = most work
    head_lo = blockIdx.y;
    head_n = 1;
    kvh = blockIdx.y / G;
  }
  const int n64 = (S + 63) / 64;
  // DKDV: q tiles i64 in [2*t128, n64) ; DQ: kv tiles j64 in [0, 2*t128+2) clipped
  const int s_lo = (MODE == MODE_DKDV) ? 2 * t128 : 0;
  const int s_hi = (MODE == MODE_DKDV) ? n64 : min(n64, 2 * t128 + 2);
  const int per_head = s_hi - s_lo;
  const int n_iter = per_head * head_n;
  const int xrow0 = b * S + t128 * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128); tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do128); tma_prefetch_desc(&tm_do64);
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); mbar_init(&t_full[i], 1);
    }
    mbar_init(w_full, 4);
    mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: T1[0] 0..63, T2[0] 64..127, T1[1] 128..191, T2[1] 192..255, acc1 256.., acc2 384..
  const uint32_t tmem_acc1 = tmem + 256, tmem_acc2 = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      // resident operands.  DKDV: per-CTA constant (K_j, V_j).  DQ: (Q_i, dO_i) of this head.
      if constexpr (MODE == MODE_DKDV) {
        mbar_arrive_expect_tx(x_full, 2 * C::X_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          tma_load_2d(sX1 + c * 16384, &tm_qkv128, x_full, (H + kvh) * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * 16384, &tm_qkv128, x_full, (H + KVH + kvh) * HD + 64 * c, xrow0);
        }
      } else {
        mbar_arrive_expect_tx(x_full, 2 * C::X_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          tma_load_2d(sX1 + c * 16384, &tm_qkv128, x_full, head_lo * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * 16384, &tm_do128, x_full, head_lo * HD + 64 * c, xrow0);
        }
      }
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        const int hh = head_lo + it / per_head;
        const int t64 = s_lo + it % per_head;
        const int yrow = b * S + t64 * 64;
        uint8_t* y1 = sY + st * 2 * C::Y_BYTES;
        uint8_t* y2 = y1 + C::Y_BYTES;
        mbar_wait(&y_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&y_full[st], 2 * C::Y_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) {
            tma_load_2d(y1 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], hh * HD + 64 * c, yrow);   // Q_i
            tma_load_2d(y2 + c * C::Y_CHUNK, &tm_do64, &y_full[st], hh * HD + 64 * c, yrow);    // dO_i
          } else {
            tma_load_2d(y1 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], (H + kvh) * HD + 64 * c, yrow);        // K_j
            tma_load_2d(y2 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], (H + KVH + kvh) * HD + 64 * c, yrow);  // V_j
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, false, false);   // scores: both K-major over HD
    constexpr uint32_t idesc_a = make_idesc_bf16(128, HD, false, true);    // accumulate: A K-major, B MN-major
    auto issue_scores = [&](int it) {
      const int st = it & 1;
      mbar_wait(&y_full[st], (it >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2);
        const uint32_t y1 = smem_u32(sY + st * 2 * C::Y_BYTES), y2 = y1 + C::Y_BYTES;
        const uint32_t t1 = tmem + (it & 1) * 128, t2 = t1 + 64;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32;
          const uint32_t yo = (kk >> 2) * C::Y_CHUNK + (kk & 3) * 32;
          umma_bf16_ss(t1, make_smem_desc(x1 + xo, 0, 1024), make_smem_desc(y1 + yo, 0, 1024), idesc_t, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32;
          const uint32_t yo = (kk >> 2) * C::Y_CHUNK + (kk & 3) * 32;
          umma_bf16_ss(t2, make_smem_desc(x2 + xo, 0, 1024), make_smem_desc(y2 + yo, 0, 1024), idesc_t, kk != 0);
        }
        umma_commit(&t_full[it & 1]);
      }
      __syncwarp();
    };
    mbar_wait(x_full, 0);
    if (n_iter > 0) issue_scores(0);
    for (int it = 0; it < n_iter; ++it) {
      if (it + 1 < n_iter) issue_scores(it + 1);
      const int st = it & 1;
      mbar_wait(w_full, it & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t y1 = smem_u32(sY + st * 2 * C::Y_BYTES), y2 = y1 + C::Y_BYTES;
        const uint32_t w1 = smem_u32(sW1), w2 = smem_u32(sW2);
        // reduction over the 64 streamed rows: 4 steps of 16; Y as MN-major B: LBO = chunk stride, SBO = 1024
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          if constexpr (MODE == MODE_DKDV) {
            umma_bf16_ss(tmem_acc1, make_smem_desc(w1 + t * 32, 0, 1024),
                         make_smem_desc(y2 + t * 2048, C::Y_CHUNK, 1024), idesc_a, (it | t) != 0);   // dV += P^T dO
          }
          umma_bf16_ss(tmem_acc2, make_smem_desc(w2 + t * 32, 0, 1024),
                       make_smem_desc(y1 + t * 2048, C::Y_CHUNK, 1024), idesc_a, (it | t) != 0);     // dK += dS^T Q | dQ += dS K
        }
        umma_commit(&y_empty[st]);
        umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------- row owners
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int tid128 = (warp - 2) * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const int x_idx = t128 * 128 + r;  // sequence position of this thread's resident row
    float row_lse2 = 0.f, row_delta = 0.f;
    for (int it = 0; it < n_iter; ++it) {
      const int hh = head_lo + it / per_head;
      const int t64 = s_lo + it % per_head;
      const int y0 = t64 * 64;
      float* stat = sStat + (it & 1) * 128;
      if constexpr (MODE == MODE_DKDV) {
        // per-column statistics of the streamed q rows
        const int qi = y0 + (tid128 & 63);
        const size_t sidx = ((size_t)b * H + hh) * S + min(qi, S - 1);
        stat[tid128] = (tid128 < 64) ? lse[sidx] * LOG2E : delta[sidx];
        named_bar_sync(1, 128);
      } else if (it == 0) {
        const size_t sidx = ((size_t)b * H + hh) * S + min(x_idx, S - 1);
        row_lse2 = lse[sidx] * LOG2E;
        row_delta = delta[sidx];
      }
      mbar_wait(&t_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      // W tiles are free once the accumulate GEMMs of the previous iteration retired
      if (it > 0) mbar_wait(acc_done, (it - 1) & 1);
      const uint32_t t1 = tmem + (it & 1) * 128 + lane_addr, t2 = t1 + 64;
#pragma unroll 1
      for (int c = 0; c < 64; c += 32) {
        uint32_t a[32], d[32];
        tmem_ld_32x32b_x32(t1 + c, a);
        tmem_ld_32x32b_x32(t2 + c, d);
        tmem_ld_wait();
        float p[32], ds[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int y_idx = y0 + c + i;
          float l2, dl;
          bool masked;
          if constexpr (MODE == MODE_DKDV) {
            l2 = stat[c + i]; dl = stat[64 + c + i];
            masked = (x_idx > y_idx) || (y_idx >= S) || (x_idx >= S);     // kv > q
          } else {
            l2 = row_lse2; dl = row_delta;
            masked = (y_idx > x_idx) || (y_idx >= S) || (x_idx >= S);
          }
          const float pv = masked ? 0.f : exp2f(__uint_as_float(a[i]) * scale_log2 - l2);
          p[i] = pv;
          ds[i] = pv * (__uint_as_float(d[i]) - dl) * scale;
        }
        const int cb = c >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          if constexpr (MODE == MODE_DKDV) {
            u.x = pack_bf16x2(p[g * 8 + 0], p[g * 8 + 1]); u.y = pack_bf16x2(p[g * 8 + 2], p[g * 8 + 3]);
            u.z = pack_bf16x2(p[g * 8 + 4], p[g * 8 + 5]); u.w = pack_bf16x2(p[g * 8 + 6], p[g * 8 + 7]);
            st_swz128(sW1, r, cb + g, u);
          }
          u.x = pack_bf16x2(ds[g * 8 + 0], ds[g * 8 + 1]); u.y = pack_bf16x2(ds[g * 8 + 2], ds[g * 8 + 3]);
          u.z = pack_bf16x2(ds[g * 8 + 4], ds[g * 8 + 5]); u.w = pack_bf16x2(ds[g * 8 + 6], ds[g * 8 + 7]);
          st_swz128(sW2, r, cb + g, u);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(w_full);
    }
    // epilogue: accumulators -> bf16 -> dqkv sections
    if (n_iter > 0) mbar_wait(acc_done, (n_iter - 1) & 1);
    tc_fence_after();
    if (x_idx < S) {
      const size_t row = (size_t)b * S + x_idx;
      const int W = (H + 2 * KVH) * HD;
      auto store_acc = [&](uint32_t tacc, int col0) {
        __nv_bfloat16* dst = dqkv + row * W + col0;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t v[32];
          if (n_iter > 0) {
            tmem_ld_32x32b_x32(tacc + lane_addr + c, v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0u;
          }
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c + g * 8) = u;
          }
        }
      };
      if constexpr (MODE == MODE_DKDV) {
        store_acc(tmem_acc2, (H + kvh) * HD);         // dK
        store_acc(tmem_acc1, (H + KVH + kvh) * HD);   // dV
      } else {
        store_acc(tmem_acc2, head_lo * HD);           // dQ
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static int g_attn_fwd_version = 2;
static int g_attn_bwd_version = 3;

// ============================================================================================ backward v2
// Same math / operand layouts as attn_bwd_kernel, re-pipelined: two row-owner warpgroups alternate iterations
// (WG p owns T[p] and its own P^T/dS^T smem tiles), the streamed (Q,dO)/(K,V) ring is 3 deep and the score GEMMs
// run two iterations ahead of the accumulate GEMMs, so neither the tensor pipe nor the row owners wait on a
// single-buffered hand-off.  384 threads: warp 0 TMA, warp 1 MMA, warps 4-7 WG0, warps 8-11 WG1.
B200_DEVINL void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int HD>
struct Bwd2Cfg {
  static constexpr int NCH = HD / 64;
  static constexpr int X_BYTES = 128 * HD * 2;
  static constexpr int Y_BYTES = 64 * HD * 2;
  static constexpr int Y_CHUNK = 64 * 128;
  static constexpr int W_BYTES = 128 * 64 * 2;
  static constexpr int STAGES = 3;
  static constexpr int STAT_BYTES = 128 * 4;   // per stage: lse2[64] | delta[64] of the streamed q rows (DKDV)
  static constexpr int SMEM = 2 * X_BYTES + STAGES * 2 * Y_BYTES + STAGES * STAT_BYTES + 1024 + 256;
};

// one 32-column chunk of the row owner's work: (S or S^T, dP or dP^T) fp32 -> P, dS as packed bf16 pairs written
// back into TENSOR MEMORY over the first 16 of the 32 columns just read (row per lane): they become the TMEM-A
// operands of the accumulate GEMMs, so neither P nor dS ever touches shared memory.
template <int MODE, bool MASK>
B200_DEVINL void bwd_chunk(const uint32_t (&a)[32], const uint32_t (&d)[32], uint32_t tP, uint32_t tdS, int c,
                           int x_idx, int y0, int S, const float* stat, float row_lse2, float row_delta,
                           float scale_log2, float scale) {
  uint32_t pk[16], dk[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float pv[2], dv[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float l2, dl;
      if constexpr (MODE == MODE_DKDV) { l2 = stat[c + i + e]; dl = stat[64 + c + i + e]; }
      else { l2 = row_lse2; dl = row_delta; }
      float p = exp2f(fmaf(__uint_as_float(a[i + e]), scale_log2, -l2));
      if constexpr (MASK) {
        const int y_idx = y0 + c + i + e;
        const bool masked = (MODE == MODE_DKDV) ? ((x_idx > y_idx) || (y_idx >= S) || (x_idx >= S))
                                                : ((y_idx > x_idx) || (y_idx >= S) || (x_idx >= S));
        if (masked) p = 0.f;
      }
      pv[e] = p;
      dv[e] = p * (__uint_as_float(d[i + e]) - dl) * scale;
    }
    pk[i >> 1] = pack_bf16x2(pv[0], pv[1]);
    dk[i >> 1] = pack_bf16x2(dv[0], dv[1]);
  }
  if constexpr (MODE == MODE_DKDV) tmem_st_32x32b_x16(tP, pk);
  tmem_st_32x32b_x16(tdS, dk);
}

#ifdef B200_ATTN_TRACE
// timeline instrumentation (scripts/attn_trace.cu only): clock64 stamps of one CTA, [iteration][16 slots]
__device__ long long* g_attn_trace = nullptr;
__device__ int g_attn_trace_cta = 0;
__device__ int g_attn_trace_mode = 0;
#define ATRACE_INIT()                                                                                      \
  long long* const atrace_ptr = (MODE == g_attn_trace_mode && (int)blockIdx.x == g_attn_trace_cta &&      \
                                 blockIdx.y == 0 && blockIdx.z == 0) ? g_attn_trace : nullptr
#define ATRACE(it, slot)                                                                                   \
  do {                                                                                                     \
    if (atrace_ptr && (it) < 64) atrace_ptr[(it) * 16 + (slot)] = clock64();                               \
  } while (0)
#else
#define ATRACE_INIT() do {} while (0)
#define ATRACE(it, slot) do {} while (0)
#endif

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                 const __grid_constant__ CUtensorMap tm_do128, const __grid_constant__ CUtensorMap tm_do64,
                 const float* __restrict__ lse2g, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                 int S, int H, int KVH, float scale, int n_t128, int ld) {
  using C = Bwd2Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem;
  uint8_t* sX2 = sX1 + C::X_BYTES;
  uint8_t* sY = sX2 + C::X_BYTES;                          // [stage][Y1 | Y2]
  float* sStat = reinterpret_cast<float*>(sY + C::STAGES * 2 * C::Y_BYTES);   // [stage][lse2 64 | delta 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + C::STAGES * C::STAT_BYTES);
  uint64_t* x_full = bars;
  uint64_t* y_full = bars + 1;     // [3]
  uint64_t* y_empty = bars + 4;    // [3]
  uint64_t* t_full = bars + 7;     // [2]
  uint64_t* w_full = bars + 9;     // [2]
  uint64_t* acc_done = bars + 11;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = H / KVH;
  const int b = blockIdx.z;
  const float scale_log2 = scale * LOG2E;
  int t128, head_lo, head_n, kvh;
  if constexpr (MODE == MODE_DKDV) {
    t128 = blockIdx.x; kvh = blockIdx.y; head_lo = kvh * G; head_n = G;
  } else {
    t128 = n_t128 - 1 - blockIdx.x; head_lo = blockIdx.y; head_n = 1; kvh = blockIdx.y / G;
  }
  const int n64 = (S + 63) / 64;
  const int s_lo = (MODE == MODE_DKDV) ? 2 * t128 : 0;
  const int s_hi = (MODE == MODE_DKDV) ? n64 : min(n64, 2 * t128 + 2);
  const int per_head = s_hi - s_lo;
  const int n_iter = per_head * head_n;
  const int xrow0 = b * S + t128 * 128;
  constexpr uint32_t Y_TX = 2 * C::Y_BYTES + (MODE == MODE_DKDV ? C::STAT_BYTES : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128); tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do128); tma_prefetch_desc(&tm_do64);
    mbar_init(x_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&y_full[i], 1); mbar_init(&y_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&w_full[i], 8); mbar_init(&acc_done[i], 1); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_acc1 = tmem + 256, tmem_acc2 = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, 2 * C::X_BYTES);
      for (int c = 0; c < C::NCH; ++c) {
        if constexpr (MODE == MODE_DKDV) {
          tma_load_2d(sX1 + c * 16384, &tm_qkv128, x_full, (H + kvh) * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * 16384, &tm_qkv128, x_full, (H + KVH + kvh) * HD + 64 * c, xrow0);
        } else {
          tma_load_2d(sX1 + c * 16384, &tm_qkv128, x_full, head_lo * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * 16384, &tm_do128, x_full, head_lo * HD + 64 * c, xrow0);
        }
      }
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < n_iter; ++it) {
        const int hh = head_lo + it / per_head;
        const int t64 = s_lo + it % per_head;
        const int yrow = b * S + t64 * 64;
        uint8_t* y1 = sY + st * 2 * C::Y_BYTES;
        uint8_t* y2 = y1 + C::Y_BYTES;
        mbar_wait(&y_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&y_full[st], Y_TX);
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) {
            tma_load_2d(y1 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], hh * HD + 64 * c, yrow);
            tma_load_2d(y2 + c * C::Y_CHUNK, &tm_do64, &y_full[st], hh * HD + 64 * c, yrow);
          } else {
            tma_load_2d(y1 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], (H + kvh) * HD + 64 * c, yrow);
            tma_load_2d(y2 + c * C::Y_CHUNK, &tm_qkv64, &y_full[st], (H + KVH + kvh) * HD + 64 * c, yrow);
          }
        }
        if constexpr (MODE == MODE_DKDV) {
          // per-column softmax statistics of the streamed q rows ride on the same barrier (two 256 B bulk copies)
          const size_t so = ((size_t)b * H + hh) * ld + (size_t)t64 * 64;
          bulk_load_1d(sStat + st * 128, lse2g + so, 256, &y_full[st]);
          bulk_load_1d(sStat + st * 128 + 64, delta + so, 256, &y_full[st]);
        }
        if (++st == 3) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, HD, false, true);
    auto issue_scores = [&](int it) {
      const int st = it % 3;
      mbar_wait(&y_full[st], (it / 3) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2);
        const uint32_t y1 = smem_u32(sY + st * 2 * C::Y_BYTES), y2 = y1 + C::Y_BYTES;
        const uint32_t t1 = tmem + (it & 1) * 128, t2 = t1 + 64;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32;
          const uint32_t yo = (kk >> 2) * C::Y_CHUNK + (kk & 3) * 32;
          umma_bf16_ss(t1, make_smem_desc(x1 + xo, 0, 1024), make_smem_desc(y1 + yo, 0, 1024), idesc_t, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t xo = (kk >> 2) * 16384 + (kk & 3) * 32;
          const uint32_t yo = (kk >> 2) * C::Y_CHUNK + (kk & 3) * 32;
          umma_bf16_ss(t2, make_smem_desc(x2 + xo, 0, 1024), make_smem_desc(y2 + yo, 0, 1024), idesc_t, kk != 0);
        }
        umma_commit(&t_full[it & 1]);
      }
      __syncwarp();
    };
    mbar_wait(x_full, 0);
    if (n_iter > 0) issue_scores(0);
    for (int it = 0; it < n_iter; ++it) {
      const int p = it & 1, st = it % 3;
      if (it + 1 < n_iter) issue_scores(it + 1);   // T[(it+1)&1] was drained before w_full(it-1) was signalled
      mbar_wait(&w_full[p], (it >> 1) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t y1 = smem_u32(sY + st * 2 * C::Y_BYTES), y2 = y1 + C::Y_BYTES;
        // P^T / dS^T (or dS) sit in TMEM over the score columns: warpgroup g wrote the 16 packed columns of its
        // 32-column half at column offset 32*g  ->  K-step t (16 streamed rows) = 8 columns at 32*(t/2) + 8*(t%2)
        const uint32_t tP = tmem + p * 128, tdS = tP + 64;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t co = 32 * (t >> 1) + 8 * (t & 1);
          if constexpr (MODE == MODE_DKDV)
            umma_bf16_ts(tmem_acc1, tP + co, make_smem_desc(y2 + t * 2048, C::Y_CHUNK, 1024), idesc_a, (it | t) != 0);
          umma_bf16_ts(tmem_acc2, tdS + co, make_smem_desc(y1 + t * 2048, C::Y_CHUNK, 1024), idesc_a, (it | t) != 0);
        }
        umma_commit(&y_empty[st]);
        umma_commit(&acc_done[p]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // 8 row-owner warps; warpgroup g owns columns [32g, 32g+32) of EVERY iteration's 128 x 64 score tiles
    const int g = (warp - 4) >> 2;
    const int p = g;                          // (epilogue split below)
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const int x_idx = t128 * 128 + r;
    float row_lse2 = 0.f, row_delta = 0.f;
    if constexpr (MODE == MODE_DQ) {
      const size_t sidx = ((size_t)b * H + head_lo) * ld + min(x_idx, S - 1);
      row_lse2 = lse2g[sidx];
      row_delta = delta[sidx];
    }
    for (int it = 0; it < n_iter; ++it) {
      const int par = it & 1;
      const int t64 = s_lo + it % per_head;
      const int y0 = t64 * 64;
      const int st = it % 3;
      const float* stat = sStat + st * 128;
      if constexpr (MODE == MODE_DKDV) mbar_wait(&y_full[st], (it / 3) & 1);   // the stats landed (TMA -> generic visibility)
      mbar_wait(&t_full[par], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t t1 = tmem + par * 128 + lane_addr + 32 * g, t2 = t1 + 64;
      uint32_t a0[32], d0[32];
      tmem_ld_32x32b_x32(t1, a0);
      tmem_ld_32x32b_x32(t2, d0);
      tmem_ld_wait();
      // only tiles touching the causal diagonal or the sequence end need per-element masking
      const bool need_mask = (y0 < t128 * 128 + 128 && y0 + 64 > t128 * 128) || (y0 + 64 > S) || (t128 * 128 + 128 > S);
      if (need_mask)
        bwd_chunk<MODE, true>(a0, d0, t1, t2, 32 * g, x_idx, y0, S, stat, row_lse2, row_delta, scale_log2, scale);
      else
        bwd_chunk<MODE, false>(a0, d0, t1, t2, 32 * g, x_idx, y0, S, stat, row_lse2, row_delta, scale_log2, scale);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&w_full[par]);
    }
    // epilogue: WG0 stores acc2 (dK | dQ), WG1 stores acc1 (dV); in DQ mode the two groups split acc2's columns
    if (n_iter > 0) mbar_wait(&acc_done[(n_iter - 1) & 1], ((n_iter - 1) >> 1) & 1);
    tc_fence_after();
    if (x_idx < S) {
      const size_t row = (size_t)b * S + x_idx;
      const int Wd = (H + 2 * KVH) * HD;
      auto store_acc = [&](uint32_t tacc, int col0, int c_lo, int c_hi) {
        __nv_bfloat16* dst = dqkv + row * Wd + col0;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tacc + lane_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
            u.y = pack_bf16x2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
            u.z = pack_bf16x2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
            u.w = pack_bf16x2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
            *reinterpret_cast<uint4*>(dst + c + g * 8) = u;
          }
        }
      };
      if constexpr (MODE == MODE_DKDV) {
        if (p == 0) store_acc(tmem_acc2, (H + kvh) * HD, 0, HD);
        else        store_acc(tmem_acc1, (H + KVH + kvh) * HD, 0, HD);
      } else {
        if (p == 0) store_acc(tmem_acc2, head_lo * HD, 0, HD / 2);
        else        store_acc(tmem_acc2, head_lo * HD, HD / 2, HD);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ================================================================================ backward, version 3
// Measured on the v2 kernel (scripts/attn_trace.cu, profiles/attn_bwd2_pipeline_trace_r1.txt): the tensor pipe idled
// ~60% of every iteration because the softmax-gradient warps (TMEM load -> exp2 -> pack -> TMEM store, ~1.8k cycles)
// sat between "scores done" and "accumulate GEMMs may start" with only ONE other GEMM pair to overlap with, and the
// N=64 score MMAs run at 48 instead of 32 cycles.  v3 streams 128-row tiles (all MMAs N=128: 64 cycles, at the floor)
// and splits the row-owner work into two phases that each hide behind a different GEMM pair:
//
//   tensor pipe :  acc1(k)   S(k+1)      acc2(k)   dP(k+1)      acc1(k+1)  S(k+2) ...
//   row owners  :  ......D(k)......|.......E(k+1)........|......D(k+1)......|.....
//
//   E(k): S(k) -> P (exp2, mask) -> packed bf16 back into the S columns (TMEM A operand of acc1)
//   D(k): dP(k), P -> dS = P * (dP - delta) -> packed bf16 into the dP columns (TMEM A operand of acc2)
//
// TMEM: S | dP | acc1 | acc2 = 4 x 128 columns.  The streamed pair is released in two halves (Y2 after acc1, Y1 after
// acc2) so the next TMA loads start as early as the data dependencies allow with only two stages of shared memory.
// The softmax scale of dS is applied once in the epilogue (dK, dQ are linear in dS).
template <int HD>
struct Bwd3Cfg {
  static constexpr int NCH = HD / 64;
  static constexpr int T_BYTES = 128 * HD * 2;   // one 128-row operand tile
  static constexpr int CHUNK = 128 * 128;        // one 64-column 128B-swizzled chunk of a tile
  static constexpr int STAT_BYTES = 256 * 4;     // per stage: lse2[128] | delta[128] of the streamed q rows (DKDV)
  static constexpr int SMEM = 2 * T_BYTES + 2 * 2 * T_BYTES + 2 * STAT_BYTES + 1024 + 256;
};

B200_DEVINL float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// E phase of one row owner: 64 fp32 scores -> P = exp2(s * scale_log2 - lse2) (masked), packed bf16 pairs
template <int MODE, bool MASK>
B200_DEVINL void bwd3_exp(const uint32_t (&a)[64], uint32_t (&pk)[32], const float* stat, float row_lse2,
                          float scale_log2, int x_idx, int yb, int S) {
#pragma unroll
  for (int i = 0; i < 64; i += 4) {
    float l2[4];
    if constexpr (MODE == MODE_DKDV) {
      const float4 v = lds128(stat + i);
      l2[0] = v.x; l2[1] = v.y; l2[2] = v.z; l2[3] = v.w;
    } else {
      l2[0] = l2[1] = l2[2] = l2[3] = row_lse2;
    }
    float p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      // (measured: routing 1/4 of these through exp2_poly -- FMA-pipe exponential -- was 2 % SLOWER; the row owners are
      // issue/latency bound here, not MUFU bound)
      p[e] = exp2f(fmaf(__uint_as_float(a[i + e]), scale_log2, -l2[e]));
      if constexpr (MASK) {
        const int y_idx = yb + i + e;
        const bool masked = (MODE == MODE_DKDV) ? ((x_idx > y_idx) || (y_idx >= S) || (x_idx >= S))
                                                : ((y_idx > x_idx) || (y_idx >= S) || (x_idx >= S));
        if (masked) p[e] = 0.f;
      }
    }
    pk[i >> 1] = pack_bf16x2(p[0], p[1]);
    pk[(i >> 1) + 1] = pack_bf16x2(p[2], p[3]);
  }
}

template <int HD, int MODE>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const float* __restrict__ lse2g, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                 int S, int H, int KVH, float scale, int n_t128, int ld, const float* __restrict__ rope) {
  using C = Bwd3Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem;
  uint8_t* sX2 = sX1 + C::T_BYTES;
  uint8_t* sY = sX2 + C::T_BYTES;                           // [stage][Y1 | Y2]
  float* sStat = reinterpret_cast<float*>(sY + 4 * C::T_BYTES);   // [stage][lse2 128 | delta 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 2 * C::STAT_BYTES);
  uint64_t* x_full = bars;
  uint64_t* y1_full = bars + 1;    // [2]
  uint64_t* y2_full = bars + 3;    // [2]
  uint64_t* y1_empty = bars + 5;   // [2]
  uint64_t* y2_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;
  uint64_t* dp_full = bars + 10;
  uint64_t* p_full = bars + 11;
  uint64_t* ds_full = bars + 12;
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  ATRACE_INIT();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = H / KVH;
  const int b = blockIdx.z;
  const float scale_log2 = scale * LOG2E;
  int t128, head_lo, head_n, kvh;
  if constexpr (MODE == MODE_DKDV) {
    t128 = blockIdx.x; kvh = blockIdx.y; head_lo = kvh * G; head_n = G;
  } else {
    t128 = n_t128 - 1 - blockIdx.x; head_lo = blockIdx.y; head_n = 1; kvh = blockIdx.y / G;
  }
  const int s_lo = (MODE == MODE_DKDV) ? t128 : 0;
  const int s_hi = (MODE == MODE_DKDV) ? n_t128 : t128 + 1;
  const int per_head = s_hi - s_lo;
  const int n_iter = per_head * head_n;
  const int xrow0 = b * S + t128 * 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do);
    mbar_init(x_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y1_full[i], 1); mbar_init(&y2_full[i], 1); mbar_init(&y1_empty[i], 1); mbar_init(&y2_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(dp_full, 1); mbar_init(p_full, 8); mbar_init(ds_full, 8); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_s = tmem, tmem_dp = tmem + 128, tmem_acc1 = tmem + 256, tmem_acc2 = tmem + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(x_full, 2 * C::T_BYTES);
      for (int c = 0; c < C::NCH; ++c) {
        if constexpr (MODE == MODE_DKDV) {
          tma_load_2d(sX1 + c * C::CHUNK, &tm_qkv, x_full, (H + kvh) * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * C::CHUNK, &tm_qkv, x_full, (H + KVH + kvh) * HD + 64 * c, xrow0);
        } else {
          tma_load_2d(sX1 + c * C::CHUNK, &tm_qkv, x_full, head_lo * HD + 64 * c, xrow0);
          tma_load_2d(sX2 + c * C::CHUNK, &tm_do, x_full, head_lo * HD + 64 * c, xrow0);
        }
      }
      for (int it = 0; it < n_iter; ++it) {
        const int st = it & 1;
        const uint32_t free_par = ((it >> 1) & 1) ^ 1;
        const int hh = head_lo + it / per_head;
        const int t = s_lo + it % per_head;
        const int yrow = b * S + t * 128;
        uint8_t* y1 = sY + st * 2 * C::T_BYTES;
        uint8_t* y2 = y1 + C::T_BYTES;
        mbar_wait(&y1_empty[st], free_par);
        mbar_arrive_expect_tx(&y1_full[st], C::T_BYTES + (MODE == MODE_DKDV ? C::STAT_BYTES : 0));
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) tma_load_2d(y1 + c * C::CHUNK, &tm_qkv, &y1_full[st], hh * HD + 64 * c, yrow);
          else tma_load_2d(y1 + c * C::CHUNK, &tm_qkv, &y1_full[st], (H + kvh) * HD + 64 * c, yrow);
        }
        if constexpr (MODE == MODE_DKDV) {
          // per-column softmax statistics of the streamed q rows ride on the Y1 barrier (two 512 B bulk copies)
          const size_t so = ((size_t)b * H + hh) * ld + (size_t)t * 128;
          bulk_load_1d(sStat + st * 256, lse2g + so, 512, &y1_full[st]);
          bulk_load_1d(sStat + st * 256 + 128, delta + so, 512, &y1_full[st]);
        }
        mbar_wait(&y2_empty[st], free_par);
        mbar_arrive_expect_tx(&y2_full[st], C::T_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) tma_load_2d(y2 + c * C::CHUNK, &tm_do, &y2_full[st], hh * HD + 64 * c, yrow);
          else tma_load_2d(y2 + c * C::CHUNK, &tm_qkv, &y2_full[st], (H + KVH + kvh) * HD + 64 * c, yrow);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, HD, false, true);
    const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2), y0 = smem_u32(sY);
    // descriptors are built once; per-MMA operands are base + constant (the start-address field counts 16-byte units)
    const uint64_t x1d = make_smem_desc(x1, 0, 1024), x2d = make_smem_desc(x2, 0, 1024);
    // scores: T = X * Y^T (both operands K-major, 128 x 128 x HD)
    auto issue_scores = [&](uint32_t tdst, uint64_t xd, uint32_t ya, uint64_t* done) {
      if (elect_one()) {
        const uint64_t yd = make_smem_desc(ya, 0, 1024);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t o = ((kk >> 2) * C::CHUNK + (kk & 3) * 32) >> 4;
          umma_bf16_ss(tdst, xd + o, yd + o, idesc_t, kk != 0);
        }
        umma_commit(done);
      }
      __syncwarp();
    };
    // accumulate: acc += W(TMEM, packed bf16 [128 x 128 streamed]) * Y (MN-major [128 streamed x HD])
    auto issue_acc = [&](uint32_t tacc, uint32_t tw, uint32_t ya, bool first, uint64_t* release, uint64_t* done) {
      if (elect_one()) {
        const uint64_t yd = make_smem_desc(ya, C::CHUNK, 1024);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const uint32_t co = 64 * (t >> 2) + 8 * (t & 3);   // warpgroup g packs its 64 columns into [64g, 64g+32)
          umma_bf16_ts(tacc, tw + co, yd + ((t * 2048) >> 4), idesc_a, !(first && t == 0));
        }
        umma_commit(release);
        if (done) umma_commit(done);
      }
      __syncwarp();
    };
    mbar_wait(x_full, 0);
    if (n_iter > 0) {
      mbar_wait(&y1_full[0], 0);
      tc_fence_after();
      issue_scores(tmem_s, x1d, y0, s_full);
      mbar_wait(&y2_full[0], 0);
      tc_fence_after();
      issue_scores(tmem_dp, x2d, y0 + C::T_BYTES, dp_full);
      if constexpr (MODE == MODE_DQ) { if (elect_one()) umma_commit(&y2_empty[0]); __syncwarp(); }
    }
    for (int it = 0; it < n_iter; ++it) {
      const int st = it & 1, nst = st ^ 1;
      const uint32_t ph = it & 1;
      const uint32_t ya1 = y0 + st * 2 * C::T_BYTES, ya2 = ya1 + C::T_BYTES;
      const uint32_t nya1 = y0 + nst * 2 * C::T_BYTES, nya2 = nya1 + C::T_BYTES;
      const bool last = (it + 1 == n_iter);
      mbar_wait(p_full, ph);                       // P(it) is in the S columns (and S(it) has been consumed)
      tc_fence_after();
      if (lane == 0) ATRACE(it, 0);
      if constexpr (MODE == MODE_DKDV) issue_acc(tmem_acc1, tmem_s, ya2, it == 0, &y2_empty[st], nullptr);
      if (!last) {
        mbar_wait(&y1_full[nst], ((it + 1) >> 1) & 1);
        tc_fence_after();
        issue_scores(tmem_s, x1d, nya1, s_full);
      }
      if (lane == 0) ATRACE(it, 1);
      mbar_wait(ds_full, ph);                      // dS(it) is in the dP columns
      tc_fence_after();
      if (lane == 0) ATRACE(it, 2);
      issue_acc(tmem_acc2, tmem_dp, ya1, it == 0, &y1_empty[st], last ? acc_done : nullptr);
      if (!last) {
        mbar_wait(&y2_full[nst], ((it + 1) >> 1) & 1);
        tc_fence_after();
        issue_scores(tmem_dp, x2d, nya2, dp_full);
        if constexpr (MODE == MODE_DQ) { if (elect_one()) umma_commit(&y2_empty[nst]); __syncwarp(); }
      }
      if (lane == 0) ATRACE(it, 3);
    }
  } else if (warp >= 4) {
    // 8 row-owner warps; warpgroup g owns columns [64g, 64g+64) of every 128 x 128 tile, thread <-> row
    const int g = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q4 * 32) << 16;
    const int x_idx = t128 * 128 + r;
    float row_lse2 = 0.f, row_delta = 0.f;
    if constexpr (MODE == MODE_DQ) {
      const size_t sidx = ((size_t)b * H + head_lo) * ld + min(x_idx, S - 1);
      row_lse2 = lse2g[sidx];
      row_delta = delta[sidx];
    }
    const uint32_t ts = tmem_s + lane_addr + 64 * g, td = tmem_dp + lane_addr + 64 * g;
    int t_in_head = 0;
    for (int it = 0; it < n_iter; ++it) {
      const int st = it & 1;
      const uint32_t ph = it & 1;
      const int t = s_lo + t_in_head;
      if (++t_in_head == per_head) t_in_head = 0;
      const int yb = t * 128 + 64 * g;             // first streamed index of this warpgroup's columns
      const float* stat = sStat + st * 256 + 64 * g;
      // only tiles on the causal diagonal or crossing the sequence end need per-element masking
      const bool need_mask = (t == t128) || (t * 128 + 128 > S) || (t128 * 128 + 128 > S);
      uint32_t pk[32];                             // P of this thread's 64 columns, packed bf16 (kept for the dS phase)
      // ---------------- E phase: S -> P
      if constexpr (MODE == MODE_DKDV) mbar_wait(&y1_full[st], (it >> 1) & 1);   // column statistics have landed
      if (lane == 0 && q4 == 0) ATRACE(it, 4 + 5 * g);
      mbar_wait(s_full, ph);
      tc_fence_after();
      if (lane == 0 && q4 == 0) ATRACE(it, 5 + 5 * g);
      {
        uint32_t a[64];
        tmem_ld_32x32b_x32(ts, *reinterpret_cast<uint32_t(*)[32]>(&a[0]));
        tmem_ld_32x32b_x32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(&a[32]));
        tmem_ld_wait();
        if (need_mask) bwd3_exp<MODE, true>(a, pk, stat, row_lse2, scale_log2, x_idx, yb, S);
        else           bwd3_exp<MODE, false>(a, pk, stat, row_lse2, scale_log2, x_idx, yb, S);
      }
      if constexpr (MODE == MODE_DKDV) {
        tmem_st_32x32b_x16(ts, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st_32x32b_x16(ts + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (lane == 0 && q4 == 0) ATRACE(it, 6 + 5 * g);
      // ---------------- D phase: dP, P -> dS (unscaled)
      mbar_wait(dp_full, ph);
      tc_fence_after();
      if (lane == 0 && q4 == 0) ATRACE(it, 7 + 5 * g);
      {
        uint32_t d[64];
        tmem_ld_32x32b_x32(td, *reinterpret_cast<uint32_t(*)[32]>(&d[0]));
        tmem_ld_32x32b_x32(td + 32, *reinterpret_cast<uint32_t(*)[32]>(&d[32]));
        tmem_ld_wait();
        uint32_t dk[32];
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          float dl[4];
          if constexpr (MODE == MODE_DKDV) {
            const float4 v = lds128(stat + 128 + i);
            dl[0] = v.x; dl[1] = v.y; dl[2] = v.z; dl[3] = v.w;
          } else {
            dl[0] = dl[1] = dl[2] = dl[3] = row_delta;
          }
          const float2 p01 = unpack_bf16x2(pk[i >> 1]), p23 = unpack_bf16x2(pk[(i >> 1) + 1]);
          const float pp[4] = {p01.x, p01.y, p23.x, p23.y};
          float dv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = pp[e] * (__uint_as_float(d[i + e]) - dl[e]);
            if (need_mask && pp[e] == 0.f) v = 0.f;   // masked / padded entries: statistics may be garbage (NaN * 0)
            dv[e] = v;
          }
          dk[i >> 1] = pack_bf16x2(dv[0], dv[1]);
          dk[(i >> 1) + 1] = pack_bf16x2(dv[2], dv[3]);
        }
        tmem_st_32x32b_x16(td, *reinterpret_cast<uint32_t(*)[16]>(&dk[0]));
        tmem_st_32x32b_x16(td + 16, *reinterpret_cast<uint32_t(*)[16]>(&dk[16]));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
      if (lane == 0 && q4 == 0) ATRACE(it, 8 + 5 * g);
    }
    // epilogue: warpgroup 0 stores acc2 (dK | dQ, times the softmax scale), warpgroup 1 acc1 (dV); in DQ mode the two
    // groups split acc2's columns
    if (n_iter > 0) mbar_wait(acc_done, 0);
    tc_fence_after();
    if (x_idx < S && n_iter > 0) {
      const size_t row = (size_t)b * S + x_idx;
      const int Wd = (H + 2 * KVH) * HD;
      // rot: apply the INVERSE rotary embedding of this row's position (the forward RoPE lives in the QKV GEMM epilogue,
      // so what leaves here is the gradient of the un-rotated projection); table [S][HD/2][cos, sin]
      auto store_acc = [&](uint32_t tacc, int col0, int c_lo, int c_hi, float mul, bool rot) {
        __nv_bfloat16* dst = dqkv + row * Wd + col0;
        const float* trow = rope + (size_t)x_idx * HD;
#pragma unroll 1
        for (int c = c_lo; c < c_hi; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tacc + lane_addr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[q * 8 + i]) * mul;
            if (rot) {
              const float4* tp = reinterpret_cast<const float4*>(trow + c + q * 8);
              const float4 t0 = tp[0], t1 = tp[1];
              const float cs[4] = {t0.x, t0.z, t1.x, t1.z}, sn[4] = {t0.y, t0.w, t1.y, t1.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float a0 = f[2 * i], a1 = f[2 * i + 1];
                f[2 * i] = a0 * cs[i] + a1 * sn[i];
                f[2 * i + 1] = a1 * cs[i] - a0 * sn[i];
              }
            }
            uint4 u;
            u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
            u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(dst + c + q * 8) = u;
          }
        }
      };
      const bool rot = rope != nullptr;
      if constexpr (MODE == MODE_DKDV) {
        if (g == 0) store_acc(tmem_acc2, (H + kvh) * HD, 0, HD, scale, rot);
        else        store_acc(tmem_acc1, (H + KVH + kvh) * HD, 0, HD, 1.f, false);
      } else {
        if (g == 0) store_acc(tmem_acc2, head_lo * HD, 0, HD / 2, scale, rot);
        else        store_acc(tmem_acc2, head_lo * HD, HD / 2, HD, scale, rot);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int HD>
static int launch_fwd2(const void* qkv, void* o, float* lse, int B, int S, int H, int KVH, float scale,
                       cudaStream_t st) {
  using C = Fwd2Cfg<HD>;
  CUtensorMap tm;
  const int W = (H + 2 * KVH) * HD;
  if (make_tmap_2d_bf16(&tm, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 128)) return -3;
  auto kern = attn_fwd2_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int n_pt = (S + 255) / 256;
  dim3 grid(n_pt, H, B);
  kern<<<grid, ATT2_THREADS, C::SMEM, st>>>(tm, (__nv_bfloat16*)o, lse, S, H, KVH, scale * LOG2E, n_pt);
  return (int)cudaGetLastError();
}

template <int HD>
static int launch_fwd(const void* qkv, void* o, float* lse, int B, int S, int H, int KVH, float scale,
                      cudaStream_t st) {
  if (g_attn_fwd_version == 2) return launch_fwd2<HD>(qkv, o, lse, B, S, H, KVH, scale, st);
  using C = FwdCfg<HD>;
  CUtensorMap tm;
  const int W = (H + 2 * KVH) * HD;
  if (make_tmap_2d_bf16(&tm, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 128)) return -3;
  auto kern = attn_fwd_kernel<HD>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int n_qt = (S + 127) / 128;
  dim3 grid(n_qt, H, B);
  kern<<<grid, ATT_THREADS, C::SMEM, st>>>(tm, (__nv_bfloat16*)o, lse, S, H, KVH, scale * LOG2E, n_qt);
  return (int)cudaGetLastError();
}

template <int HD>
static int launch_bwd(const void* dout, const void* qkv, const void* o, const float* lse, void* dqkv, float* delta,
                      int B, int S, int H, int KVH, float scale, const float* rope, cudaStream_t st) {
  using C = BwdCfg<HD>;
  const int W = (H + 2 * KVH) * HD;
  CUtensorMap q128, q64, d128, d64;
  if (make_tmap_2d_bf16(&q128, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 128)) return -3;
  if (make_tmap_2d_bf16(&q64, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 64)) return -3;
  if (make_tmap_2d_bf16(&d128, dout, (uint64_t)H * HD, (uint64_t)B * S, (uint64_t)H * HD, 64, 128)) return -3;
  if (make_tmap_2d_bf16(&d64, dout, (uint64_t)H * HD, (uint64_t)B * S, (uint64_t)H * HD, 64, 64)) return -3;
  {
    const long long warps = (long long)B * S * H;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    // v2: rows padded to a multiple of 64 so a 64-float TMA bulk copy never leaves the row; plane 1 = lse * log2(e)
    const int ld = (g_attn_bwd_version >= 2) ? ((S + 127) / 128) * 128 : S;
    float* lse2 = (g_attn_bwd_version >= 2) ? delta + (size_t)B * H * ld : nullptr;
    attn_delta_kernel<<<(unsigned)blocks, threads, 0, st>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)o, delta,
                                                            lse, lse2, B, S, H, HD, ld);
  }
  const int n_t = (S + 127) / 128;
  if (g_attn_bwd_version == 3) {
    using C3 = Bwd3Cfg<HD>;
    auto j1 = attn_bwd3_kernel<HD, MODE_DKDV>;
    auto j2 = attn_bwd3_kernel<HD, MODE_DQ>;
    static bool configured3 = false;
    if (!configured3) {
      cudaError_t e = cudaFuncSetAttribute(j1, cudaFuncAttributeMaxDynamicSharedMemorySize, C3::SMEM);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(j2, cudaFuncAttributeMaxDynamicSharedMemorySize, C3::SMEM);
      if (e != cudaSuccess) return (int)e;
      configured3 = true;
    }
    const int ld3 = ((S + 127) / 128) * 128;
    const float* lse2p = delta + (size_t)B * H * ld3;
    j1<<<dim3(n_t, KVH, B), ATT2_THREADS, C3::SMEM, st>>>(q128, d128, lse2p, delta, (__nv_bfloat16*)dqkv, S, H, KVH, scale,
                                                         n_t, ld3, rope);
    j2<<<dim3(n_t, H, B), ATT2_THREADS, C3::SMEM, st>>>(q128, d128, lse2p, delta, (__nv_bfloat16*)dqkv, S, H, KVH, scale,
                                                       n_t, ld3, rope);
    return (int)cudaGetLastError();
  }
  if (rope) return -9;   // only the v3 kernels fuse the inverse RoPE
  if (g_attn_bwd_version == 2) {
    using C2 = Bwd2Cfg<HD>;
    auto j1 = attn_bwd2_kernel<HD, MODE_DKDV>;
    auto j2 = attn_bwd2_kernel<HD, MODE_DQ>;
    static bool configured2 = false;
    if (!configured2) {
      cudaError_t e = cudaFuncSetAttribute(j1, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM);
      if (e != cudaSuccess) return (int)e;
      e = cudaFuncSetAttribute(j2, cudaFuncAttributeMaxDynamicSharedMemorySize, C2::SMEM);
      if (e != cudaSuccess) return (int)e;
      configured2 = true;
    }
    const int ld2 = ((S + 127) / 128) * 128;
    const float* lse2p = delta + (size_t)B * H * ld2;
    j1<<<dim3(n_t, KVH, B), ATT2_THREADS, C2::SMEM, st>>>(q128, q64, d128, d64, lse2p, delta, (__nv_bfloat16*)dqkv, S, H,
                                                         KVH, scale, n_t, ld2);
    j2<<<dim3(n_t, H, B), ATT2_THREADS, C2::SMEM, st>>>(q128, q64, d128, d64, lse2p, delta, (__nv_bfloat16*)dqkv, S, H, KVH,
                                                       scale, n_t, ld2);
    return (int)cudaGetLastError();
  }
  auto k1 = attn_bwd_kernel<HD, MODE_DKDV>;
  auto k2 = attn_bwd_kernel<HD, MODE_DQ>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  k1<<<dim3(n_t, KVH, B), ATT_THREADS, C::SMEM, st>>>(q128, q64, d128, d64, lse, delta, (__nv_bfloat16*)dqkv, S, H, KVH,
                                                     scale, n_t);
  k2<<<dim3(n_t, H, B), ATT_THREADS, C::SMEM, st>>>(q128, q64, d128, d64, lse, delta, (__nv_bfloat16*)dqkv, S, H, KVH,
                                                   scale, n_t);
  return (int)cudaGetLastError();
}

}  // namespace b200

extern "C" void b200_attn_set_fwd_version(int v) { b200::g_attn_fwd_version = v; }
extern "C" void b200_attn_set_bwd_version(int v) { b200::g_attn_bwd_version = v; }
extern "C" int b200_attn_fwd(const void* qkv, void* o, float* lse, int B, int S, int H, int KVH, int HD, float scale,
                             cudaStream_t st) {
  if (H % KVH) return -1;
  if (HD == 128) return b200::launch_fwd<128>(qkv, o, lse, B, S, H, KVH, scale, st);
  if (HD == 64) return b200::launch_fwd<64>(qkv, o, lse, B, S, H, KVH, scale, st);
  return -2;
}
// rope (optional): [S][HD/2][cos, sin] table; dq and dk leave the kernel with the inverse rotation applied
extern "C" int b200_attn_bwd(const void* dout, const void* qkv, const void* o, const float* lse, void* dqkv,
                             float* delta, int B, int S, int H, int KVH, int HD, float scale, const float* rope,
                             cudaStream_t st) {
  if (H % KVH) return -1;
  if (HD == 128) return b200::launch_bwd<128>(dout, qkv, o, lse, dqkv, delta, B, S, H, KVH, scale, rope, st);
  if (HD == 64) return b200::launch_bwd<64>(dout, qkv, o, lse, dqkv, delta, B, S, H, KVH, scale, rope, st);
  return -2;
}
