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

// ---- packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2: two fp32 lanes per issue slot) and an exponential that runs on
// the FMA pipe.  Measured (ncu source page of the round-1 kernels, profiles/ncu_attention_r1.txt): the softmax /
// softmax-gradient warps stall on MUFU.EX2 behind the MIO queue -- one 128 x 128 tile costs ~2.1k cycles of exponentials
// against 1k (forward) / 2k (backward) cycles of tensor-core work, so the tensor pipe idles at ~45 %.  Moving a
// compile-time fraction of every group of 8 exponentials onto the FMA pipe (Cody-Waite range reduction + degree-3
// minimax polynomial, 7.5e-5 relative error -- 50x below bf16 rounding of P) lets both pipes work in parallel.
typedef unsigned long long f32x2_t;
B200_DEVINL f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
B200_DEVINL void f2_unpack(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
B200_DEVINL f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
B200_DEVINL f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
B200_DEVINL f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// 2^x for two values, FMA + ALU pipes only.  x is clamped at -126 (result ~1e-38, i.e. 0 after the bf16 pack).
B200_DEVINL f32x2_t exp2_fma2(f32x2_t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  x = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const f32x2_t t = f2_add(x, f2_pack(12582912.f, 12582912.f));      // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const f32x2_t r = f2_add(t, f2_pack(-12582912.f, -12582912.f));    // round(x)
  const f32x2_t f = f2_fma(r, f2_pack(-1.f, -1.f), x);               // in [-0.5, 0.5]
  f32x2_t p = f2_fma(f, f2_pack(0.05517159402370453f, 0.05517159402370453f), f2_pack(0.2426111400127411f, 0.2426111400127411f));
  p = f2_fma(p, f, f2_pack(0.6932610273361206f, 0.6932610273361206f));
  p = f2_fma(p, f, f2_pack(0.9999280571937561f, 0.9999280571937561f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  return f2_pack(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)),     // * 2^round(x) through the exponent
                 __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}
// exponentials of 8 values (4 packed pairs): the first PP pairs on the FMA pipe, the others on the MUFU
template <int PP>
B200_DEVINL void exp2_group8(const f32x2_t (&x)[4], f32x2_t (&y)[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < PP) {
      y[q] = exp2_fma2(x[q]);
    } else {
      float a, b;
      f2_unpack(x[q], a, b);
      y[q] = f2_pack(exp2f(a), exp2f(b));
    }
  }
}
B200_DEVINL float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
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

template <int HD, int PP>
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
      for (int i = 0; i < 128; i += 4) {   // 3-input max: two elements per instruction
        mx0 = fmax3(mx0, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
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
      // P = 2^(x - m_ref) packed two bf16 per column, written over S's own columns (in order => no hazard).
      // Packed FFMA2 for the scale / shift, PP of every 4 pairs exponentiated on the FMA pipe, the rest on the MUFU.
      {
        const f32x2_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(-m_ref, -m_ref);
        f32x2_t ls2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[16];
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            f32x2_t x[4], y[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              x[q] = f2_fma(f2_pack(__uint_as_float(v[c * 32 + g8 * 8 + 2 * q]), __uint_as_float(v[c * 32 + g8 * 8 + 2 * q + 1])),
                            sc2, nm2);
            exp2_group8<PP>(x, y);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              ls2 = f2_add(ls2, y[q]);
              float p0, p1;
              f2_unpack(y[q], p0, p1);
              w[g8 * 4 + q] = pack_bf16x2(p0, p1);
            }
          }
          tmem_st_32x32b_x16(tS + c * 16, w);
        }
        float s0, s1;
        f2_unpack(ls2, s0, s1);
        l_sum += s0 + s1;
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
// -delta[b,h,s] = -sum_d dO[b,s,h,d] * O[b,s,h,d]  and  -lse * log2(e)  (both planes NEGATED: the backward kernels use
// them as the addend of a packed FFMA2 / FADD2).  HD/8 lanes per (row, head), one 16-byte load of dO and of O per lane
// (the round-1 kernel issued 4-byte loads from one warp per row: 78 us for 134 MB, now bandwidth-bound).
template <int HD>
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ o,
                                                         float* __restrict__ delta, const float* __restrict__ lse,
                                                         float* __restrict__ lse2, int B, int S, int H, int ld) {
  constexpr int G = HD / 8;                              // lanes per (row, head)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long item = gid / G;                        // (b*S + s) * H + h
  const int sub = (int)(gid % G);
  const bool live = item < (long long)B * S * H;
  float acc = 0.f;
  if (live) {
    const uint4 x = *reinterpret_cast<const uint4*>(dout + item * HD + sub * 8);
    const uint4 y = *reinterpret_cast<const uint4*>(o + item * HD + sub * 8);
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = unpack_bf16x2(xs[i]), c = unpack_bf16x2(ys[i]);
      acc = fmaf(a.x, c.x, fmaf(a.y, c.y, acc));
    }
  }
#pragma unroll
  for (int m = G / 2; m > 0; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
  if (live && sub == 0) {
    const int h = (int)(item % H);
    const long long bs = item / H;
    const int b = (int)(bs / S), sq = (int)(bs % S);
    delta[((size_t)b * H + h) * ld + sq] = -acc;
    if (lse2) lse2[((size_t)b * H + h) * ld + sq] = -lse[((size_t)b * H + h) * S + sq] * LOG2E;
  }
}

// ============================================================================================ backward
enum { MODE_DKDV = 0, MODE_DQ = 1 };

B200_DEVINL void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
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
  static constexpr int STAT_BYTES = 256 * 4;     // per stage: -lse2[128] | -delta[128] of the streamed q rows (DKDV)
  // Y1 (K in the dQ kernel, Q in the dK/dV kernel) is held from the score GEMM of an iteration until its LAST accumulate
  // GEMM, so with two stages the refill of a stage (global -> smem, ~1.7k cycles measured) sat on the critical path of
  // every second iteration (profiles/attn_bwd3_pipeline_trace_r2.txt: 2.7k cycles per iteration against 1.5k of tensor
  // work in the dQ kernel).  Y1 gets a third stage; Y2 (released early) keeps two.  At HD = 128 that is 7 x 32 KiB of
  // tiles: the carve-up only fits the 227 KiB of the SM without the usual 1 KiB alignment slack, so the kernel asks for
  // the maximum and traps if the (in practice 1 KiB-aligned) dynamic base leaves too little room.
  static constexpr int Y1_STAGES = 3;
  static constexpr int NEED = (2 + Y1_STAGES + 2) * T_BYTES + 2 * STAT_BYTES + 256;
  static constexpr int SMEM = (NEED + 1024 <= 232448) ? NEED + 1024 : 232448;
};

B200_DEVINL float4 lds128(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// E phase of one row owner: 64 fp32 scores -> P = exp2(s * scale_log2 - lse2) (masked), packed bf16 pairs.
// ``nstat`` / ``row_nlse2`` hold -lse2 (attn_delta_kernel stores the statistics negated).
template <int MODE, bool MASK, int PP>
B200_DEVINL void bwd3_exp(const uint32_t (&a)[64], uint32_t (&pk)[32], const float* nstat, float row_nlse2,
                          float scale_log2, int x_idx, int yb, int S) {
  const f32x2_t sc2 = f2_pack(scale_log2, scale_log2), rn2 = f2_pack(row_nlse2, row_nlse2);
#pragma unroll
  for (int i = 0; i < 64; i += 8) {
    f32x2_t nl[4];
    if constexpr (MODE == MODE_DKDV) {
      const float4 u = lds128(nstat + i), w = lds128(nstat + i + 4);
      nl[0] = f2_pack(u.x, u.y); nl[1] = f2_pack(u.z, u.w); nl[2] = f2_pack(w.x, w.y); nl[3] = f2_pack(w.z, w.w);
    } else {
      nl[0] = nl[1] = nl[2] = nl[3] = rn2;
    }
    f32x2_t x[4], y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      x[q] = f2_fma(f2_pack(__uint_as_float(a[i + 2 * q]), __uint_as_float(a[i + 2 * q + 1])), sc2, nl[q]);
    exp2_group8<PP>(x, y);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float p[2];
      f2_unpack(y[q], p[0], p[1]);
      if constexpr (MASK) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int y_idx = yb + i + 2 * q + e;
          const bool masked = (MODE == MODE_DKDV) ? ((x_idx > y_idx) || (y_idx >= S) || (x_idx >= S))
                                                  : ((y_idx > x_idx) || (y_idx >= S) || (x_idx >= S));
          if (masked) p[e] = 0.f;
        }
      }
      pk[(i >> 1) + q] = pack_bf16x2(p[0], p[1]);
    }
  }
}

// D phase: dS = P * (dP - delta) (unscaled), packed bf16 pairs.  ``nstat`` / ``row_ndelta`` hold -delta.
template <int MODE, bool MASK>
B200_DEVINL void bwd3_ds(const uint32_t (&d)[64], const uint32_t (&pk)[32], uint32_t (&dk)[32], const float* nstat,
                         float row_ndelta) {
  const f32x2_t rd2 = f2_pack(row_ndelta, row_ndelta);
#pragma unroll
  for (int i = 0; i < 64; i += 8) {
    f32x2_t nd[4];
    if constexpr (MODE == MODE_DKDV) {
      const float4 u = lds128(nstat + i), w = lds128(nstat + i + 4);
      nd[0] = f2_pack(u.x, u.y); nd[1] = f2_pack(u.z, u.w); nd[2] = f2_pack(w.x, w.y); nd[3] = f2_pack(w.z, w.w);
    } else {
      nd[0] = nd[1] = nd[2] = nd[3] = rd2;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t pp = pk[(i >> 1) + q];
      const f32x2_t p2 = f2_pack(__uint_as_float(pp << 16), __uint_as_float(pp & 0xffff0000u));
      const f32x2_t v2 = f2_mul(p2, f2_add(f2_pack(__uint_as_float(d[i + 2 * q]), __uint_as_float(d[i + 2 * q + 1])), nd[q]));
      float v0, v1;
      f2_unpack(v2, v0, v1);
      if constexpr (MASK) {   // masked / padded entries: the statistics may be garbage (NaN * 0)
        if ((pp & 0xffffu) == 0) v0 = 0.f;
        if ((pp >> 16) == 0) v1 = 0.f;
      }
      dk[(i >> 1) + q] = pack_bf16x2(v0, v1);
    }
  }
}

template <int HD, int MODE, int PP>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                 const float* __restrict__ lse2g, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv,
                 int S, int H, int KVH, float scale, int n_t128, int ld, const float* __restrict__ rope) {
  using C = Bwd3Cfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sX1 = smem;
  uint8_t* sX2 = sX1 + C::T_BYTES;
  uint8_t* sY1 = sX2 + C::T_BYTES;                          // [3 stages]
  uint8_t* sY2 = sY1 + C::Y1_STAGES * C::T_BYTES;           // [2 stages]
  float* sStat = reinterpret_cast<float*>(sY2 + 2 * C::T_BYTES);   // [2 stages][-lse2 128 | -delta 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sStat) + 2 * C::STAT_BYTES);
  if (reinterpret_cast<uint8_t*>(bars) + 256 > smem_raw + C::SMEM) __trap();   // dynamic smem base less aligned than assumed
  uint64_t* x_full = bars;
  uint64_t* y1_full = bars + 1;    // [3]
  uint64_t* y1_empty = bars + 4;   // [3]
  uint64_t* y2_full = bars + 7;    // [2]
  uint64_t* y2_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;
  uint64_t* dp_full = bars + 12;
  uint64_t* p_full = bars + 13;
  uint64_t* ds_full = bars + 14;
  uint64_t* acc_done = bars + 15;
  uint64_t* stat_full = bars + 16;   // [2]  (dK/dV kernel: column statistics of the streamed q rows)
  uint64_t* stat_empty = bars + 18;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
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
    for (int i = 0; i < C::Y1_STAGES; ++i) { mbar_init(&y1_full[i], 1); mbar_init(&y1_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&y2_full[i], 1); mbar_init(&y2_empty[i], 1);
      mbar_init(&stat_full[i], 1); mbar_init(&stat_empty[i], 8);   // 8 row-owner warps release a statistics stage
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
        const int s1 = it % C::Y1_STAGES, s2 = it & 1;
        const uint32_t free1 = ((it / C::Y1_STAGES) & 1) ^ 1, free2 = ((it >> 1) & 1) ^ 1;
        const int hh = head_lo + it / per_head;
        const int t = s_lo + it % per_head;
        const int yrow = b * S + t * 128;
        uint8_t* y1 = sY1 + s1 * C::T_BYTES;
        uint8_t* y2 = sY2 + s2 * C::T_BYTES;
        mbar_wait(&y1_empty[s1], free1);
        mbar_arrive_expect_tx(&y1_full[s1], C::T_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) tma_load_2d(y1 + c * C::CHUNK, &tm_qkv, &y1_full[s1], hh * HD + 64 * c, yrow);
          else tma_load_2d(y1 + c * C::CHUNK, &tm_qkv, &y1_full[s1], (H + kvh) * HD + 64 * c, yrow);
        }
        if constexpr (MODE == MODE_DKDV) {
          // per-column softmax statistics of the streamed q rows: two 512 B bulk copies, own 2-stage ring (a stage is
          // released by the row owners after the dS phase has read it)
          const size_t so = ((size_t)b * H + hh) * ld + (size_t)t * 128;
          mbar_wait(&stat_empty[s2], free2);
          mbar_arrive_expect_tx(&stat_full[s2], C::STAT_BYTES);
          bulk_load_1d(sStat + s2 * 256, lse2g + so, 512, &stat_full[s2]);
          bulk_load_1d(sStat + s2 * 256 + 128, delta + so, 512, &stat_full[s2]);
        }
        mbar_wait(&y2_empty[s2], free2);
        mbar_arrive_expect_tx(&y2_full[s2], C::T_BYTES);
        for (int c = 0; c < C::NCH; ++c) {
          if constexpr (MODE == MODE_DKDV) tma_load_2d(y2 + c * C::CHUNK, &tm_do, &y2_full[s2], hh * HD + 64 * c, yrow);
          else tma_load_2d(y2 + c * C::CHUNK, &tm_qkv, &y2_full[s2], (H + KVH + kvh) * HD + 64 * c, yrow);
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_t = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t idesc_a = make_idesc_bf16(128, HD, false, true);
    const uint32_t x1 = smem_u32(sX1), x2 = smem_u32(sX2), y1b = smem_u32(sY1), y2b = smem_u32(sY2);
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
      issue_scores(tmem_s, x1d, y1b, s_full);
      mbar_wait(&y2_full[0], 0);
      tc_fence_after();
      issue_scores(tmem_dp, x2d, y2b, dp_full);
      if constexpr (MODE == MODE_DQ) { if (elect_one()) umma_commit(&y2_empty[0]); __syncwarp(); }
    }
    for (int it = 0; it < n_iter; ++it) {
      const int s1 = it % C::Y1_STAGES, s2 = it & 1, ns1 = (it + 1) % C::Y1_STAGES, ns2 = s2 ^ 1;
      const uint32_t ph = it & 1;
      const uint32_t ya1 = y1b + s1 * C::T_BYTES, ya2 = y2b + s2 * C::T_BYTES;
      const uint32_t nya1 = y1b + ns1 * C::T_BYTES, nya2 = y2b + ns2 * C::T_BYTES;
      const bool last = (it + 1 == n_iter);
      mbar_wait(p_full, ph);                       // P(it) is in the S columns (and S(it) has been consumed)
      tc_fence_after();
      if (lane == 0) ATRACE(it, 0);
      if constexpr (MODE == MODE_DKDV) issue_acc(tmem_acc1, tmem_s, ya2, it == 0, &y2_empty[s2], nullptr);
      if (!last) {
        mbar_wait(&y1_full[ns1], ((it + 1) / C::Y1_STAGES) & 1);
        tc_fence_after();
        issue_scores(tmem_s, x1d, nya1, s_full);
      }
      if (lane == 0) ATRACE(it, 1);
      mbar_wait(ds_full, ph);                      // dS(it) is in the dP columns
      tc_fence_after();
      if (lane == 0) ATRACE(it, 2);
      issue_acc(tmem_acc2, tmem_dp, ya1, it == 0, &y1_empty[s1], last ? acc_done : nullptr);
      if (!last) {
        mbar_wait(&y2_full[ns2], ((it + 1) >> 1) & 1);
        tc_fence_after();
        issue_scores(tmem_dp, x2d, nya2, dp_full);
        if constexpr (MODE == MODE_DQ) { if (elect_one()) umma_commit(&y2_empty[ns2]); __syncwarp(); }
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
      if constexpr (MODE == MODE_DKDV) mbar_wait(&stat_full[st], (it >> 1) & 1);   // column statistics have landed
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 4);
      mbar_wait(s_full, ph);
      tc_fence_after();
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 5);
      {
        uint32_t a[64];
        tmem_ld_32x32b_x32(ts, *reinterpret_cast<uint32_t(*)[32]>(&a[0]));
        tmem_ld_32x32b_x32(ts + 32, *reinterpret_cast<uint32_t(*)[32]>(&a[32]));
        tmem_ld_wait();
        if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 9);
        if (need_mask) bwd3_exp<MODE, true, PP>(a, pk, stat, row_lse2, scale_log2, x_idx, yb, S);
        else           bwd3_exp<MODE, false, PP>(a, pk, stat, row_lse2, scale_log2, x_idx, yb, S);
      }
#ifdef B200_ATTN_TRACE
      asm volatile("" ::"r"(pk[0]), "r"(pk[31]) : "memory");   // the exponentials are done before the stamp
#endif
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 10);
      if constexpr (MODE == MODE_DKDV) {
        tmem_st_32x32b_x16(ts, *reinterpret_cast<uint32_t(*)[16]>(&pk[0]));
        tmem_st_32x32b_x16(ts + 16, *reinterpret_cast<uint32_t(*)[16]>(&pk[16]));
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 6);
      // ---------------- D phase: dP, P -> dS (unscaled)
      mbar_wait(dp_full, ph);
      tc_fence_after();
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 7);
      {
        uint32_t d[64];
        tmem_ld_32x32b_x32(td, *reinterpret_cast<uint32_t(*)[32]>(&d[0]));
        tmem_ld_32x32b_x32(td + 32, *reinterpret_cast<uint32_t(*)[32]>(&d[32]));
        tmem_ld_wait();
        if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 11);
        uint32_t dk[32];
        if (need_mask) bwd3_ds<MODE, true>(d, pk, dk, stat + 128, row_delta);
        else           bwd3_ds<MODE, false>(d, pk, dk, stat + 128, row_delta);
        tmem_st_32x32b_x16(td, *reinterpret_cast<uint32_t(*)[16]>(&dk[0]));
        tmem_st_32x32b_x16(td + 16, *reinterpret_cast<uint32_t(*)[16]>(&dk[16]));
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(ds_full);
        if constexpr (MODE == MODE_DKDV) mbar_arrive(&stat_empty[st]);   // this warp is done with the stage's statistics
      }
      if (lane == 0 && q4 == 0 && g == 0) ATRACE(it, 8);
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

// pairs (of every 4) whose exponential runs on the FMA pipe instead of the MUFU: 0 = all MUFU (round-1 behaviour)
static int g_attn_poly_fwd = 0;
static int g_attn_poly_bwd = 1;

template <int HD, int PP>
static int launch_fwd2(const void* qkv, void* o, float* lse, int B, int S, int H, int KVH, float scale,
                       cudaStream_t st) {
  using C = Fwd2Cfg<HD>;
  CUtensorMap tm;
  const int W = (H + 2 * KVH) * HD;
  if (make_tmap_2d_bf16(&tm, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 128)) return -3;
  auto kern = attn_fwd2_kernel<HD, PP>;
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
  switch (g_attn_poly_fwd) {
    case 0: return launch_fwd2<HD, 0>(qkv, o, lse, B, S, H, KVH, scale, st);
    case 1: return launch_fwd2<HD, 1>(qkv, o, lse, B, S, H, KVH, scale, st);
    case 3: return launch_fwd2<HD, 3>(qkv, o, lse, B, S, H, KVH, scale, st);
    default: return launch_fwd2<HD, 2>(qkv, o, lse, B, S, H, KVH, scale, st);
  }
}

template <int HD, int PP>
static int launch_bwd3(const CUtensorMap& q128, const CUtensorMap& d128, const float* lse2p, const float* delta, void* dqkv,
                       int B, int S, int H, int KVH, float scale, int n_t, int ld3, const float* rope, cudaStream_t st) {
  using C3 = Bwd3Cfg<HD>;
  auto j1 = attn_bwd3_kernel<HD, MODE_DKDV, PP>;
  auto j2 = attn_bwd3_kernel<HD, MODE_DQ, PP>;
  static bool configured3 = false;
  if (!configured3) {
    cudaError_t e = cudaFuncSetAttribute(j1, cudaFuncAttributeMaxDynamicSharedMemorySize, C3::SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(j2, cudaFuncAttributeMaxDynamicSharedMemorySize, C3::SMEM);
    if (e != cudaSuccess) return (int)e;
    configured3 = true;
  }
  j1<<<dim3(n_t, KVH, B), ATT2_THREADS, C3::SMEM, st>>>(q128, d128, lse2p, delta, (__nv_bfloat16*)dqkv, S, H, KVH, scale,
                                                       n_t, ld3, rope);
  j2<<<dim3(n_t, H, B), ATT2_THREADS, C3::SMEM, st>>>(q128, d128, lse2p, delta, (__nv_bfloat16*)dqkv, S, H, KVH, scale,
                                                     n_t, ld3, rope);
  return (int)cudaGetLastError();
}

template <int HD>
static int launch_bwd(const void* dout, const void* qkv, const void* o, const float* lse, void* dqkv, float* delta,
                      int B, int S, int H, int KVH, float scale, const float* rope, cudaStream_t st) {
  const int W = (H + 2 * KVH) * HD;
  CUtensorMap q128, d128;
  if (make_tmap_2d_bf16(&q128, qkv, (uint64_t)W, (uint64_t)B * S, (uint64_t)W, 64, 128)) return -3;
  if (make_tmap_2d_bf16(&d128, dout, (uint64_t)H * HD, (uint64_t)B * S, (uint64_t)H * HD, 64, 128)) return -3;
  // rows padded to a multiple of 128 so a 128-float TMA bulk copy never leaves the row; plane 0 = -delta,
  // plane 1 = -lse * log2(e)
  const int ld = ((S + 127) / 128) * 128;
  {
    const long long lanes = (long long)B * S * H * (HD / 8);
    const long long blocks = (lanes + 255) / 256;
    float* lse2 = delta + (size_t)B * H * ld;
    attn_delta_kernel<HD><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)o, delta,
                                                            lse, lse2, B, S, H, ld);
  }
  const int n_t = (S + 127) / 128;
  const float* lse2p = delta + (size_t)B * H * ld;
  switch (g_attn_poly_bwd) {
    case 0: return launch_bwd3<HD, 0>(q128, d128, lse2p, delta, dqkv, B, S, H, KVH, scale, n_t, ld, rope, st);
    case 1: return launch_bwd3<HD, 1>(q128, d128, lse2p, delta, dqkv, B, S, H, KVH, scale, n_t, ld, rope, st);
    case 3: return launch_bwd3<HD, 3>(q128, d128, lse2p, delta, dqkv, B, S, H, KVH, scale, n_t, ld, rope, st);
    default: return launch_bwd3<HD, 2>(q128, d128, lse2p, delta, dqkv, B, S, H, KVH, scale, n_t, ld, rope, st);
  }
}

}  // namespace b200

// (superseded kernel generations were removed in round 2; the selectors are kept as no-ops for old scripts)
extern "C" void b200_attn_set_fwd_version(int) {}
extern "C" void b200_attn_set_bwd_version(int) {}
extern "C" void b200_attn_set_poly(int fwd_pairs, int bwd_pairs) {
  if (fwd_pairs >= 0 && fwd_pairs <= 3) b200::g_attn_poly_fwd = fwd_pairs;
  if (bwd_pairs >= 0 && bwd_pairs <= 3) b200::g_attn_poly_bwd = bwd_pairs;
}
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
