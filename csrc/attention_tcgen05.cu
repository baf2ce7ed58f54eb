// Causal flash attention for sm_100a on tcgen05 tensor cores with TMEM-resident S / P / O.
//
// Forward (attn_fwd_kernel<DH>): one CTA = one (batch, head, 128-query tile).
//   warp 4 : TMA producer  — Q once, then a 2-stage ring of K / V tiles straight out of the fused
//            [B*S, 3*d] QKV projection output (no head split / transpose copies)
//   warp 5 : MMA issuer    — S = Q K^T (SS, both K-major) into TMEM; O += P V with P read FROM TMEM
//            (TS form) and V consumed MN-major as stored
//   warps 0-3: softmax      — one thread per query row (TMEM lane): tcgen05.ld S, online max/sum in
//            the log2 domain, P written back to TMEM as packed bf16 (tcgen05.st), lazy O rescale
// For d_head = 64 a CTA needs 256 TMEM columns and 80 KiB smem, so two CTAs share an SM and one CTA's
// softmax overlaps the other's MMAs.
//
// Backward (attn_bwd_kernel<64>): one CTA = one (batch, head, 128-key tile) holding K_j, V_j in smem and
// dK_j, dV_j in TMEM; loops over the query tiles i >= j: S = Q K^T and dP = dO V^T (SS), P / dS computed by
// the row threads and staged in smem ONCE in a layout that is simultaneously K-major (for dQ = dS K) and
// MN-major (for dV += P^T dO, dK += dS^T Q); dQ partial tiles leave through TMA reduce-add (fp32).
//
// Replaces flash-attn 2.x (mma.sync/cp.async) the reference calls through llm-foundry.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>
#include <type_traits>

#include "attention_tcgen05.h"
#include "gemm_tcgen05.h"
#include "ptx.cuh"

namespace pb {

namespace {

constexpr int BQ = 128, BKV = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

// ======================================================================================= forward
template <int DH>
struct FwdCfg {
  static constexpr int CH = DH / 64;              // 64-column swizzle chunks per tile
  static constexpr int TILE = 128 * 128 * CH;     // bytes of one [128 x DH] bf16 tile
  static constexpr int STAGES = (DH == 64) ? 4 : 2;
  static constexpr int SMEM = TILE * (1 + 2 * STAGES) + 256 + 2048 + 1024;   // tiles + barriers + row-max exchange + align
  // two S buffers (128 fp32 columns each); P_j (packed bf16, 64 columns) aliases the head of S_j once the row
  // threads hold S_j in registers; O after them
  static constexpr int COL_S = 0, COL_O = 256;
  static constexpr int TMEM_COLS = 512;
};

template <int DH>
// 320 threads: warps 0-7 = softmax (warp w and w+4 own the same 32 query rows / TMEM lanes and split the 128 key
// columns 64/64, so every SMSP has two softmax warps to interleave; row max / row sum halves meet through smem),
// warp 8 = TMA producer, warp 9 = MMA issuer + TMEM allocator.
__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int S, int H,
                float scale, int causal) {
  using C = FwdCfg<DH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + C::TILE;                   // [STAGES]
  uint8_t* sV = sK + C::STAGES * C::TILE;       // [STAGES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + C::STAGES * C::TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;                  // [STAGES]
  uint64_t* v_full = bars + 5;                  // [STAGES]
  uint64_t* kv_empty = bars + 9;                // [STAGES]
  uint64_t* s_full = bars + 13;                 // [2]
  uint64_t* p_ready = bars + 15;                // [2]
  uint64_t* o_done = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* xch = reinterpret_cast<float*>(bars + 20);   // [2 tiles][2 halves][128 rows] row-max exchange (+ reused for the final row sums)

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_qt = S / BQ;
  const int qt = n_qt - 1 - blockIdx.x;  // heaviest (longest causal prefix) tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int d_model = H * DH;
  const int n_kv = causal ? qt + 1 : S / BKV;
  const int row0 = b * S + qt * BQ;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmQKV);
    mbar_init(q_full, 1);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 256);
    }
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, C::TILE);
#pragma unroll
      for (int c = 0; c < C::CH; ++c) tma_load_2d(sQ + c * 16384, &tmQKV, q_full, h * DH + c * 64, row0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % C::STAGES;
        mbar_wait(&kv_empty[st], ((j / C::STAGES) & 1) ^ 1);
        mbar_expect_tx(&k_full[st], C::TILE);
#pragma unroll
        for (int c = 0; c < C::CH; ++c)
          tma_load_2d(sK + st * C::TILE + c * 16384, &tmQKV, &k_full[st], d_model + h * DH + c * 64, b * S + j * BKV);
        mbar_expect_tx(&v_full[st], C::TILE);
#pragma unroll
        for (int c = 0; c < C::CH; ++c)
          tma_load_2d(sV + st * C::TILE + c * 16384, &tmQKV, &v_full[st], 2 * d_model + h * DH + c * 64, b * S + j * BKV);
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BQ, BKV, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(BQ, DH, 0, 1);
      const uint32_t q_base = smem_u32(sQ);
      mbar_wait(q_full, 0);
      auto issue_s = [&](int j) {  // S_j = Q K_j^T into S buffer (j & 1)
        const int st = j % C::STAGES;
        const uint32_t k_base = smem_u32(sK + st * C::TILE);
        mbar_wait(&k_full[st], (j / C::STAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk) {
          const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
          tc_mma_f16_ss(tmem + C::COL_S + (j & 1) * 128, make_smem_desc_sw128(q_base + off, 16, 1024),
                        make_smem_desc_sw128(k_base + off, 16, 1024), idesc_s, kk != 0);
        }
        tc_commit(&s_full[j & 1]);
      };
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        // S_{j+1} is issued BEFORE waiting for softmax_j: it runs on the tensor pipe while the row threads work.
        // Buffer (j+1)&1 is free: softmax_{j-1} finished reading S_{j-1} (p_ready) and P_{j-1} is consumed by
        // P V_{j-1}, which was issued earlier on the in-order tensor pipe.
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j % C::STAGES;
        const uint32_t v_base = smem_u32(sV + st * C::TILE);
        mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        mbar_wait(&v_full[st], (j / C::STAGES) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk) {
          tc_mma_f16_ts(tmem + C::COL_O, tmem + C::COL_S + (j & 1) * 128 + kk * 8, make_smem_desc_sw128(v_base + kk * 2048, 16384, 1024),
                        idesc_o, (j | kk) != 0);
        }
        tc_commit(&kv_empty[st]);
        tc_commit(o_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax rows (warps 0..7)
    const int half = warp >> 2;                         // which 64 key columns of the row this thread owns
    const int row = (warp & 3) * 32 + lane;             // query row inside the tile == TMEM lane
    const uint32_t tl = tmem + (uint32_t((warp & 3) * 32) << 16);
    const float sc = scale * LOG2E;
    float m = -INFINITY, l = 0.f;                        // l = partial row sum over this thread's columns
    auto tile = [&](int j, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t ts = tl + C::COL_S + (j & 1) * 128;
      uint32_t r0[32], r1[32];                           // this thread's 64 S values, read from TMEM once
      tmem_ld_32x32(ts + half * 64, r0);
      tmem_ld_32x32(ts + half * 64 + 32, r1);
      tmem_ld_wait();
      if (DIAG) {  // causal mask: columns beyond this row -> -inf (exp2 -> 0)
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          if (half * 64 + t > row) r0[t] = 0xff800000u;
          if (half * 64 + 32 + t > row) r1[t] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(r0[t]), __uint_as_float(r0[t + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(r1[t]), __uint_as_float(r1[t + 1])));
      }
      // the two halves of a row agree on the row max through smem (double-buffered per tile parity)
      float* xb = xch + (j & 1) * 256;
      const float mine = fmaxf(mx0, mx1);
      xb[half * 128 + row] = mine;
      named_bar_sync(2, 256);
      const float mx = fmaxf(mine, xb[(half ^ 1) * 128 + row]);
      // lazy rescale: keep the old reference max unless it grew by more than 2^8 (bounded overflow, exact result)
      const float m_cand = fmaxf(m, mx * sc);
      const bool bump = (m_cand - m) > 8.0f;   // also true on the first tile (m = -inf)
      const float m_new = bump ? m_cand : m;
      const float alpha = bump ? exp2f(m - m_new) : 1.0f;
      float rs0 = 0.f, rs1 = 0.f;
      auto emit = [&](const uint32_t (&r)[32], int cc) {
        uint32_t pk[16];
#pragma unroll
        for (int t = 0; t < 32; t += 2) {
          const float p0 = exp2f(fmaf(__uint_as_float(r[t]), sc, -m_new));
          const float p1 = exp2f(fmaf(__uint_as_float(r[t + 1]), sc, -m_new));
          rs0 += p0;
          rs1 += p1;
          pk[t >> 1] = pack_bf16(p0, p1);
        }
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
                ts + half * 32 + cc * 16),
            "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]), "r"(pk[8]), "r"(pk[9]),
            "r"(pk[10]), "r"(pk[11]), "r"(pk[12]), "r"(pk[13]), "r"(pk[14]), "r"(pk[15])
            : "memory");
      };
      // NOTE: P (64 packed columns) aliases S columns [0,64): the OTHER half of this row may still be reading its S
      // values from columns [64,128) -> only columns [0,64) are overwritten, and half 0 has them in registers already;
      // half 1's stores land in [32,64), which half 0 read before the named barrier above.
      emit(r0, 0);
      emit(r1, 1);
      l = l * alpha + (rs0 + rs1);
      m = m_new;
      // rescale the running O only when some row of this warp actually moved its reference max (DH/2 columns per thread)
      if (j > 0) {
        if (__any_sync(0xffffffff, bump)) {
          mbar_wait(o_done, (j - 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = half * (DH / 64); c < (half + 1) * (DH / 64); ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tl + C::COL_O + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int t = 0; t < 32; ++t) r[t] = __float_as_uint(__uint_as_float(r[t]) * alpha);
            tmem_st_32x32(tl + C::COL_O + c * 32, r);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[j & 1]);
    };
    const int n_plain = causal ? n_kv - 1 : n_kv;
    for (int j = 0; j < n_plain; ++j) tile(j, std::false_type{});
    if (causal) tile(n_kv - 1, std::true_type{});
    // ---- epilogue: row sum = both halves; O / l -> bf16 -> global (DH/2 columns per thread); lse
    named_bar_sync(2, 256);                 // everyone is past its last max exchange: xch can be reused
    xch[half * 128 + row] = l;
    named_bar_sync(2, 256);
    const float l_tot = l + xch[(half ^ 1) * 128 + row];
    mbar_wait(o_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l_tot;
    __nv_bfloat16* orow = out + (long long)(row0 + row) * d_model + h * DH;
#pragma unroll 1
    for (int c = half * (DH / 64); c < (half + 1) * (DH / 64); ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tl + C::COL_O + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int t = 0; t < 32; t += 8) {
        uint4 o;
        o.x = pack_bf16(__uint_as_float(r[t]) * inv_l, __uint_as_float(r[t + 1]) * inv_l);
        o.y = pack_bf16(__uint_as_float(r[t + 2]) * inv_l, __uint_as_float(r[t + 3]) * inv_l);
        o.z = pack_bf16(__uint_as_float(r[t + 4]) * inv_l, __uint_as_float(r[t + 5]) * inv_l);
        o.w = pack_bf16(__uint_as_float(r[t + 6]) * inv_l, __uint_as_float(r[t + 7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c * 32 + t) = o;
      }
    }
    if (half == 0) lse[((long long)b * H + h) * S + qt * BQ + row] = m * LN2 + __logf(l_tot);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<C::TMEM_COLS>(tmem);
}

// ======================================================================================= backward
// delta[b,h,q] = sum_d dO * O   (row-wise; feeds dS = P * (dP - delta))
__global__ void attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout, float* __restrict__ delta,
                                      int S, int H, int DH, long long rows) {
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;  // one warp per (row, head)
  if (gw >= rows * H) return;
  const long long r = gw / H;
  const int h = int(gw % H);
  const __nv_bfloat16* po = o + r * (long long)(H * DH) + h * DH;
  const __nv_bfloat16* pd = dout + r * (long long)(H * DH) + h * DH;
  float acc = 0.f;
  for (int i = lane * 2; i < DH; i += 64) {
    const float2 a = unpack_bf16(*reinterpret_cast<const uint32_t*>(po + i));
    const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(pd + i));
    acc += a.x * g.x + a.y * g.y;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffff, acc, off);
  if (lane == 0) {
    const long long bb = r / S, s = r % S;
    delta[(bb * H + h) * S + s] = acc;
  }
}

// dq_acc (fp32, [B*S, H*DH]) * scale -> bf16 into the q third of dqkv
__global__ void attn_bwd_dq_convert_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, long long rows, int d, float scale) {
  const long long n4 = rows * d / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(dq_acc)[i];
    const long long e = i * 4, r = e / d, c = e % d;
    uint2 o;
    o.x = pack_bf16(v.x * scale, v.y * scale), o.y = pack_bf16(v.z * scale, v.w * scale);
    *reinterpret_cast<uint2*>(dqkv + r * 3 * d + c) = o;
  }
}

struct BwdCfg {
  static constexpr int DH = 64;
  static constexpr int TILE = 128 * 128;                    // [128 x 64] bf16
  static constexpr int PS_TILE = 2 * 128 * 128;             // [128 x 128] bf16 as two 64-col chunks
  static constexpr int DQ_STAGE = 128 * 64 * 4;             // fp32 [128 x 64] staging (two 32-col chunks)
  static constexpr int SMEM = 2 * TILE /*K,V*/ + 2 * 2 * TILE /*Q,dO x2 stages*/ + 2 * PS_TILE /*P,dS*/ + DQ_STAGE + 256 + 1024;
  static constexpr int COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 320, COL_DQ = 384;
  static constexpr int TMEM_COLS = 512;
};

// 320 threads: warps 0-7 = row threads (two per SMSP: warp w and w+4 own the same 32 query rows / TMEM lanes and
// split the 128 key columns 64/64 - no cross-thread exchange is needed because lse and delta are per-row inputs),
// warp 8 = TMA producer, warp 9 = MMA issuer + TMEM allocator.
__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmDQ,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int S, int H, float scale,
                int causal) {
  using C = BwdCfg;
  constexpr int DH = C::DH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + C::TILE;
  uint8_t* sQ = sV + C::TILE;            // [2]
  uint8_t* sDO = sQ + 2 * C::TILE;       // [2]
  uint8_t* sP = sDO + 2 * C::TILE;       // bf16 [128 q][128 kv] as 2 chunks of [128][64]
  uint8_t* sDS = sP + C::PS_TILE;
  uint8_t* sDQ = sDS + C::PS_TILE;       // fp32 staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDQ + C::DQ_STAGE);
  uint64_t* kv_full = bars;              // K_j and V_j
  uint64_t* qd_full = bars + 1;          // [2] Q_i + dO_i
  uint64_t* qd_empty = bars + 3;         // [2]
  uint64_t* sdp_full = bars + 5;         // S and dP ready
  uint64_t* pds_ready = bars + 6;        // P, dS in smem (128 arrivals)
  uint64_t* dq_full = bars + 7;          // dQ partial tile in TMEM
  uint64_t* dq_free = bars + 8;          // dQ TMEM columns drained (128 arrivals)
  uint64_t* mma_done = bars + 9;         // dV/dK/dQ MMAs of the tile retired -> P/dS smem reusable
  uint64_t* final_done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_t = S / 128;
  const int jt = blockIdx.x;  // key tile (small j = most query tiles, launched first)
  const int h = blockIdx.y, b = blockIdx.z;
  const int d_model = H * DH;
  const int i0 = causal ? jt : 0;
  const int n_it = n_t - i0;

  if (warp == 8 && lane == 0) {
    prefetch_tmap(&tmQKV);
    prefetch_tmap(&tmDO);
    prefetch_tmap(&tmDQ);
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qd_full[s], 1);
      mbar_init(&qd_empty[s], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 256);
    mbar_init(mma_done, 1);
    mbar_init(final_done, 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * C::TILE);
      tma_load_2d(sK, &tmQKV, kv_full, d_model + h * DH, b * S + jt * 128);
      tma_load_2d(sV, &tmQKV, kv_full, 2 * d_model + h * DH, b * S + jt * 128);
      for (int t = 0; t < n_it; ++t) {
        const int st = t & 1;
        mbar_wait(&qd_empty[st], ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(&qd_full[st], 2 * C::TILE);
        tma_load_2d(sQ + st * C::TILE, &tmQKV, &qd_full[st], h * DH, b * S + (i0 + t) * 128);
        tma_load_2d(sDO + st * C::TILE, &tmDO, &qd_full[st], h * DH, b * S + (i0 + t) * 128);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);   // S, dP : both K-major, N = 128
      constexpr uint32_t idesc_mm = make_idesc_bf16(128, DH, 1, 1);    // dV, dK: A (P/dS) MN-major, B (dO/Q) MN-major
      constexpr uint32_t idesc_km = make_idesc_bf16(128, DH, 0, 1);    // dQ    : A (dS) K-major, B (K_j) MN-major
      const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV), p_base = smem_u32(sP), ds_base = smem_u32(sDS);
      mbar_wait(kv_full, 0);
      auto issue_sdp = [&](int t) {  // S_t = Q_t K^T ; dP_t = dO_t V^T  (contraction over d_head = 64: 4 k-steps in one swizzle atom)
        const int st = t & 1;
        const uint32_t q_base = smem_u32(sQ + st * C::TILE), do_base = smem_u32(sDO + st * C::TILE);
        mbar_wait(&qd_full[st], (t >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          tc_mma_f16_ss(tmem + C::COL_S, make_smem_desc_sw128(q_base + kk * 32, 16, 1024), make_smem_desc_sw128(k_base + kk * 32, 16, 1024),
                        idesc_kk, kk != 0);
#pragma unroll
        for (int kk = 0; kk < DH / 16; ++kk)
          tc_mma_f16_ss(tmem + C::COL_DP, make_smem_desc_sw128(do_base + kk * 32, 16, 1024), make_smem_desc_sw128(v_base + kk * 32, 16, 1024),
                        idesc_kk, kk != 0);
        tc_commit(sdp_full);
      };
      issue_sdp(0);
      for (int t = 0; t < n_it; ++t) {
        const int st = t & 1;
        const uint32_t q_base = smem_u32(sQ + st * C::TILE), do_base = smem_u32(sDO + st * C::TILE);
        mbar_wait(pds_ready, t & 1);   // row threads consumed S_t / dP_t and staged P_t / dS_t in smem
        // next tile's S / dP go first: the row threads start on them while dV / dK / dQ of this tile run
        if (t + 1 < n_it) issue_sdp(t + 1);
        if (t > 0) mbar_wait(dq_free, (t - 1) & 1);
        tc_fence_after();
        // contraction over the 128 query rows of this tile: 8 k-steps, 16 rows (2048 B) each
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // dV[kv, dh] += P^T dO : A = P (MN-major: kv contiguous, 2 chunks of 64 -> LBO 16 KiB), B = dO (MN-major, N = 64)
          tc_mma_f16_ss(tmem + C::COL_DV, make_smem_desc_sw128(p_base + kk * 2048, 16384, 1024),
                        make_smem_desc_sw128(do_base + kk * 2048, 16384, 1024), idesc_mm, (t | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // dK[kv, dh] += dS^T Q
          tc_mma_f16_ss(tmem + C::COL_DK, make_smem_desc_sw128(ds_base + kk * 2048, 16384, 1024),
                        make_smem_desc_sw128(q_base + kk * 2048, 16384, 1024), idesc_mm, (t | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          // dQ[q, dh] = dS K_j : A = dS K-major (kv contiguous: chunk = kk/4, 32 B per k-step), B = K_j MN-major over kv rows
          tc_mma_f16_ss(tmem + C::COL_DQ, make_smem_desc_sw128(ds_base + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024),
                        make_smem_desc_sw128(k_base + kk * 2048, 16384, 1024), idesc_km, kk != 0);
        }
        tc_commit(&qd_empty[st]);
        tc_commit(dq_full);
        tc_commit(mma_done);
      }
      tc_commit(final_done);
    }
  } else {
    // ------------------------------------------------------------------ row threads (warps 0..7)
    const int half = warp >> 2;                      // which 64 key columns of the row this thread owns
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tl = tmem + (uint32_t((warp & 3) * 32) << 16);
    const float sc = scale * LOG2E;
    const uint32_t swz = row & 7;
    const bool issuer = (threadIdx.x == 0);
    auto drain_dq = [&](int t) {  // dQ partial of tile t: TMEM -> fp32 smem staging -> TMA reduce-add into dq_acc
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      if (issuer) tma_wait_read<0>();
      named_bar_sync(1, 256);
      {
        const int c = half;  // 32 of the 64 dQ columns per thread
        uint32_t r[32];
        tmem_ld_32x32(tl + C::COL_DQ + c * 32, r);
        tmem_ld_wait();
        uint8_t* qrow = sDQ + c * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(qrow + ((uint32_t(g) ^ swz) << 4)) = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
      }
      tc_fence_before();
      mbar_arrive(dq_free);
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (issuer) {
        tma_reduce_add_2d(&tmDQ, sDQ, h * DH, b * S + (i0 + t) * 128);
        tma_reduce_add_2d(&tmDQ, sDQ + 16384, h * DH + 32, b * S + (i0 + t) * 128);
        tma_commit();
      }
    };
    auto row_tile = [&](int t, auto diag_tag) {
      constexpr bool DIAG = decltype(diag_tag)::value;
      const int it = i0 + t;
      const long long grow = (long long)(b * H + h) * S + it * 128 + row;
      const float lse2 = lse[grow] * LOG2E;
      const float dl = delta[grow];
      mbar_wait(sdp_full, t & 1);
      tc_fence_after();
      uint32_t pk[32], dk[32];  // packed bf16 P and dS of this thread's 64 columns, kept in registers
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;  // 32-column block of the row
        uint32_t rs[32], rp[32];
        tmem_ld_32x32(tl + C::COL_S + c * 32, rs);
        tmem_ld_32x32(tl + C::COL_DP + c * 32, rp);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float p0 = exp2f(fmaf(__uint_as_float(rs[e]), sc, -lse2));
          float p1 = exp2f(fmaf(__uint_as_float(rs[e + 1]), sc, -lse2));
          if (DIAG) {
            if (c * 32 + e > row) p0 = 0.f;
            if (c * 32 + e + 1 > row) p1 = 0.f;
          }
          const float s0 = p0 * (__uint_as_float(rp[e]) - dl);
          const float s1 = p1 * (__uint_as_float(rp[e + 1]) - dl);
          pk[cc * 16 + (e >> 1)] = pack_bf16(p0, p1);
          dk[cc * 16 + (e >> 1)] = pack_bf16(s0, s1);
        }
      }
      // S_t / dP_t are consumed (the MMA warp may overwrite them); the smem staging is reusable once the
      // previous tile's dV / dK / dQ MMAs retired
      if (t > 0) mbar_wait(mma_done, (t - 1) & 1);
      {
        // this thread's 64 columns = swizzle chunk `half` of row `row`: eight 16-byte groups
        uint8_t* prow = sP + half * 16384 + row * 128;
        uint8_t* drow = sDS + half * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t slot = (uint32_t(g) ^ swz) << 4;
          *reinterpret_cast<uint4*>(prow + slot) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          *reinterpret_cast<uint4*>(drow + slot) = make_uint4(dk[4 * g], dk[4 * g + 1], dk[4 * g + 2], dk[4 * g + 3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_ready);
      if (t > 0) drain_dq(t - 1);   // overlaps with S_{t+1}/dP_{t+1} and dV/dK/dQ_t on the tensor pipe
    };
    // causal: only the first query tile (it == jt) touches the diagonal; every other tile runs the mask-free body
    for (int t = 0; t < n_it; ++t) {
      if (causal && t == 0) row_tile(t, std::true_type{});
      else row_tile(t, std::false_type{});
    }
    drain_dq(n_it - 1);
    // ---- epilogue: dK (x scale), dV -> bf16 -> dqkv
    mbar_wait(final_done, 0);
    tc_fence_after();
    __nv_bfloat16* krow = dqkv + (long long)(b * S + jt * 128 + row) * (3 * d_model) + d_model + h * DH;
    __nv_bfloat16* vrow = krow + d_model;
    {
      const int c = half;
      uint32_t rk[32], rv[32];
      tmem_ld_32x32(tl + C::COL_DK + c * 32, rk);
      tmem_ld_32x32(tl + C::COL_DV + c * 32, rv);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; e += 8) {
        uint4 ok, ov;
        ok.x = pack_bf16(__uint_as_float(rk[e]) * scale, __uint_as_float(rk[e + 1]) * scale);
        ok.y = pack_bf16(__uint_as_float(rk[e + 2]) * scale, __uint_as_float(rk[e + 3]) * scale);
        ok.z = pack_bf16(__uint_as_float(rk[e + 4]) * scale, __uint_as_float(rk[e + 5]) * scale);
        ok.w = pack_bf16(__uint_as_float(rk[e + 6]) * scale, __uint_as_float(rk[e + 7]) * scale);
        ov.x = pack_bf16(__uint_as_float(rv[e]), __uint_as_float(rv[e + 1]));
        ov.y = pack_bf16(__uint_as_float(rv[e + 2]), __uint_as_float(rv[e + 3]));
        ov.z = pack_bf16(__uint_as_float(rv[e + 4]), __uint_as_float(rv[e + 5]));
        ov.w = pack_bf16(__uint_as_float(rv[e + 6]), __uint_as_float(rv[e + 7]));
        *reinterpret_cast<uint4*>(krow + c * 32 + e) = ok;
        *reinterpret_cast<uint4*>(vrow + c * 32 + e) = ov;
      }
    }
    if (issuer) tma_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<C::TMEM_COLS>(tmem);
}

template <typename K>
void set_smem(K kern, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) throw std::runtime_error(std::string("attention smem attr: ") + cudaGetErrorString(e));
}

float* g_dq_acc = nullptr;
size_t g_dq_acc_bytes = 0;

}  // namespace

void attention_fwd_launch(const void* qkv, void* out, float* lse, int B, int S, int H, int dh, float scale, bool causal, int /*num_sms*/,
                          cudaStream_t st) {
  if (S % 128) throw std::runtime_error("photon_b200 attention: sequence length must be a multiple of 128");
  if (dh != 64 && dh != 128) throw std::runtime_error("photon_b200 attention: d_head must be 64 or 128");
  const int d = H * dh;
  CUtensorMap tm = make_tmap_2d(qkv, 2, false, uint64_t(3) * d, uint64_t(B) * S, uint64_t(3) * d * 2, 64, 128);
  dim3 grid(S / 128, H, B);
  if (dh == 64) {
    static bool once = (set_smem(attn_fwd_kernel<64>, FwdCfg<64>::SMEM), true);
    (void)once;
    attn_fwd_kernel<64><<<grid, 320, FwdCfg<64>::SMEM, st>>>(tm, (__nv_bfloat16*)out, lse, S, H, scale, causal ? 1 : 0);
  } else {
    static bool once = (set_smem(attn_fwd_kernel<128>, FwdCfg<128>::SMEM), true);
    (void)once;
    attn_fwd_kernel<128><<<grid, 320, FwdCfg<128>::SMEM, st>>>(tm, (__nv_bfloat16*)out, lse, S, H, scale, causal ? 1 : 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("attention fwd launch: ") + cudaGetErrorString(e));
}

int attention_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* delta, int B, int S, int H,
                         int dh, float scale, bool causal, int /*num_sms*/, cudaStream_t st) {
  if (S % 128) throw std::runtime_error("photon_b200 attention: sequence length must be a multiple of 128");
  if (dh != 64) throw std::runtime_error("photon_b200 attention backward: d_head must be 64 (use kernels.attention=torch otherwise)");
  const int d = H * dh;
  const long long rows = (long long)B * S;
  const size_t need = size_t(rows) * d * sizeof(float);
  if (need > g_dq_acc_bytes) {
    if (g_dq_acc) cudaFree(g_dq_acc);
    if (cudaMalloc(&g_dq_acc, need) != cudaSuccess) throw std::runtime_error("attention bwd: dq accumulator allocation failed");
    g_dq_acc_bytes = need;
  }
  cudaMemsetAsync(g_dq_acc, 0, need, st);
  {
    const long long warps = rows * H;
    const int threads = 256;
    attn_bwd_delta_kernel<<<int((warps * 32 + threads - 1) / threads), threads, 0, st>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout,
                                                                                         delta, S, H, dh, rows);
  }
  CUtensorMap tmQKV = make_tmap_2d(qkv, 2, false, uint64_t(3) * d, uint64_t(rows), uint64_t(3) * d * 2, 64, 128);
  CUtensorMap tmDO = make_tmap_2d(dout, 2, false, uint64_t(d), uint64_t(rows), uint64_t(d) * 2, 64, 128);
  CUtensorMap tmDQ = make_tmap_2d(g_dq_acc, 4, true, uint64_t(d), uint64_t(rows), uint64_t(d) * 4, 32, 128);
  static bool once = (set_smem(attn_bwd_kernel, BwdCfg::SMEM), true);
  (void)once;
  dim3 grid(S / 128, H, B);
  attn_bwd_kernel<<<grid, 320, BwdCfg::SMEM, st>>>(tmQKV, tmDO, tmDQ, lse, delta, (__nv_bfloat16*)dqkv, S, H, scale, causal ? 1 : 0);
  attn_bwd_dq_convert_kernel<<<148 * 8, 256, 0, st>>>(g_dq_acc, (__nv_bfloat16*)dqkv, rows, d, scale);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("attention bwd launch: ") + cudaGetErrorString(e));
  return 3;
}

}  // namespace pb
