// Persistent, warp-specialised bf16 GEMM for sm_100a:  D[M,N] = A[M,K] * B[N,K]^T  (+ fused epilogue)
//
//   * operands staged global->shared by TMA (cp.async.bulk.tensor, 128B swizzle), 3-stage mbarrier ring
//   * tcgen05.mma (kind::f16, M=128 N=256 K=16) issued by ONE thread, accumulators in TMEM
//     (2 x 256 fp32 columns, double-buffered so tile i's epilogue overlaps tile i+1's MMAs)
//   * epilogue warps: tcgen05.ld TMEM->registers, fused bias / residual / GELU / dGELU, pack,
//     swizzled st.shared, TMA store (or TMA reduce-add for fp32 weight-gradient accumulation)
//   * either operand may be K-major ([rows,K] row-major) or MN-major ([K,rows] row-major) so forward,
//     dgrad (B = W as stored) and wgrad (A = dY^T, B = X^T as stored) need no transposes.
//
// Replaces the cuBLASLt calls the reference reaches through torch (SURVEY §2.5 K3,K5-K8,K10).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "gemm_tcgen05.h"
#include "ptx.cuh"

namespace pb {

namespace {

constexpr int BM = 128, BN = 256;
// A stage holds 128 BYTES of K per row for either operand type: 64 bf16 (4 UMMA k-steps of 16) or 128 fp8 (4 k-steps of 32),
// so the shared-memory plan, the TMA box bytes and the descriptor arithmetic are the same for kind::f16 and kind::f8f6f4.
constexpr int ROW_BYTES = 128;
constexpr int A_STAGE = BM * ROW_BYTES;  // 16 KiB
constexpr int EPI_BUF = 128 * 128;    // 128 rows x 128 B (one swizzle atom wide)
constexpr int NUM_THREADS = 384;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue (two groups of four)
constexpr int TMEM_COLS = 512;

// Shared-memory plan per (epilogue, cluster) variant. A CTA of a pair stages only HALF of the B tile, so a stage is
// 32 instead of 48 KiB and the same 227 KiB hold a 5-deep instead of a 3-deep TMA pipeline (plus 64 KiB of epilogue
// staging): ~2500 instead of ~1500 tensor-core clocks of loads in flight per SM.
template <int EPI, int CL>
struct Cfg {
  static constexpr int B_STAGE = (BN / CL) * ROW_BYTES;   // 32 KiB (CL=1) / 16 KiB (CL=2)
  static constexpr int STAGE_BYTES = A_STAGE + B_STAGE;
  static constexpr int STORES = (EPI == EPI_GELU_DUAL || EPI == EPI_GELU_GRAD) ? 2 : 1;   // staging buffers consumed per epilogue chunk
  // two epilogue groups (even / odd column chunks of a tile), two staging buffers each: the exact-erf GELU epilogues
  // measured ALU-bound with one epilogue warp per SM sub-partition (dGELU dgrad 634 TFLOP/s against 1400+ for the
  // plain epilogue on the same shape)
  static constexpr int EPI_BUFS_G = 2;
  static constexpr int EPI_BUFS = 2 * EPI_BUFS_G;
  static constexpr int STAGES = (CL == 1) ? 3 : 5;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BUFS * EPI_BUF + 256 + 1024;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB dynamic shared memory of sm_100");
};

struct Params {
  int M, N, K;
  int tiles_m, tiles_n, m_fastest;
  const float* bias;            // [N] fp32 or nullptr
  const __nv_bfloat16* aux;     // residual (EPI_RESIDUAL) or pre-activation z (EPI_DGELU)
  long long ld_aux;             // row stride of aux, elements
  int accumulate;               // EPI_F32: reduce-add into D instead of overwrite
  float alpha;                  // scale applied to the accumulator before the epilogue
  int splits, kb_per_split;     // split-K (EPI_F32 + accumulate): work item = (tile, k-range)
  // fp8 (kind::f8f6f4) operands: element formats (0 = E4M3, 1 = E5M2) and the per-tensor de-scales, read from device memory
  // (delayed scaling keeps them on the device so the whole step stays CUDA-graph capturable)
  int a_fmt, b_fmt;
  const float* sinv_a;
  const float* sinv_b;
  // EPI_GELU_GRAD_Q8: the activation leaves as E4M3 (scaled by *out_scale, amax recorded) for the next fp8 GEMM
  uint8_t* d8;
  long long ld_d8;
  const float* out_scale;
  float* out_amax;
};

// CL = 1: one CTA per 128x256 tile (tcgen05 cta_group::1).
// CL = 2: a CTA PAIR (two SMs of one TPC, cluster 2x1x1) computes a 256x256 tile with ONE tcgen05.mma.cta_group::2
//         per k-step, issued by the leader CTA: each CTA stages its own 128 A rows and HALF of the B tile (128 of the
//         256 N rows), the tensor cores of both SMs read both halves, each SM accumulates its 128 rows in its own TMEM.
//         Per SM this halves the B traffic through shared memory (the 1-CTA form is smem-bandwidth bound: 96 B/clk of
//         MMA operand reads + 96 B/clk of TMA writes against a 128 B/clk port) and needs 32 instead of 48 KiB per stage.
template <int A_MN, int B_MN, int EPI, int CL, int F8>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2, const Params p) {
  using C = Cfg<EPI, CL>;
  constexpr int STAGES = C::STAGES, B_STAGE = C::B_STAGE, STAGE_BYTES = C::STAGE_BYTES, EPI_BUFS = C::EPI_BUFS;
  constexpr int BK = F8 ? 128 : 64;            // K elements per stage (128 bytes per row)
  constexpr int UK = F8 ? 32 : 16;             // K elements per tcgen05.mma
  constexpr int MNC = F8 ? 128 : 64;           // MN elements in one 128-byte row of an MN-major tile
  constexpr int CHUNK = BK * ROW_BYTES;        // bytes of one MN-major chunk: BK k-rows of 128 bytes
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_STAGE;
  uint8_t* sE = sB + STAGES * B_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(sE + EPI_BUFS * EPI_BUF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 24);   // own 64-byte line, away from the mbarrier words (<= 18 of them)

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmD);
    if (EPI == EPI_GELU_DUAL || EPI == EPI_GELU_GRAD || EPI == EPI_GELU_GRAD_Q8) prefetch_tmap(&tmD2);
  }
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], CL);   // CL = 2: leader's arrive.expect_tx + the peer's remote arrive (bytes of both land here)
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8 * CL);  // one arrival per epilogue warp (of both CTAs: the leader's MMA warp waits for the pair)
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CL > 1) tmem_alloc_2sm<TMEM_COLS>(tmem_slot);
    else tmem_alloc<TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peer barriers initialised before any remote arrive / 2-SM TMA / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work decomposition: units of CL M-adjacent tiles; every CTA of a cluster walks the same unit sequence
  const int tiles_mu = (p.tiles_m + CL - 1) / CL;
  const int num_tiles = tiles_mu * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  const int num_work = num_tiles * p.splits;
  const int w0 = blockIdx.x / CL, wstride = gridDim.x / CL;

  if (warp == 0) {
    // ======================= TMA producer (one thread) =======================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = w0; w < num_work; w += wstride) {
        const int tile = w % num_tiles, sp = w / num_tiles;
        const int m_blk = (p.m_fastest ? tile % tiles_mu : tile / p.tiles_n) * CL + crank;
        const int n_blk = p.m_fastest ? tile / tiles_mu : tile % p.tiles_n;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          if (CL > 1) {
            // CTA pair: stage = own A rows (16 KiB) + own half of B (16 KiB); all bytes are credited to the LEADER's
            // full barrier, which expects 2 x 32 KiB; the peer additionally arrives remotely
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* a_dst = sA + stage * A_STAGE;
            uint8_t* b_dst = sB + stage * B_STAGE;
            if (crank == 0) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / MNC; ++j)
                tma_load_2d_2sm(a_dst + j * CHUNK, &tmA, &full[stage], m_blk * BM + j * MNC, kb * BK);
            } else {
              tma_load_2d_2sm(a_dst, &tmA, &full[stage], kb * BK, m_blk * BM);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN / 2 / MNC; ++j)
                tma_load_2d_2sm(b_dst + j * CHUNK, &tmB, &full[stage], n_blk * BN + (crank * (BN / 2 / MNC) + j) * MNC, kb * BK);
            } else {
              tma_load_2d_2sm(b_dst, &tmB, &full[stage], kb * BK, n_blk * BN + crank * (BN / 2));
            }
            if (crank != 0) mbar_arrive_remote(&full[stage], 0);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
            continue;
          }
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          uint8_t* a_dst = sA + stage * A_STAGE;
          uint8_t* b_dst = sB + stage * B_STAGE;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / MNC; ++j)
              tma_load_2d(a_dst + j * CHUNK, &tmA, &full[stage], m_blk * BM + j * MNC, kb * BK);
          } else {
            tma_load_2d(a_dst, &tmA, &full[stage], kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN / MNC; ++j)
              tma_load_2d(b_dst + j * CHUNK, &tmB, &full[stage], n_blk * BN + j * MNC, kb * BK);
          } else {
            tma_load_2d(b_dst, &tmB, &full[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ======================= MMA issuer =======================
    // The whole warp walks the loop (waits included) so every operand is warp-uniform; one elected lane issues the
    // tcgen05 instructions. (`if (lane == 0)` around the loop costs a ~20-instruction elect/broadcast round trip per MMA.)
    if (crank == 0) {
      const uint32_t idesc = F8 ? make_idesc_f8(BM * CL, BN, A_MN, B_MN, p.a_fmt, p.b_fmt) : make_idesc_bf16(BM * CL, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int w = w0; w < num_work; w += wstride) {
        const int sp = w / num_tiles;
        const int kb0 = sp * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_base = smem_u32(sA + stage * A_STAGE);
          const uint32_t b_base = smem_u32(sB + stage * B_STAGE);
          const uint32_t ad0 = smem_desc_lo(a_base, A_MN ? CHUNK : 16);
          const uint32_t bd0 = smem_desc_lo(b_base, B_MN ? CHUNK : 16);
#pragma unroll
          for (int k = 0; k < 4; ++k) {   // BK / UK = 4 for both operand types
            // k-step inside the stage: UK rows of 128 B (MN-major) or 32 bytes of K (K-major), in 16-byte descriptor units
            const uint32_t adesc = ad0 + uint32_t(A_MN ? (k * UK * 128) >> 4 : (k * 32) >> 4);
            const uint32_t bdesc = bd0 + uint32_t(B_MN ? (k * UK * 128) >> 4 : (k * 32) >> 4);
            if (elect_one()) {
              const uint32_t accum = (kb > kb0 || k > 0) ? 1u : 0u;
              if (F8) {
                if (CL > 1) tc_mma_f8_ss_2sm_lo(d_tmem, adesc, bdesc, idesc, accum);
                else tc_mma_f8_ss_lo(d_tmem, adesc, bdesc, idesc, accum);
              } else {
                if (CL > 1) tc_mma_f16_ss_2sm_lo(d_tmem, adesc, bdesc, idesc, accum);
                else tc_mma_f16_ss_lo(d_tmem, adesc, bdesc, idesc, accum);
              }
            }
          }
          if (elect_one()) {
            if (CL > 1) tc_commit_2sm(&empty[stage], 0x3);  // release the stage in BOTH CTAs
            else tc_commit(&empty[stage]);                  // frees the smem slot once these MMAs retire
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) {
          if (CL > 1) tc_commit_2sm(&tfull[acc], 0x3);  // accumulator complete -> both CTAs' epilogues
          else tc_commit(&tfull[acc]);
        }
        __syncwarp();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ======================= epilogue (2 groups x 4 warps, one TMEM lane quadrant each) =======================
    // group g takes the column chunks c = g, g+2, ... of every tile with its own staging buffers, named barrier and
    // TMA-store issuer: two epilogue warps per SM sub-partition
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int row_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128 + grp * 128);
    uint8_t* sEg = sE + grp * (C::EPI_BUFS_G * EPI_BUF);
    constexpr int EPI_BUFS_G = C::EPI_BUFS_G;
    constexpr int CW = (EPI == EPI_F32) ? 32 : 64;             // columns per staged chunk (128 B wide)
    constexpr int STORES = C::STORES;                           // staging buffers consumed per chunk
    int ebuf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t swz = (row_in_tile & 7);
    // fp8 operands: fold the per-tensor de-scales (device scalars written by the quantising producers) into alpha
    float alpha = p.alpha;
    if (F8) alpha *= (p.sinv_a ? __ldg(p.sinv_a) : 1.0f) * (p.sinv_b ? __ldg(p.sinv_b) : 1.0f);
    const float q8_scale = (EPI == EPI_GELU_GRAD_Q8 && p.out_scale) ? __ldg(p.out_scale) : 1.0f;
    float q8_amax = 0.f;
    for (int w = w0; w < num_work; w += wstride) {
      const int tile = w % num_tiles;
      const int m_blk = (p.m_fastest ? tile % tiles_mu : tile / p.tiles_n) * CL + crank;
      const int n_blk = p.m_fastest ? tile / tiles_mu : tile % p.tiles_n;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const long long row = (long long)m_blk * BM + row_in_tile;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c = grp; c < BN / CW; c += 2) {
        const int col0 = n_blk * BN + c * CW;
        if (col0 >= p.N) break;  // whole chunk out of range (uniform)
        // aux tile (residual / pre-activation) for this chunk: issue the global loads FIRST so their latency
        // overlaps the TMEM read below
        uint4 ax[8];
        if (EPI == EPI_RESIDUAL || EPI == EPI_DGELU || EPI == EPI_MUL) {
          const __nv_bfloat16* arow = p.aux + row * p.ld_aux + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ax[j] = (row_ok && col0 + 8 * j < p.N) ? __ldg(reinterpret_cast<const uint4*>(arow + 8 * j)) : make_uint4(0, 0, 0, 0);
        }
        float v[CW];
        {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + c * CW;
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * alpha;
          if (CW == 64) {
            tmem_ld_32x32(taddr + 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[(CW == 64 ? 32 : 0) + j] = __uint_as_float(r[j]) * alpha;
          }
        }
        if (EPI != EPI_F32 && p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            if (col0 + j < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
        }
        uint8_t* buf0 = sEg + ebuf * EPI_BUF;
        uint8_t* buf1 = sEg + ((ebuf + 1) % EPI_BUFS_G) * EPI_BUF;
        if (issuer) tma_wait_read<EPI_BUFS_G / STORES - 1>();
        named_bar_sync(1 + grp, 128);
        if (EPI == EPI_F32) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            *reinterpret_cast<float4*>(buf0 + row_in_tile * 128 + ((j ^ swz) << 4)) = o;
          }
        } else {
          if (EPI == EPI_RESIDUAL || EPI == EPI_DGELU || EPI == EPI_MUL) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              {
                const uint4 a = ax[j];
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  const float2 f = unpack_bf16(w[t]);
                  if (EPI == EPI_RESIDUAL) {
                    v[8 * j + 2 * t] += f.x;
                    v[8 * j + 2 * t + 1] += f.y;
                  } else if (EPI == EPI_MUL) {
                    v[8 * j + 2 * t] *= f.x;
                    v[8 * j + 2 * t + 1] *= f.y;
                  } else {
                    v[8 * j + 2 * t] *= gelu_grad(f.x);
                    v[8 * j + 2 * t + 1] *= gelu_grad(f.y);
                  }
                }
              }
            }
          }
          if (EPI == EPI_GELU_GRAD) {
            // gelu and its derivative share the CDF / PDF evaluation; the derivative replaces the pre-activation as the
            // tensor saved for backward, so the dgrad epilogue is a plain multiply
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float d[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                float c, pd;
                gelu_cdf_pdf(v[8 * j + t], c, pd);
                d[t] = fmaf(v[8 * j + t], pd, c);
                v[8 * j + t] *= c;
              }
              uint4 o;
              o.x = pack_bf16(d[0], d[1]), o.y = pack_bf16(d[2], d[3]), o.z = pack_bf16(d[4], d[5]), o.w = pack_bf16(d[6], d[7]);
              *reinterpret_cast<uint4*>(buf1 + row_in_tile * 128 + ((j ^ swz) << 4)) = o;   // gelu'(z)
            }
          }
          if (EPI == EPI_GELU_GRAD_Q8) {
            // gelu'(z) -> v (staged + TMA-stored as bf16 below, through tmD2); gelu(z) -> E4M3 straight to global memory:
            // each thread owns 64 consecutive bytes of one output row (two full 32-byte sectors)
            uint32_t q[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float a4[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                float c, pd;
                const float z = v[4 * j + t];
                gelu_cdf_pdf(z, c, pd);
                v[4 * j + t] = fmaf(z, pd, c);
                a4[t] = z * c;
                if (row_ok && col0 + 4 * j + t < p.N) q8_amax = fmaxf(q8_amax, fabsf(a4[t]));
              }
              q[j] = pack_e4m3x4(a4[0] * q8_scale, a4[1] * q8_scale, a4[2] * q8_scale, a4[3] * q8_scale);
            }
            if (row_ok) {
              uint8_t* drow = p.d8 + row * p.ld_d8 + col0;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (col0 + 16 * j < p.N) *reinterpret_cast<uint4*>(drow + 16 * j) = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
            }
          }
          if (EPI == EPI_GELU_DUAL) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 o;
              o.x = pack_bf16(v[8 * j], v[8 * j + 1]);
              o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
              o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
              *reinterpret_cast<uint4*>(buf1 + row_in_tile * 128 + ((j ^ swz) << 4)) = o;  // z (pre-activation)
            }
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = gelu_exact(v[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 o;
            o.x = pack_bf16(v[8 * j], v[8 * j + 1]);
            o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(buf0 + row_in_tile * 128 + ((j ^ swz) << 4)) = o;
          }
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (issuer) {
          if (EPI == EPI_F32 && (p.accumulate || p.splits > 1)) {
            tma_reduce_add_2d(&tmD, buf0, col0, m_blk * BM);
          } else if (EPI == EPI_GELU_GRAD_Q8) {
            tma_store_2d(&tmD2, buf0, col0, m_blk * BM);
          } else {
            tma_store_2d(&tmD, buf0, col0, m_blk * BM);
          }
          if (EPI == EPI_GELU_DUAL || EPI == EPI_GELU_GRAD) tma_store_2d(&tmD2, buf1, col0, m_blk * BM);
          tma_commit();
        }
        ebuf = (ebuf + STORES) % EPI_BUFS_G;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CL > 1 && crank != 0) mbar_arrive_remote(&tempty[acc], 0);
        else mbar_arrive(&tempty[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (issuer) tma_wait_all<0>();
    if (EPI == EPI_GELU_GRAD_Q8 && p.out_amax) {
      q8_amax = warp_max_f(q8_amax);
      if (lane == 0 && q8_amax > 0.f) atomic_max_pos_f32(p.out_amax, q8_amax);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody leaves (or frees TMEM) while the pair's MMAs / remote arrives are in flight
  if (warp == 2) {
    if (CL > 1) tmem_dealloc_2sm<TMEM_COLS>(tmem_base);
    else tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------ host
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
      throw std::runtime_error("photon_b200: cuTensorMapEncodeTiled not available from the driver");
    fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

}  // namespace

CUtensorMap make_tmap_2d(const void* ptr, int elem_bytes, bool is_float32, uint64_t inner, uint64_t outer,
                         uint64_t ld_bytes, uint32_t box_inner, uint32_t box_outer) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld_bytes & 15))
    throw std::runtime_error("photon_b200 TMA: base pointer and row stride must be 16-byte aligned");
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = is_float32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
  CUresult r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw std::runtime_error("photon_b200: cuTensorMapEncodeTiled failed, code " + std::to_string(int(r)));
  return m;
}

template <int A_MN, int B_MN, int EPI, int CL, int F8>
static void launch_cl(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& td2,
                      const Params& p, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, EPI, CL, F8>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<EPI, CL>::SMEM_BYTES);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm smem attr: ") + cudaGetErrorString(e));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = Cfg<EPI, CL>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, td, td2, p);
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm launch: ") + cudaGetErrorString(e));
}

template <int A_MN, int B_MN, int EPI, int F8>
static void launch_inst(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& td2,
                        const Params& p, int grid, int cluster, cudaStream_t stream) {
  if (cluster == 2) launch_cl<A_MN, B_MN, EPI, 2, F8>(ta, tb, td, td2, p, grid, stream);
  else launch_cl<A_MN, B_MN, EPI, 1, F8>(ta, tb, td, td2, p, grid, stream);
}

// Operands are bf16 (kind::f16) or, with g.fp8, 8-bit floats (kind::f8f6f4: E4M3 / E5M2 chosen per operand, fp32 accumulate,
// per-tensor de-scales applied in the epilogue). Everything else — pipeline, tile shape, epilogues — is shared.
void gemm_launch(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return;
  if (g.N % 8) throw std::runtime_error("photon_b200 gemm: N must be a multiple of 8");
  const bool f32 = g.epi == EPI_F32;
  const int eb = g.fp8 ? 1 : 2;                 // operand element bytes
  const int BKE = ROW_BYTES / eb;               // K elements per stage
  const int MNC = ROW_BYTES / eb;               // MN elements per 128-byte row of an MN-major tile
  // A: K-major -> dims {K, M}, box {128 B of K, 128 rows};  MN-major -> dims {M, K}, box {128 B of M, BKE k-rows}
  CUtensorMap ta = g.a_mn ? make_tmap_2d(g.A, eb, false, g.M, g.K, g.lda * eb, MNC, BKE)
                          : make_tmap_2d(g.A, eb, false, g.K, g.M, g.lda * eb, BKE, BM);
  // 2-CTA clusters (each CTA stages half of the B tile) whenever there are >= 2 M tiles
  const int cluster = (g.cluster > 0 ? g.cluster : ((g.M + BM - 1) / BM >= 2 ? 2 : 1));
  CUtensorMap tb = g.b_mn ? make_tmap_2d(g.B, eb, false, g.N, g.K, g.ldb * eb, MNC, BKE)
                          : make_tmap_2d(g.B, eb, false, g.K, g.N, g.ldb * eb, BKE, BN / cluster);
  const bool two_out = g.epi == EPI_GELU_DUAL || g.epi == EPI_GELU_GRAD || g.epi == EPI_GELU_GRAD_Q8;
  CUtensorMap td2{};
  if (two_out) {
    if (!g.D2) throw std::runtime_error("photon_b200 gemm: the two-output GELU epilogues need the second output");
    td2 = make_tmap_2d(g.D2, 2, false, g.N, g.M, g.ldd2 * 2, 64, BM);
  }
  CUtensorMap td;
  if (g.epi == EPI_GELU_GRAD_Q8) {
    if (!g.fp8 || !g.D8 || (g.ldd8 % 16) || (reinterpret_cast<uintptr_t>(g.D8) & 15))
      throw std::runtime_error("photon_b200 gemm: the fp8-output epilogue needs fp8 operands and a 16-byte aligned fp8 output");
    td = td2;   // the primary output leaves by direct stores
  } else {
    td = f32 ? make_tmap_2d(g.D, 4, true, g.N, g.M, g.ldd * 4, 32, BM) : make_tmap_2d(g.D, 2, false, g.N, g.M, g.ldd * 2, 64, BM);
  }
  if (!two_out) td2 = td;
  if ((g.epi == EPI_RESIDUAL || g.epi == EPI_DGELU || g.epi == EPI_MUL) && (!g.aux || (g.ld_aux % 8) || (reinterpret_cast<uintptr_t>(g.aux) & 15)))
    throw std::runtime_error("photon_b200 gemm: aux operand missing or not 16-byte aligned");
  if (g.bias && (reinterpret_cast<uintptr_t>(g.bias) & 15)) throw std::runtime_error("photon_b200 gemm: bias must be 16-byte aligned");
  Params p{};
  p.M = g.M, p.N = g.N, p.K = g.K;
  p.tiles_m = (g.M + BM - 1) / BM;
  p.tiles_n = (g.N + BN - 1) / BN;
  // keep the larger operand streaming once: iterate the other dimension fastest
  p.m_fastest = (double(g.N) > double(g.M)) ? 1 : 0;
  p.bias = g.bias;
  p.aux = reinterpret_cast<const __nv_bfloat16*>(g.aux);
  p.ld_aux = g.ld_aux;
  p.accumulate = g.accumulate;
  p.alpha = g.alpha;
  p.a_fmt = g.a_fmt, p.b_fmt = g.b_fmt, p.sinv_a = g.sinv_a, p.sinv_b = g.sinv_b;
  p.d8 = reinterpret_cast<uint8_t*>(g.D8), p.ld_d8 = g.ldd8, p.out_scale = g.out_scale, p.out_amax = g.out_amax;
  int sms = g.num_sms;
  if (sms <= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (g.K + BKE - 1) / BKE;
  p.splits = 1;
  if (f32 && g.accumulate && tiles < sms && num_kb >= 16) {
    // few output tiles, long K (weight gradients): split K across CTAs, partials meet in the fp32
    // gradient bucket through TMA reduce-add. Aim for ~2 waves, >= 8 k-blocks per split.
    int s = (2 * sms + tiles / 2) / tiles;
    if (s > num_kb / 8) s = num_kb / 8;
    if (s < 1) s = 1;
    p.splits = s;
    if (const char* e = std::getenv("PB_GEMM_SPLITS")) {  // tuning knob (scripts/bench_kernels.py sweeps it)
      const int v = std::atoi(e);
      if (v >= 1 && v <= num_kb) p.splits = v;
    }
  }
  p.kb_per_split = (num_kb + p.splits - 1) / p.splits;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;  // drop empty tails
  const int work_units = ((p.tiles_m + cluster - 1) / cluster) * p.tiles_n * p.splits;
  int grid = work_units * cluster < sms ? work_units * cluster : (sms / cluster) * cluster;

#define PB_CASE(AM, BMJ, E, F)                                              \
  if (g.a_mn == AM && g.b_mn == BMJ && g.epi == E && int(g.fp8) == F) {     \
    launch_inst<AM, BMJ, E, F>(ta, tb, td, td2, p, grid, cluster, stream);  \
    return;                                                                 \
  }
  PB_CASE(0, 0, EPI_BF16, 0)
  PB_CASE(0, 0, EPI_RESIDUAL, 0)
  PB_CASE(0, 0, EPI_GELU_DUAL, 0)
  PB_CASE(0, 0, EPI_F32, 0)
  PB_CASE(0, 1, EPI_BF16, 0)
  PB_CASE(0, 1, EPI_DGELU, 0)
  PB_CASE(0, 1, EPI_MUL, 0)
  PB_CASE(0, 0, EPI_GELU_GRAD, 0)
  PB_CASE(0, 1, EPI_F32, 0)
  PB_CASE(1, 1, EPI_F32, 0)
  PB_CASE(1, 1, EPI_BF16, 0)
  PB_CASE(1, 0, EPI_F32, 0)
  PB_CASE(1, 0, EPI_BF16, 0)
  // fp8 operands: forward (K-major x K-major), dgrad (dY K-major x W MN-major), wgrad (both MN-major)
  PB_CASE(0, 0, EPI_BF16, 1)
  PB_CASE(0, 0, EPI_RESIDUAL, 1)
  PB_CASE(0, 0, EPI_GELU_GRAD, 1)
  PB_CASE(0, 0, EPI_GELU_GRAD_Q8, 1)
  PB_CASE(0, 0, EPI_F32, 1)
  PB_CASE(0, 1, EPI_BF16, 1)
  PB_CASE(0, 1, EPI_MUL, 1)
  PB_CASE(0, 1, EPI_F32, 1)
  PB_CASE(1, 1, EPI_F32, 1)
  PB_CASE(1, 1, EPI_BF16, 1)
#undef PB_CASE
  throw std::runtime_error("photon_b200 gemm: unsupported (a_major, b_major, epilogue, operand type) combination");
}

void gemm_bf16_launch(const GemmArgs& g, cudaStream_t stream) { gemm_launch(g, stream); }

}  // namespace pb
