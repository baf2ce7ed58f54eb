// Block-scaled FP8 GEMM for sm_100a:  D[M,N] = (A .* SFA)[M,K] * (B .* SFB)[N,K]^T (+ bias), bf16 out
//
//   tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale — OCP MXFP8: E4M3 elements, one E8M0 scale per 32 consecutive K elements
//   of every row of A and of B, applied INSIDE the tensor core (the scale factors are operands of the instruction, read from TMEM).
//
//   * A / B tiles: TMA (128-byte swizzle) into a 3-stage mbarrier ring, 128 K-elements per stage, as in gemm_tcgen05.cu;
//   * scale factors: stored in global memory already in the layout the tensor core wants — per (128 rows x 128 K) block a 512-byte
//     atom [32][4][4]: byte (r % 32) * 16 + (r / 32) * 4 + (k / 32) holds the scale of row r, K-group k/32 — so a stage's scales
//     are ONE contiguous bulk copy per operand (cp.async.bulk, counted on the same mbarrier as the tiles); from shared memory they
//     go to TMEM with tcgen05.cp.32x128b.warpx4 (4 columns per 128 rows), issued by the MMA thread right before the stage's four
//     MMAs (cp and mma execute in issue order, so one TMEM scale slot serves every stage);
//   * instruction descriptor: scale format E8M0, a_sf_id / b_sf_id = which of the 4 bytes of the TMEM word (K-group inside the stage);
//   * accumulator: 256 fp32 TMEM columns (the scale factors take 12 more, so there is ONE accumulator stage here — the per-tensor
//     fp8 kernel keeps the double-buffered, CTA-pair pipeline); epilogue: tcgen05.ld -> bias -> bf16 -> swizzled smem -> TMA store.
//
// mx_quantize_kernel produces both the E4M3 data and the scale atoms from a bf16 matrix in one pass (scale = 2^(floor(log2 amax) - 8)).
// North-star item "block-scaled fp8 for the GEMMs"; numerics oracle: de-quantise and multiply in fp32 (tests/test_fp8_gpu.py).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "gemm_tcgen05.h"
#include "ptx.cuh"

namespace pb {

namespace {

constexpr int BM = 128, BN = 256, BK = 128;             // BK in elements = bytes
constexpr int A_STAGE = BM * 128, B_STAGE = BN * 128;    // 16 KiB, 32 KiB
constexpr int SFA_BYTES = 512, SFB_BYTES = 1024;         // one / two 128-row scale atoms per stage
constexpr int STAGE_BYTES = A_STAGE + B_STAGE + SFA_BYTES + SFB_BYTES;
constexpr int STAGES = 3;
constexpr int EPI_BUF = 128 * 128, EPI_BUFS = 2;
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) + STAGES * (SFA_BYTES + SFB_BYTES) + EPI_BUFS * EPI_BUF + 256 + 1024;
constexpr int NUM_THREADS = 256;   // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int TMEM_COLS = 512;
constexpr int COL_SFA = 256, COL_SFB = 260;

struct MxParams {
  int M, N, K, tiles_m, tiles_n, mblk128, nblk128;
  const uint8_t* sfa;   // [K/128][ceil(M/128)][512]
  const uint8_t* sfb;   // [K/128][ceil(N/128) padded to even][512]
  const float* bias;
};

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared-memory descriptor of a scale atom for tcgen05.cp: 32 rows of 16 bytes, no swizzle, 8-row groups 128 bytes apart
__device__ __forceinline__ uint64_t sf_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(16 >> 4) << 16;     // leading byte offset (one 16-byte column only)
  d |= static_cast<uint64_t>(128 >> 4) << 32;    // stride byte offset between 8-row groups
  d |= static_cast<uint64_t>(1) << 46;           // descriptor version 1 (Blackwell); layout type 0 = no swizzle
  return d;
}
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void tc_mma_mxf8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                               uint32_t sfa_tmem, uint32_t sfb_tmem) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa_tmem), "r"(sfb_tmem)
      : "memory");
}
// kind::mxf8f6f4.block_scale instruction descriptor: E4M3 x E4M3, both K-major, E8M0 scales
__device__ __forceinline__ uint32_t make_idesc_mx(int M, int N, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4)                   // B scale-factor id (byte of the TMEM word)
         | (0u << 7) | (0u << 10)         // A, B format = E4M3
         | (uint32_t(N >> 3) << 17)       // N / 8
         | (1u << 23)                     // scale format = E8M0
         | (uint32_t(M >> 4) << 24)       // M / 16
         | (a_sf_id << 29);               // A scale-factor id
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_mx_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
               const MxParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + STAGES * A_STAGE;
  uint8_t* sE = sB + STAGES * B_STAGE;
  uint8_t* sSFA = sE + EPI_BUFS * EPI_BUF;
  uint8_t* sSFB = sSFA + STAGES * SFA_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sSFB + STAGES * SFB_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 16);

  const int warp = __shfl_sync(0xffffffff, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.tiles_m * p.tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          tma_load_2d(sA + stage * A_STAGE, &tmA, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(sB + stage * B_STAGE, &tmB, &full[stage], kb * BK, n_blk * BN);
          bulk_load(sSFA + stage * SFA_BYTES, p.sfa + ((long long)kb * p.mblk128 + m_blk) * 512, SFA_BYTES, &full[stage]);
          bulk_load(sSFB + stage * SFB_BYTES, p.sfb + ((long long)kb * p.nblk128 + 2 * n_blk) * 512, SFB_BYTES, &full[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      mbar_wait(tempty, acc_phase ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * A_STAGE), b_base = smem_u32(sB + stage * B_STAGE);
        if (elect_one()) {
          // this stage's scale factors: shared memory -> TMEM (ordered before the MMAs below by the tcgen05 pipe)
          tmem_cp_32x128b_warpx4(tmem_base + COL_SFA, sf_smem_desc(smem_u32(sSFA + stage * SFA_BYTES)));
          tmem_cp_32x128b_warpx4(tmem_base + COL_SFB, sf_smem_desc(smem_u32(sSFB + stage * SFB_BYTES)));
          tmem_cp_32x128b_warpx4(tmem_base + COL_SFB + 4, sf_smem_desc(smem_u32(sSFB + stage * SFB_BYTES + 512)));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = (uint64_t(kDescHiSw128) << 32) | (smem_desc_lo(a_base, 16) + uint32_t((k * 32) >> 4));
            const uint64_t bdesc = (uint64_t(kDescHiSw128) << 32) | (smem_desc_lo(b_base, 16) + uint32_t((k * 32) >> 4));
            tc_mma_mxf8_ss(tmem_base, adesc, bdesc, make_idesc_mx(BM, BN, k, k), (kb > 0 || k > 0) ? 1u : 0u, tmem_base + COL_SFA,
                           tmem_base + COL_SFB);
          }
          tc_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) tc_commit(tfull);
      __syncwarp();
      acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    const uint32_t swz = (row_in_tile & 7);
    int ebuf = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m_blk = t % p.tiles_m, n_blk = t / p.tiles_m;
      mbar_wait(tfull, acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 64; ++c) {
        const int col0 = n_blk * BN + c * 64;
        if (col0 >= p.N) break;
        float v[64];
        {
          uint32_t r[32];
          const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + c * 64;
          tmem_ld_32x32(taddr, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          tmem_ld_32x32(taddr + 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[32 + j] = __uint_as_float(r[j]);
        }
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            if (col0 + j < p.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
        }
        uint8_t* buf = sE + ebuf * EPI_BUF;
        if (issuer) tma_wait_read<EPI_BUFS - 1>();
        named_bar_sync(1, 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_bf16(v[8 * j], v[8 * j + 1]);
          o.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
          o.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
          o.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
          *reinterpret_cast<uint4*>(buf + row_in_tile * 128 + ((j ^ swz) << 4)) = o;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          tma_store_2d(&tmD, buf, col0, m_blk * BM);
          tma_commit();
        }
        ebuf = (ebuf + 1) % EPI_BUFS;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
      acc_phase ^= 1;
    }
    if (issuer) tma_wait_all<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// bf16 [R, K] (row stride ld) -> E4M3 [R, K] + E8M0 scale atoms [K/128][ceil(R/128) (padded to rblk)][512]; one thread = 8 elements,
// four neighbouring threads = one 32-element scale block.
__global__ void __launch_bounds__(256) mx_quantize_kernel(const __nv_bfloat16* __restrict__ x, long long ld, uint8_t* __restrict__ q,
                                                          uint8_t* __restrict__ sf, long long R, int K, int rblk) {
  const long long groups_per_row = K / 8;
  const long long total = R * groups_per_row;
  for (long long g = blockIdx.x * 256ll + threadIdx.x; g < ((total + 3) & ~3ll); g += gridDim.x * 256ll) {
    const bool ok = g < total;
    const long long r = ok ? g / groups_per_row : 0;
    const int k8 = ok ? int(g % groups_per_row) : 0;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (ok) {
      const uint4 u = *reinterpret_cast<const uint4*>(x + r * ld + k8 * 8);
      const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
      f[0] = a.x, f[1] = a.y, f[2] = b.x, f[3] = b.y, f[4] = c.x, f[5] = c.y, f[6] = d.x, f[7] = d.y;
    }
    float am = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) am = fmaxf(am, fabsf(f[i]));
    am = fmaxf(am, __shfl_xor_sync(0xffffffff, am, 1));
    am = fmaxf(am, __shfl_xor_sync(0xffffffff, am, 2));
    // OCP MX: shared exponent = floor(log2(amax)) - emax(E4M3 = 8); stored biased by 127; amax = 0 -> smallest scale
    int e = am > 0.f ? (int((__float_as_uint(am) >> 23) & 0xFF) - 127 - 8) : -127;
    e = e < -127 ? -127 : (e > 127 ? 127 : e);
    const float inv = __uint_as_float(uint32_t(127 - e) << 23);   // 2^-e (e in [-127, 127] -> exponent field 0..254; 0 -> denormal 0 handled below)
    const float s = (127 - e) == 0 ? 0.f : inv;
    if (ok) {
      uint2 o;
      o.x = pack_e4m3x4(f[0] * s, f[1] * s, f[2] * s, f[3] * s);
      o.y = pack_e4m3x4(f[4] * s, f[5] * s, f[6] * s, f[7] * s);
      *reinterpret_cast<uint2*>(q + r * (long long)K + k8 * 8) = o;
      if ((k8 & 3) == 0) {
        const int kg = k8 / 4;                   // 32-element group index along K
        const long long blk = (long long)(kg / 4) * rblk + r / 128;
        const int rr = int(r % 128);
        sf[blk * 512 + (rr % 32) * 16 + (rr / 32) * 4 + (kg % 4)] = uint8_t(e + 127);
      }
    }
  }
}

}  // namespace

void mx_quantize(const void* x_bf16, long long ld, void* q8, void* sf, long long R, int K, int rblk, cudaStream_t st) {
  if (K % 128) throw std::runtime_error("mx_quantize: K must be a multiple of 128");
  const long long groups = R * (K / 8);
  long long blocks = (groups + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mx_quantize_kernel<<<int(blocks), 256, 0, st>>>((const __nv_bfloat16*)x_bf16, ld, (uint8_t*)q8, (uint8_t*)sf, R, K, rblk);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("mx_quantize launch: ") + cudaGetErrorString(e));
}

void gemm_mxfp8_launch(const void* A, const void* B, void* D, const void* sfa, const void* sfb, const float* bias, int M, int N, int K,
                       long long ldd, int num_sms, cudaStream_t stream) {
  if (K % 128 || N % 8) throw std::runtime_error("gemm_mxfp8: K must be a multiple of 128 and N of 8");
  CUtensorMap ta = make_tmap_2d(A, 1, false, K, M, K, BK, BM);
  CUtensorMap tb = make_tmap_2d(B, 1, false, K, N, K, BK, BN);
  CUtensorMap td = make_tmap_2d(D, 2, false, N, M, ldd * 2, 64, BM);
  MxParams p{};
  p.M = M, p.N = N, p.K = K;
  p.tiles_m = (M + BM - 1) / BM, p.tiles_n = (N + BN - 1) / BN;
  p.mblk128 = p.tiles_m, p.nblk128 = 2 * p.tiles_n;
  p.sfa = reinterpret_cast<const uint8_t*>(sfa), p.sfb = reinterpret_cast<const uint8_t*>(sfb), p.bias = bias;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_mxfp8 smem attr: ") + cudaGetErrorString(e));
    configured = true;
  }
  int sms = num_sms > 0 ? num_sms : 148;
  const int tiles = p.tiles_m * p.tiles_n;
  gemm_mx_kernel<<<tiles < sms ? tiles : sms, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, td, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("gemm_mxfp8 launch: ") + cudaGetErrorString(e));
}

}  // namespace pb
