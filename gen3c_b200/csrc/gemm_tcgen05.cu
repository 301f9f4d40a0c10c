// Path D — dense bf16 GEMM on the 5th-gen tensor cores:  D[M,N] = A[M,K] · B[N,K]^T  (fp32 accum)
//
// Every Linear of the DiT (reference: cosmos_predict1/diffusion/module/attention.py:263-266,289,
// 91-102; blocks.py:153-163,222-242) is `y = x · W^T` with x [tokens, in] and W [out, in], i.e.
// both operands K-major — exactly the layout tcgen05.mma consumes from 128-byte-swizzled shared
// memory.  V^T for the attention kernel is produced by the same kernel with the operands swapped
// (A = W_v, B = x), so no transpose pass exists anywhere.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor A/B tiles -> smem ring (kStages), mbarrier tx
//   warp 1      MMA issuer     : one thread, tcgen05.mma 128 x BN x 16, accumulators in TMEM,
//                                tcgen05.commit releases smem slots / publishes accumulators
//   warp 2      TMEM allocator
//   warps 4-7   epilogue       : tcgen05.ld (lane = row) -> fused epilogue -> global
// Two TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace g3c {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle span
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;

struct GemmParams {
  int M, N, K;
  int ldd;             // leading dimension of D in elements
  void* D;             // bf16 or f32
  const float* gate;   // [N] for EPI_GATED_RESIDUAL
  int num_m_blk, num_n_blk, num_k_blk;
  int super_n;         // n-blocks per super-column (L2 reuse of the B operand)
  int n_peer;          // bf16 epilogues: additional destinations (peer GPUs), same offsets as D
  void* peer[7];
  // EPI_NORM_ROPE_BF16: per-head (128 columns) RMSNorm gain [128], cos|sin table [M][128] (or NULL: no rotation), eps
  const float* nr_gamma;
  const float* nr_cs;
  float nr_eps;
};

// internal epilogue (not part of the C ABI enum; reached through gemm_bf16(..., norm_rope)):
// D (bf16) = RoPE(RMSNorm_head(acc) * gamma) — the to_q / to_k Sequential(Linear, RMSNorm) of the reference followed by
// the rotate-half RoPE, applied to the fp32 accumulators of one head while they are still in TMEM.
constexpr int EPI_NORM_ROPE_BF16 = 4;

template <int BN>
struct GemmSmem {
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kStageF32 = 4 * 2 * 4096;  // fp32 epilogue: per warp 2 x (32 rows x 128 B) TMA-store tiles
  static constexpr int kTotal = kStages * kStageBytes + kStageF32 + kBarBytes + 1024;  // + alignment slack
};

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
  // super-columns of `super_n` n-blocks; inside one, n fastest so that concurrently running CTAs
  // share A row-blocks and the B super-column stays L2 resident.
  int per_super = p.num_m_blk * p.super_n;
  int sc = tile / per_super;
  int rem = tile - sc * per_super;
  int n0 = sc * p.super_n;
  int width = p.num_n_blk - n0 < p.super_n ? p.num_n_blk - n0 : p.super_n;
  // the last super-column may be narrower
  if (width != p.super_n) {
    m_blk = rem / width;
    n_blk = n0 + rem - m_blk * width;
  } else {
    m_blk = rem / p.super_n;
    n_blk = n0 + rem - m_blk * p.super_n;
  }
}

// Fused Linear -> per-head RMSNorm -> rotate-half RoPE for one head (128 accumulator columns at TMEM address `taddr`,
// lane = row).  Pass 1 reduces the sum of squares, pass 2 re-reads the columns in (c, c + 64) pairs — the rotate-half
// partners — scales, rotates and stores two 64-byte row segments.  reference: module/attention.py:263-266 (to_q/to_k
// = Linear + RMSNorm) and :268-283 (apply_rotary_pos_emb); same arithmetic as k_rmsnorm_rope (dit_elementwise.cu).
__device__ __forceinline__ void epilogue_norm_rope_head(uint32_t taddr, const GemmParams& p, int row, int col0) {
  float ss = 0.0f;
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    uint32_t r[32];
    tmem_ld32(taddr + cc * 32, r);
    tc_wait_ld();
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) s4[i & 3] = fmaf(__uint_as_float(r[i]), __uint_as_float(r[i]), s4[i & 3]);
    ss += (s4[0] + s4[1]) + (s4[2] + s4[3]);
  }
  const float rstd = rsqrtf(ss * (1.0f / 128.0f) + p.nr_eps);
  const bool row_ok = row < p.M;
  const float* cs = p.nr_cs ? p.nr_cs + (size_t)(row_ok ? row : 0) * 128 : nullptr;
  __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col0;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    uint32_t ra[32], rb[32];
    tmem_ld32(taddr + cc * 32, ra);
    tmem_ld32(taddr + 64 + cc * 32, rb);
    tc_wait_ld();
    uint4 qa[4], qb[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // 4 columns per step
      const int c = cc * 32 + 4 * j;
      const float4 ga = __ldg(reinterpret_cast<const float4*>(p.nr_gamma + c));
      const float4 gb = __ldg(reinterpret_cast<const float4*>(p.nr_gamma + 64 + c));
      float a[4] = {__uint_as_float(ra[4 * j]) * (rstd * ga.x), __uint_as_float(ra[4 * j + 1]) * (rstd * ga.y),
                    __uint_as_float(ra[4 * j + 2]) * (rstd * ga.z), __uint_as_float(ra[4 * j + 3]) * (rstd * ga.w)};
      float b[4] = {__uint_as_float(rb[4 * j]) * (rstd * gb.x), __uint_as_float(rb[4 * j + 1]) * (rstd * gb.y),
                    __uint_as_float(rb[4 * j + 2]) * (rstd * gb.z), __uint_as_float(rb[4 * j + 3]) * (rstd * gb.w)};
      if (cs) {
        const float4 co = *reinterpret_cast<const float4*>(cs + c);
        const float4 si = *reinterpret_cast<const float4*>(cs + 64 + c);
        const float cv[4] = {co.x, co.y, co.z, co.w}, sv[4] = {si.x, si.y, si.z, si.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float n = a[e] * cv[e] - b[e] * sv[e], m = b[e] * cv[e] + a[e] * sv[e];
          a[e] = n;
          b[e] = m;
        }
      }
      uint32_t* pa = reinterpret_cast<uint32_t*>(&qa[j >> 1]) + (j & 1) * 2;
      uint32_t* pb = reinterpret_cast<uint32_t*>(&qb[j >> 1]) + (j & 1) * 2;
      pa[0] = pack_bf16x2(a[0], a[1]);
      pa[1] = pack_bf16x2(a[2], a[3]);
      pb[0] = pack_bf16x2(b[0], b[1]);
      pb[1] = pack_bf16x2(b[2], b[3]);
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        reinterpret_cast<uint4*>(dptr + cc * 32)[i] = qa[i];
        reinterpret_cast<uint4*>(dptr + 64 + cc * 32)[i] = qb[i];
      }
    }
  }
}

// One 32-column chunk of the epilogue for the 32 rows of a warp (lane = row).  `row0w` = first row of the warp.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], const GemmParams& p, const CUtensorMap* tmD_ptr,
                                               int row0w, int col0, uint32_t lane, uint8_t* stage_warp,
                                               uint32_t& ebuf) {
  const int row = row0w + (int)lane;
  const bool row_ok = row < p.M;
        if constexpr (EPI == G3C_EPI_BF16 || EPI == G3C_EPI_GELU_BF16) {
          // lane = row: each thread stores 64 contiguous bytes of its own row (full 32-B sectors)
          if (row_ok && col0 < p.N) {
            const bool full_chunk = col0 + 32 <= p.N;
            __nv_bfloat16* dptr = reinterpret_cast<__nv_bfloat16*>(p.D) + (size_t)row * p.ldd + col0;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = __uint_as_float(r[i]);
              if constexpr (EPI == G3C_EPI_GELU_BF16) v[i] = gelu_erf(v[i]);
            }
            if (full_chunk) {
              uint4 q[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                q[i].x = pack_bf16x2(v[8 * i + 0], v[8 * i + 1]);
                q[i].y = pack_bf16x2(v[8 * i + 2], v[8 * i + 3]);
                q[i].z = pack_bf16x2(v[8 * i + 4], v[8 * i + 5]);
                q[i].w = pack_bf16x2(v[8 * i + 6], v[8 * i + 7]);
                reinterpret_cast<uint4*>(dptr)[i] = q[i];
              }
              // fused all-gather: the same 64 bytes go to the peers' copies over NVLink (posted writes)
#pragma unroll
              for (int pd = 0; pd < 7; ++pd) {
                if (pd < p.n_peer) {
                  uint4* rp = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.peer[pd]) +
                                                       (size_t)row * p.ldd + col0);
#pragma unroll
                  for (int i = 0; i < 4; ++i) rp[i] = q[i];
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (col0 + i < p.N) dptr[i] = __float2bfloat16_rn(v[i]);
            }
          }
        } else {
          // fp32 output / gated residual: the 32x32 fp32 chunk goes to a 128-byte-swizzled shared tile and one
          // elected lane hands it to the TMA: plain store, or reduce-add into the fp32 residual stream in L2
          // (x += gate * acc without ever loading x into the SM).  TMA clips rows >= M / columns >= N.
          if (col0 < p.N) {  // warp-uniform
            uint8_t* stg = stage_warp + (ebuf & 1) * 4096;
            if (lane == 0) tma_store_wait_read<1>();  // the store that last used this buffer has read it
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                     __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
              if constexpr (EPI == G3C_EPI_GATED_RESIDUAL_F32) {
                const int cg = col0 + 4 * j;
                float4 g;
                if (cg + 4 <= p.N) {
                  g = *reinterpret_cast<const float4*>(p.gate + cg);
                } else {  // ragged right edge: columns >= N are clipped by the TMA
                  g.x = cg + 0 < p.N ? p.gate[cg + 0] : 0.f;
                  g.y = cg + 1 < p.N ? p.gate[cg + 1] : 0.f;
                  g.z = cg + 2 < p.N ? p.gate[cg + 2] : 0.f;
                  g.w = 0.f;
                }
                v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w;
              }
              // row = lane, 16-byte chunk j -> swizzled chunk j ^ (row % 8)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              const int row0 = row0w;
              if constexpr (EPI == G3C_EPI_GATED_RESIDUAL_F32) tma_reduce_add_2d(tmD_ptr, stg, col0, row0);
              else tma_store_2d(tmD_ptr, stg, col0, row0);
              tma_store_commit();
            }
            ++ebuf;
          }
        }
      }

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    k_gemm(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
           const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                 : (2 * BN <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* stage_f32 = smem + kStages * S::kStageBytes;  // 1024-byte aligned (stage bytes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_f32 + S::kStageF32);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + kStages;         // [kStages]
  uint64_t* tfull = bars + 2 * kStages;     // [2]
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.num_m_blk * p.num_n_blk;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(p, tile, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blk; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], S::kStageBytes);
          tma_load_2d(smem_a + stage * S::kABytes, &tmA, &full[stage], kb * BK, m_blk * BM);
          tma_load_2d(smem_b + stage * S::kBBytes, &tmB, &full[stage], kb * BK, n_blk * BN);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      // ===== MMA issuer: the warp runs converged, one elected lane issues (operands stay in uniform registers; inside
      // `if (lane == 0)` ptxas put 5 R2UR + a broadcast loop between consecutive UTCHMMAs — see attn_tcgen05.cu) =====
      const bool issuer = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tbase + as * BN;
        for (int kb = 0; kb < p.num_k_blk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sdesc_sw128(smem_u32(smem_a + stage * S::kABytes));
          const uint64_t db = make_sdesc_sw128(smem_u32(smem_b + stage * S::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (issuer)
              umma_ss(d_tmem, sdesc_advance(da, k * UMMA_K * 2), sdesc_advance(db, k * UMMA_K * 2),
                      idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (issuer) umma_commit(&empty[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (issuer) umma_commit(&tfull[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const uint32_t ew = warp - 4;  // == warp % 4 -> TMEM lanes [32*ew, 32*ew+32)
    uint32_t as = 0, aphase = 0, ebuf = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if constexpr (EPI == EPI_NORM_ROPE_BF16) {
#pragma unroll 1
        for (int hd = 0; hd < BN / 128; ++hd)
          if (n_blk * BN + hd * 128 < p.N)
            epilogue_norm_rope_head(tmem_base + ((ew * 32u) << 16) + as * BN + hd * 128, p,
                                    m_blk * BM + (int)ew * 32 + (int)lane, n_blk * BN + hd * 128);
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((ew * 32u) << 16) + as * BN + c * 32, r);
          tc_wait_ld();
          epilogue_chunk<EPI>(r, p, &tmD, m_blk * BM + (int)ew * 32, n_blk * BN + c * 32, lane,
                              stage_f32 + ew * 2 * 4096, ebuf);
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();  // all bulk stores of this warp are complete before exit
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------------
// CTA-pair variant: 256 x 256 output tile per cluster of two CTAs (`tcgen05.mma.cta_group::2`, UMMA M = 256).
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 rows of W); the tensor cores of both
// SMs read both halves, so the shared-memory operand traffic per MMA drops from 12 KB to 8 KB per SM — the
// single-CTA kernel above is bound there (ncu: l1tex 75 %, tensor pipe 78 %, profiles/r01_gemm_ncu_summary.txt).
// Barriers: `full` lives in the leader (count 2: its own expect_tx arrive + the peer's remote arrive, both CTAs'
// TMA bytes are credited to it); `empty` / `tfull` exist in both CTAs and are hit by multicast commits;
// `tempty` lives in the leader and collects one elected arrive per epilogue warp of both CTAs.
// ------------------------------------------------------------------------------------------------------------
struct Gemm2Smem {
  static constexpr int kStages = 6;
  static constexpr int kABytes = 128 * BK * 2;   // 16 KB: this CTA's 128 rows of A
  static constexpr int kBBytes = 128 * BK * 2;   // 16 KB: this CTA's half of the 256-row B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStageF32 = 4 * 2 * 4096;
  static constexpr int kTotal = kStages * kStageBytes + kStageF32 + 256 + 1024;
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
    k_gemm2(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  using S = Gemm2Smem;
  constexpr int kStages = S::kStages;
  constexpr int BN2 = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * S::kABytes;
  uint8_t* stage_f32 = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_f32 + S::kStageF32);
  uint64_t* full = bars;                      // [kStages] (used in the leader)
  uint64_t* empty = bars + kStages;           // [kStages] (both CTAs)
  uint64_t* tfull = bars + 2 * kStages;       // [2]       (both CTAs)
  uint64_t* tempty = bars + 2 * kStages + 2;  // [2]       (used in the leader)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const uint32_t cta = cluster_ctarank();
  const bool leader = cta == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.num_m_blk * p.num_n_blk;  // num_m_blk counts 256-row blocks here
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs) =====
      uint32_t stage = 0, phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        int m_blk, n_blk;
        tile_coords(p, tile, m_blk, n_blk);
        for (int kb = 0; kb < p.num_k_blk; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full[stage], 2 * S::kStageBytes);
          else mbar_arrive_leader(&full[stage]);
          tma_load_2d_2sm(smem_a + stage * S::kABytes, &tmA, &full[stage], kb * BK, m_blk * 256 + (int)cta * 128);
          tma_load_2d_2sm(smem_b + stage * S::kBBytes, &tmB, &full[stage], kb * BK, n_blk * BN2 + (int)cta * 128);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===== MMA issuer (leader CTA; warp converged, one elected lane issues) =====
      const bool issuer = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_bf16(256, BN2);
      uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tbase + as * BN2;
        for (int kb = 0; kb < p.num_k_blk; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sdesc_sw128(smem_u32(smem_a + stage * S::kABytes));
          const uint64_t db = make_sdesc_sw128(smem_u32(smem_b + stage * S::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (issuer)
              umma_ss_2sm(d_tmem, sdesc_advance(da, k * UMMA_K * 2), sdesc_advance(db, k * UMMA_K * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
          }
          if (issuer) umma_commit_2sm(&empty[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (issuer) umma_commit_2sm(&tfull[as]);
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, each for its own 128 rows) =====
    const uint32_t ew = warp - 4;
    uint32_t as = 0, aphase = 0, ebuf = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      if constexpr (EPI == EPI_NORM_ROPE_BF16) {
#pragma unroll 1
        for (int hd = 0; hd < BN2 / 128; ++hd)
          epilogue_norm_rope_head(tmem_base + ((ew * 32u) << 16) + as * BN2 + hd * 128, p,
                                  m_blk * 256 + (int)cta * 128 + (int)ew * 32 + (int)lane, n_blk * BN2 + hd * 128);
      } else {
#pragma unroll 1
        for (int c = 0; c < BN2 / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((ew * 32u) << 16) + as * BN2 + c * 32, r);
          tc_wait_ld();
          epilogue_chunk<EPI>(r, p, &tmD, m_blk * 256 + (int)cta * 128 + (int)ew * 32, n_blk * BN2 + c * 32, lane,
                              stage_f32 + ew * 2 * 4096, ebuf);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tempty[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's tensor core may still read this CTA's shared memory until both are done
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

template <int EPI>
static int launch_gemm2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                        const GemmParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    G3C_CUDA(cudaFuncSetAttribute(k_gemm2<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Smem::kTotal));
    configured = true;
  }
  int tiles = p.num_m_blk * p.num_n_blk;
  int clusters = sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  k_gemm2<EPI><<<2 * clusters, GEMM_THREADS, Gemm2Smem::kTotal, st>>>(tmA, tmB, tmD, p);
  G3C_CUDA(cudaGetLastError());
  return G3C_OK;
}

static int dispatch_epi2(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                         const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case G3C_EPI_BF16: return launch_gemm2<G3C_EPI_BF16>(tmA, tmB, tmD, p, st);
    case G3C_EPI_GELU_BF16: return launch_gemm2<G3C_EPI_GELU_BF16>(tmA, tmB, tmD, p, st);
    case G3C_EPI_GATED_RESIDUAL_F32: return launch_gemm2<G3C_EPI_GATED_RESIDUAL_F32>(tmA, tmB, tmD, p, st);
    case G3C_EPI_F32: return launch_gemm2<G3C_EPI_F32>(tmA, tmB, tmD, p, st);
    case EPI_NORM_ROPE_BF16: return launch_gemm2<EPI_NORM_ROPE_BF16>(tmA, tmB, tmD, p, st);
  }
  set_error("gemm: unknown epilogue %d", epi);
  return G3C_EINVAL;
}

template <int BN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                       const GemmParams& p, cudaStream_t st) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  if (!configured) {
    G3C_CUDA(cudaFuncSetAttribute(k_gemm<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  S::kTotal));
    configured = true;
  }
  int tiles = p.num_m_blk * p.num_n_blk;
  int grid = tiles < sm_count() ? tiles : sm_count();
  k_gemm<BN, EPI><<<grid, GEMM_THREADS, S::kTotal, st>>>(tmA, tmB, tmD, p);
  G3C_CUDA(cudaGetLastError());
  return G3C_OK;
}

template <int BN>
static int dispatch_epi(int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                        const GemmParams& p, cudaStream_t st) {
  switch (epi) {
    case G3C_EPI_BF16: return launch_gemm<BN, G3C_EPI_BF16>(tmA, tmB, tmD, p, st);
    case G3C_EPI_GELU_BF16: return launch_gemm<BN, G3C_EPI_GELU_BF16>(tmA, tmB, tmD, p, st);
    case G3C_EPI_GATED_RESIDUAL_F32: return launch_gemm<BN, G3C_EPI_GATED_RESIDUAL_F32>(tmA, tmB, tmD, p, st);
    case G3C_EPI_F32: return launch_gemm<BN, G3C_EPI_F32>(tmA, tmB, tmD, p, st);
    case EPI_NORM_ROPE_BF16:
      if constexpr (BN >= 128) return launch_gemm<BN, EPI_NORM_ROPE_BF16>(tmA, tmB, tmD, p, st);
      break;
  }
  set_error("gemm: unknown epilogue %d", epi);
  return G3C_EINVAL;
}

// Host entry used by the engine and by the C ABI.
int gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb, int ldd,
              int epilogue, const float* gate, int block_n, cudaStream_t st, const PeerDst* peers,
              const NormRope* norm_rope) {
  G3C_REQUIRE(A && B && D, "gemm: null operand");
  if (norm_rope) {
    G3C_REQUIRE(epilogue == G3C_EPI_BF16 && N % 128 == 0 && norm_rope->gamma && (!peers || peers->n == 0),
                "gemm: the RMSNorm/RoPE epilogue needs the bf16 epilogue, N %% 128 == 0, a gain vector and no peers");
    G3C_REQUIRE((reinterpret_cast<uintptr_t>(norm_rope->gamma) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(norm_rope->cs) & 15) == 0,
                "gemm: RMSNorm gain / RoPE table must be 16-byte aligned");
    epilogue = EPI_NORM_ROPE_BF16;
  }
  G3C_REQUIRE(!peers || peers->n == 0 || (epilogue == G3C_EPI_BF16 && N % 32 == 0 && peers->n <= 7),
              "gemm: peer destinations need the bf16 epilogue, N %% 32 == 0 and at most 7 peers");
  G3C_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: bad shape %dx%dx%d", M, N, K);
  G3C_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "gemm: K/lda/ldb must be multiples of 8");
  G3C_REQUIRE(lda >= K && ldb >= K && ldd >= N, "gemm: leading dimension smaller than extent");
  if (epilogue == G3C_EPI_BF16 || epilogue == G3C_EPI_GELU_BF16 || epilogue == EPI_NORM_ROPE_BF16)
    G3C_REQUIRE(ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0,
                "gemm: bf16 output needs ldd %% 8 == 0 and 16-byte aligned base");
  else
    G3C_REQUIRE(ldd % 4 == 0 && (reinterpret_cast<uintptr_t>(D) & 15) == 0,
                "gemm: f32 output needs ldd %% 4 == 0 and 16-byte aligned base");
  G3C_REQUIRE(epilogue != G3C_EPI_GATED_RESIDUAL_F32 ||
                  (gate && (reinterpret_cast<uintptr_t>(gate) & 15) == 0),
              "gemm: gated-residual epilogue needs a 16-byte aligned gate vector");
  int bn = block_n;
  static int pair_default = -1;
  if (pair_default < 0) {
    const char* e = getenv("G3C_GEMM_2CTA");
    pair_default = e ? atoi(e) : 1;
  }
  // block_n 512 = CTA-pair kernel (256 x 256 tile over two SMs); chosen automatically for large, aligned problems
  // (G3C_GEMM_2CTA=0 disables it)
  if (bn == 0 && pair_default && N % 256 == 0 && M >= 1024 && K >= 256) {
    // wave quantisation: a pair tile is 256 x 256 on two SMs, a single tile 128 x 256 on one, i.e. the same work per
    // SM and wave; the pair kernel is ~6 % faster per wave (half the B traffic) unless it needs an extra, mostly
    // empty wave (e.g. M = 7 040 = 27.5 pair rows at cp = 8: 7 waves against 6).
    const int sms = sm_count();
    const long long tiles_p = (long long)((M + 255) / 256) * (N / 256), tiles_s = (long long)((M + 127) / 128) * (N / 256);
    const long long waves_p = (tiles_p + sms / 2 - 1) / (sms / 2), waves_s = (tiles_s + sms - 1) / sms;
    bn = (100 * waves_p <= 106 * waves_s) ? 512 : 256;
  }
  if (bn == 0) bn = (N >= 256 && N % 256 == 0) ? 256 : (N > 64 ? 128 : 64);
  G3C_REQUIRE(epilogue != EPI_NORM_ROPE_BF16 || bn >= 128, "gemm: the RMSNorm/RoPE epilogue needs tiles of whole heads");
  G3C_REQUIRE(bn == 64 || bn == 128 || bn == 256 || bn == 512, "gemm: block_n %d unsupported", bn);
  const bool pair = bn == 512;
  if (pair) {
    G3C_REQUIRE(N % 256 == 0, "gemm: the CTA-pair kernel needs N %% 256 == 0 (N=%d)", N);
    bn = 256;
  }

  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {(uint64_t)K, (uint64_t)M}, strA[1] = {(uint64_t)lda * 2};
  uint32_t boxA[2] = {BK, 128};
  int rc = make_tmap_bf16_sw128(&tmA, A, 2, dimsA, strA, boxA);
  if (rc) return rc;
  uint64_t dimsB[2] = {(uint64_t)K, (uint64_t)N}, strB[1] = {(uint64_t)ldb * 2};
  uint32_t boxB[2] = {BK, (uint32_t)(pair ? 128 : bn)};
  rc = make_tmap_bf16_sw128(&tmB, B, 2, dimsB, strB, boxB);
  if (rc) return rc;

  CUtensorMap tmD;
  memset(&tmD, 0, sizeof(tmD));
  if (epilogue == G3C_EPI_GATED_RESIDUAL_F32 || epilogue == G3C_EPI_F32) {
    uint64_t dimsD[2] = {(uint64_t)N, (uint64_t)M}, strD[1] = {(uint64_t)ldd * 4};
    uint32_t boxD[2] = {32, 32};
    rc = make_tmap_f32_sw128(&tmD, D, dimsD, strD, boxD);
    if (rc) return rc;
  }

  GemmParams p;
  p.M = M;
  p.N = N;
  p.K = K;
  p.ldd = ldd;
  p.D = D;
  p.gate = gate;
  p.n_peer = peers ? peers->n : 0;
  for (int i = 0; i < 7; ++i) p.peer[i] = (peers && i < peers->n) ? peers->ptr[i] : nullptr;
  p.nr_gamma = norm_rope ? norm_rope->gamma : nullptr;
  p.nr_cs = norm_rope ? norm_rope->cs : nullptr;
  p.nr_eps = norm_rope ? norm_rope->eps : 0.0f;
  p.num_m_blk = pair ? (M + 255) / 256 : (M + BM - 1) / BM;
  p.num_n_blk = (N + bn - 1) / bn;
  p.num_k_blk = (K + BK - 1) / BK;
  // keep one super-column of B (super_n * bn * K * 2 bytes) well inside the 126 MB L2
  long long col_bytes = (long long)bn * K * 2;
  int sn = (int)((48ll << 20) / (col_bytes > 0 ? col_bytes : 1));
  if (sn < 1) sn = 1;
  if (sn > p.num_n_blk) sn = p.num_n_blk;
  p.super_n = sn;
  if (pair) return dispatch_epi2(epilogue, tmA, tmB, tmD, p, st);
  switch (bn) {
    case 64: return dispatch_epi<64>(epilogue, tmA, tmB, tmD, p, st);
    case 128: return dispatch_epi<128>(epilogue, tmA, tmB, tmD, p, st);
    default: return dispatch_epi<256>(epilogue, tmA, tmB, tmD, p, st);
  }
}

}  // namespace g3c

extern "C" int g3c_gemm_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda,
                             int ldb, int ldd, int epilogue, const float* gate, int block_n,
                             void* stream) {
  return g3c::gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, epilogue, gate, block_n,
                        (cudaStream_t)stream, nullptr);
}

extern "C" int g3c_gemm_norm_rope_bf16(const void* A, const void* B, void* D, int M, int N, int K, int lda, int ldb,
                                       int ldd, const float* gamma, const float* cos_sin, float eps, void* stream) {
  g3c::NormRope nr;
  nr.gamma = gamma;
  nr.cs = cos_sin;
  nr.eps = eps;
  return g3c::gemm_bf16(A, B, D, M, N, K, lda, ldb, ldd, G3C_EPI_BF16, nullptr, 0, (cudaStream_t)stream, nullptr, &nr);
}
