// Path D — non-causal multi-head attention forward, head_dim 128, on tcgen05 / TMEM.
// Three kernels live here, in the order they were written:
//   k_attn_fwd     two query tiles per CTA, per-tile S buffers with P aliasing S, A/B tile alternation (round 1 / early
//                  round 2; today: key ranges <= 1 024 in its exact 1-CTA form, and the A/B baseline G3C_ATTN_1T=0)
//   k_attn_fwd16   the same with 16 softmax warps (measured: no gain)
//   k_attn_fwd1t   DEFAULT for self-attention: one query tile per CTA, three S buffers, software-pipelined over KV steps
//                  (description above its definition; 96 % tensor-pipe activity)
// Variants built and measured slower: profiles/r01_attention_variants.txt, profiles/r02_attention_variants.txt.
// The text below describes k_attn_fwd.
//   O = softmax(Q K^T * scale) V        (reference: cosmos_predict1/diffusion/module/attention.py
//   :282-297 `cal_attn` -> transformer_engine DotProductAttention(sbhd, no_mask, dropout 0);
//   self-attention Lq = Lk = 56 320, cross-attention Lk = 512)
//
// Layouts (all bf16, produced by the projection GEMMs of gemm_tcgen05.cu):
//   Q  [Lq, heads*128]  token-major          K [Lk, heads*128] token-major
//   Vt [chunks][heads*128][chunk_len]        (V transposed, keys contiguous -> K-major B operand;
//                                             `chunks` = context-parallel ranks after the KV
//                                             all-gather, 1 otherwise)
//   O  [Lq, heads*128]
//
// One CTA = 256 query rows (two 128-row tiles A/B) of one head, 320 threads:
//   warps 0-3  softmax of tile A   (thread = one query row; S row read from TMEM into registers)
//   warps 4-7  softmax of tile B
//   warp  8    TMA producer: Q once, then K_j / V_j through a 4-slot ring of 32 KB tiles
//   warp  9    TMEM allocator + single-thread MMA issuer
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,384) O_B [384,512); P (bf16) overwrites
// the first 64 columns of its S tile and feeds the P·V MMA straight from TMEM.
// MMA order per KV step j:  PV_A(j) ; S_A(j+1) ; PV_B(j) ; S_B(j+1)  — the S MMA of one tile and
// the whole PV/S pair of the other overlap with that tile's softmax.  P is released to the MMA warp
// in two 64-key halves, so the first four P·V k-steps run under the second half of the exponentials.
// Per tile the dependent chain is  S MMA -> softmax -> PV MMA -> (P columns free) -> next S MMA, so
// the softmax latency of one tile sets the step period (clock64 timelines: profiles/r01_attn_v1_timeline.txt),
// and the GPU is power-capped under this kernel: what pays is fewer instructions per score.  Hence the
// default softmax (kMode 2, see below): no per-tile row max, no scale-subtract when Q carries the scale,
// FADD2 row sums.  CTA pairs share every K / V tile through TMA multicast (kCluster).
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "kernels.h"

namespace g3c {
namespace v1 {

#ifndef G3C_ATTN_POLY_DEFAULT
#define G3C_ATTN_POLY_DEFAULT 0  // fraction 1/n of the exponential pairs on the FMA pipe in the default path (0 = none)
#endif
constexpr int ATT_THREADS = 320;
constexpr int ATT_TILE = 128;             // rows per Q tile, keys per KV tile, head dim
constexpr int ATT_HALF_BYTES = 128 * 128; // one 64-column half of a 128x128 bf16 tile
constexpr int ATT_TILE_BYTES = 2 * ATT_HALF_BYTES;
constexpr int ATT_SLOTS = 4;
constexpr int ATT_SMEM = 2 * ATT_TILE_BYTES + ATT_SLOTS * ATT_TILE_BYTES + 256 + 1024;

struct AttnParams {
  int Lq, Lk, heads;
  int ldo;
  int vt_chunk_len;
  __nv_bfloat16* O;
  float scale_log2;  // softmax scale * log2(e)
  const uint32_t* chunk_flags;  // context-parallel gate (or NULL): chunk c readable once chunk_flags[c] >= flag_seq
  uint32_t flag_seq;
  int first_chunk;
  unsigned long long peer_timeout_ns;  // bound of the wait for a peer's chunk flag (and of this CTA's barrier waits
                                       // while such a wait may be pending): an inter-process dependency, not a protocol bug
  unsigned long long* wait_ns;         // optional profiling counter: ns spent polling chunk flags, summed over CTAs
  int unit_scale;             // 1: scale_log2 == 1 (the caller folded softmax scale * log2 e into Q): S is in log2 units
  int p_halves;               // 1: P is released to the MMA warp per 64-key half, 0: per 128-key tile
  int st_overlap;             // 1: the first P half is released after the first 16 exponentials of the second half
  int p_quarters;             // 1: reference-free fast tiles release P in four 32-key quarters (software-pipelined stores)
  int split_s;                // 2-CTA kernel: S = Q K^T as two 64-key UMMAs; the upper one is issued for step j+1 as soon as
                              // the softmax has read columns 64..127 of S(j) (they do not alias P), i.e. under the
                              // exponentials and before P.V(j), which shortens the per-tile dependent chain by half an S MMA
  const __nv_bfloat16* Q;     // k_attn_fwd1t<kQT>: Q rows are read straight from global memory into TMEM
  int ldq;
  int big_boxes;              // k_attn_fwd1t: K / V^T parts arrive as ONE 16 KB TMA box each (tensor maps with the 64-element halves as an extra dimension)
  int kv_rotate;              // k_attn_fwd1t: every CTA pair visits the KV tiles of a chunk from its own start offset
  int dbg_dup_loads;          // unused (round-2 experiment: every K / V part fetched twice cost +9 % of the step, three times +23 %)
  unsigned long long* trace;  // kTrace only: [3 roles][64 steps][8 slots] clock64 stamps of CTA (0,0)
};

#define ATT_TR(role, slot)                                                                       \
  do {                                                                                           \
    if constexpr (kTrace) {                                                                      \
      if (blockIdx.x == 0 && blockIdx.y == 0 && j < 64) p.trace[((role) * 64 + j) * 8 + (slot)] = clock64(); \
    }                                                                                            \
  } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {  // FADD2: two fp32 adds in one instruction
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {  // FFMA2: two fp32 fmas in one instruction
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// 2^x for a PAIR of scores on the FMA pipe, packed fp32x2 arithmetic (FA4-style MUFU offload): round-to-nearest split
// by the magic-number add, degree-3 minimax polynomial on [-0.5, 0.5] (rel. err 7.5e-5, far below the bf16 rounding of
// P), exponent inserted with one shift-add per element.  8 FMA-pipe / ALU instructions per pair against 2 MUFU slots of
// 8 clk each: the 16 ex2/clk/SM MUFU is as slow as the two MMAs of a KV step, so moving a fraction of the exponentials
// here takes the softmax off the tensor pipe's critical path.
__device__ __forceinline__ void ex2_poly2(float xa, float xb, float& ya, float& yb) {
  const uint64_t magic = pack2(12582912.0f, 12582912.0f);  // 1.5 * 2^23
  const uint64_t x = pack2(fmaxf(xa, -125.0f), fmaxf(xb, -125.0f));
  const uint64_t xr = fadd2(x, magic);                                           // low mantissa bits = round(x)
  const uint64_t n = fadd2(xr, pack2(-12582912.0f, -12582912.0f));
  const uint64_t f = ffma2(n, pack2(-1.0f, -1.0f), x);                           // x - n in [-0.5, 0.5]
  uint64_t pl = ffma2(pack2(0.05517165f, 0.05517165f), f, pack2(0.24261113f, 0.24261113f));
  pl = ffma2(pl, f, pack2(0.69326097f, 0.69326097f));
  pl = ffma2(pl, f, pack2(0.99992806f, 0.99992806f));
  float pa, pb, ra, rb;
  unpack2(pl, pa, pb);
  unpack2(xr, ra, rb);
  ya = __int_as_float(__float_as_int(pa) + (__float_as_int(ra) << 23));
  yb = __int_as_float(__float_as_int(pb) + (__float_as_int(rb) << 23));
}

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial, rel. err 7.5e-5, far below the bf16
// rounding of P): the MUFU unit delivers only 16 ex2/clk/SM, which is exactly as slow as the two MMAs of a
// KV step; computing every kPolyEvery-th exponential here takes the softmax off the critical path.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;  // 1.5 * 2^23: low mantissa bits now hold round(x)
  const float n = xr - 12582912.0f;
  const float f = x - n;  // [-0.5, 0.5]
  const float p = fmaf(fmaf(fmaf(0.05517165f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

// kMode 0: exact row max of every KV tile -> lazy rescale -> exponentials.
// kMode 2: the row max is only reduced for the first KV tile.  Afterwards the exponentials simply keep using the
//          current reference exponent: fp32 (and bf16 P) carry 8 exponent bits, so a stale reference costs no
//          precision, only range.  Range is guarded by the row sums that are computed anyway: when a tile's sum
//          exceeds 2^16 the reference is shifted by that sum's exponent before the next tile (O and the running sum
//          are scaled by an exact power of two); if a row sum still ends up non-finite — a jump of > 2^100 inside one
//          tile — the CTA repeats its work once in the exact mode (second pass of the role loops below).
//          The ~300 clk max reduction leaves the per-tile dependent chain (see the timeline in profiles/).
// kCluster: the CTAs of two neighbouring query blocks of one head form a cluster; each TMA-loads HALF of every K / V
//           tile and multicasts it into both CTAs' rings, so every K / V byte leaves L2 once per 512 query rows.
// k2Cta:   (implies kCluster) the pair issues ONE 256-row UMMA per tile (`tcgen05.mma.cta_group::2`): tile t of the
//           leader CTA and tile t of its peer share every S = Q K^T and P.V instruction.  Each CTA stages only HALF of
//           the B operand — 64 of the 128 keys of a K tile, 64 of the 128 head dimensions of a V^T tile — so the
//           shared-memory operand traffic of the tensor pipe and the TMA fill per SM are halved against the multicast
//           variant (whose MMAs ran at ~82 % of their nominal rate: 1 388 clk per PV+S pair instead of 1 134, clock64
//           timeline in profiles/r02_attn_trace.txt), and the ring holds 8 instead of 4 tiles in the same 128 KB.
//           The leader's MMA thread issues for both SMs; `kv_full` / `p_part` / `q_full` live in the leader and collect
//           both CTAs' arrivals, `s_full` / `kv_empty` are hit in both CTAs by multicast commits.
// kSharedS: (implies k2Cta) ONE S buffer (128 TMEM columns) shared by both tiles, P in its own 64 columns per tile:
//           S_t(j+1) no longer waits for P.V_t(j) to drain the P columns that alias S, only for the OTHER tile's softmax
//           to have pulled its S row into registers.  With MMAs now issued at their native rate (63 clk per 128x128x16,
//           clock64 trace profiles/r02_attn_trace.txt) the per-tile chain softmax -> P.V -> S -> softmax was the limiter
//           (tensor pipe 62 % busy, period 3 160-3 260 clk): here the next S is computed while the tile's own softmax
//           still runs, so the softmax warps run back to back.  TMEM: S [0,128) P_A [128,192) P_B [192,256) O_A O_B.
template <int kPolyEvery, int kTrace, int kMode, bool kCluster, bool k2Cta = false, bool kSharedS = false>  // kTrace 1: stamps in every role, 2: MMA warp only
__global__ void __launch_bounds__(ATT_THREADS, 1)
    k_attn_fwd(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  static_assert(!k2Cta || kCluster, "2-CTA MMAs need the cluster launch");
  static_assert(!kSharedS || k2Cta, "the shared-S variant is built on the 2-CTA kernel");
  constexpr int kSlots = k2Cta ? 2 * ATT_SLOTS : ATT_SLOTS;            // ring entries (one K or V^T tile each)
  constexpr int kSlotBytes = k2Cta ? ATT_TILE_BYTES / 2 : ATT_TILE_BYTES;  // this CTA's part of a tile
  constexpr int kKvHalf = kSlotBytes / 2;                               // one 64-element K-dimension half of it
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;                        // [2 tiles][2 halves][128 x 128 B]
  uint8_t* smem_kv = smem + 2 * ATT_TILE_BYTES;  // [slots][2 halves][rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + kSlots * kSlotBytes);
  uint64_t* q_full = bars;                        // [1]
  uint64_t* kv_full = bars + 1;                   // [slots]
  uint64_t* kv_empty = bars + 1 + kSlots;         // [slots]
  uint64_t* s_full = bars + 1 + 2 * kSlots;       // [2]
  uint64_t* p_part = bars + 3 + 2 * kSlots;       // [tile][key quarter]: P columns of 32 keys stored
  uint64_t* hi_free = bars + 11 + 2 * kSlots;     // [tile]: S columns 64..127 have been read (split_s)
  uint64_t* s_free = hi_free;                     // kSharedS: the S buffer has been read into registers (either tile)
  uint64_t* p_free = bars + 13 + 2 * kSlots;      // kSharedS [tile]: P.V_t(j) has consumed P_t (and updated O_t)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15 + 2 * kSlots);
  uint32_t* redo_flag = tmem_ptr + 1;  // kMode 2: some row sum left the fp32 range, repeat in the exact mode

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 2 * ATT_TILE;
  const int n_kv = p.Lk / ATT_TILE;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    // 2-CTA: the leader's barrier collects its own expect_tx arrive and the peer's remote arrive (both CTAs' TMA bytes
    // are credited to it); multicast variant: the slot is rewritten in both CTAs, both consumers release it
    mbar_init(q_full, k2Cta ? 2 : 1);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&kv_full[i], k2Cta ? 2 : 1);
      mbar_init(&kv_empty[i], (kCluster && !k2Cta) ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      for (int q = 0; q < 4; ++q) mbar_init(&p_part[4 * i + q], k2Cta ? 8 : 4);  // one elected arrive per softmax warp (of both CTAs)
      mbar_init(&hi_free[i], k2Cta ? 8 : 4);
      mbar_init(&p_free[i], 1);
    }
    *redo_flag = 0u;
    fence_barrier_init();
  }
  if (warp == 9) {
    if constexpr (k2Cta) tmem_alloc_2sm(tmem_ptr, 512);
    else tmem_alloc(tmem_ptr, 512);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t crank = kCluster ? cluster_ctarank() : 0u;

  // Pipeline state of every role lives outside the pass loop: kMode 2 may run the KV sweep a second time.
  uint32_t slot = 0, phase = 0;  // KV ring position (TMA warp: producer side, MMA warp: consumer side)
  uint32_t pph = 0;              // MMA warp: parity of the p_part barriers
  uint32_t sfp = 0;              // MMA warp (kSharedS): parity of s_free
  uint32_t sphase = 0;           // softmax warps: parity of s_full
  int pass = 0;
  for (;;) {
  const bool exact = kMode != 2 || pass == 1;
  if (warp == 8) {
    if (lane == 0) {
      // ===== TMA producer =====
      if (pass == 0) {
        if constexpr (k2Cta) {
          if (crank == 0) mbar_expect_tx(q_full, 4 * ATT_TILE_BYTES);  // both CTAs' two Q tiles
          else mbar_arrive_leader(q_full);
        } else {
          mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if constexpr (k2Cta)
              tma_load_2d_2sm(smem_q + t * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmQ, q_full, head * 128 + h * 64,
                              q0 + t * ATT_TILE);
            else
              tma_load_2d(smem_q + t * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmQ, q_full,
                          head * 128 + h * 64, q0 + t * ATT_TILE);
          }
      }
      const int tiles_per_chunk = p.vt_chunk_len / ATT_TILE;
      const int n_chunks = p.Lk / p.vt_chunk_len;
      for (int j = 0; j < n_kv; ++j) {
        // KV tiles are visited chunk by chunk starting with `first_chunk` (the local one under context
        // parallelism); a remote chunk is only touched after its producer rank has published it.
        int chunk = p.first_chunk + j / tiles_per_chunk;
        if (chunk >= n_chunks) chunk -= n_chunks;
        const int within = j % tiles_per_chunk;
        if (p.chunk_flags && within == 0 && chunk != p.first_chunk) {  // the local chunk is ordered by the stream
          uint32_t v, spins = 0;
          uint64_t t0 = 0;
          for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.chunk_flags + chunk) : "memory");
            if ((int)(v - p.flag_seq) >= 0) break;
            if (t0 == 0) t0 = global_timer_ns();
            if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > p.peer_timeout_ns) asm volatile("trap;\n");
          }
          if (t0 != 0 && p.wait_ns) atomicAdd(p.wait_ns, (unsigned long long)(global_timer_ns() - t0));
          asm volatile("fence.proxy.async.global;\n" ::: "memory");  // peer-written data is read by the TMA next
        }
        const int kv0 = chunk * p.vt_chunk_len + within * ATT_TILE;
        // K_j
        mbar_wait_ns(&kv_empty[slot], phase ^ 1, p.peer_timeout_ns);
        if constexpr (k2Cta) {
          // this CTA's 64 keys of the tile (box 64 x 64, both 64-dim halves); bytes credited to the leader's barrier
          if (crank == 0) mbar_expect_tx(&kv_full[slot], 2 * kSlotBytes);
          else mbar_arrive_leader(&kv_full[slot]);
          if (p.split_s) {
            // rows 0..31 of this CTA's part = keys 32*rank.. of the lower 64 keys, rows 32..63 = of the upper 64:
            // each 64-key UMMA (32 B rows per CTA) then covers a contiguous key range (TMA box 64 x 32)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int u = 0; u < 2; ++u)
                tma_load_2d_2sm(smem_kv + slot * kSlotBytes + h * kKvHalf + u * (kKvHalf / 2), &tmK, &kv_full[slot],
                                head * 128 + h * 64, kv0 + u * 64 + (int)crank * 32);
          } else {
#pragma unroll
            for (int h = 0; h < 2; ++h)
              tma_load_2d_2sm(smem_kv + slot * kSlotBytes + h * kKvHalf, &tmK, &kv_full[slot], head * 128 + h * 64,
                              kv0 + (int)crank * 64);
          }
        } else if constexpr (kCluster) {
          mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
          tma_load_2d_mc(smem_kv + slot * ATT_TILE_BYTES + crank * ATT_HALF_BYTES, &tmK, &kv_full[slot],
                         head * 128 + crank * 64, kv0, 3);
        } else {
          mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d(smem_kv + slot * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmK, &kv_full[slot],
                        head * 128 + h * 64, kv0);
        }
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
        // V_j  (transposed: rows = head dim, columns = keys)
        mbar_wait_ns(&kv_empty[slot], phase ^ 1, p.peer_timeout_ns);
        const int koff = within * ATT_TILE;
        if constexpr (k2Cta) {
          // this CTA's 64 head dimensions of the V^T tile, all 128 keys (two 64-key halves)
          if (crank == 0) mbar_expect_tx(&kv_full[slot], 2 * kSlotBytes);
          else mbar_arrive_leader(&kv_full[slot]);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_3d_2sm(smem_kv + slot * kSlotBytes + h * kKvHalf, &tmV, &kv_full[slot], koff + h * 64,
                            head * 128 + (int)crank * 64, chunk);
        } else if constexpr (kCluster) {
          mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
          tma_load_3d_mc(smem_kv + slot * ATT_TILE_BYTES + crank * ATT_HALF_BYTES, &tmV, &kv_full[slot],
                         koff + crank * 64, head * 128, chunk, 3);
        } else {
          mbar_expect_tx(&kv_full[slot], ATT_TILE_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_3d(smem_kv + slot * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmV, &kv_full[slot],
                        koff + h * 64, head * 128, chunk);
        }
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    if (!k2Cta || crank == 0) {
      // ===== MMA issuer (2-CTA: the leader issues for both SMs) =====
      // The WHOLE warp runs this control flow converged and one elected lane issues the tcgen05 instructions.  With
      // the loop inside `if (lane == 0)` every operand (descriptors, TMEM addresses) was lane-divergent for ptxas: 5
      // R2UR + an ELECT/broadcast loop between consecutive UTCHMMA (11 SASS instructions, ~86 clk per MMA measured with
      // clock64 — more than the 68 clk the 128x128x16 MMA needs), i.e. the tensor pipe was fed by an issue-bound
      // thread.  Warp-uniform values live in uniform registers and the UTCHMMAs go out back to back.
      const bool issuer = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);   // tells the compiler the address is uniform
      constexpr uint32_t idesc = make_idesc_bf16(k2Cta ? 256 : 128, 128);
      const uint32_t tS[2] = {tbase, kSharedS ? tbase : tbase + 128};                       // S destination of tile t
      const uint32_t tPa[2] = {kSharedS ? tbase + 128 : tbase, kSharedS ? tbase + 192 : tbase + 128};  // P (A operand)
      const uint32_t tO[2] = {tbase + 256, tbase + 384};
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1; } };
      auto commit = [&](uint64_t* bar) {   // s_full: seen by the softmax warps of every CTA the MMAs wrote to
        if (issuer) {
          if constexpr (k2Cta) umma_commit_2sm(bar);
          else umma_commit(bar);
        }
      };
      auto release_slot = [&](uint32_t sl) {
        if (issuer) {
          if constexpr (k2Cta) umma_commit_2sm(&kv_empty[sl]);
          else if constexpr (kCluster) umma_commit_mc(&kv_empty[sl], 3);
          else umma_commit(&kv_empty[sl]);
        }
      };
      auto mma_s = [&](int t, uint32_t kslot) {
        // S_t = Q_t K^T : 8 k-steps over the head dimension
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t da = make_sdesc_sw128(smem_u32(smem_q + t * ATT_TILE_BYTES + half * ATT_HALF_BYTES));
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + kslot * kSlotBytes + half * kKvHalf));
          if (issuer) {
            if constexpr (k2Cta) umma_ss_2sm(tS[t], sdesc_advance(da, off), sdesc_advance(db, off), idesc, k != 0 ? 1u : 0u);
            else umma_ss(tS[t], sdesc_advance(da, off), sdesc_advance(db, off), idesc, k != 0 ? 1u : 0u);
          }
        }
      };
      // split_s (2-CTA only): the 64 keys [64u, 64u+64) of the tile -> S columns [64u, 64u+64); B rows 32u..32u+31 of
      // this CTA's K part (and of the peer's)
      auto mma_s_part = [&](int t, uint32_t kslot, int u) {
        constexpr uint32_t idesc64 = make_idesc_bf16(256, 64);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t da = make_sdesc_sw128(smem_u32(smem_q + t * ATT_TILE_BYTES + half * ATT_HALF_BYTES));
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + kslot * kSlotBytes + half * kKvHalf + u * (kKvHalf / 2)));
          if (issuer) umma_ss_2sm(tS[t] + u * 64, sdesc_advance(da, off), sdesc_advance(db, off), idesc64, k != 0 ? 1u : 0u);
        }
      };
      auto mma_pv = [&](int t, uint32_t vslot, bool first, int qq) {
        // O_t += P_t V for the 32 keys of quarter qq: 2 k-steps; A = P from TMEM (bf16 pairs per column)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = qq * 2 + kk;
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + vslot * kSlotBytes + half * kKvHalf));
          if (issuer) {
            if constexpr (k2Cta) umma_ts_2sm(tO[t], tPa[t] + k * 8, sdesc_advance(db, off), idesc, (first && k == 0) ? 0u : 1u);
            else umma_ts(tO[t], tPa[t] + k * 8, sdesc_advance(db, off), idesc, (first && k == 0) ? 0u : 1u);
          }
        }
      };
      if (pass == 0) mbar_wait_ns(q_full, 0, p.peer_timeout_ns);
      mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);
      tc_fence_after();
      uint32_t kslot = slot;
      advance();
      if constexpr (kSharedS) {
        // Dynamic issue order.  The tensor pipe executes in issue order, and what the softmax warps wait for is the next
        // S; a static order (S_A, P.V_A quarters, S_B, P.V_B quarters) parks S behind P.V work and blocks the issuing
        // thread on one tile's P while the other tile's work is ready (measured: period 4 095 clk, the two softmaxes
        // strictly alternating).  Here the thread polls (mbarrier.test_wait) and issues whatever is ready, S first:
        //   S #i (tile i & 1, step i >> 1): K of that step loaded, and S #(i-1) read by its softmax (s_free);
        //   P.V quarter q of tile t, step j: V_j loaded and p_part[t][q] of that step arrived.
        // K_j sits in ring entry 2g, V_j in 2g + 1 with g = pass * n_kv + j (the loader's order).
        (void)kslot;
        const int g0 = pass * n_kv;
        // a phase only moves forward: if any lane saw it complete it is complete (and the vote keeps the flow uniform)
        auto done = [&](uint64_t* bar, uint32_t parity) { return __any_sync(0xffffffffu, mbar_test(bar, parity)) != 0; };
        auto ring_ready = [&](int r) { return done(&kv_full[r % kSlots], (uint32_t)((r / kSlots) & 1)); };
        int s_idx = 0;                    // next S computation of this pass
        int pv_j[2] = {0, 0}, pv_q[2] = {0, 0};
        int v_done[2] = {0, 0};           // steps whose P.V is fully issued, per tile
        const int n_s = 2 * n_kv;
        uint32_t spins = 0;
        uint64_t t_first = 0;
        // consumed the prologue wait on K_0 above through `slot`/`phase`; the dynamic loop indexes the ring itself
        while (s_idx < n_s || pv_j[0] < n_kv || pv_j[1] < n_kv) {
          bool progressed = false;
          if (s_idx < n_s) {
            const int tile = s_idx & 1, step = s_idx >> 1;
            const bool first_ever = (pass == 0 && s_idx == 0);
            if (ring_ready(2 * (g0 + step)) && (first_ever || done(s_free, sfp))) {
              if (!first_ever) sfp ^= 1;
              tc_fence_after();
              mma_s(tile, (uint32_t)((2 * (g0 + step)) % kSlots));
              commit(&s_full[tile]);
              if (tile == 1) release_slot((uint32_t)((2 * (g0 + step)) % kSlots));
              ++s_idx;
              progressed = true;
            }
          }
          if (!progressed) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              if (pv_j[t] < n_kv) {
                const int j = pv_j[t], q = pv_q[t];
                const int rv = 2 * (g0 + j) + 1;
                if (ring_ready(rv) && done(&p_part[4 * t + q], (uint32_t)((g0 + j) & 1))) {
                  tc_fence_after();
                  mma_pv(t, (uint32_t)(rv % kSlots), j == 0, q);
                  if (q == 3) {
                    commit(&p_free[t]);
                    if (j + 1 == n_kv) commit(&s_full[t]);   // epilogue: O_t complete
                    v_done[t] = j + 1;
                    if (v_done[t ^ 1] > j) release_slot((uint32_t)(rv % kSlots));  // both tiles are through with V_j
                    pv_q[t] = 0;
                    pv_j[t] = j + 1;
                  } else {
                    pv_q[t] = q + 1;
                  }
                  progressed = true;
                }
              }
            }
          }
          if (!progressed && (++spins & 0x3FFFu) == 0) {
            const uint64_t now = global_timer_ns();
            if (t_first == 0) t_first = now;
            else if (now - t_first > p.peer_timeout_ns) asm volatile("trap;\n");
          }
          if (progressed) t_first = 0;
        }
        // keep the (unused) in-order ring cursor consistent for a possible second pass
        for (int r = 0; r < 2 * n_kv - 1; ++r) advance();
        pph = (uint32_t)((g0 + n_kv) & 1);
      } else {
      const bool split = k2Cta && p.split_s;
      if (split) {
        if constexpr (k2Cta) {
          mma_s_part(0, kslot, 0);
          mma_s_part(0, kslot, 1);
          commit(&s_full[0]);
          mma_s_part(1, kslot, 0);
          mma_s_part(1, kslot, 1);
        }
      } else {
        mma_s(0, kslot);
        commit(&s_full[0]);
        mma_s(1, kslot);
      }
      commit(&s_full[1]);
      release_slot(kslot);
      for (int j = 0; j < n_kv; ++j) {
        const bool more = j + 1 < n_kv;
        mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);  // V_j
        const uint32_t vslot = slot;
        advance();
        // ---- tile A
        ATT_TR(0, 0);
        // the P·V MMAs of the first 64 keys start while the softmax still exponentiates the second 64
        mbar_wait_ns(&p_part[0], pph, p.peer_timeout_ns);
        ATT_TR(0, 1);
        tc_fence_after();
        mma_pv(0, vslot, j == 0, 0);
        mbar_wait_ns(&p_part[1], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(0, vslot, j == 0, 1);
        if (split && more) {
          if constexpr (k2Cta) {
            mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);  // K_{j+1}
            kslot = slot;
            advance();
            mbar_wait_ns(&hi_free[0], pph, p.peer_timeout_ns);       // columns 64..127 of S_A(j) are in registers
            tc_fence_after();
            mma_s_part(0, kslot, 1);
          }
        }
        mbar_wait_ns(&p_part[2], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(0, vslot, j == 0, 2);
        mbar_wait_ns(&p_part[3], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(0, vslot, j == 0, 3);
        if (more) {
          if (split) {
            if constexpr (k2Cta) mma_s_part(0, kslot, 0);
          } else {
            mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);  // K_{j+1}
            tc_fence_after();
            kslot = slot;
            advance();
            mma_s(0, kslot);
          }
        }
        commit(&s_full[0]);
        ATT_TR(0, 2);
        // ---- tile B
        mbar_wait_ns(&p_part[4], pph, p.peer_timeout_ns);
        ATT_TR(0, 3);
        tc_fence_after();
        mma_pv(1, vslot, j == 0, 0);
        mbar_wait_ns(&p_part[5], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(1, vslot, j == 0, 1);
        if (split && more) {
          if constexpr (k2Cta) {
            mbar_wait_ns(&hi_free[1], pph, p.peer_timeout_ns);
            tc_fence_after();
            mma_s_part(1, kslot, 1);
          }
        }
        mbar_wait_ns(&p_part[6], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(1, vslot, j == 0, 2);
        mbar_wait_ns(&p_part[7], pph, p.peer_timeout_ns);
        tc_fence_after();
        mma_pv(1, vslot, j == 0, 3);
        release_slot(vslot);
        if (more) {
          if (split) {
            if constexpr (k2Cta) mma_s_part(1, kslot, 0);
          } else {
            mma_s(1, kslot);
          }
          commit(&s_full[1]);
          release_slot(kslot);
        } else {
          commit(&s_full[1]);
        }
        pph ^= 1;
        ATT_TR(0, 4);
      }
      }  // !kSharedS
    }
  } else {
    // ===== softmax warpgroups (warps 0-3: tile A, warps 4-7: tile B) =====
    const int t = warp >> 2;
    const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
    const uint32_t tSr = tmem_base + lane_base + (kSharedS ? 0 : t * 128);             // S is read from here
    const uint32_t tS = tmem_base + lane_base + (kSharedS ? 128 + t * 64 : t * 128);   // P is stored here
    const uint32_t tO = tmem_base + lane_base + 256 + t * 128;
    bool pf_waited = true;  // kSharedS: P.V_t(j-1) known complete (P_t reusable, O_t up to date)
    int cur_j = 0;
    // kSharedS: before P_t is overwritten or O_t touched in step j >= 1, P.V_t(j-1) must have completed (p_free is
    // committed behind it; completion number pass * n_kv + j - 1 -> its parity).  Taken as late as possible: the last
    // P.V quarter of step j-1 is only issued when this warp has finished step j-1.
    auto ensure_pfree = [&]() {
      if constexpr (kSharedS) {
        if (!pf_waited) {
          mbar_wait_ns(&p_free[t], (uint32_t)((pass * n_kv + cur_j - 1) & 1), p.peer_timeout_ns);
          tc_fence_after();
          pf_waited = true;
        }
      }
    };
    const float c = p.scale_log2;
    float ref = 0.0f;   // reference exponent (log2 units): the stored exponentials are 2^(s*c - ref)
    float l = 0.0f;     // running row sum (relative to ref)
    bool ovf = false;   // kMode 2: the running row sum left the safe range at some tile
    bool plain = false; // kMode 2: ref == 0 in every row of this warp and S is in log2 units
    float pend = 0.0f;  // kMode 2: exponent shift to apply to ref / O / l before the next tile (0 = none)
    const bool tr = kTrace == 1 && (warp & 3) == 0 && lane == 0;
    auto rescale = [&](float alpha) {
      // s_full(j) was committed after PV(j-1): O already holds every earlier contribution (kSharedS: wait for it).
      ensure_pfree();
      l *= alpha;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t o[32];
        tmem_ld32(tO + cc * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tO + cc * 32, o);
      }
      tc_wait_st();
    };
    // exponentials of 64 keys (S values in sv) -> bf16 pairs in pk, partial row sums in ls
    auto exp64 = [&](const uint32_t* sv, float neg, uint32_t* pk, float* ls) {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xa = fmaf(__uint_as_float(sv[2 * i]), c, neg);
        const float xb = fmaf(__uint_as_float(sv[2 * i + 1]), c, neg);
        const float a = ex2_approx(xa);
        const float b = ex2_approx(xb);
        ls[(2 * i) & 3] += a;
        ls[(2 * i + 1) & 3] += b;
        pk[i] = pack_bf16x2(a, b);
      }
    };
    // P columns of half hh are in TMEM: make them visible to the tensor pipe and tell the MMA warp
    auto release_half = [&](int hh) {
      if (p.p_halves || hh == 1) {
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          // half hh = quarters 2hh and 2hh+1 (p_halves == 0: the whole tile at hh == 1)
          const int q0 = p.p_halves ? 2 * hh : 0;
          for (int q = q0; q <= 2 * hh + 1; ++q) {
            if constexpr (k2Cta) mbar_arrive_leader(&p_part[4 * t + q]);  // the leader's MMA thread waits for both CTAs
            else mbar_arrive(&p_part[4 * t + q]);
          }
        }
      }
    };
    // the 16 P columns of key quarter q are in TMEM: make them visible to the tensor pipe and tell the MMA warp
    auto release_part = [&](int q) {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (k2Cta) mbar_arrive_leader(&p_part[4 * t + q]);
        else mbar_arrive(&p_part[4 * t + q]);
      }
    };
    // split_s: columns 64..127 of this tile's S are in registers -> the MMA warp may compute S(j+1) of the upper 64 keys
    // into them (call once per tile, after the tcgen05.wait::ld that covered those columns)
    auto hi_read = [&]() {
      if constexpr (kSharedS) {   // the whole S row is in registers: the MMA warp may overwrite the shared S buffer
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(s_free);
      } else if constexpr (k2Cta) {
        if (p.split_s) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&hi_free[t]);
        }
      }
    };
    auto wait_s = [&](int j) {
      cur_j = j;
      pf_waited = (j == 0);
      if (tr) ATT_TR(1 + t, 0);
      mbar_wait_ns(&s_full[t], sphase, p.peer_timeout_ns);
      if (tr) ATT_TR(1 + t, 1);
      sphase ^= 1;
      tc_fence_after();
    };
    // exact tile: whole S row in registers, row max, lazy rescale, exponentials
    auto tile_exact = [&](int j) {
      wait_s(j);
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t s[128];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) tmem_ld32(tSr + cc * 32, s + cc * 32);
      tc_wait_ld();
      hi_read();
      if (tr) ATT_TR(1 + t, 2);
      // 8 independent max chains (a single dependent FMNMX chain costs ~6 clk per link)
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < 128; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(s[i]));
      const float mxl = c * fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                                  fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      if (j == 0) {
        // kMode 2 with S already in log2 units: if every first-tile row max of this warp is within 2^+-40 the
        // reference stays 0 and the fast tiles need no subtraction at all (p = 2^s)
        plain = !exact && p.unit_scale && __all_sync(0xffffffffu, fabsf(mxl) <= 40.0f);
        ref = plain ? 0.0f : mxl;
      } else if (__any_sync(0xffffffffu, mxl - ref > 8.0f)) {  // lazy: only when the max grew by > 2^8
        const float nref = fmaxf(ref, mxl);
        rescale(ex2_approx(ref - nref));
        ref = nref;
      }
      if (tr) ATT_TR(1 + t, 3);
      ensure_pfree();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t pk[32];
        exp64(s + hh * 64, -ref, pk, ls);
        tmem_st32(tS + hh * 32, pk);
        release_half(hh);
        if (tr && hh == 0) ATT_TR(1 + t, 4);
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      if (tr) ATT_TR(1 + t, 5);
      if (tr) ATT_TR(1 + t, 6);
    };
    // fast tile (kMode 2, j > 0): no row max; S is read 64 columns at a time so that only half a row is live.
    // Row sums exceeding kBig shift the reference before the next tile (see the kMode comment).
    auto fast_tail = [&](float tsum, int j) {
      l += tsum;
      // sticky: O <= l * max|v|, so a running sum that stays below 2^90 keeps O finite as well; a transient excursion
      // (later scaled away by a shift) would otherwise leave inf in O behind a harmless-looking final l
      ovf |= !(l < 1e27f);
      // inf: exponent field 255 -> shift 100, l is non-finite by then and the CTA takes the exact second pass
      const int e = ((__float_as_int(tsum) >> 23) & 0xff) - 127;
      pend = tsum > 1.0995116e12f /* 2^40 */ ? (float)(e < 100 ? e : 100) : 0.0f;
      if (tr) ATT_TR(1 + t, 5);
      if (tr) ATT_TR(1 + t, 6);
    };
    auto tile_fast = [&](int j) {
      wait_s(j);
      if (__any_sync(0xffffffffu, pend != 0.0f)) {
        rescale(__int_as_float((127 - (int)pend) << 23));  // exact power of two
        ref += pend;
        pend = 0.0f;
        plain = false;
      }
      uint32_t s[64], pk[32];
      tmem_ld32(tSr, s);
      tmem_ld32(tSr + 32, s + 32);
      tc_wait_ld();
      if (tr) ATT_TR(1 + t, 2);
      if (tr) ATT_TR(1 + t, 3);
      if (plain) {
        // p = 2^s: one MUFU per element, one FADD2 and one F2FP per pair
        uint64_t ls2[2] = {0ull, 0ull};
        auto exp_pairs = [&](auto lo, auto hi) {   // pairs [lo, hi) of the 32 pairs of a 64-key half
#pragma unroll
          for (int i = decltype(lo)::value; i < decltype(hi)::value; ++i) {
            float a, b;
            // kPolyEvery = n > 0: every n-th PAIR of exponentials runs on the FMA pipe (packed polynomial)
            if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery > 0 ? kPolyEvery - 1 : 0)) {
              ex2_poly2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), a, b);
            } else {
              a = ex2_approx(__uint_as_float(s[2 * i]));
              b = ex2_approx(__uint_as_float(s[2 * i + 1]));
            }
            ls2[i & 1] = fadd2(ls2[i & 1], pack2(a, b));
            pk[i] = pack_bf16x2(a, b);
          }
        };
        using I0 = std::integral_constant<int, 0>;
        using I8 = std::integral_constant<int, 8>;
        using I16 = std::integral_constant<int, 16>;
        using I24 = std::integral_constant<int, 24>;
        using I32 = std::integral_constant<int, 32>;
        if (p.p_quarters) {
          // P leaves in four 32-key quarters; every store completes under the next quarter's first exponentials, the
          // second half of S is loaded into the registers the first half has already vacated, and the last release
          // leaves only two P.V k-steps (instead of four) between the end of the softmax and the next S MMA
          exp_pairs(I0{}, I16{});
          ensure_pfree();
          tmem_st16(tS, pk);
          tmem_ld32(tSr + 64, s);            // s[0..31] are dead: S columns 64..95
          exp_pairs(I16{}, I24{});
          release_part(0);
          if (tr) ATT_TR(1 + t, 4);
          exp_pairs(I24{}, I32{});
          tmem_st16(tS + 16, pk + 16);
          tmem_ld32(tSr + 96, s + 32);       // S columns 96..127
          tc_wait_ld();
          hi_read();
          exp_pairs(I0{}, I8{});
          release_part(1);
          exp_pairs(I8{}, I16{});
          tmem_st16(tS + 32, pk);
          exp_pairs(I16{}, I24{});
          release_part(2);
          exp_pairs(I24{}, I32{});
          tmem_st16(tS + 48, pk + 16);
          release_part(3);
          float s0, s1, s2, s3;
          unpack2(ls2[0], s0, s1);
          unpack2(ls2[1], s2, s3);
          fast_tail((s0 + s1) + (s2 + s3), j);
          return;
        }
        exp_pairs(I0{}, I32{});
        ensure_pfree();
        tmem_st32(tS, pk);
        tmem_ld32(tSr + 64, s);  // second half of the row
        tmem_ld32(tSr + 96, s + 32);
        if (p.st_overlap) {
          // the store of the first P half completes under the first 16 exponentials of the second half: its
          // tcgen05.wait::st (~150 clk when taken right after the store) no longer sits in this warp's critical path
          tc_wait_ld();
          hi_read();
          exp_pairs(I0{}, I8{});
          release_half(0);
          if (tr) ATT_TR(1 + t, 4);
          exp_pairs(I8{}, I32{});
        } else {
          release_half(0);
          if (tr) ATT_TR(1 + t, 4);
          tc_wait_ld();
          hi_read();
          exp_pairs(I0{}, I32{});
        }
        tmem_st32(tS + 32, pk);
        release_half(1);
        float s0, s1, s2, s3;
        unpack2(ls2[0], s0, s1);
        unpack2(ls2[1], s2, s3);
        fast_tail((s0 + s1) + (s2 + s3), j);
      } else {
        const float neg = -ref;
        float ls[4] = {0.f, 0.f, 0.f, 0.f};
        exp64(s, neg, pk, ls);
        ensure_pfree();
        tmem_st32(tS, pk);
        tmem_ld32(tSr + 64, s);
        tmem_ld32(tSr + 96, s + 32);
        release_half(0);
        if (tr) ATT_TR(1 + t, 4);
        tc_wait_ld();
        hi_read();
        exp64(s, neg, pk, ls);
        tmem_st32(tS + 32, pk);
        release_half(1);
        fast_tail((ls[0] + ls[1]) + (ls[2] + ls[3]), j);
      }
    };
    if (exact) {
#pragma unroll 1
      for (int j = 0; j < n_kv; ++j) tile_exact(j);
    } else {
      tile_exact(0);
#pragma unroll 1
      for (int j = 1; j < n_kv; ++j) tile_fast(j);
    }
    // final: PV(n_kv-1) complete
    mbar_wait_ns(&s_full[t], sphase, p.peer_timeout_ns);
    sphase ^= 1;
    tc_fence_after();
    if constexpr (kMode == 2) {
      if (pass == 0 && (ovf || !(l < 1e27f))) *reinterpret_cast<volatile uint32_t*>(redo_flag) = 1u;
    }
    const int row = q0 + t * ATT_TILE + (warp & 3) * 32 + lane;
    const float inv = 1.0f / l;
    __nv_bfloat16* optr = p.O + (size_t)row * p.ldo + head * 128;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t o[32];
      tmem_ld32(tO + cc * 32, o);
      tc_wait_ld();
      if (row < p.Lq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          reinterpret_cast<uint4*>(optr + cc * 32)[i] = q;
        }
      }
    }
  }
    if constexpr (kMode != 2) {
      break;
    } else {
      // did any row of this CTA leave the fp32 range?  (never for RMS-normalised q/k; the generic kernel must cope)
      tc_fence_before();
      __syncthreads();
      if constexpr (kCluster) {
        // both CTAs of a cluster share the K / V ring protocol: they repeat the sweep together or not at all
        if (threadIdx.x == 0 && pass == 0 && *reinterpret_cast<volatile uint32_t*>(redo_flag) != 0u)
          st_shared_cluster_u32(redo_flag, crank ^ 1u, 1u);
        cluster_sync_all();
      }
      tc_fence_after();
      if (pass == 1 || *reinterpret_cast<volatile uint32_t*>(redo_flag) == 0u) break;
      pass = 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kCluster) cluster_sync_all();  // no arrive / multicast may target a CTA that has already exited
  if (warp == 9) {
    tc_fence_after();
    if constexpr (k2Cta) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================================
// k_attn_fwd16 — the 2-CTA kernel above with SIXTEEN softmax warps: every 128x128 S tile is exponentiated by two warps
// per TMEM lane quadrant, each owning 64 of the 128 keys of its 32 rows.
//
// Why: the exponentials are MUFU-bound, and ONE warp per sub-partition cannot saturate that pipe — 10.8 clk per element for
// the loop `2 MUFU.EX2, FADD2, F2FP` issued by a lone warp against 8.7 clk when two warps of the sub-partition interleave
// (profiles/r01_issue_rates_microbench.txt).  In k_attn_fwd the two tiles alternate strictly (tile A exponentiates while
// tile B's MMAs run), so at any time a sub-partition had exactly one exponentiating warp: softmax(tile) ~ 1 500 clk against
// the 1 024 clk of MMA work it overlaps, tensor pipe 62 % busy.  Two warps per quadrant bring a tile's softmax to
// ~64 x 8.7 + load/store latency, and each warp's registers hold half a row (no 64-column software pipelining needed).
//
// What it takes:
//  * TMEM lane quadrant q is only accessible to warps with warp % 4 == q, so the split is by COLUMNS: warp (t, h, q) owns
//    keys [64h, 64h + 64) of rows [32q, 32q + 32) of tile t.  Its partner (t, h ^ 1, q) holds the same rows.
//  * P (bf16, 64 columns) sits at columns [32, 96) of its S tile: the lower-half warp overwrites S columns 32..63 (its
//    own, already in registers), the upper-half warp S columns 64..95 (its own) — no cross-warp hazard, no barrier.
//  * the reference-free fast tiles need nothing from the partner.  The sum guard (row sum of a tile > 2^40 -> shift the
//    reference by its exponent before the next tile) must take the same decision in both warps of a row: the warp that
//    sees its partial sum exceed the bound posts (tile << 8 | exponent) with atomicMax into a per-row slot (two slots by
//    tile parity; posted BEFORE the warp's last P release, read after the next s_full wait, so the mbarrier chain
//    P release -> MMA -> commit -> s_full orders it), and both warps read the slot at the start of every fast tile.
//  * exact tiles (tile 0 of every CTA, the exact mode, the second pass) exchange their partial row maxima through
//    shared memory with a 64-thread named barrier per (tile, quadrant) pair; the final row sums likewise.
//  * O rescales (rare) and the epilogue are split by columns as well (64 of the 128 head dimensions per warp); after a
//    rescale the pair synchronises before either warp releases P, because a P.V MMA updates all 128 columns of O.
// The loader and the MMA issuer are those of k_attn_fwd<.., k2Cta = true> (P released in four 32-key quarters: quarters
// 0, 1 come from the lower-half warps, 2, 3 from the upper-half warps, 8 arrivals per barrier as before).
constexpr int ATT16_THREADS = 576;           // warps 0-15 softmax, 16 loader, 17 MMA issuer / TMEM allocator
constexpr int ATT16_XCH = 8192;              // row-max / row-sum exchange + sum-guard slots
constexpr int ATT16_SMEM = ATT_SMEM + ATT16_XCH;

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory"); }

template <int kTrace, int kMode>
__global__ void __launch_bounds__(ATT16_THREADS, 1)
    k_attn_fwd16(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int kSlots = 2 * ATT_SLOTS;
  constexpr int kSlotBytes = ATT_TILE_BYTES / 2;
  constexpr int kKvHalf = kSlotBytes / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + 2 * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + kSlots * kSlotBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;
  uint64_t* kv_empty = bars + 1 + kSlots;
  uint64_t* s_full = bars + 1 + 2 * kSlots;
  uint64_t* p_part = bars + 3 + 2 * kSlots;       // [tile][key quarter]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11 + 2 * kSlots);
  uint32_t* redo_flag = tmem_ptr + 1;
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [step parity][tile][half][128]
  float* xsum = xmax + 2 * 2 * 2 * 128;                                            // [tile][half][128]
  int* guard = reinterpret_cast<int*>(xsum + 2 * 2 * 128);                         // [step parity][tile][128]

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 2 * ATT_TILE;
  const int n_kv = p.Lk / ATT_TILE;

  if (warp == 16 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 2);
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(&kv_full[i], 2);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      for (int q = 0; q < 4; ++q) mbar_init(&p_part[4 * i + q], 8);  // 4 warps of the owning half, in both CTAs
    }
    *redo_flag = 0u;
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc_2sm(tmem_ptr, 512);
  if (threadIdx.x < 2 * 2 * 128) guard[threadIdx.x] = 0;
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t crank = cluster_ctarank();

  uint32_t slot = 0, phase = 0;
  uint32_t pph = 0;
  uint32_t sphase = 0;
  int pass = 0;
  for (;;) {
  const bool exact = kMode != 2 || pass == 1;
  if (warp == 16) {
    if (lane == 0) {
      // ===== TMA producer (as k_attn_fwd, 2-CTA) =====
      if (pass == 0) {
        if (crank == 0) mbar_expect_tx(q_full, 4 * ATT_TILE_BYTES);
        else mbar_arrive_leader(q_full);
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d_2sm(smem_q + t * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmQ, q_full, head * 128 + h * 64,
                            q0 + t * ATT_TILE);
      }
      const int tiles_per_chunk = p.vt_chunk_len / ATT_TILE;
      const int n_chunks = p.Lk / p.vt_chunk_len;
      for (int j = 0; j < n_kv; ++j) {
        int chunk = p.first_chunk + j / tiles_per_chunk;
        if (chunk >= n_chunks) chunk -= n_chunks;
        const int within = j % tiles_per_chunk;
        if (p.chunk_flags && within == 0 && chunk != p.first_chunk) {
          uint32_t v, spins = 0;
          uint64_t t0 = 0;
          for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.chunk_flags + chunk) : "memory");
            if ((int)(v - p.flag_seq) >= 0) break;
            if (t0 == 0) t0 = global_timer_ns();
            if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > p.peer_timeout_ns) asm volatile("trap;\n");
          }
          if (t0 != 0 && p.wait_ns) atomicAdd(p.wait_ns, (unsigned long long)(global_timer_ns() - t0));
          asm volatile("fence.proxy.async.global;\n" ::: "memory");
        }
        const int kv0 = chunk * p.vt_chunk_len + within * ATT_TILE;
        mbar_wait_ns(&kv_empty[slot], phase ^ 1, p.peer_timeout_ns);
        if (crank == 0) mbar_expect_tx(&kv_full[slot], 2 * kSlotBytes);
        else mbar_arrive_leader(&kv_full[slot]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_2d_2sm(smem_kv + slot * kSlotBytes + h * kKvHalf, &tmK, &kv_full[slot], head * 128 + h * 64,
                          kv0 + (int)crank * 64);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
        mbar_wait_ns(&kv_empty[slot], phase ^ 1, p.peer_timeout_ns);
        const int koff = within * ATT_TILE;
        if (crank == 0) mbar_expect_tx(&kv_full[slot], 2 * kSlotBytes);
        else mbar_arrive_leader(&kv_full[slot]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_3d_2sm(smem_kv + slot * kSlotBytes + h * kKvHalf, &tmV, &kv_full[slot], koff + h * 64,
                          head * 128 + (int)crank * 64, chunk);
        if (++slot == kSlots) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 17) {
    if (crank == 0) {
      // ===== MMA issuer (converged warp, elected lane; as k_attn_fwd 2-CTA with P at S + 32) =====
      const bool issuer = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_bf16(256, 128);
      const uint32_t tS[2] = {tbase, tbase + 128};
      const uint32_t tPa[2] = {tbase + 32, tbase + 128 + 32};
      const uint32_t tO[2] = {tbase + 256, tbase + 384};
      auto advance = [&]() { if (++slot == kSlots) { slot = 0; phase ^= 1; } };
      auto commit = [&](uint64_t* bar) { if (issuer) umma_commit_2sm(bar); };
      auto mma_s = [&](int t, uint32_t kslot) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t da = make_sdesc_sw128(smem_u32(smem_q + t * ATT_TILE_BYTES + half * ATT_HALF_BYTES));
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + kslot * kSlotBytes + half * kKvHalf));
          if (issuer) umma_ss_2sm(tS[t], sdesc_advance(da, off), sdesc_advance(db, off), idesc, k != 0 ? 1u : 0u);
        }
      };
      auto mma_pv = [&](int t, uint32_t vslot, bool first, int qq) {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const int k = qq * 2 + kk;
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + vslot * kSlotBytes + half * kKvHalf));
          if (issuer) umma_ts_2sm(tO[t], tPa[t] + k * 8, sdesc_advance(db, off), idesc, (first && k == 0) ? 0u : 1u);
        }
      };
      if (pass == 0) mbar_wait_ns(q_full, 0, p.peer_timeout_ns);
      mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);
      tc_fence_after();
      uint32_t kslot = slot;
      advance();
      mma_s(0, kslot);
      commit(&s_full[0]);
      mma_s(1, kslot);
      commit(&s_full[1]);
      commit(&kv_empty[kslot]);
      for (int j = 0; j < n_kv; ++j) {
        const bool more = j + 1 < n_kv;
        mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);  // V_j
        const uint32_t vslot = slot;
        advance();
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (t == 0) ATT_TR(0, 0);
          // quarters 0/1 (lower-half warps) and 2/3 (upper-half warps) become ready pairwise at about the same time
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            mbar_wait_ns(&p_part[4 * t + q], pph, p.peer_timeout_ns);
            if (t == 0 && q == 0) ATT_TR(0, 1);
            if (t == 1 && q == 0) ATT_TR(0, 3);
            tc_fence_after();
            mma_pv(t, vslot, j == 0, q);
          }
          if (t == 1) commit(&kv_empty[vslot]);
          if (more) {
            if (t == 0) {
              mbar_wait_ns(&kv_full[slot], phase, p.peer_timeout_ns);  // K_{j+1}
              tc_fence_after();
              kslot = slot;
              advance();
            }
            mma_s(t, kslot);
          }
          commit(&s_full[t]);
          if (t == 0) ATT_TR(0, 2);
          if (t == 1 && more) commit(&kv_empty[kslot]);
        }
        pph ^= 1;
        ATT_TR(0, 4);
      }
    }
  } else {
    // ===== softmax: warp = 8 t + 4 h + quadrant =====
    const int t = warp >> 3, h = (warp >> 2) & 1, quad = warp & 3;
    const uint32_t lane_base = ((uint32_t)quad * 32u) << 16;
    const int rowl = quad * 32 + (int)lane;                                      // row inside the tile
    const uint32_t tSr = tmem_base + lane_base + t * 128 + h * 64;              // this thread's 64 S columns
    const uint32_t tP = tmem_base + lane_base + t * 128 + 32 + h * 32;          // its 32 P columns (keys 64h..64h+63)
    const uint32_t tO = tmem_base + lane_base + 256 + t * 128 + h * 64;         // its 64 O columns
    const int pair_id = 1 + t * 4 + quad;                                       // named barrier of the (lower, upper) pair
    const float c = p.scale_log2;
    float ref = 0.0f, l = 0.0f;
    bool ovf = false, plain = false;
    const bool tr = kTrace == 1 && quad == 0 && h == 0 && lane == 0;
    auto rescale = [&](float alpha) {
      // both warps of the pair take this path together (same per-row decision); a P.V MMA of this step updates all 128
      // columns of O, so neither may release P before the other has finished its columns
      l *= alpha;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t o[32];
        tmem_ld32(tO + cc * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tO + cc * 32, o);
      }
      tc_wait_st();
      tc_fence_before();
      pair_barrier(pair_id);
      tc_fence_after();
    };
    auto release_part = [&](int q) {
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&p_part[4 * t + q]);
    };
    auto wait_s = [&](int j) {
      if (tr) ATT_TR(1 + t, 0);
      mbar_wait_ns(&s_full[t], sphase, p.peer_timeout_ns);
      if (tr) ATT_TR(1 + t, 1);
      sphase ^= 1;
      tc_fence_after();
    };
    // exponentials of this thread's 64 keys (s) -> P (two 32-key quarters) ; returns the partial row sum.  `before_last`
    // runs between the last P store and its release (sum guard: must be ordered before the release).
    auto tile_body = [&](auto plain_c, uint32_t* s, float neg, int j, auto&& before_last) {
      constexpr bool kPlain = decltype(plain_c)::value;
      uint64_t ls2[2] = {0ull, 0ull};
      uint32_t pk[32];
      auto exp_pairs = [&](auto lo, auto hi) {
#pragma unroll
        for (int i = decltype(lo)::value; i < decltype(hi)::value; ++i) {
          float a, b;
          if constexpr (kPlain) {
            a = ex2_approx(__uint_as_float(s[2 * i]));
            b = ex2_approx(__uint_as_float(s[2 * i + 1]));
          } else {
            a = ex2_approx(fmaf(__uint_as_float(s[2 * i]), c, neg));
            b = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), c, neg));
          }
          ls2[i & 1] = fadd2(ls2[i & 1], pack2(a, b));
          pk[i] = pack_bf16x2(a, b);
        }
      };
      using I0 = std::integral_constant<int, 0>;
      using I16 = std::integral_constant<int, 16>;
      using I24 = std::integral_constant<int, 24>;
      using I32 = std::integral_constant<int, 32>;
      exp_pairs(I0{}, I16{});
      tmem_st16(tP, pk);
      exp_pairs(I16{}, I24{});          // the store completes under these
      release_part(2 * h);
      if (tr) ATT_TR(1 + t, 4);
      exp_pairs(I24{}, I32{});
      tmem_st16(tP + 16, pk + 16);
      float s0, s1, s2, s3;
      unpack2(ls2[0], s0, s1);
      unpack2(ls2[1], s2, s3);
      const float tsum = (s0 + s1) + (s2 + s3);
      before_last(tsum);
      if (tr) ATT_TR(1 + t, 5);
      release_part(2 * h + 1);
      if (tr) ATT_TR(1 + t, 6);
      return tsum;
    };
    auto nothing = [](float) {};
    // exact tile: partial row max of 64 keys, exchanged with the partner; lazy rescale; exponentials
    auto tile_exact = [&](int j) {
      wait_s(j);
      uint32_t s[64];
      tmem_ld32(tSr, s);
      tmem_ld32(tSr + 32, s + 32);
      tc_wait_ld();
      if (tr) ATT_TR(1 + t, 2);
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < 64; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(s[i]));
      const float part = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                               fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      // buffers alternate with the step parity: the next write to this one (step j + 2) lies behind the pair barrier of
      // step j + 1, which the partner only reaches after this step's read
      float* xb = xmax + (((j & 1) * 2 + t) * 2) * 128;
      xb[h * 128 + rowl] = part;
      pair_barrier(pair_id);
      const float mxl = c * fmaxf(part, xb[(h ^ 1) * 128 + rowl]);
      if (j == 0) {
        plain = !exact && p.unit_scale && __all_sync(0xffffffffu, fabsf(mxl) <= 40.0f);
        ref = plain ? 0.0f : mxl;
      } else if (__any_sync(0xffffffffu, mxl - ref > 8.0f)) {
        const float nref = fmaxf(ref, mxl);
        rescale(ex2_approx(ref - nref));
        ref = nref;
        tmem_ld32(tSr, s);            // S is reloaded instead of being kept live across the rescale (register budget)
        tmem_ld32(tSr + 32, s + 32);
        tc_wait_ld();
      }
      if (tr) ATT_TR(1 + t, 3);
      l += tile_body(std::false_type{}, s, -ref, j, nothing);
    };
    // fast tile (kMode 2, j > 0): no row max
    auto tile_fast = [&](int j) {
      wait_s(j);
      uint32_t s[64];
      tmem_ld32(tSr, s);
      tmem_ld32(tSr + 32, s + 32);
      // sum guard of the previous tile (posted by either warp of this row)
      const int gv = *reinterpret_cast<volatile int*>(&guard[(((j - 1) & 1) * 2 + t) * 128 + rowl]);
      const float pend = (gv >> 8) == j - 1 ? (float)(gv & 0xff) : 0.0f;
      tc_wait_ld();
      if (tr) ATT_TR(1 + t, 2);
      if (__any_sync(0xffffffffu, pend != 0.0f)) {
        rescale(__int_as_float((127 - (int)pend) << 23));  // exact power of two
        ref += pend;
        plain = false;
        tmem_ld32(tSr, s);            // reloaded: keeping 64 S registers live across the rescale spills in the hot path
        tmem_ld32(tSr + 32, s + 32);
        tc_wait_ld();
      }
      if (tr) ATT_TR(1 + t, 3);
      auto post_guard = [&](float tsum) {
        if (tsum > 1.0995116e12f /* 2^40 */) {
          const int e = ((__float_as_int(tsum) >> 23) & 0xff) - 127;   // inf -> 128 -> clamp 100: l is non-finite, second pass
          atomicMax(&guard[((j & 1) * 2 + t) * 128 + rowl], (j << 8) | (e < 100 ? e : 100));
        }
      };
      const float tsum = plain ? tile_body(std::true_type{}, s, 0.0f, j, post_guard)
                               : tile_body(std::false_type{}, s, -ref, j, post_guard);
      l += tsum;
      ovf |= !(l < 1e27f);
    };
    if (exact) {
#pragma unroll 1
      for (int j = 0; j < n_kv; ++j) tile_exact(j);
    } else {
      tile_exact(0);
#pragma unroll 1
      for (int j = 1; j < n_kv; ++j) tile_fast(j);
    }
    // final: PV(n_kv-1) complete; total row sum = both halves
    mbar_wait_ns(&s_full[t], sphase, p.peer_timeout_ns);
    sphase ^= 1;
    tc_fence_after();
    xsum[(t * 2 + h) * 128 + rowl] = l;
    pair_barrier(pair_id);
    const float lt = l + xsum[(t * 2 + (h ^ 1)) * 128 + rowl];
    if constexpr (kMode == 2) {
      if (pass == 0 && (ovf || !(lt < 1e27f))) *reinterpret_cast<volatile uint32_t*>(redo_flag) = 1u;
    }
    const int row = q0 + t * ATT_TILE + rowl;
    const float inv = 1.0f / lt;
    __nv_bfloat16* optr = p.O + (size_t)row * p.ldo + head * 128 + h * 64;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t o[32];
      tmem_ld32(tO + cc * 32, o);
      tc_wait_ld();
      if (row < p.Lq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          reinterpret_cast<uint4*>(optr + cc * 32)[i] = q;
        }
      }
    }
  }
    if constexpr (kMode != 2) {
      break;
    } else {
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0 && pass == 0 && *reinterpret_cast<volatile uint32_t*>(redo_flag) != 0u)
        st_shared_cluster_u32(redo_flag, crank ^ 1u, 1u);
      cluster_sync_all();
      tc_fence_after();
      if (pass == 1 || *reinterpret_cast<volatile uint32_t*>(redo_flag) == 0u) break;
      pass = 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// =====================================================================================================================
// k_attn_fwd1t — ONE 128-row query tile per CTA (256 rows per 2-CTA pair), THREE S buffers in TMEM, software-pipelined:
// the S MMA of KV step j + 3 is issued right behind the P.V MMA of step j, so the scores of the next steps are already
// in TMEM when the softmax warps finish a step.
//
// Why: in the two-tile kernels above, P aliases its S tile and TMEM is full (2 S + 2 O), so per tile the chain
// softmax(j) -> P.V(j) -> S(j+1) -> softmax(j+1) is strictly serial and the two tiles alternate: the tensor pipe idles
// while a softmax runs longer than the other tile's MMAs (clock64 timelines in profiles/: period ~3 150 clk for
// 2 x 1 024 clk of MMA work and 2 x 1 090 clk of MUFU work), and at any time only one tile's warps feed the MUFU pipe.
// With one O accumulator (128 columns) three S buffers fit (384 columns): the softmax warps go from step to step
// without waiting for any MMA in the steady state, the tensor pipe always has P.V(j-1) and S(j+2) queued, and the
// step period tends to max(MUFU time, MMA time) of ONE tile (measured: 1 028 clk for 1 024 clk of MMAs).  The price is
// K / V traffic per query row (each CTA pair now covers 256 rows instead of 512): 32 KB per CTA and step from L2, which
// runs at 25 % of its throughput (12 % before); a cluster of four would halve it but strands 16 of the 148 SMs.
//
// Softmax: kSplit warps per TMEM lane quadrant, each owning 128 / kSplit keys of its 32 rows (two warps per
// sub-partition saturate the MUFU pipe, one does not: profiles/r01_issue_rates_microbench.txt).  P (bf16) of key quarter
// q lives in columns [32 q, 32 q + 16) of its S buffer — inside the S columns of the warp that produces it, so no
// warp overwrites scores another warp still has to read; the P.V MMA takes one A address per 16-key k-step anyway.
// Reference handling (kMode 2): tile 0 fixes the reference exponent (0 if all its row maxima are within 2^+-40 and S is
// in log2 units); afterwards p = 2^(s - ref) with no row max.  A tile whose partial row sum exceeds 2^40 (or a running
// sum beyond 1e27) makes the CTA pair repeat its sweep in the exact mode (row max per tile, lazy rescale) — the
// two-tile kernels shift the reference instead, which needs a per-step rendezvous between the warps of a row that this
// pipeline does not have.  Exact tiles exchange partial row maxima through shared memory (named barrier per quadrant).
// A rescale of O waits for P.V(j-1) through the NEXT completion of s_full[(j-1) % 3]: the issuer commits that barrier
// behind every P.V (with or without a new S MMA in front of it).
// kQT: Q lives in TMEM (the A operand of the S MMA comes from TMEM like P does for P.V) instead of shared memory: its
// 32 KB buy a sixth ring stage, at the price of the third S buffer (TMEM: O [0,128) Q [128,192) S [256,512)).  On most
// boxes the K / V delivery — not the MMAs or the exponentials — set the step period of the 5-stage variant (issuer
// waiting ~850 of 1 560 clk per step for the stage, i.e. ~4 us from TMA issue to arrival with 4 stages in flight).
template <int kSplit, bool kQT = false> struct Att1 {
  static constexpr int kSoftmaxWarps = 4 * kSplit;
  static constexpr int kThreads = 32 * (kSoftmaxWarps + 2);
  static constexpr int kBufs = kQT ? 2 : 3;                      // S buffers = look-ahead of the S MMAs in KV steps
  static constexpr int kStages = kQT ? 6 : 5;                    // ring of {K_{j+kBufs} half, V_j half} pairs, 16 + 16 KB
  static constexpr int kPartBytes = ATT_TILE_BYTES / 2;          // this CTA's half of a K / V^T tile
  static constexpr int kStageBytes = 2 * kPartBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kXchBytes = (2 * kSplit + kSplit) * 128 * 4;  // row-max exchange (two step parities) + row sums
  static constexpr int kQBytes = kQT ? 0 : ATT_TILE_BYTES;
  static constexpr int kSBase = kQT ? 256 : 128;                 // first TMEM column of the S buffers
  static constexpr int kSmem = kQBytes + kStages * kStageBytes + kBarBytes + kXchBytes + 1024;
};

__device__ __forceinline__ void group_barrier(int id, int threads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}

template <int kSplit, int kMode, int kTrace = 0, int kPoly = 0, bool kQT = false>   // kPoly = n > 0: every n-th pair of exponentials of the plain tiles on the FMA pipe
__global__ void __launch_bounds__(Att1<kSplit, kQT>::kThreads, 1)
    k_attn_fwd1t(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using C = Att1<kSplit, kQT>;
  constexpr int kStages = C::kStages;
  constexpr int kBufs = C::kBufs;
  constexpr uint32_t kSBase = C::kSBase;
  constexpr int kStageBytes = C::kStageBytes;
  constexpr int kPartBytes = C::kPartBytes;
  constexpr int kKvHalf = kPartBytes / 2;
  constexpr int kLoader = C::kSoftmaxWarps, kIssuer = C::kSoftmaxWarps + 1;
  constexpr int kKeys = ATT_TILE / kSplit;      // keys per softmax warp and step
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;                        // [2 halves][128 x 128 B]
  uint8_t* smem_kv = smem + C::kQBytes;          // [stages]{K part [2 halves][64 x 128 B], V^T part [2 halves][64 x 128 B]}
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + kStages * kStageBytes);
  uint64_t* q_full = bars;                        // [1]
  uint64_t* st_full = bars + 1;                   // [stages]
  uint64_t* st_empty = bars + 1 + kStages;        // [stages]
  uint64_t* s_full = bars + 1 + 2 * kStages;      // [3 buffers]: S ready, and every earlier MMA (P.V) complete
  uint64_t* p_full = bars + 4 + 2 * kStages;      // [3 buffers]: P of the step stored by every softmax warp of the pair
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7 + 2 * kStages);
  uint32_t* redo_flag = tmem_ptr + 1;
  static_assert((7 + 2 * kStages + 1) * 8 <= C::kBarBytes, "barrier area");
  float* xmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + C::kBarBytes);  // [step parity][kSplit][128]
  float* xsum = xmax + 2 * kSplit * 128;                                                     // [kSplit][128]

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * ATT_TILE;
  const int n_kv = p.Lk / ATT_TILE;

  if (warp == kLoader && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, kQT ? 2 * C::kSoftmaxWarps : 2);   // kQT: one arrive per softmax warp of the pair (Q stored to TMEM)
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&st_full[i], 2);
      mbar_init(&st_empty[i], 1);
    }
    for (int b = 0; b < 3; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&p_full[b], 2 * C::kSoftmaxWarps);   // one elected arrive per softmax warp, of both CTAs
    }
    *redo_flag = 0u;
    fence_barrier_init();
  }
  if (warp == kIssuer) tmem_alloc_2sm(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t crank = cluster_ctarank();

  uint32_t slot = 0, phase = 0;  // KV ring position (loader: producer side, issuer: consumer side)
  uint32_t pph = 0;              // issuer: bit b = parity of the next p_full[b] phase
  uint32_t sph = 0;              // softmax warps: bit b = parity of the next s_full[b] phase
  int pass = 0;
  for (;;) {
  const bool exact = kMode != 2 || pass == 1;
  if (warp == kLoader) {
    {
      // ===== TMA producer: Q once, then K_0 .. K_{kBufs-1}, then {V_j, K_{j+kBufs}} per step (the issuer's consumption order).
      // The whole warp runs this loop converged and ONE elected lane executes the TMA / mbarrier instructions: inside an
      // `if (lane == 0)` region every UTMALDG operand is lane-varying for ptxas, and each load became an ELECT + 5 x
      // R2UR.BROADCAST loop — together with integer divisions on the (exponential-saturated) XU pipe this thread needed
      // ~1 300 clk per ring stage and paced the whole kernel (clock64 trace, profiles/r02_attn_trace_1t.txt) =====
      const bool issuer = elect_one();
      if (!kQT && pass == 0) {
        if (issuer) {
          if (crank == 0) mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);  // both CTAs' Q tiles
          else mbar_arrive_leader(q_full);
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d_2sm(smem_q + h * ATT_HALF_BYTES, &tmQ, q_full, head * 128 + h * 64, q0);
        }
      }
      const int tiles_per_chunk = p.vt_chunk_len / ATT_TILE;
      const int n_chunks = p.Lk / p.vt_chunk_len;
      // The order of the KV tiles inside a chunk is free (softmax sums commute): kv_rotate starts every CTA pair at its own
      // offset (measured: no effect; off by default).
      const int rot = p.kv_rotate ? (int)(((blockIdx.x >> 1) * 37u + blockIdx.y * 11u) % (unsigned)tiles_per_chunk) : 0;
      // Two cursors (K stream, V stream) advanced tile by tile: no integer division in this loop (I2F / MUFU.RCP / F2I
      // queue behind the exponentials on the XU pipe).
      int k_chunk = p.first_chunk, k_within = 0, v_chunk = p.first_chunk, v_within = 0;
      auto fill = [&](bool has_k, bool has_v, int j) {
        if (crank == 0) ATT_TR(1, 4);
        mbar_wait_ns(&st_empty[slot], phase ^ 1, p.peer_timeout_ns);
        if (crank == 0) ATT_TR(1, 5);
        uint8_t* st = smem_kv + slot * kStageBytes;
        if (issuer) {
          const uint32_t parts = (has_k ? 1u : 0u) + (has_v ? 1u : 0u);
          if (crank == 0) mbar_expect_tx(&st_full[slot], parts * 2 * kPartBytes);   // both CTAs' halves
          else mbar_arrive_leader(&st_full[slot]);
        }
        if (has_k) {
          const int chunk = k_chunk;
          int within = k_within + rot;
          if (within >= tiles_per_chunk) within -= tiles_per_chunk;
          const bool first_of_chunk = k_within == 0;
          if (++k_within == tiles_per_chunk) { k_within = 0; if (++k_chunk == n_chunks) k_chunk = 0; }
          if (p.chunk_flags && first_of_chunk && chunk != p.first_chunk) {  // first visit of a remote chunk: wait for its producer rank
            uint32_t v, spins = 0;
            uint64_t t0 = 0;
            for (;;) {
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p.chunk_flags + chunk) : "memory");
              if (__all_sync(0xffffffffu, (int)(v - p.flag_seq) >= 0)) break;
              if (t0 == 0) t0 = global_timer_ns();
              if ((++spins & 0x3FFu) == 0 && global_timer_ns() - t0 > p.peer_timeout_ns) asm volatile("trap;\n");
            }
            if (issuer && t0 != 0 && p.wait_ns) atomicAdd(p.wait_ns, (unsigned long long)(global_timer_ns() - t0));
            asm volatile("fence.proxy.async.global;\n" ::: "memory");
          }
          const int kv0 = chunk * p.vt_chunk_len + within * ATT_TILE;
          // one 16 KB box {64 head dimensions, 64 keys, 2 halves}: shared memory [half][key][128 B]
          if (issuer) tma_load_3d_2sm(st, &tmK, &st_full[slot], 0, kv0 + (int)crank * 64, head * 2);
        }
        if (has_v) {   // V_j's chunk flag was checked when K_j was loaded (kBufs stages earlier)
          const int chunk = v_chunk;
          int within = v_within + rot;
          if (within >= tiles_per_chunk) within -= tiles_per_chunk;
          if (++v_within == tiles_per_chunk) { v_within = 0; if (++v_chunk == n_chunks) v_chunk = 0; }
          // one 16 KB box {64 keys, 64 head dimensions, 2 key halves, 1 chunk}
          if (issuer)
            tma_load_4d_2sm(st + kPartBytes, &tmV, &st_full[slot], 0, head * 128 + (int)crank * 64, within * 2, chunk);
        }
        if (crank == 0) ATT_TR(1, 6);
        if (++slot == kStages) { slot = 0; phase ^= 1; }
      };
      for (int j = 0; j < kBufs && j < n_kv; ++j) fill(true, false, 63);
      for (int j = 0; j < n_kv; ++j) fill(j + kBufs < n_kv, true, j);
    }
  } else if (warp == kIssuer) {
    if (crank == 0) {
      // ===== MMA issuer: converged warp, one elected lane issues for both SMs =====
      const bool issuer = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = make_idesc_bf16(256, 128);
      const uint32_t tO = tbase;
      auto advance = [&]() { if (++slot == kStages) { slot = 0; phase ^= 1; } };
      auto commit = [&](uint64_t* bar) { if (issuer) umma_commit_2sm(bar); };
      auto mma_s = [&](uint32_t b, uint32_t st) {   // S buffer b = Q K^T : 8 k-steps over the head dimension
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + st * kStageBytes + half * kKvHalf));
          if constexpr (kQT) {   // A = Q from TMEM: 16 head dimensions = 8 packed columns per k-step
            if (issuer) umma_ts_2sm(tbase + kSBase + b * 128, tbase + 128 + k * 8, sdesc_advance(db, off), idesc, k != 0 ? 1u : 0u);
          } else {
            uint64_t da = make_sdesc_sw128(smem_u32(smem_q + half * ATT_HALF_BYTES));
            if (issuer) umma_ss_2sm(tbase + kSBase + b * 128, sdesc_advance(da, off), sdesc_advance(db, off), idesc, k != 0 ? 1u : 0u);
          }
        }
      };
      auto mma_pv = [&](uint32_t b, uint32_t st, bool first) {   // O += P V : 8 k-steps of 16 keys; P quarter q at column 32 q
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + st * kStageBytes + kPartBytes + half * kKvHalf));
          if (issuer)
            umma_ts_2sm(tO, tbase + kSBase + b * 128 + (k >> 1) * 32 + (k & 1) * 8, sdesc_advance(db, off), idesc,
                        (first && k == 0) ? 0u : 1u);
        }
      };
      if (pass == 0) mbar_wait_ns(q_full, 0, p.peer_timeout_ns);
      for (int j = 0; j < kBufs && j < n_kv; ++j) {
        mbar_wait_ns(&st_full[slot], phase, p.peer_timeout_ns);  // K_j
        tc_fence_after();
        mma_s((uint32_t)j, slot);
        commit(&s_full[j]);
        commit(&st_empty[slot]);
        advance();
      }
      // steady state, per KV step: two waits, 8 P.V + 8 S MMAs back to back, two commits back to back (a tcgen05.commit
      // costs the issuing thread ~80 clk: with five per step — per-tile ring entries, per-quarter P barriers — this loop
      // took 1 395 clk per step against 1 024 clk of MMA work, clock64 timeline in profiles/r02_attn_trace_1t.txt)
      uint32_t b = 0;
      for (int j = 0; j < n_kv; ++j) {
        ATT_TR(0, 0);
        mbar_wait_ns(&st_full[slot], phase, p.peer_timeout_ns);  // V_j (and K_{j+3})
        ATT_TR(0, 1);
        mbar_wait_ns(&p_full[b], (pph >> b) & 1u, p.peer_timeout_ns);
        pph ^= 1u << b;
        ATT_TR(0, 2);
        tc_fence_after();
        mma_pv(b, slot, j == 0);
        ATT_TR(0, 3);
        if (j + kBufs < n_kv) mma_s(b, slot);
        ATT_TR(0, 4);
        commit(&s_full[b]);     // S(j+3) ready / P.V(j) complete (consumed by a rescale of step j+1 and by the epilogue)
        commit(&st_empty[slot]);
        ATT_TR(0, 5);
        advance();
        b = b == kBufs - 1 ? 0 : b + 1;
      }
    }
  } else {
    // ===== softmax: warp = 4 h + quadrant, keys [kKeys h, kKeys h + kKeys) of rows [32 quadrant, +32) =====
    const int h = warp >> 2, quad = warp & 3;
    const uint32_t lane_base = ((uint32_t)quad * 32u) << 16;
    const int rowl = quad * 32 + (int)lane;
    const uint32_t tSw = tmem_base + lane_base + kSBase + h * kKeys;  // + 128 b: this thread's S columns (P: same base)
    const uint32_t tO = tmem_base + lane_base + h * kKeys;            // its O columns
    const int group_id = 1 + quad;
    const float c = p.scale_log2;
    float ref = 0.0f, l = 0.0f;
    bool bad = false, plain = false;
    auto wait_s = [&](uint32_t b) {
      mbar_wait_ns(&s_full[b], (sph >> b) & 1u, p.peer_timeout_ns);
      sph ^= 1u << b;
      tc_fence_after();
    };
    auto release = [&](uint32_t b) {   // this warp's P columns of the step are in TMEM
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&p_full[b]);
    };
    auto load_s = [&](uint32_t b, uint32_t* s) {
#pragma unroll
      for (int cc = 0; cc < kKeys / 32; ++cc) tmem_ld32(tSw + b * 128 + cc * 32, s + cc * 32);
      tc_wait_ld();
    };
    // O *= alpha (this warp's columns) at step j >= 1: P.V(j-1) must be complete — the next completion of
    // s_full[(j-1) % 3], waited for without consuming it.  The warps of the quadrant then synchronise: P.V(j) updates
    // all 128 columns of O, so none may release P before all have rescaled.
    auto rescale = [&](float alpha, uint32_t b) {
      const uint32_t bp = b == 0 ? (uint32_t)(kBufs - 1) : b - 1u;
      mbar_wait_ns(&s_full[bp], (sph >> bp) & 1u, p.peer_timeout_ns);
      tc_fence_after();
      l *= alpha;
#pragma unroll 1
      for (int cc = 0; cc < kKeys / 32; ++cc) {
        uint32_t o[32];
        tmem_ld32(tO + cc * 32, o);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
        tmem_st32(tO + cc * 32, o);
      }
      tc_wait_st();
      tc_fence_before();
      group_barrier(group_id, 32 * kSplit);
      tc_fence_after();
    };
    // exponentials of this thread's keys -> P (key quarter q at column 32 q of the S buffer); returns the partial row sum
    auto tile_body = [&](auto plain_c, uint32_t* s, float neg, uint32_t b) {
      constexpr bool kPlain = decltype(plain_c)::value;
      uint64_t ls2[2] = {0ull, 0ull};
      uint32_t pk[kKeys / 2];
      auto exp_pairs = [&](auto lo, auto hi) {
#pragma unroll
        for (int i = decltype(lo)::value; i < decltype(hi)::value; ++i) {
          float a, b2;
          if constexpr (kPlain) {
            if (kPoly > 0 && (i % (kPoly > 0 ? kPoly : 1)) == (kPoly > 0 ? kPoly - 1 : 0)) {
              ex2_poly2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), a, b2);
            } else {
              a = ex2_approx(__uint_as_float(s[2 * i]));
              b2 = ex2_approx(__uint_as_float(s[2 * i + 1]));
            }
          } else {
            a = ex2_approx(fmaf(__uint_as_float(s[2 * i]), c, neg));
            b2 = ex2_approx(fmaf(__uint_as_float(s[2 * i + 1]), c, neg));
          }
          ls2[i & 1] = fadd2(ls2[i & 1], pack2(a, b2));
          pk[i] = pack_bf16x2(a, b2);
        }
      };
      using I0 = std::integral_constant<int, 0>;
      using I16 = std::integral_constant<int, 16>;
      const uint32_t tPw = tSw + b * 128;
      exp_pairs(I0{}, I16{});
      tmem_st16(tPw, pk);
      if constexpr (kSplit == 2) {
        using I32 = std::integral_constant<int, 32>;
        exp_pairs(I16{}, I32{});          // the first store completes under these
        tmem_st16(tPw + 32, pk + 16);
      }
      release(b);
      float s0, s1, s2, s3;
      unpack2(ls2[0], s0, s1);
      unpack2(ls2[1], s2, s3);
      return (s0 + s1) + (s2 + s3);
    };
    auto tile_exact = [&](int j, uint32_t b) {
      wait_s(b);
      uint32_t s[kKeys];
      load_s(b, s);
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < kKeys; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(s[i]));
      float part = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                         fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      // exchange buffers alternate with the step parity: the next write to this one (step j + 2) lies behind the group
      // barrier of step j + 1, which every warp of the quadrant reaches only after this step's reads
      float* xb = xmax + (j & 1) * kSplit * 128;
      xb[h * 128 + rowl] = part;
      group_barrier(group_id, 32 * kSplit);
#pragma unroll
      for (int o = 1; o < kSplit; ++o) part = fmaxf(part, xb[((h + o) % kSplit) * 128 + rowl]);
      const float mxl = c * part;
      if (j == 0) {
        plain = !exact && p.unit_scale && __all_sync(0xffffffffu, fabsf(mxl) <= 40.0f);
        ref = plain ? 0.0f : mxl;
      } else if (__any_sync(0xffffffffu, mxl - ref > 8.0f)) {   // same row values in every warp of the quadrant
        const float nref = fmaxf(ref, mxl);
        rescale(ex2_approx(ref - nref), b);
        ref = nref;
        load_s(b, s);   // reloaded rather than kept live across the rescale
      }
      l += tile_body(std::false_type{}, s, -ref, b);
    };
    const bool tr = kTrace && quad == 0 && lane == 0 && (h == 0 || h == kSplit - 1);
    const int trole = h == 0 ? 1 : 2;
    auto tile_fast = [&](uint32_t b, int j) {
      if (tr) ATT_TR(trole, 0);
      wait_s(b);
      if (tr) ATT_TR(trole, 1);
      uint32_t s[kKeys];
      load_s(b, s);
      if (tr) ATT_TR(trole, 2);
      const float tsum = plain ? tile_body(std::true_type{}, s, 0.0f, b) : tile_body(std::false_type{}, s, -ref, b);
      if (tr) ATT_TR(trole, 3);
      l += tsum;
      bad |= !(tsum < 1.0995116e12f /* 2^40 */);
    };
    if constexpr (kQT) {
      if (pass == 0) {
        // this thread's share of its Q row -> TMEM (bf16 pairs per column: the K-major A operand layout of the S MMA)
        uint32_t qv[kKeys / 2];
        const int qrow = q0 + rowl;
        if (qrow < p.Lq) {
          const uint4* src = reinterpret_cast<const uint4*>(p.Q + (size_t)qrow * p.ldq + head * 128 + h * kKeys);
#pragma unroll
          for (int i = 0; i < kKeys / 8; ++i) {
            const uint4 v = __ldg(src + i);
            qv[4 * i + 0] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < kKeys / 2; ++i) qv[i] = 0u;
        }
        const uint32_t tQw = tmem_base + lane_base + 128 + h * (kKeys / 2);
        if constexpr (kKeys == 32) tmem_st16(tQw, qv);
        else tmem_st32(tQw, qv);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(q_full);
      }
    }
    {
      uint32_t b = 0;
      if (exact) {
#pragma unroll 1
        for (int j = 0; j < n_kv; ++j) { tile_exact(j, b); b = b == kBufs - 1 ? 0 : b + 1; }
      } else {
        tile_exact(0, 0);
        b = 1;
#pragma unroll 1
        for (int j = 1; j < n_kv; ++j) { tile_fast(b, j); b = b == kBufs - 1 ? 0 : b + 1; }
      }
    }
    // the issuer commits s_full behind every P.V: consume the completions of the last (up to three) steps; the last one
    // says that O is final
    for (int j = n_kv > kBufs ? n_kv - kBufs : 0; j < n_kv; ++j) wait_s((uint32_t)(j % kBufs));
    xsum[h * 128 + rowl] = l;
    group_barrier(group_id, 32 * kSplit);
    float lt = l;
#pragma unroll
    for (int o = 1; o < kSplit; ++o) lt += xsum[((h + o) % kSplit) * 128 + rowl];
    if constexpr (kMode == 2) {
      if (pass == 0 && (bad || !(lt < 1e27f))) *reinterpret_cast<volatile uint32_t*>(redo_flag) = 1u;
    }
    const int row = q0 + rowl;
    const float inv = 1.0f / lt;
    __nv_bfloat16* optr = p.O + (size_t)row * p.ldo + head * 128 + h * kKeys;
#pragma unroll 1
    for (int cc = 0; cc < kKeys / 32; ++cc) {
      uint32_t o[32];
      tmem_ld32(tO + cc * 32, o);
      tc_wait_ld();
      if (row < p.Lq) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(o[8 * i + 0]) * inv, __uint_as_float(o[8 * i + 1]) * inv);
          q.y = pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv);
          q.z = pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv);
          q.w = pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv);
          reinterpret_cast<uint4*>(optr + cc * 32)[i] = q;
        }
      }
    }
  }
    if constexpr (kMode != 2) {
      break;
    } else {
      tc_fence_before();
      __syncthreads();
      if (threadIdx.x == 0 && pass == 0 && *reinterpret_cast<volatile uint32_t*>(redo_flag) != 0u)
        st_shared_cluster_u32(redo_flag, crank ^ 1u, 1u);
      cluster_sync_all();
      tc_fence_after();
      if (pass == 1 || *reinterpret_cast<volatile uint32_t*>(redo_flag) == 0u) break;
      pass = 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kIssuer) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

unsigned long long* g_attn_trace = nullptr;  // set through g3c_attn_set_trace (profiling aid, not a product path)

int attn_fwd_v1(const void* q, const void* k, const void* vt, void* o, int Lq, int Lk, int heads,
                int ldq, int ldk, int ldo, int vt_chunk_len, float scale, cudaStream_t st, const ChunkGate* gate) {
  G3C_REQUIRE(q && k && vt && o, "attn: null operand");
  G3C_REQUIRE(Lq > 0 && Lk > 0 && heads > 0, "attn: bad sizes");
  G3C_REQUIRE(Lk % ATT_TILE == 0, "attn: Lk=%d must be a multiple of 128", Lk);
  if (vt_chunk_len <= 0) vt_chunk_len = Lk;
  G3C_REQUIRE(Lk % vt_chunk_len == 0 && vt_chunk_len % ATT_TILE == 0,
              "attn: vt_chunk_len=%d must divide Lk=%d and be a multiple of 128", vt_chunk_len, Lk);
  G3C_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldo % 8 == 0 && ldq >= heads * 128 &&
                  ldk >= heads * 128 && ldo >= heads * 128,
              "attn: leading dimensions must be >= heads*128 and multiples of 8");
  G3C_REQUIRE((reinterpret_cast<uintptr_t>(o) & 15) == 0, "attn: O must be 16-byte aligned");
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[2] = {(uint64_t)heads * 128, (uint64_t)Lq}, str[1] = {(uint64_t)ldq * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16_sw128(&tmQ, q, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)heads * 128, (uint64_t)Lk}, str[1] = {(uint64_t)ldk * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap_bf16_sw128(&tmK, k, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const int chunks = Lk / vt_chunk_len;
    uint64_t dims[3] = {(uint64_t)vt_chunk_len, (uint64_t)heads * 128, (uint64_t)chunks};
    uint64_t str[2] = {(uint64_t)vt_chunk_len * 2, (uint64_t)vt_chunk_len * 2 * heads * 128};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_bf16_sw128(&tmV, vt, 3, dims, str, box);
    if (rc) return rc;
  }
  // G3C_ATTN_MODE: softmax variant (see k_attn_fwd), 2 = sum-guarded reference (default), 0 = exact max per tile;
  // G3C_ATTN_POLY=4: every 4th exponential on the FMA pipe (measured slower, kept for A/B runs)
  static int poly = -1, mode = 2, cluster = 1, two_cta = 1, shared_s = 1, w16 = 0, one_tile = 4, poly1t = 4, short_1t = 0, q_tmem = 0;
  if (poly < 0) {
    const char* e = getenv("G3C_ATTN_POLY");
    poly = e ? atoi(e) : G3C_ATTN_POLY_DEFAULT;
    if (poly != 0) poly = 4;  // profiles/r02_attention_variants.txt: 1/6 .. 1/2 all measured slower; 1/4 kept for A/B
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<4, 0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 0, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 1, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 0, 2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 1, 2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 2, 2, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    e = getenv("G3C_ATTN_SHAREDS");
    shared_s = e ? atoi(e) != 0 : 0;
    e = getenv("G3C_ATTN_W16");
    w16 = e ? atoi(e) != 0 : 0;
    e = getenv("G3C_ATTN_1T");   // one query tile per CTA, three S buffers: softmax warps per lane quadrant (0 = off)
    one_tile = e ? atoi(e) : 4;
    if (one_tile != 0 && one_tile != 2 && one_tile != 4) one_tile = 4;
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<2>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4>::kSmem));
    e = getenv("G3C_ATTN_QT");       // 1: Q in TMEM, two S buffers, six ring stages (k_attn_fwd1t<.., kQT>)
    q_tmem = e ? atoi(e) != 0 : 0;
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 0, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4, true>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 1, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4, true>::kSmem));
    e = getenv("G3C_ATTN_SHORT1T");  // 1: short key ranges (cross-attention) on the pipelined kernel too
    short_1t = e ? atoi(e) != 0 : 0;
    e = getenv("G3C_ATTN_POLY1T");   // 2 | 3 | 4: every n-th pair of exponentials on the FMA pipe (k_attn_fwd1t<4>)
    poly1t = e ? atoi(e) : 4;   // in-step A/B on one box: 1/4 -> 0.2900 steps/s, none 0.2871, 1/3 0.2867, 1/2 0.2685
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<2, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<2>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd1t<4, 2, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Att1<4>::kSmem));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd16<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT16_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd16<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT16_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd16<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT16_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 2, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    e = getenv("G3C_ATTN_2CTA");
    two_cta = e ? atoi(e) != 0 : 1;
    e = getenv("G3C_ATTN_MODE");
    mode = e ? (atoi(e) != 0 ? 2 : 0) : 2;
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 0, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 0, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<4, 0, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 1, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 1, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 2, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0, 0, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    e = getenv("G3C_ATTN_CLUSTER");
    cluster = e ? atoi(e) != 0 : 1;
  }
  AttnParams p;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.ldo = ldo;
  p.vt_chunk_len = vt_chunk_len;
  p.O = reinterpret_cast<__nv_bfloat16*>(o);
  p.Q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.ldq = ldq;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.unit_scale = fabsf(p.scale_log2 - 1.0f) < 1e-6f;  // scale = ln 2: the caller already folded scale * log2(e) into Q
  if (p.unit_scale) p.scale_log2 = 1.0f;
  dim3 grid((Lq + 2 * ATT_TILE - 1) / (2 * ATT_TILE), heads);
  p.chunk_flags = gate ? gate->flags : nullptr;
  p.peer_timeout_ns = gate ? peer_timeout_ns() : G3C_MBAR_TIMEOUT_NS;
  p.wait_ns = gate ? gate->wait_ns : nullptr;
  p.flag_seq = gate ? gate->seq : 0;
  p.first_chunk = gate ? gate->first : 0;
  static int halves = -1;
  if (halves < 0) {
    const char* e = getenv("G3C_ATTN_PHALF");
    halves = e ? (atoi(e) != 0) : 1;
  }
  p.p_halves = halves;
  static int stovl = -1;
  if (stovl < 0) {
    const char* e = getenv("G3C_ATTN_STOVL");
    stovl = e ? (atoi(e) != 0) : 1;
  }
  p.st_overlap = stovl;
  static int splits = -1;
  if (splits < 0) {
    const char* e = getenv("G3C_ATTN_SPLITS");
    splits = e ? (atoi(e) != 0) : 0;  // measured 977 vs 1 235 TFLOP/s: 64-key UMMAs cost as much as 128-key ones
  }
  p.split_s = splits;
  static int quarters = -1;
  if (quarters < 0) {
    const char* e = getenv("G3C_ATTN_PQUARTERS");
    quarters = e ? (atoi(e) != 0) : 1;
  }
  p.p_quarters = quarters && halves;
  G3C_REQUIRE(p.first_chunk >= 0 && p.first_chunk < Lk / vt_chunk_len, "attn: first chunk %d out of range", p.first_chunk);
  p.dbg_dup_loads = 0;
  static int kv_rotate = -1;
  if (kv_rotate < 0) {
    const char* e = getenv("G3C_ATTN_ROTATE");
    kv_rotate = e ? atoi(e) != 0 : 0;   // measured: no effect on the step (2 596-2 613 ms of self-attention either way)
  }
  p.kv_rotate = kv_rotate;
  p.trace = g_attn_trace;
  const char* tmo = getenv("G3C_ATTN_TRACE_MMA_ONLY");
  const int trace_level = g_attn_trace ? ((tmo && atoi(tmo)) ? 2 : 1) : 0;
  const bool trace_2cta = trace_level && mode && two_cta && cluster && !poly && Lk > 8 * ATT_TILE;
  if (g_attn_trace && !trace_2cta) {
    const char* mo = getenv("G3C_ATTN_TRACE_MMA_ONLY");  // stamps of the MMA warp only: no perturbation of the softmax warps
    if (mode && mo && atoi(mo)) k_attn_fwd<0, 2, 2, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
    else if (mode) k_attn_fwd<0, 1, 2, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
    else k_attn_fwd<0, 1, 0, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  } else if (mode == 0 || (Lk <= 8 * ATT_TILE && !short_1t)) {
    // also the choice for short key ranges (cross-attention: 4 KV tiles): the CTA is prologue-bound there and the
    // cluster launch / second-pass agreement of the default path only add latency (0.72 vs 0.83 ms at 56 320 x 512)
    k_attn_fwd<0, 0, 0, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  } else if (cluster && (grid.x % 2 == 0 || grid.x < 16 || (two_cta && !poly))) {
    // CTA pairs along the query dimension.  An odd number of query blocks gets a padding CTA (zero-filled Q rows,
    // stores nothing): 1 / grid.x extra work (1.8 % for the 55 blocks of cp = 4) buys the 2-CTA MMAs; the multicast-only
    // variant pads small grids only (there it keeps the path covered by the unit tests).
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((grid.x + 1) & ~1u, grid.y);
    cfg.blockDim = dim3(ATT_THREADS);
    cfg.dynamicSmemBytes = ATT_SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (poly) {  // G3C_ATTN_POLY: a quarter of the exponential pairs on the FMA pipe (measured slower)
      G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<4, 0, 2, true>, tmQ, tmK, tmV, p));
    } else if (two_cta) {
      // default: one 256-row UMMA per tile for the CTA pair; each CTA stages 64 keys of a K tile / 64 head dimensions
      // of a V^T tile (TMA boxes of 64 rows)
      CUtensorMap tmK2, tmV2;
      {
        uint64_t dims[2] = {(uint64_t)heads * 128, (uint64_t)Lk}, str[1] = {(uint64_t)ldk * 2};
        uint32_t box[2] = {64, p.split_s ? 32u : 64u};   // split_s: two 32-key boxes per 64-dim half (see the loader)
        int rc = make_tmap_bf16_sw128(&tmK2, k, 2, dims, str, box);
        if (rc) return rc;
      }
      {
        const int chunks = Lk / vt_chunk_len;
        uint64_t dims[3] = {(uint64_t)vt_chunk_len, (uint64_t)heads * 128, (uint64_t)chunks};
        uint64_t str[2] = {(uint64_t)vt_chunk_len * 2, (uint64_t)vt_chunk_len * 2 * heads * 128};
        uint32_t box[3] = {64, 64, 1};
        int rc = make_tmap_bf16_sw128(&tmV2, vt, 3, dims, str, box);
        if (rc) return rc;
      }
      p.big_boxes = 0;
      if (one_tile && !p.split_s && !shared_s) {
        // K as {64 elements, Lk keys, 2 * heads halves of 64 head dimensions}: box {64, 64, 2} = both halves of 64 keys;
        // V^T as {64 keys, heads * 128 rows, chunk_len / 64 key halves, chunks}: box {64, 64, 2, 1}
        {
          uint64_t dims[3] = {64, (uint64_t)Lk, (uint64_t)heads * 2}, str[2] = {(uint64_t)ldk * 2, 128};
          uint32_t box[3] = {64, 64, 2};
          int rc = make_tmap_bf16_sw128(&tmK2, k, 3, dims, str, box);
          if (rc) return rc;
        }
        {
          const int chunks = Lk / vt_chunk_len;
          uint64_t dims[4] = {64, (uint64_t)heads * 128, (uint64_t)vt_chunk_len / 64, (uint64_t)chunks};
          uint64_t str[3] = {(uint64_t)vt_chunk_len * 2, 128, (uint64_t)vt_chunk_len * 2 * heads * 128};
          uint32_t box[4] = {64, 64, 2, 1};
          int rc = make_tmap_bf16_sw128(&tmV2, vt, 4, dims, str, box);
          if (rc) return rc;
        }
        p.big_boxes = 1;
      }
      if (one_tile && trace_level && !p.split_s && !shared_s) {
        cfg.gridDim = dim3((((unsigned)(Lq + ATT_TILE - 1) / ATT_TILE) + 1) & ~1u, grid.y);
        if (one_tile == 4 && q_tmem) {
          cfg.blockDim = dim3(Att1<4, true>::kThreads);
          cfg.dynamicSmemBytes = Att1<4, true>::kSmem;
          G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 1, 4, true>, tmQ, tmK2, tmV2, p));
        } else if (one_tile == 4) {
          cfg.blockDim = dim3(Att1<4>::kThreads);
          cfg.dynamicSmemBytes = Att1<4>::kSmem;
          G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 1, 4>, tmQ, tmK2, tmV2, p));
        } else {
          cfg.blockDim = dim3(Att1<2>::kThreads);
          cfg.dynamicSmemBytes = Att1<2>::kSmem;
          G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<2, 2, 1>, tmQ, tmK2, tmV2, p));
        }
      } else if (one_tile == 4 && q_tmem && !p.split_s && !shared_s && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
        cfg.gridDim = dim3((((unsigned)(Lq + ATT_TILE - 1) / ATT_TILE) + 1) & ~1u, grid.y);
        cfg.blockDim = dim3(Att1<4, true>::kThreads);
        cfg.dynamicSmemBytes = Att1<4, true>::kSmem;
        G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 0, 4, true>, tmQ, tmK2, tmV2, p));
      } else if (one_tile == 4 && poly1t && !p.split_s && !shared_s) {
        cfg.gridDim = dim3((((unsigned)(Lq + ATT_TILE - 1) / ATT_TILE) + 1) & ~1u, grid.y);
        cfg.blockDim = dim3(Att1<4>::kThreads);
        cfg.dynamicSmemBytes = Att1<4>::kSmem;
        if (poly1t == 3) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 0, 3>, tmQ, tmK2, tmV2, p));
        else if (poly1t == 2) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 0, 2>, tmQ, tmK2, tmV2, p));
        else G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2, 0, 4>, tmQ, tmK2, tmV2, p));
      } else if (one_tile && !p.split_s && !shared_s) {
        // default: one 128-row tile per CTA, three S buffers, software-pipelined (k_attn_fwd1t)
        cfg.gridDim = dim3((((unsigned)(Lq + ATT_TILE - 1) / ATT_TILE) + 1) & ~1u, grid.y);
        if (one_tile == 4) {
          cfg.blockDim = dim3(Att1<4>::kThreads);
          cfg.dynamicSmemBytes = Att1<4>::kSmem;
          G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<4, 2>, tmQ, tmK2, tmV2, p));
        } else {
          cfg.blockDim = dim3(Att1<2>::kThreads);
          cfg.dynamicSmemBytes = Att1<2>::kSmem;
          G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd1t<2, 2>, tmQ, tmK2, tmV2, p));
        }
      } else if (w16 && !p.split_s && !shared_s) {
        // default: sixteen softmax warps (two per TMEM lane quadrant and tile, 64 keys each)
        cfg.blockDim = dim3(ATT16_THREADS);
        cfg.dynamicSmemBytes = ATT16_SMEM;
        if (trace_level == 1) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd16<1, 2>, tmQ, tmK2, tmV2, p));
        else if (trace_level == 2) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd16<2, 2>, tmQ, tmK2, tmV2, p));
        else G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd16<0, 2>, tmQ, tmK2, tmV2, p));
      } else if (shared_s && !p.split_s) {
        if (trace_level == 1) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 1, 2, true, true, true>, tmQ, tmK2, tmV2, p));
        else if (trace_level == 2) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 2, 2, true, true, true>, tmQ, tmK2, tmV2, p));
        else G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 0, 2, true, true, true>, tmQ, tmK2, tmV2, p));
      } else if (trace_level == 1) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 1, 2, true, true>, tmQ, tmK2, tmV2, p));
      else if (trace_level == 2) G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 2, 2, true, true>, tmQ, tmK2, tmV2, p));
      else G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 0, 2, true, true>, tmQ, tmK2, tmV2, p));
    } else {
      G3C_CUDA(cudaLaunchKernelEx(&cfg, k_attn_fwd<0, 0, 2, true>, tmQ, tmK, tmV, p));
    }
  } else if (poly) {
    k_attn_fwd<4, 0, 2, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  } else {
    k_attn_fwd<0, 0, 2, false><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p);
  }
  G3C_CUDA(cudaGetLastError());
  return G3C_OK;
}

}  // namespace v1
}  // namespace g3c

extern "C" int g3c_attn_set_trace(unsigned long long* device_buffer) {
  g3c::v1::g_attn_trace = device_buffer;
  return G3C_OK;
}

namespace g3c {

// Host entry used by the engine and by the C ABI.
int attn_fwd(const void* q, const void* k, const void* vt, void* o, int Lq, int Lk, int heads,
             int ldq, int ldk, int ldo, int vt_chunk_len, float scale, cudaStream_t st, const ChunkGate* gate) {
  return v1::attn_fwd_v1(q, k, vt, o, Lq, Lk, heads, ldq, ldk, ldo, vt_chunk_len, scale, st, gate);
}

}  // namespace g3c

extern "C" int g3c_attn_fwd(const void* q, const void* k, const void* vt, void* o, int Lq, int Lk,
                            int heads, int ldq, int ldk, int ldo, int vt_chunk_len, float scale,
                            void* stream) {
  return g3c::attn_fwd(q, k, vt, o, Lq, Lk, heads, ldq, ldk, ldo, vt_chunk_len, scale,
                       (cudaStream_t)stream, nullptr);
}
