// Path D — attention forward, variant 2 (experimental, selected with G3C_ATTN_V2=1): 64-key KV tiles with
// DOUBLE-BUFFERED S per query tile.  See attn_tcgen05.cu for the default kernel and the shared description.
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
//   warp  8    TMA producer: Q once, then K_j / V_j (64 keys each) through an 8-slot ring of 16 KB tiles
//   warp  9    TMEM allocator + single-thread MMA issuer
// TMEM (512 columns), per Q tile t: S[t][0] S[t][1] (2 x 64 columns, DOUBLE-BUFFERED) and O[t] (128).
// P (bf16) overwrites the first 32 columns of its S buffer and feeds the P·V MMA straight from TMEM.
// Round-1 profiling (profiles/r01_attn_ncu_summary.txt) showed the single-buffered version latency-bound on
// the chain softmax(j) -> PV(j) -> QK(j+1) -> softmax(j+1) (tensor pipe 56 %, softmax warps idle 49 %).
// With two S buffers QK(j+2) only depends on PV(j) (which frees the buffer), so S(j+1) is already in TMEM
// when softmax(j) finishes: the softmax warpgroups run back to back and the MMA issue order per KV step is
//   PV_A(j) ; QK_A(j+2) ; PV_B(j) ; QK_B(j+2).
// Online softmax keeps a lazily updated reference max: O / row-sum are only rescaled when the row max grew
// by more than 2^8 (the rescale first waits for PV(j-1), tracked by its own commit barrier).
// A fraction of the exponentials runs on the FMA pipe (ex2_poly) because MUFU (16 ex2/clk/SM) is exactly as
// slow as the two MMAs of a step.
#include <cstdlib>

#include "kernels.h"

namespace g3c {
namespace v2 {

constexpr int ATT_THREADS = 320;
constexpr int ATT_QTILE = 128;                   // rows per Q tile
constexpr int ATT_KV = 64;                       // keys per KV tile
constexpr int ATT_QTILE_BYTES = 128 * 128 * 2;   // 32 KB: two 64-column halves of 128 x 128 B
constexpr int ATT_QHALF_BYTES = 128 * 128;       // 16 KB
constexpr int ATT_SLOT_BYTES = 64 * 128 * 2;     // 16 KB: K tile (2 halves of 64 x 128 B) or V^T tile (128 x 128 B)
constexpr int ATT_KHALF_BYTES = 64 * 128;        // 8 KB
constexpr int ATT_SLOTS = 8;
constexpr int ATT_SMEM = 2 * ATT_QTILE_BYTES + ATT_SLOTS * ATT_SLOT_BYTES + 512 + 1024;

struct AttnParams {
  int Lq, Lk, heads;
  int ldo;
  int vt_chunk_len;
  __nv_bfloat16* O;
  float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe: Cody-Waite split + degree-3 minimax polynomial (rel. err 7.5e-5, far below the bf16
// rounding of P).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;  // 1.5 * 2^23: low mantissa bits now hold round(x)
  const float n = xr - 12582912.0f;
  const float f = x - n;  // [-0.5, 0.5]
  const float p = fmaf(fmaf(fmaf(0.05517165f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

template <int kPolyEvery>
__global__ void __launch_bounds__(ATT_THREADS, 1)
    k_attn_fwd(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;                          // [2 tiles][2 halves][128 x 128 B]
  uint8_t* smem_kv = smem + 2 * ATT_QTILE_BYTES;   // [slots][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_kv + ATT_SLOTS * ATT_SLOT_BYTES);
  uint64_t* q_full = bars;                           // [1]
  uint64_t* kv_full = bars + 1;                      // [slots]
  uint64_t* kv_empty = bars + 1 + ATT_SLOTS;         // [slots]
  uint64_t* s_full = bars + 1 + 2 * ATT_SLOTS;       // [tile][buf] -> 4
  uint64_t* p_full = s_full + 4;                     // [tile][buf] -> 4
  uint64_t* pv_done = p_full + 4;                    // [tile][buf] -> 4
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 4);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * 2 * ATT_QTILE;
  const int n_kv = p.Lk / ATT_KV;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    if (lane == 0) {
      // ===== TMA producer =====  order: K0, K1, then per step j: V_j, K_{j+2}
      mbar_expect_tx(q_full, 2 * ATT_QTILE_BYTES);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_2d(smem_q + t * ATT_QTILE_BYTES + h * ATT_QHALF_BYTES, &tmQ, q_full, head * 128 + h * 64,
                      q0 + t * ATT_QTILE);
      uint32_t slot = 0, phase = 0;
      auto load_k = [&](int j) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_expect_tx(&kv_full[slot], ATT_SLOT_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          tma_load_2d(smem_kv + slot * ATT_SLOT_BYTES + h * ATT_KHALF_BYTES, &tmK, &kv_full[slot],
                      head * 128 + h * 64, j * ATT_KV);
        if (++slot == ATT_SLOTS) { slot = 0; phase ^= 1; }
      };
      auto load_v = [&](int j) {
        mbar_wait(&kv_empty[slot], phase ^ 1);
        mbar_expect_tx(&kv_full[slot], ATT_SLOT_BYTES);
        const int kv0 = j * ATT_KV;
        const int chunk = kv0 / p.vt_chunk_len;
        tma_load_3d(smem_kv + slot * ATT_SLOT_BYTES, &tmV, &kv_full[slot], kv0 - chunk * p.vt_chunk_len,
                    head * 128, chunk);
        if (++slot == ATT_SLOTS) { slot = 0; phase ^= 1; }
      };
      load_k(0);
      if (n_kv > 1) load_k(1);
      for (int j = 0; j < n_kv; ++j) {
        load_v(j);
        if (j + 2 < n_kv) load_k(j + 2);
      }
    }
  } else if (warp == 9) {
    if (lane == 0) {
      // ===== MMA issuer =====
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_KV);
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128);
      uint32_t slot = 0, phase = 0;
      auto take = [&]() {  // wait for the next ring slot to be filled, return its index
        mbar_wait(&kv_full[slot], phase);
        const uint32_t s = slot;
        if (++slot == ATT_SLOTS) { slot = 0; phase ^= 1; }
        return s;
      };
      auto mma_qk = [&](int t, int b, uint32_t kslot) {
        // S[t][b] = Q_t K^T : 8 k-steps over the head dimension (two 64-wide halves)
        const uint32_t d = tmem_base + t * 256 + b * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          const uint64_t da = make_sdesc_sw128(smem_u32(smem_q + t * ATT_QTILE_BYTES + half * ATT_QHALF_BYTES));
          const uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + kslot * ATT_SLOT_BYTES + half * ATT_KHALF_BYTES));
          umma_ss(d, sdesc_advance(da, off), sdesc_advance(db, off), idesc_qk, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t * 2 + b]);
      };
      auto mma_pv = [&](int t, int b, uint32_t vslot, bool first) {
        // O[t] += P[t][b] V : 4 k-steps over the 64 keys; A = P from TMEM (bf16 pairs per column)
        const uint32_t d = tmem_base + t * 256 + 128;
        const uint32_t a = tmem_base + t * 256 + b * 64;
        const uint64_t db = make_sdesc_sw128(smem_u32(smem_kv + vslot * ATT_SLOT_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ts(d, a + k * 8, sdesc_advance(db, k * 32), idesc_pv, (first && k == 0) ? 0u : 1u);
        umma_commit(&pv_done[t * 2 + b]);
      };
      mbar_wait(q_full, 0);
      {
        const uint32_t k0 = take();
        tc_fence_after();
        mma_qk(0, 0, k0);
        mma_qk(1, 0, k0);
        umma_commit(&kv_empty[k0]);
        if (n_kv > 1) {
          const uint32_t k1 = take();
          tc_fence_after();
          mma_qk(0, 1, k1);
          mma_qk(1, 1, k1);
          umma_commit(&kv_empty[k1]);
        }
      }
      for (int j = 0; j < n_kv; ++j) {
        const int b = j & 1;
        const uint32_t par = (j >> 1) & 1;
        const bool more = j + 2 < n_kv;
        const uint32_t vslot = take();  // V_j
        // ---- tile A
        mbar_wait(&p_full[0 * 2 + b], par);
        tc_fence_after();
        mma_pv(0, b, vslot, j == 0);
        uint32_t kslot = 0;
        if (more) {
          kslot = take();  // K_{j+2}
          tc_fence_after();
          mma_qk(0, b, kslot);
        }
        // ---- tile B
        mbar_wait(&p_full[1 * 2 + b], par);
        tc_fence_after();
        mma_pv(1, b, vslot, j == 0);
        umma_commit(&kv_empty[vslot]);
        if (more) {
          mma_qk(1, b, kslot);
          umma_commit(&kv_empty[kslot]);
        }
      }
    }
  } else {
    // ===== softmax warpgroups (warps 0-3: tile A, warps 4-7: tile B) =====
    const int t = warp >> 2;
    const uint32_t lane_base = ((warp & 3u) * 32u) << 16;
    const uint32_t tS = tmem_base + lane_base + t * 256;
    const uint32_t tO = tS + 128;
    const float c = p.scale_log2;
    float m_used = 0.0f;  // reference max (raw score units) the stored exponentials are relative to
    float l = 0.0f;       // running row sum (relative to m_used)
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[t * 2 + b], (j >> 1) & 1);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32(tS + b * 64, s);
      tmem_ld32(tS + b * 64 + 32, s + 32);
      tc_wait_ld();
      // 8 independent max chains (a single dependent FMNMX chain costs ~6 clk per link)
      float mxs[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mxs[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < 64; ++i) mxs[i & 7] = fmaxf(mxs[i & 7], __uint_as_float(s[i]));
      const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                             fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      if (j == 0) {
        m_used = mx;
      } else {
        const bool grow = (mx - m_used) * c > 8.0f;
        if (__any_sync(0xffffffffu, grow)) {
          // O may still be accumulating PV(j-1): wait for its commit before the read-modify-write
          mbar_wait(&pv_done[t * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1);
          tc_fence_after();
          const float m_new = fmaxf(m_used, mx);
          const float alpha = ex2_approx((m_used - m_new) * c);
          m_used = m_new;
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
        }
      }
      const float neg = -m_used * c;
      uint32_t pk[32];
      float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float xa = fmaf(__uint_as_float(s[2 * i]), c, neg);
        const float xb = fmaf(__uint_as_float(s[2 * i + 1]), c, neg);
        const float a = ex2_approx(xa);
        const float bb = (kPolyEvery > 0 && (i % (kPolyEvery / 2 > 0 ? kPolyEvery / 2 : 1)) == 0) ? ex2_poly(xb)
                                                                                                    : ex2_approx(xb);
        ls[(2 * i) & 3] += a;
        ls[(2 * i + 1) & 3] += bb;
        pk[i] = pack_bf16x2(a, bb);
      }
      l += (ls[0] + ls[1]) + (ls[2] + ls[3]);
      tmem_st32(tS + b * 64, pk);
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[t * 2 + b]);
    }
    // final: PV(n_kv-1) complete
    mbar_wait(&pv_done[t * 2 + ((n_kv - 1) & 1)], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    const int row = q0 + t * ATT_QTILE + (warp & 3) * 32 + lane;
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

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_fwd_v2(const void* q, const void* k, const void* vt, void* o, int Lq, int Lk, int heads,
             int ldq, int ldk, int ldo, int vt_chunk_len, float scale, cudaStream_t st) {
  G3C_REQUIRE(q && k && vt && o, "attn: null operand");
  G3C_REQUIRE(Lq > 0 && Lk > 0 && heads > 0, "attn: bad sizes");
  G3C_REQUIRE(Lk % 128 == 0, "attn: Lk=%d must be a multiple of 128", Lk);
  if (vt_chunk_len <= 0) vt_chunk_len = Lk;
  G3C_REQUIRE(Lk % vt_chunk_len == 0 && vt_chunk_len % 128 == 0,
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
    uint32_t box[2] = {64, ATT_KV};
    int rc = make_tmap_bf16_sw128(&tmK, k, 2, dims, str, box);
    if (rc) return rc;
  }
  {
    const int chunks = Lk / vt_chunk_len;
    uint64_t dims[3] = {(uint64_t)vt_chunk_len, (uint64_t)heads * 128, (uint64_t)chunks};
    uint64_t str[2] = {(uint64_t)vt_chunk_len * 2, (uint64_t)vt_chunk_len * 2 * heads * 128};
    uint32_t box[3] = {ATT_KV, 128, 1};
    int rc = make_tmap_bf16_sw128(&tmV, vt, 3, dims, str, box);
    if (rc) return rc;
  }
  // fraction of exponentials evaluated on the FMA pipe: 1/kPolyEvery (0 = none); G3C_ATTN_POLY overrides
  static int poly = -1;
  if (poly < 0) {
    const char* e = getenv("G3C_ATTN_POLY");
    poly = e ? atoi(e) : 4;
    if (poly != 0 && poly != 2 && poly != 4 && poly != 8) poly = 4;
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    G3C_CUDA(cudaFuncSetAttribute(k_attn_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
  }
  AttnParams p;
  p.Lq = Lq;
  p.Lk = Lk;
  p.heads = heads;
  p.ldo = ldo;
  p.vt_chunk_len = vt_chunk_len;
  p.O = reinterpret_cast<__nv_bfloat16*>(o);
  p.scale_log2 = scale * 1.4426950408889634f;
  dim3 grid((Lq + 2 * ATT_QTILE - 1) / (2 * ATT_QTILE), heads);
  switch (poly) {
    case 0: k_attn_fwd<0><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
    case 2: k_attn_fwd<2><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
    case 8: k_attn_fwd<8><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
    default: k_attn_fwd<4><<<grid, ATT_THREADS, ATT_SMEM, st>>>(tmQ, tmK, tmV, p); break;
  }
  G3C_CUDA(cudaGetLastError());
  return G3C_OK;
}

}  // namespace v2
}  // namespace g3c
