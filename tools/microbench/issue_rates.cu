// Single-warp / two-warp issue-cost microbenchmark for the softmax instruction mix of k_attn_fwd (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rates issue_rates.cu ; run: ./issue_rates
// Every test works on 64 independent registers per thread (64-way ILP), repeats the block ITER times in a
// non-unrolled loop and reports clk per block for warp 0 with 1, 2 and 4 warps resident per SM sub-partition.
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#define ITER 64

__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t cvt2(float a, float b) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a)); return r; }
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float xr = x + 12582912.0f;
  const float n = xr - 12582912.0f;
  const float f = x - n;
  const float p = fmaf(fmaf(fmaf(0.05517165f, f, 0.24261113f), f, 0.69326097f), f, 0.99992806f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}

template <int TEST>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, float c, float neg) {
  float s[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) s[i] = out[threadIdx.x * 64 + i];
  uint32_t pk[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) pk[i] = 0;
  float ls[4] = {0.f, 0.f, 0.f, 0.f};
  float mx[4] = {-1e30f, -1e30f, -1e30f, -1e30f};
  uint64_t ls2[2] = {0ull, 0ull};
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    if constexpr (TEST == 0) {  // 64 MUFU.EX2
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = ex2(s[i]);
    } else if constexpr (TEST == 1) {  // 64 FFMA reg,reg,reg
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = fmaf(s[i], c, neg);
    } else if constexpr (TEST == 2) {  // 32 FFMA2
      const uint64_t c2 = pack2(c, c), n2 = pack2(neg, neg);
#pragma unroll
      for (int i = 0; i < 32; ++i) { uint64_t r = ffma2(pack2(s[2 * i], s[2 * i + 1]), c2, n2); unpack2(r, s[2 * i], s[2 * i + 1]); }
    } else if constexpr (TEST == 3) {  // 64 FADD
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = s[i] + neg;
    } else if constexpr (TEST == 4) {  // 32 FADD2
      const uint64_t n2 = pack2(neg, neg);
#pragma unroll
      for (int i = 0; i < 32; ++i) { uint64_t r = fadd2(pack2(s[2 * i], s[2 * i + 1]), n2); unpack2(r, s[2 * i], s[2 * i + 1]); }
    } else if constexpr (TEST == 5) {  // 64 FMNMX
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = fmaxf(s[i], s[(i + 1) & 63] * 1.0f);
    } else if constexpr (TEST == 6) {  // 32 FMNMX3 folding 64 elements into 4 chains
#pragma unroll
      for (int i = 0; i < 32; ++i) mx[i & 3] = fmax3(mx[i & 3], s[2 * i], s[2 * i + 1]);
      s[it & 63] += mx[0];
    } else if constexpr (TEST == 7) {  // 64 FMNMX folding 64 elements into 8 chains
      float m8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m8[i] = s[i];
#pragma unroll
      for (int i = 8; i < 64; ++i) m8[i & 7] = fmaxf(m8[i & 7], s[i]);
      s[it & 63] += fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
    } else if constexpr (TEST == 8) {  // 32 F2FP (cvt.rn.bf16x2.f32)
#pragma unroll
      for (int i = 0; i < 32; ++i) { pk[i] = cvt2(s[2 * i], s[2 * i + 1]); s[2 * i] = __uint_as_float(pk[i]); }
    } else if constexpr (TEST == 9) {  // mode-0 exp loop on 64 elements: FFMA, MUFU, FADD, F2FP/2
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = ex2(fmaf(s[2 * i], c, neg)), b = ex2(fmaf(s[2 * i + 1], c, neg));
        ls[(2 * i) & 3] += a; ls[(2 * i + 1) & 3] += b;
        pk[i] = cvt2(a, b); s[2 * i] = a; s[2 * i + 1] = b;
      }
    } else if constexpr (TEST == 10) {  // packed exp loop: FFMA2, 2 MUFU, FADD2, F2FP
      const uint64_t c2 = pack2(c, c), n2 = pack2(neg, neg);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float xa, xb; unpack2(ffma2(pack2(s[2 * i], s[2 * i + 1]), c2, n2), xa, xb);
        const float a = ex2(xa), b = ex2(xb);
        ls2[i & 1] = fadd2(ls2[i & 1], pack2(a, b));
        pk[i] = cvt2(a, b); s[2 * i] = a; s[2 * i + 1] = b;
      }
    } else if constexpr (TEST == 11) {  // 64 MUFU + 64 FMNMX (max of a second array folded meanwhile)
      float m8[8] = {mx[0], mx[1], mx[2], mx[3], mx[0], mx[1], mx[2], mx[3]};
#pragma unroll
      for (int i = 0; i < 64; ++i) { m8[i & 7] = fmaxf(m8[i & 7], s[i]); s[i] = ex2(s[i]); }
      mx[0] = fmaxf(fmaxf(m8[0], m8[4]), mx[0]); mx[1] = fmaxf(m8[1], m8[5]); mx[2] = fmaxf(m8[2], m8[6]); mx[3] = fmaxf(m8[3], m8[7]);
    } else if constexpr (TEST == 12) {  // 64 MUFU + 64 FFMA independent
#pragma unroll
      for (int i = 0; i < 32; ++i) { s[i] = ex2(s[i]); s[32 + i] = fmaf(s[32 + i], c, neg); s[i] = ex2(s[i]); s[32 + i] = fmaf(s[32 + i], c, neg); }
    } else if constexpr (TEST == 13) {  // 64 poly exp2 (FMA/ALU pipes only)
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = ex2_poly(s[i]);
    } else if constexpr (TEST == 14) {  // exp loop with every 4th exponential on the FMA pipe
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = ex2(fmaf(s[2 * i], c, neg));
        const float xb = fmaf(s[2 * i + 1], c, neg);
        const float b = (i & 1) ? ex2_poly(xb) : ex2(xb);
        ls[(2 * i) & 3] += a; ls[(2 * i + 1) & 3] += b;
        pk[i] = cvt2(a, b); s[2 * i] = a; s[2 * i + 1] = b;
      }
    } else if constexpr (TEST == 15) {  // 64 FMUL by immediate (imm-form)
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = s[i] * 1.0009765625f;
    } else if constexpr (TEST == 16) {  // 64 FFMA imm-form: x*imm + imm
#pragma unroll
      for (int i = 0; i < 64; ++i) s[i] = fmaf(s[i], 1.0009765625f, -0.25f);
    }
  }
  long long t1 = clock64();
  float acc = ls[0] + ls[1] + ls[2] + ls[3] + mx[0] + mx[1] + mx[2] + mx[3];
  float q0, q1, q2, q3; unpack2(ls2[0], q0, q1); unpack2(ls2[1], q2, q3);
  acc += q0 + q1 + q2 + q3;
#pragma unroll
  for (int i = 0; i < 64; ++i) acc += s[i];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += __uint_as_float(pk[i]);
  out[threadIdx.x * 64] = acc;
  if (threadIdx.x == 0) clk[0] = (t1 - t0);
}

template <int TEST>
void run(const char* name, float* out, long long* clk) {
  printf("%-58s", name);
  for (int threads : {128, 256, 512}) {
    cudaMemset(out, 0, 512 * 64 * 4);
    k<TEST><<<1, threads>>>(out, clk, 1.0001f, -0.5f);
    k<TEST><<<1, threads>>>(out, clk, 1.0001f, -0.5f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" error %s\n", cudaGetErrorString(e)); return; }
    long long h;
    cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
    printf("  %dw/smsp: %7.1f", threads / 128, (double)h / ITER);
  }
  printf("   clk per block of 64 elements\n");
}

int main() {
  float* out; long long* clk;
  cudaMalloc(&out, 512 * 64 * 4); cudaMalloc(&clk, 8);
  run<0>("64 MUFU.EX2", out, clk);
  run<1>("64 FFMA (3 reg)", out, clk);
  run<16>("64 FFMA (imm, imm)", out, clk);
  run<15>("64 FMUL (imm)", out, clk);
  run<2>("32 FFMA2 (64 elements)", out, clk);
  run<3>("64 FADD", out, clk);
  run<4>("32 FADD2 (64 elements)", out, clk);
  run<5>("64 FMNMX (+64 FMUL)", out, clk);
  run<7>("64 FMNMX into 8 chains", out, clk);
  run<6>("32 FMNMX3 into 4 chains (64 elements)", out, clk);
  run<8>("32 F2FP.BF16 pack (64 elements)", out, clk);
  run<9>("exp loop scalar: 64x(FFMA,MUFU,FADD)+32 F2FP", out, clk);
  run<10>("exp loop packed: 32x(FFMA2,2 MUFU,FADD2,F2FP)", out, clk);
  run<11>("64 MUFU + 64 FMNMX interleaved", out, clk);
  run<12>("64 MUFU + 64 FFMA interleaved", out, clk);
  run<13>("64 poly exp2 (no MUFU)", out, clk);
  run<14>("exp loop scalar, 1/4 poly", out, clk);
  return 0;
}
