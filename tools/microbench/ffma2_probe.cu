// Does packed fma.rn.f32x2 (FFMA2) reach the scalar FMA rate in the correlation's inner-loop shape -- accumulator and one
// operand changing every instruction, the other reused, 16-byte shared-memory loads interleaved?  (tools/microbench/ffma.cu
// measured it with both non-accumulator operands constant.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

constexpr int TY = 6, WIN = 9, HROWS = TY + WIN - 1;

// MODE bit0: packed (FFMA2) / scalar; bit1: operands from shared memory (LDS.128 per row) / from registers
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 1) k(float* out, int iters, long long* cyc) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < (HROWS + TY) * 40 * 16; i += NT) sm[i] = 1.0f + 1e-6f * i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* ap = sm + HROWS * 40 * 16 + lane * 16 + (((0 ^ (lane >> 1)) & 3) << 2);
  const float* bp = sm + ((lane + warp % 9) * 16) + (((0 ^ ((lane + warp % 9) >> 1)) & 3) << 2);
  constexpr bool PACK = MODE & 1, LDS = MODE & 2;
  unsigned long long acc2[TY][WIN];
  float acc1[TY][WIN];
#pragma unroll
  for (int p = 0; p < TY; ++p)
#pragma unroll
    for (int q = 0; q < WIN; ++q) { acc2[p][q] = 0ull; acc1[p][q] = 0.f; }
  ulonglong2 breg[HROWS];
  if (!LDS) {
#pragma unroll
    for (int j = 0; j < HROWS; ++j) breg[j] = *reinterpret_cast<const ulonglong2*>(bp + j * 40 * 16);
  }
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    ulonglong2 a[TY];
#pragma unroll
    for (int p = 0; p < TY; ++p) a[p] = *reinterpret_cast<const ulonglong2*>(ap + p * 32 * 16 + (it & 3) * 4 * 0);
#pragma unroll
    for (int j = 0; j < HROWS; ++j) {
      ulonglong2 b = LDS ? *reinterpret_cast<const ulonglong2*>(bp + j * 40 * 16) : breg[j];
      if (MODE & 4) {
        // component-major order: consecutive instructions never touch the same accumulator
#pragma unroll
        for (int kk = 0; kk < (PACK ? 2 : 4); ++kk) {
#pragma unroll
          for (int p = 0; p < TY; ++p) {
            const int d = j - p;
            if (d >= 0 && d < WIN) {
              if (PACK) {
                acc2[p][d] = ffma2(kk ? a[p].y : a[p].x, kk ? b.y : b.x, acc2[p][d]);
              } else {
                const unsigned long long av = kk < 2 ? a[p].x : a[p].y, bv = kk < 2 ? b.x : b.y;
                const float af = __uint_as_float((unsigned)((kk & 1) ? (av >> 32) : av)), bf = __uint_as_float((unsigned)((kk & 1) ? (bv >> 32) : bv));
                acc1[p][d] = fmaf(af, bf, acc1[p][d]);
              }
            }
          }
        }
      } else
#pragma unroll
      for (int p = 0; p < TY; ++p) {
        const int d = j - p;
        if (d >= 0 && d < WIN) {
          if (PACK) {
            acc2[p][d] = ffma2(a[p].x, b.x, acc2[p][d]);
            acc2[p][d] = ffma2(a[p].y, b.y, acc2[p][d]);
          } else {
            float s = acc1[p][d];
            s = fmaf(__uint_as_float((unsigned)a[p].x), __uint_as_float((unsigned)b.x), s);
            s = fmaf(__uint_as_float((unsigned)(a[p].x >> 32)), __uint_as_float((unsigned)(b.x >> 32)), s);
            s = fmaf(__uint_as_float((unsigned)a[p].y), __uint_as_float((unsigned)b.y), s);
            s = fmaf(__uint_as_float((unsigned)(a[p].y >> 32)), __uint_as_float((unsigned)(b.y >> 32)), s);
            acc1[p][d] = s;
          }
        }
      }
    }
    if (LDS) asm volatile("" ::: "memory");
  }
  const long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) cyc[0] = t1 - t0;
  float r = 0;
#pragma unroll
  for (int p = 0; p < TY; ++p)
#pragma unroll
    for (int q = 0; q < WIN; ++q) r += acc1[p][q] + __uint_as_float((unsigned)acc2[p][q]) + __uint_as_float((unsigned)(acc2[p][q] >> 32));
  out[blockIdx.x * NT + threadIdx.x] = r;
}

template <int MODE, int NT>
void run(const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * NT * 4);
  cudaMalloc(&cyc, 8);
  const size_t smem = (HROWS + TY) * 40 * 16 * 4;
  cudaFuncSetAttribute(k<MODE, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int iters = 20000;
  k<MODE, NT><<<148, NT, smem>>>(out, 10, cyc);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE, NT><<<148, NT, smem>>>(out, iters, cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = 148.0 * NT * (double)TY * WIN * 4 * iters;
  long long hc = 0;
  cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %2d warps  %.3f ms  %6.1f FMA/clk/SM by clock64 (SM clock %.0f MHz; %.1f at a nominal 1.965 GHz)  (%s)\n", name, NT / 32, ms,
         (double)NT * TY * WIN * 4 * iters / (double)hc, hc / (ms * 1e3), fma / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<4, 256>("scalar FFMA, registers, component-major");
  run<5, 256>("FFMA2,       registers, component-major");
  run<6, 256>("scalar FFMA, LDS.128,   component-major");
  run<7, 256>("FFMA2,       LDS.128,   component-major");
  run<6, 384>("scalar FFMA, LDS.128,   component-major");
  run<7, 384>("FFMA2,       LDS.128,   component-major");
  run<0, 256>("scalar FFMA, operands in registers");
  run<1, 256>("FFMA2,       operands in registers");
  run<2, 256>("scalar FFMA, LDS.128 per window row");
  run<3, 256>("FFMA2,       LDS.128 per window row");
  run<2, 288>("scalar FFMA, LDS.128 per window row");
  run<3, 288>("FFMA2,       LDS.128 per window row");
  run<2, 384>("scalar FFMA, LDS.128 per window row");
  run<3, 384>("FFMA2,       LDS.128 per window row");
  run<3, 512>("FFMA2,       LDS.128 per window row");
  return 0;
}
