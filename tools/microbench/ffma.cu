// FFMA vs FFMA2 (fma.rn.f32x2) issue-rate microbenchmark for sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma ffma.cu && ./ffma
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  // 32 independent accumulators per thread
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = threadIdx.x * 0.001f + i;
  float x = s, y = s * 0.5f;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[i], x, y);            // 2 distinct non-acc operands (reuse-friendly)
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 32; ++i) a[i] = fmaf(a[(i + 1) & 31], x, a[i]);  // 3 register operands, all different
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        unsigned long long acc, xx, yy;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(a[i]), "f"(a[i + 1]));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(xx) : "f"(x), "f"(x));
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(yy) : "f"(y), "f"(y));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(xx), "l"(yy));
        asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(acc));
      }
    }
  }
  float r = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4);
  int iters = 4000;
  k<MODE><<<148 * 8, 256>>>(out, 10, 1.0001f);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = 148.0 * 8 * 256 * 32.0 * iters;
  printf("%-28s %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, ms, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  run<0>("FFMA acc*x+y");
  run<1>("FFMA 3 distinct regs");
  run<2>("FFMA2 (fma.rn.f32x2)");
  return 0;
}
