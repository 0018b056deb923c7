// What one producer thread pays per pipeline step (B200, sm_100a): issue cost (back to back, one thread) of the mbarrier and
// TMA instructions a ring-buffer producer executes, and the round trip of one 16 KB operand tile as a tensor box vs as
// a bulk copy.  Behind DESIGN.md 3.3's "the K loops are bound by the producer's issue chain" and corr_planar.cu's rotating
// requester.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../upflow_pytorch_b200/csrc -o tma_issue_probe tma_issue_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace upf;
namespace upf {
void set_error(const char*, ...) {}
void count_launch(int) {}
void note_kernel(const char*) {}
int g_tc_pdl = 0;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// out[i] = cycles per operation
__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap map32, const __grid_constant__ CUtensorMap map128,
                                            const float* src, long long* out, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint32_t b0 = smem_u32(&bar[0]), b1 = smem_u32(&bar[1]), b2 = smem_u32(&bar[2]), b3 = smem_u32(&bar[3]);
  const uint32_t dst = smem_u32(base);
  long long t0, t1;
  // (a) try_wait on a completed phase: complete phase 0 of bar0 once, then wait on parity 0 repeatedly
  mbar_expect_tx(b0, 0);
  t0 = clock64();
  for (int r = 0; r < reps; ++r) mbar_wait(b0, 0);
  t1 = clock64(); out[0] = (t1 - t0) / reps;
  // (b) arrive.expect_tx (tx = 0 completes the phase every time)
  t0 = clock64();
  for (int r = 0; r < reps; ++r) mbar_expect_tx(b1, 0);
  t1 = clock64(); out[1] = (t1 - t0) / reps;
  // (c) fence.proxy.async
  t0 = clock64();
  for (int r = 0; r < reps; ++r) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  t1 = clock64(); out[2] = (t1 - t0) / reps;
  // (d) issue of bulk copies (4 KB each, 8 rotating slots), completion awaited once at the end
  uint32_t ph = 0;
  mbar_expect_tx(b2, (uint32_t)(reps * 4096));
  t0 = clock64();
  for (int r = 0; r < reps; ++r) bulk_g2s(dst + (r & 7) * 4096, src + (size_t)(r & 63) * 1024, 4096, b2);
  t1 = clock64(); out[3] = (t1 - t0) / reps;
  mbar_wait(b2, ph); ph ^= 1;
  long long t2 = clock64(); out[4] = (t2 - t0) / reps;              // incl. drain: throughput bound of 4 KB bulk copies
  // (e) issue of tensor boxes {32 k, 32 rows} = 4 KB
  mbar_expect_tx(b2, (uint32_t)(reps * 4096));
  t0 = clock64();
  for (int r = 0; r < reps; ++r) tma_load_2d(dst + (r & 7) * 4096, &map32, b2, 0, (r & 63) * 32);
  t1 = clock64(); out[5] = (t1 - t0) / reps;
  mbar_wait(b2, ph); ph ^= 1;
  t2 = clock64(); out[6] = (t2 - t0) / reps;
  // (f) round trip of ONE 16 KB tile: bulk copy
  uint32_t p3 = 0;
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    mbar_expect_tx(b3, 16384);
    bulk_g2s(dst, src + (size_t)(r & 15) * 4096, 16384, b3);
    mbar_wait(b3, p3); p3 ^= 1;
  }
  t1 = clock64(); out[7] = (t1 - t0) / reps;
  // (g) round trip of ONE 16 KB tile: tensor box {32 k, 128 rows}
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    mbar_expect_tx(b3, 16384);
    tma_load_2d(dst, &map128, b3, 0, (r & 15) * 128);
    mbar_wait(b3, p3); p3 ^= 1;
  }
  t1 = clock64(); out[8] = (t1 - t0) / reps;
  // (h) the whole producer step as conv_tc.cu runs it: wait(empty, complete) + expect_tx + two 16 KB tensor boxes, 8 slots deep
  mbar_expect_tx(b0, 0);                                             // bar0 phase 1 complete -> parity 1 returns at once
  uint32_t p2 = ph;
  t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    mbar_wait(b0, 1);
    mbar_expect_tx(b2, 32768);
    tma_load_2d(dst + (r & 3) * 32768, &map128, b2, 0, (r & 15) * 128);
    tma_load_2d(dst + (r & 3) * 32768 + 16384, &map128, b2, 0, ((r + 5) & 15) * 128);
    if ((r & 3) == 3) { /* drain every 4 steps so that shared memory is not overwritten while in flight */ }
    mbar_wait(b2, p2); p2 ^= 1;                                      // (completion awaited: upper bound, includes the round trip)
  }
  t1 = clock64(); out[9] = (t1 - t0) / reps;
  // (i) the same, issue only (completion awaited every 4th step through one accumulating barrier)
  t0 = clock64();
  for (int r = 0; r < reps; r += 4) {
    mbar_expect_tx(b2, 4 * 32768);
    for (int q = 0; q < 4; ++q) {
      mbar_wait(b0, 1);
      tma_load_2d(dst + q * 32768, &map128, b2, 0, ((r + q) & 15) * 128);
      tma_load_2d(dst + q * 32768 + 16384, &map128, b2, 0, ((r + q + 5) & 15) * 128);
    }
    mbar_wait(b2, p2); p2 ^= 1;
  }
  t1 = clock64(); out[10] = (t1 - t0) / reps;
}

int main() {
  const int rows = 4096;                       // [rows][32] floats = 512 KB, L2 resident
  float* d; cudaMalloc(&d, (size_t)rows * 32 * 4); cudaMemset(d, 0, (size_t)rows * 32 * 4);
  long long* out; cudaMalloc(&out, 16 * 8); cudaMemset(out, 0, 16 * 8);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  CUtensorMap m32, m128;
  const cuuint64_t dims[2] = {32, (cuuint64_t)rows}; const cuuint64_t str[1] = {128}; const cuuint32_t es[2] = {1, 1};
  const cuuint32_t box32[2] = {32, 32}, box128[2] = {32, 128};
  enc(&m32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box32, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&m128, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int pass = 0; pass < 2; ++pass) {
    probe<<<1, 32, 160 * 1024>>>(m32, m128, d, out, 128);   // 128 x 4 KB stays below the 2^20 - 1 tx-count limit of an mbarrier
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  long long h[16]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"mbarrier.try_wait on a completed phase", "mbarrier.arrive.expect_tx", "fence.proxy.async.shared::cta",
                         "cp.async.bulk 4 KB: issue", "cp.async.bulk 4 KB: issue + drain (throughput)",
                         "tensor box {32,32} 4 KB: issue", "tensor box {32,32} 4 KB: issue + drain (throughput)",
                         "round trip 16 KB bulk copy (expect + copy + wait)", "round trip 16 KB tensor box {32,128}",
                         "producer step: wait + expect + 2 x 16 KB boxes + completion", "producer step, issue only (completion every 4th)"};
  printf("cycles per operation, one thread, L2-resident source (SM clock by clock64)\n");
  for (int i = 0; i < 11; ++i) printf("%-62s %6lld\n", names[i], h[i]);
  return 0;
}
