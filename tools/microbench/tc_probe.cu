// Microbenchmarks behind DESIGN.md's conv-kernel decisions (B200, sm_100a):
//   (1) round-trip latency of ONE TMA box load (L2-resident source) vs box rows
//   (2) cost of back-to-back tcgen05.mma kind::tf32 (M=128, K=8) vs N, operands in shared memory
//   (3) the same with the pixel operand addressed as a shifted window (unaligned start, SBO 2048)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../upflow_pytorch_b200/csrc -o tc_probe tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_common.cuh"

using namespace upf;
namespace upf {
void set_error(const char*, ...) {}
void count_launch(int) {}
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(128) tma_rtt_kernel(const __grid_constant__ CUtensorMap map, long long* out, int bytes, int reps, int inflight) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[16];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < inflight; ++i) {
        mbar_expect_tx(smem_u32(&bar[i]), bytes);
        tma_load_4d(smem_u32(base + (size_t)i * bytes), &map, smem_u32(&bar[i]), 0, 0, (r * inflight + i) % 7, blockIdx.x % 2);
      }
      for (int i = 0; i < inflight; ++i) mbar_wait(smem_u32(&bar[i]), r & 1);
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (t1 - t0) / reps;
  }
}

__global__ void __launch_bounds__(128) mma_cost_kernel(long long* out, int N, int reps, int shifted) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  // zero the operands (avoid NaN slow paths)
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_addr = smem_u32(base);
    const uint32_t b_addr = smem_u32(base + 16384) + (shifted ? (1 * 16 + 1) * 128 : 0);
    const uint64_t da = umma_desc_sw128(a_addr);
    const uint64_t db = shifted ? umma_desc_sw128_ex(b_addr, 2048, 0) : umma_desc_sw128(b_addr);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int k = 0; k < 4; ++k) umma_tf32(tmem, da + k * 2, db + k * 2, idesc, 1u);
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}


// (4) the conv kernels' issue pattern: per "tap" a group of 4*UNITS MMAs (4 K-steps x UNITS accumulators) followed by
//     tcgen05.commit(s), issued by one elected lane of a full warp exactly as conv_win.cu / conv_halo.cu do (warp-uniform
//     control flow, elect.sync).  ISSUERS = 1 or 2 warps; with 2, warp w owns accumulators w, w+2, ...
//     Reports cycles per group (thread clock around the whole loop incl. the final drain).
template <int UNITS, int ISSUERS>
__global__ void __launch_bounds__(128) mma_group_kernel(long long* out, int N, int groups, int commit_every, int ncommit, int poll) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar[4], bar2, done[2];
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<float*>(base)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1);
    mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&done[0]), 1); mbar_init(smem_u32(&done[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp < ISSUERS) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_base = smem_u32(base);              // 96 KB of "halo"
    const uint32_t b_base = smem_u32(base + 96 * 1024);  // 4 x 16 KB of "weights"
    long long t0 = clock64();
    for (int g = 0; g < groups; ++g) {
      if (poll) { mbar_wait(smem_u32(&bar2), 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }   // completes at once
      if (elect_one()) {
        const int tap = g % 9, ky = tap / 3, kx = tap % 3;
        const uint64_t db = umma_desc_sw128(b_base + (uint32_t)((N > 128 ? (g & 1) : (g & 3)) * 16384));
#pragma unroll
        for (int u = warp; u < UNITS; u += ISSUERS) {
          const uint64_t da = umma_desc_sw128(a_base + (uint32_t)(((4 * u + ky) * 32 + kx) * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_tf32(tmem + u * N, da + k * 2, db + k * 2, idesc, 1u);
        }
        if (commit_every && (g % commit_every) == commit_every - 1)
          for (int c = 0; c < ncommit; ++c) umma_commit(smem_u32(&bar[c]));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(smem_u32(&done[warp]));
    __syncwarp();
    mbar_wait(smem_u32(&done[warp]), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = (t1 - t0);
  }
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int UNITS, int ISSUERS>
static void run_groups(long long* out, int grid, int n) {
  cudaFuncSetAttribute(mma_group_kernel<UNITS, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170 * 1024);
  const int G = 180;
  struct { const char* name; int commit_every, ncommit, poll; } V[] = {
      {"no commit", 0, 0, 0}, {"1 commit/group", 1, 1, 0}, {"1 commit/group + poll/fence", 1, 1, 1}, {"2 commits/group + poll/fence", 1, 2, 1},
      {"1 commit/3 groups + poll/fence", 3, 1, 1}, {"3 commits/3 groups + poll/fence", 3, 3, 1}};
  for (auto& v : V) {
    mma_group_kernel<UNITS, ISSUERS><<<grid, 128, 170 * 1024>>>(out, n, G, v.commit_every, v.ncommit, v.poll);
    cudaError_t e = cudaDeviceSynchronize();
    printf("grid %3d N=%3d units %d issuers %d %-34s: %7.1f cycles / group = %6.1f / MMA (%s)\n", grid, n, UNITS, ISSUERS, v.name,
           out[0] / (double)G, out[0] / (double)G / (4 * UNITS), cudaGetErrorString(e));
  }
}

__global__ void __launch_bounds__(128) store_bw_kernel(float* out, long long* res, int steps, int pitch, int tiles_per_sm) {
  // 4 warps, each writes one 512-byte pixel row (128 channels) per step, like the conv epilogue
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int t = 0; t < tiles_per_sm; ++t) {
    float* base = out + ((size_t)(blockIdx.x * tiles_per_sm + t) * 256) * pitch;
    for (int i = 0; i < steps; ++i) {
      const int px = warp + 4 * i;
      *reinterpret_cast<float4*>(base + (size_t)px * pitch + lane * 4) = make_float4(1.f, 2.f, 3.f, (float)i);
    }
  }
  long long t1 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) res[0] = t1 - t0;
}

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  const int ld = 576, W = 311, H = 94, N = 2;
  float* x;
  cudaMalloc(&x, (size_t)N * H * W * ld * 4);
  cudaMemset(x, 0, (size_t)N * H * W * ld * 4);
  long long* out;
  cudaMallocManaged(&out, 64);
  cudaFuncSetAttribute(tma_rtt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(mma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);

  if (getenv("TC_GROUPS_ONLY")) {
    printf("== conv-like issue pattern (elected lane of a full warp): cycles per group of 4*units MMAs (M=128, K=8, tf32)\n");
    for (int grid : {148})
      for (int n : {32, 128, 256}) {
        run_groups<1, 1>(out, grid, n);
        if (n <= 256) run_groups<2, 1>(out, grid, n);
        if (n <= 128) run_groups<4, 1>(out, grid, n);
        if (n <= 256) run_groups<2, 2>(out, grid, n);
        if (n <= 128) run_groups<4, 2>(out, grid, n);
      }
    return 0;
  }
  printf("== TMA box round trip (cycles @ SM clock), source L2-resident, rows of 128 B at pitch %d B\n", ld * 4);
  for (int grid : {1, 148}) {
    for (int rows : {16, 32, 64, 128, 256, 544}) {
      for (int inflight : {1, 4}) {
        if ((size_t)rows * 128 * inflight > 190 * 1024) continue;
        int bw = rows >= 16 ? 16 : rows, bh = rows / bw;
        CUtensorMap m;
        cuuint64_t dims[4] = {(cuuint64_t)ld, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {(cuuint64_t)ld * 4, (cuuint64_t)W * ld * 4, (cuuint64_t)H * W * ld * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        tma_rtt_kernel<<<grid, 128, 195 * 1024>>>(m, out, rows * 128, 2, inflight);   // warm L2
        cudaDeviceSynchronize();
        tma_rtt_kernel<<<grid, 128, 195 * 1024>>>(m, out, rows * 128, 20, inflight);
        cudaError_t e = cudaDeviceSynchronize();
        printf("grid %3d  box %3d rows (%5.1f KB) x %d in flight: %6lld cycles per round  (%s)\n", grid, rows, rows * 128 / 1024.0, inflight,
               out[0], cudaGetErrorString(e));
      }
    }
  }
  {
    float* big;
    cudaMalloc(&big, (size_t)148 * 8 * 256 * 576 * 4);
    cudaFuncSetAttribute(store_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 206 * 1024);
    printf("== epilogue-like stores: 4 warps x 64 steps x 512 B per 256-pixel tile (cycles for thread 0 to ISSUE them)\n");
    for (int grid : {1, 148})
      for (int pitch : {128, 576})
        for (int smem : {0, 206 * 1024})
          for (int tiles : {1, 4}) {
            store_bw_kernel<<<grid, 128, smem>>>(big, out, 64, pitch, tiles);
            cudaDeviceSynchronize();
            store_bw_kernel<<<grid, 128, smem>>>(big, out, 64, pitch, tiles);
            cudaError_t e = cudaDeviceSynchronize();
            printf("grid %3d pitch %3d floats, dyn smem %3d KB, %d tile(s): %6lld cycles per tile (%s)\n", grid, pitch, smem / 1024, tiles, out[0] / tiles, cudaGetErrorString(e));
          }
  }
  printf("== tcgen05.mma kind::tf32 M=128 K=8, SS operands: cycles per instruction (400 back-to-back, 1 CTA and 148 CTAs)\n");
  for (int grid : {1, 148})
    for (int shifted : {0, 1})
      for (int n : {16, 32, 64, 128, 256}) {
        mma_cost_kernel<<<grid, 128, 100 * 1024>>>(out, n, 100, shifted);
        cudaError_t e = cudaDeviceSynchronize();
        printf("grid %3d  N=%3d %s: %7.1f cycles / MMA  (%s)\n", grid, n, shifted ? "shifted-window B" : "aligned B        ", out[0] / 400.0,
               cudaGetErrorString(e));
      }
  return 0;
}
