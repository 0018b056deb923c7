// Tensor-core weight gradient of a 3x3 convolution: the nine taps along N, output channels in parts of <= 32.
//
// conv_tc.cu's weight-gradient mode runs the nine tap GEMMs dW[tap] = XT(shifted by the tap) . GT^T as nine sets of CTAs:
// every tap streams its own 128-row tile of the planar input, 162 KB of operand tiles per 32 pixels of K for an M tile, and
// ncu puts the stall at the MMA issuers' wait for operands (160->16 at 8x64x208: 368 us, tensor pipe 4 % active, L2
// delivering 22 GB/s per SM).  Here the roles are exchanged: the input tile XT[k block j] is loaded ONCE and
// multiplied with the nine gradient tiles that meet it,
//     dW[ky,kx][ci][co] += XT[ci][j] . GT_kx[co][j - koff(ky)]^T ,
// (the vertical part of a tap's shift is a whole number of 32-pixel k blocks, the horizontal part lives in the three
// pre-shifted copies GT_kx, as in backward.cu) into NINE accumulators side by side in tensor memory (9 x BN <= 288
// columns): 16 KB + 9 x BN x 128 B per 32 pixels of K and M tile.  K is split over the GRID (not a cluster: the split can be
// as wide as the chip), partial sums go to a workspace and are added in split order by reduce_splits_kernel (deterministic).
// Wider layers run as N parts of 32 channels in separate CTAs (nine 32-column accumulators each): measured against the
// per-tap GEMMs at 8x64x208 -- 128->128 139 vs 192 us, 96->64 80 vs 187, 480->64 244 vs 378, 256->128 246 vs 377,
// 576->128 573 vs 580, 384->96 285 vs 271.  INSIDE the training step, though, serving the 96- and 128-wide layers here costs
// 2 ms (49.5 vs 46.7-48.1 ms per step, same process, UPF_WGRAD_TAPS = 128 / 64 / 32): the default limit is 64.
// Operands are the blocked, pre-swizzled planar tensors of backward.cu ([k block][row][32]: one bulk copy per tile).
// Roles: warp 0 = producer (one thread), warps 1, 6, 7 = MMA issuers (one thread each: a thread issues ~one tcgen05 instruction
// per 100 cycles, 36 MMAs per ring slot would take it 3700 cycles against 1440 of tensor work at N = 32; issuer i takes the
// three taps of kernel row i -- 12 MMAs per slot into its own three accumulators, so every accumulator keeps ONE issuer and a
// fixed order), warps 2..5 = epilogue.
#include "tc_common.cuh"

namespace upf {

constexpr int WT_THREADS = 256;
constexpr int WT_ISSUERS = 3;
constexpr int WT_A_BYTES = 128 * 128;

struct WtParams {
  const float* xg; int xrows;         // XT blocked: block j, row r at xg[(j * xrows + r) * 32]
  const float* gg; int grows;         // GT blocked (block 0 of copy 0; zero blocks in front / behind): gg[kx * gcopy + (j * grows + r) * 32]
  long long gcopy;
  int koff[3];                        // k-block offset of the vertical tap shift, (ky - 1) * dil * Wp / 32
  int kblocks, bps;                   // K blocks in all / per split
  int Cin, Cout, BN, a_bytes;
  int nstage, tmem_cols;
  float* part;                        // [splits][9][Cin][Cout]
};

__device__ __forceinline__ void wt_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(WT_THREADS)
wgrad_taps_kernel(const WtParams p) {
  extern __shared__ __align__(1024) uint8_t wt_smem_raw[];
  uint8_t* base = wt_smem_raw + ((1024u - (smem_u32(wt_smem_raw) & 1023u)) & 1023u);
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = WT_A_BYTES + 9u * b_bytes;              // BN is 16 or 32: a multiple of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)p.nstage * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.nstage;
  uint64_t* accum_full = bars + 2 * p.nstage;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;                                      // first input channel (row of XT) of this M tile
  const int split = blockIdx.y;
  const int co0 = blockIdx.z * p.BN;                                    // first output channel of this N part
  const int kb0 = split * p.bps;
  const int iters = (p.kblocks - kb0 < p.bps ? p.kblocks - kb0 : p.bps);   // >= 1 by construction

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), WT_ISSUERS);
    }
    mbar_init(smem_u32(accum_full), WT_ISSUERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer =====================
    if (elect_one()) {
      for (int it = 0; it < iters; ++it) {
        const int j = kb0 + it;
        const int s = it % p.nstage;
        const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
        const uint32_t a_dst = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t fb = smem_u32(&full[s]);
        mbar_expect_tx(fb, (uint32_t)p.a_bytes + 9u * b_bytes);
        wt_bulk_g2s(a_dst, p.xg + ((size_t)j * p.xrows + m0) * 32, (uint32_t)p.a_bytes, fb);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int ky = t / 3, kx = t - ky * 3;
          wt_bulk_g2s(a_dst + WT_A_BYTES + (uint32_t)t * b_bytes,
                      p.gg + (size_t)kx * p.gcopy + ((long long)(j - p.koff[ky]) * p.grows + co0) * 32, b_bytes, fb);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp >= 6) {
    // ===================== MMA issuers =====================
    const int ky = warp == 1 ? 0 : warp - 5;                  // kernel row whose three taps this warp accumulates
    // instruction descriptor: D=f32 (bit4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    for (int it = 0; it < iters; ++it) {
      const int s = it % p.nstage;
      const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
      mbar_wait(smem_u32(&full[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(base + (size_t)s * stage_bytes);
        const uint64_t da = umma_desc_sw128(a_addr);
#pragma unroll
        for (int k = 0; k < 4; ++k)        // K steps outside, the nine accumulators inside: independent MMAs back to back
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {
            const int t = ky * 3 + tx;
            const uint64_t db = umma_desc_sw128(a_addr + WT_A_BYTES + (uint32_t)t * b_bytes);
            umma_tf32(tmem_base + (uint32_t)(t * p.BN), da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        umma_commit(smem_u32(&empty[s]));
        if (it + 1 == iters) umma_commit(smem_u32(accum_full));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                    // TMEM lane quarter this warp may read
    const int ci = m0 + q * 32 + lane;
    mbar_wait(smem_u32(accum_full), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* dst = p.part + (size_t)split * 9 * p.Cin * p.Cout;
    for (int t = 0; t < 9; ++t)
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.BN + c0), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ci < p.Cin) {
          float* o = dst + ((size_t)t * p.Cin + ci) * p.Cout;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            if (co0 + c0 + jj < p.Cout) o[co0 + c0 + jj] = __uint_as_float(v[jj]);
        }
      }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int launch_reduce_splits(const float* part, float* out, long long n, int splits, cudaStream_t st);   // backward.cu

static int g_wgrad_taps = 64;         // largest padded Cout served here (upf_debug_wgrad_taps; 0 routes every shape to conv_tc's weight-gradient mode)

// xt: block 0 of the blocked input (rows [row0, row0 + Cin) of a buffer with xt_rows rows per block: pass xt + row0 * 32);
// gt: block 0 of copy 0 (the caller guarantees max|koff| zero blocks in front of and behind every copy).
// Returns 1 in *taken when this kernel handled the call.
int wgrad_taps_gemm(const float* xt, int xt_rows, const float* gt, long long gcopy, int cout_pad, float* part, float* gw,
                    int Cin, int Cout, int kblocks, const int* koff3, cudaStream_t st, int* taken) {
  *taken = 0;
  if (cout_pad > g_wgrad_taps || kblocks < 8) return 0;
  *taken = 1;
  WtParams p;
  p.xg = xt; p.xrows = xt_rows; p.gg = gt; p.grows = cout_pad; p.gcopy = gcopy;
  for (int i = 0; i < 3; ++i) p.koff[i] = koff3[i];
  p.kblocks = kblocks;
  p.Cin = Cin; p.Cout = Cout; p.BN = cout_pad <= 16 ? 16 : 32;
  const int nparts = (cout_pad + p.BN - 1) / p.BN;             // N parts of 32 channels: nine accumulators of 32 columns each per CTA
  int tw = 8;
  while (tw < 128 && tw < Cin) tw <<= 1;
  p.a_bytes = tw * 128;
  const int mtiles = (Cin + 127) / 128;
  int splits = UPF_NUM_SMS / (mtiles * nparts);
  if (splits > kblocks / 8) splits = kblocks / 8;              // at least 8 k blocks per split
  if (splits < 1) splits = 1;
  p.bps = (kblocks + splits - 1) / splits;
  splits = (kblocks + p.bps - 1) / p.bps;                      // no empty split
  const int stage_bytes = WT_A_BYTES + 9 * p.BN * 128;
  int nstage = (200 * 1024) / stage_bytes;
  if (nstage > 8) nstage = 8;
  p.nstage = nstage;
  p.tmem_cols = p.BN == 16 ? 256 : 512;                        // 9 x BN columns, power of two
  p.part = part;
  const size_t smem = (size_t)nstage * stage_bytes + (2 * nstage + 1) * 8 + 16 + 1024;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("wgrad_taps smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set.mark();
  }
  wgrad_taps_kernel<<<dim3((unsigned)mtiles, (unsigned)splits, (unsigned)nparts), WT_THREADS, smem, st>>>(p);
  int e = check_launch("wgrad_taps");
  if (e) return e;
  return launch_reduce_splits(part, gw, 9ll * Cin * Cout, splits, st);
}

long long wgrad_taps_part_elems(int Cin, int Cout) { return (long long)UPF_NUM_SMS * 9 * Cin * Cout; }

}  // namespace upf

extern "C" int upf_debug_wgrad_taps(int max_cout_pad) {
  upf::g_wgrad_taps = max_cout_pad < 0 ? 64 : max_cout_pad;
  return 0;
}
