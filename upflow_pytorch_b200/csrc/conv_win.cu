// 3x3 stride-1 convolution, dilation 1/2/4, tensor cores (tcgen05, TF32): pixels as the M dimension, the
// activation halo loaded ONCE per 32-channel block, the nine taps read it as LINEARLY SHIFTED WINDOWS, one weight
// fetch shared by up to four accumulators, and TWO MMA-issuing warps.
//
// Measured facts that shape it (B200; tools/microbench/tc_probe.cu "conv-like issue pattern", tools/bench_win.py):
//  (1) a single issuing thread cannot keep the tensor pipe busy with this loop shape: per tap it issues 4*m MMAs, one
//      or two tcgen05.commit and one mbarrier poll, and the pipe idles for ~280 cycles around every commit
//      (M=128, N=128: 4 MMAs + commit + poll = 634 cycles against 259 of MMA work; N=256: 635 against 515).
//      conv_halo.cu hides most of it behind N=256 instructions; with N <= 128 it does not hide.  TWO issuing warps,
//      each with its own accumulators, overlap one warp's commit with the other's MMAs: 16 MMAs + commits in 1029
//      cycles = 64.3 per N=128 MMA, the full tensor rate.
//  (2) MMAs that accumulate into the same TMEM tile back to back cost >= 87 cycles each whatever N is; interleaved
//      with another accumulator they cost 65 (N=128) / 40 (N<=32).  So K steps are the OUTER loop, units the inner.
//  (3) with the pixels as M a layer with few output channels is cheap (40 cycles per 128 pixels and K=8 at N<=32,
//      against 64 in the transposed form of conv_halo.cu, which pays for M = 128 output channels whatever Cout is)
//      -- and 84% of the estimator's K volume has Cout < 128.
// Layout.  A CTA owns a tile of 4*m rows x (32 - 2d) columns (m = 1, 2 or 4 "units" of 4 rows).  Per channel block it
// loads one halo box {32 ch, 32 positions, 4m + 2d rows}: position (r, c) lies at byte (r*32 + c)*128 of the stage
// (SWIZZLE_128B).  Because a halo row is exactly 32 positions, the operand of unit u and tap (ky, kx) -- M index
// j = h*32 + w, h < 4 -- is the CONTIGUOUS range of 128 positions starting at (4u + ky*d)*32 + kx*d: a plain K-major
// SWIZZLE_128B descriptor (SBO = 1024 B) with a shifted start address (base_offset 0: measured, the swizzle is a
// function of absolute shared-memory address bits).  Columns w >= 32 - 2d of every row wrap into the next halo row:
// they are computed and never stored (6% / 12% / 25% of the MMA work at d = 1 / 2 / 4), which buys a halo of
// 1.33 loaded bytes per useful pixel-channel at m = 4, d = 1 (conv_halo.cu: 2.1) and ONE weight tile per tap feeding
// m accumulators.
// Work split between the two issuers (both deterministic: every accumulator has ONE issuer and a fixed order):
//   tap split  (2*m*BN <= 512 TMEM columns): issuer w takes the taps with (kb*9 + tap) % 2 == w into its own set of
//              m accumulators; the epilogue adds the two sets.  A weight slot is released by the one issuer that read it.
//   unit split (m = 4, BN > 64): issuer w takes units w, w+2 of every tap; a weight slot is released by both.
//   accumulators: [128 positions x BN channels] fp32 tiles in TMEM; TMEM lane = position, so the epilogue thread of
//   lane j owns one pixel and stores its channels as 16-byte vectors straight from registers.
//   warps: 0 = weight TMA, 1 and 7 = MMA issuers, 2-5 = epilogue, 6 = halo TMA; all eight in the epilogue when m >= 2.
#include "tc_common.cuh"

namespace upf {

constexpr int CW_THREADS = 256;
constexpr int CW_POS = 32;                         // halo positions per row
constexpr int CW_ROW_BYTES = CW_POS * 128;         // 4096

struct WinParams {
  float* out; int ldo;
  const float* res; int ldr;
  const float* bias;
  int H, W, Cout, BN;
  int m, dil, kblocks;
  int tiles_x, tiles_y;
  int na, nb;
  int a_bytes, b_stage_bytes;
  int tmem_cols;
  int tap_split;              // 1: issuers alternate taps (two accumulator sets), 0: issuers alternate units
  int epi_helpers;            // 1: warps 0, 1, 6, 7 take the odd units of the epilogue
  int mc;                     // 1: clusters of two CTAs share every weight tile (each fetches half of its rows, TMA multicast)
  int kxn;                    // 1: the three horizontal taps along N (one MMA per ky and K step, N = 3*BN; header note (4))
  long long* probe;
  float slope;
  int flags;
};

// half a weight slab, delivered to the same shared-memory offset (and counted on the same mbarrier offset) of BOTH CTAs of a pair
__device__ __forceinline__ void win_tma_load_3d_mc2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  const uint16_t mask = 3;
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in BOTH CTAs of the pair (a shared weight slot is free when both have read it)
__device__ __forceinline__ void win_umma_commit_mc2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

template <bool PROBE>
__global__ void __launch_bounds__(CW_THREADS)
conv_win_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const WinParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps shared-space addressing
  uint8_t* a_ring = base;
  uint8_t* b_ring = base + (size_t)p.na * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.nb * p.b_stage_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + p.na;
  uint64_t* fullB = emptyA + p.na;
  uint64_t* emptyB = fullB + p.nb;
  uint64_t* accum_full = emptyB + p.nb;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + ((2 * p.na + 2 * p.nb + 2) & ~1));   // 16-byte aligned
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);                        // 128 floats, 16-byte aligned

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int n = tile / p.tiles_y;
  const int d = p.dil;
  const int useful = CW_POS - 2 * d;
  const int x0 = tx * useful, y0 = ty * 4 * p.m;
  const int ntap = p.kxn == 2 ? 1 : p.kxn ? 3 : 9;                    // ring items per channel block: one / kernel rows / taps
  const int nslab = p.kxn == 2 ? 9 : p.kxn ? 3 : 1;                   // [BN][32] weight slabs (taps) per ring item
  const int NB = p.kxn ? 3 * p.BN : p.BN;                             // accumulator width of one unit
  const uint32_t b_bytes = (uint32_t)(nslab * p.BN) * 128u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    // a halo slot is released by both issuers -- except where a ring item is a whole channel block and the issuers alternate items
    for (int s = 0; s < p.na; ++s) { mbar_init(smem_u32(&fullA[s]), 1); mbar_init(smem_u32(&emptyA[s]), (p.kxn == 2 && p.tap_split) ? 1 : 2); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(smem_u32(&fullB[s]), 1); mbar_init(smem_u32(&emptyB[s]), (p.tap_split ? 1 : 2) * (p.mc ? 2 : 1)); }
    mbar_init(smem_u32(accum_full), 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2 && warp < 6) {                                  // bias is a constant weight: safe before the dependency wait
    const int c = (int)threadIdx.x - 64;
    s_bias[c] = c < p.Cout ? __ldg(p.bias + c) : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (p.mc) cluster_sync_all();        // the peer's barriers exist before anything is multicast into this CTA
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int crank = p.mc ? (int)cluster_ctarank() : 0;
  // programmatic dependent launch: everything above touched only this CTA's state and constant weights
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer: weight tiles (one per tap and channel block) =====================
    if (elect_one()) {
      long long w_eb = 0, t_start = PROBE ? clock64() : 0;
      const int total = p.kblocks * ntap;
      for (int ib = 0; ib < total; ++ib) {
        const int kb = ib / ntap, tap = ib - kb * ntap;
        const int sb = ib % p.nb;
        const long long t0 = PROBE ? clock64() : 0;
        mbar_wait(smem_u32(&emptyB[sb]), (((uint32_t)(ib / p.nb)) & 1u) ^ 1u);
        if (PROBE) w_eb += clock64() - t0;
        const uint32_t fb = smem_u32(&fullB[sb]);
        mbar_expect_tx(fb, b_bytes);
        const uint32_t b_dst = smem_u32(b_ring + (size_t)sb * p.b_stage_bytes);
        if (p.mc) {
          // this CTA fetches rows [rank BN/2, +BN/2) of every tap slab for both CTAs; the peer's halves arrive on the same barrier
          const int hr = p.BN >> 1;
          for (int t = 0; t < nslab; ++t)
            win_tma_load_3d_mc2(b_dst + (uint32_t)((t * p.BN + crank * hr) * 128), &map_w, fb, kb * 32, crank * hr, tap * nslab + t);
        } else {
          // kxn: the box holds the three taps of kernel row `tap`, rows (kx, co) -- 3*BN operand rows of 128 bytes (kxn = 2: all nine)
          tma_load_3d(b_dst, &map_w, fb, kb * 32, 0, tap * nslab);
        }
      }
      if (PROBE && p.probe && blockIdx.x == 0) { p.probe[1] = w_eb; p.probe[2] = clock64() - t_start; }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ===================== TMA producer: halo boxes (one per channel block) =====================
    if (elect_one()) {
      long long w_ea = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int sa = kb % p.na;
        const long long t0 = PROBE ? clock64() : 0;
        mbar_wait(smem_u32(&emptyA[sa]), (((uint32_t)(kb / p.na)) & 1u) ^ 1u);
        if (PROBE) w_ea += clock64() - t0;
        const uint32_t fa = smem_u32(&fullA[sa]);
        mbar_expect_tx(fa, (uint32_t)p.a_bytes);
        tma_load_4d(smem_u32(a_ring + (size_t)sa * p.a_bytes), &map_x, fa, kb * 32, x0 - d, y0 - d, n);
      }
      if (PROBE && p.probe && blockIdx.x == 0) p.probe[0] = w_ea;
    }
    __syncwarp();
  } else if (warp == 1 || warp == 7) {
    // ===================== MMA issuers: M = 128 positions, N = BN output channels =====================
    const int wi = warp == 1 ? 0 : 1;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);
    const int total = p.kblocks * ntap;
    const int step = p.tap_split ? 2 : 1;
    const int u0 = p.tap_split ? 0 : wi, ustep = p.tap_split ? 1 : 2;
    const uint32_t tset = tmem_base + (uint32_t)(p.tap_split ? wi * p.m * NB : 0);
    int kb_ready = -1;
    uint32_t acc = 0;
    long long w_fa = 0, w_fb = 0, t_start = PROBE ? clock64() : 0;
    for (int ib = p.tap_split ? wi : 0; ib < total; ib += step) {
      const int kb = ib / ntap, tap = ib - kb * ntap;
      const int sa = kb % p.na, sb = ib % p.nb;
      if (kb != kb_ready) {
        const long long t0 = PROBE ? clock64() : 0;
        mbar_wait(smem_u32(&fullA[sa]), ((uint32_t)(kb / p.na)) & 1u);
        if (PROBE) w_fa += clock64() - t0;
        kb_ready = kb;
      }
      {
        const long long t0 = PROBE ? clock64() : 0;
        mbar_wait(smem_u32(&fullB[sb]), ((uint32_t)(ib / p.nb)) & 1u);
        if (PROBE) w_fb += clock64() - t0;
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const int ky0 = p.kxn ? tap : tap / 3, kx = p.kxn ? 0 : tap - ky0 * 3;
        const int nky = p.kxn == 2 ? 3 : 1;                  // kxn = 2: the item is the whole channel block, the three kernel rows in turn
        for (int kyi = 0; kyi < nky; ++kyi) {
          const int ky = ky0 + kyi;
          const uint64_t dw = umma_desc_sw128(smem_u32(b_ring + (size_t)sb * p.b_stage_bytes) + (uint32_t)(kyi * 3 * p.BN * 128));
          const uint64_t dx = umma_desc_sw128(smem_u32(a_ring + (size_t)sa * p.a_bytes) + (uint32_t)(((ky * d) * CW_POS + kx * d) * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            for (int u = u0; u < p.m; u += ustep)      // unit u: 4 halo rows = 16 KB further (1024 descriptor units)
              umma_tf32(tset + (uint32_t)(u * NB), dx + (uint64_t)(u * 1024 + k * 2), dw + (uint64_t)(k * 2), idesc, acc | (uint32_t)(k | kyi));
          }
        }
        if (p.mc) win_umma_commit_mc2(smem_u32(&emptyB[sb]));
        else umma_commit(smem_u32(&emptyB[sb]));
        if ((ib + step) / ntap != kb) umma_commit(smem_u32(&emptyA[sa]));  // this issuer's last tap of the channel block
        if (ib + step >= total) umma_commit(smem_u32(accum_full));
      }
      __syncwarp();
      acc = 1;
    }
    if (PROBE && p.probe && blockIdx.x == 0 && lane == 0 && wi == 0) { p.probe[3] = w_fa; p.probe[4] = w_fb; p.probe[5] = clock64() - t_start; }
  }
  // ===================== epilogue: lane = position, one pixel per thread and unit =====================
  // warps 2..5 from the start; the producer and issuer warps (0, 1, 6, 7 -- one per TMEM lane quarter as well) join when
  // their loops are done and take every other unit: the phase is a chain of TMEM-load / store latencies per unit, two
  // groups of four warps run two such chains at once
  const bool primary = warp >= 2 && warp < 6;
  const bool helpers = p.epi_helpers && p.m >= 2;
  if (primary || helpers) {
    const int q = warp & 3;                                    // TMEM lane quarter = row h of the unit
    const int u_first = (helpers && !primary) ? 1 : 0, u_step = helpers ? 2 : 1;
    const bool vec_out = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const size_t img = (size_t)n * p.H * p.W;
    const int x = x0 + lane;
    const bool col_ok = lane < useful && x < p.W;
    const uint32_t set2 = (uint32_t)(p.m * NB);
    mbar_wait(smem_u32(accum_full), 0);
    const long long t_e1 = PROBE ? clock64() : 0;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int u = u_first; u < p.m; u += u_step) {
      const int y = y0 + 4 * u + q;
      const bool ok = col_ok && y < p.H;
      const size_t pix = img + (size_t)y * p.W + x;
      float* orow = p.out + pix * p.ldo;
      const float* rrow = p.res ? p.res + pix * p.ldr : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * NB);
      for (int c0 = 0; c0 < p.Cout; c0 += 32) {
        uint32_t v[32];
        const int nv = (c0 + 16 < p.BN) ? 32 : 16;                               // warp-uniform
        if (p.kxn) {
          // column group kx of lane j holds the products of tap (., kx) with the input at position j: the output at
          // position j needs them from position j + kx*d -- lane + kx*d of the same warp (a unit row is one warp's 32
          // lanes, and the lanes whose neighbour would wrap, w >= 32 - 2d, are the ones never stored)
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            uint32_t t[32];
            const uint32_t ta = taddr + (uint32_t)(g * p.BN + c0);
            tmem_ld16(ta, t);
            if (nv == 32) tmem_ld16(ta + 16, t + 16);
            if (p.tap_split) {
              uint32_t t2[32];
              tmem_ld16(ta + set2, t2);
              if (nv == 32) tmem_ld16(ta + set2 + 16, t2 + 16);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) t[j] = __float_as_uint(__uint_as_float(t[j]) + __uint_as_float(t2[j]));
            } else {
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nv) {
                if (g == 0) v[j] = t[j];
                else v[j] = __float_as_uint(__uint_as_float(v[j]) + __shfl_down_sync(0xffffffffu, __uint_as_float(t[j]), g * d));
              }
            }
          }
        } else {
        tmem_ld16(taddr + (uint32_t)c0, v);
        if (c0 + 16 < p.BN) tmem_ld16(taddr + (uint32_t)(c0 + 16), v + 16);      // warp-uniform
        if (p.tap_split) {                                                       // second accumulator set (odd taps)
          uint32_t v2[32];
          tmem_ld16(taddr + set2 + (uint32_t)c0, v2);
          if (c0 + 16 < p.BN) tmem_ld16(taddr + set2 + (uint32_t)(c0 + 16), v2 + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        }
        if (!ok) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int c = c0 + j;
          if (c >= p.Cout) break;
          const float4 bv = *reinterpret_cast<const float4*>(s_bias + c);
          float f[4] = {lrelu(__uint_as_float(v[j]) + bv.x, p.slope), lrelu(__uint_as_float(v[j + 1]) + bv.y, p.slope),
                        lrelu(__uint_as_float(v[j + 2]) + bv.z, p.slope), lrelu(__uint_as_float(v[j + 3]) + bv.w, p.slope)};
          if (rrow) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c + k < p.Cout) f[k] += __ldg(rrow + c + k);
          }
          if (p.flags & UPF_FLAG_ROUND_TF32) {
#pragma unroll
            for (int k = 0; k < 4; ++k) f[k] = round_tf32(f[k]);
          }
          if (vec_out && c + 4 <= p.Cout) {
            *reinterpret_cast<float4*>(orow + c) = make_float4(f[0], f[1], f[2], f[3]);
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c + k < p.Cout) orow[c + k] = f[k];
          }
        }
      }
    }
    if (PROBE && p.probe && blockIdx.x == 0 && threadIdx.x == 64) { p.probe[6] = 0; p.probe[7] = clock64() - t_e1; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (p.mc) cluster_sync_all();        // nobody leaves while the peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

extern int g_tc_pdl;
extern long long* g_halo_probe;
static int g_win_enabled = 1;
static int g_win_min_cin = 0;
static int g_win_max_cout = 64;   // measured (tools/bench_win.py, 1/4- and 1/8-res KITTI): faster than conv_halo.cu for Cout <= 64
                                  // (544->32: 97 vs 139 us, 480->64: 108 vs 130, 576->2: 91 vs 131), slower for 96..128
static int g_win_force_m = 0;
static int g_win_kxn = 1;         // the three horizontal taps along N (header note (4)); upf_debug_conv_win bit 2 switches it off
static int g_win_kxn2_min_kb16 = 3;         // BN = 16: channel-block items (one CTA per SM) from this many channel blocks on, two CTAs with kernel-row items
                                            // below (A/B: upf_debug_conv_win min_cin >= 1000 sets it; KITTI forward 2.366 -> 2.348 ms)
// ... with ONE ring item per channel block (all nine weight slabs in a stage).  OFF by default: alone it is faster and parity-green
// (KITTI forward 2.406 -> 2.348 ms), but with four forwards replaying concurrently (pipeline.PipelinedInference lanes, bench.py)
// the run hung or died with a launch failure in 5 of 7 bench runs with it on and in none of 6 with it off
// (profiles/r2_triage_lanes.txt); cause not found.  upf_debug_conv_win bit 4 (16) switches it ON (A/B, tests).
static int g_win_kxn2 = 0;

// returns with *taken = 1 when the launch was made, 0 when the shape is not eligible (caller falls through)
int conv2d_fwd_win(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                   const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                   float slope, int flags, cudaStream_t st, int* taken) {
  *taken = 0;
  if (!g_win_enabled || ks != 3 || stride != 1 || !(dil == 1 || dil == 2 || dil == 4) || Cout > 128) return 0;
  if (Cin < g_win_min_cin || Cout > g_win_max_cout) return 0;
  const int BN = (Cout + 15) & ~15;
  const int kblocks = (Cin + 31) / 32;
  const int cin_pad = kblocks * 32;
  const int useful = CW_POS - 2 * dil;
  const int tiles_x = (W + useful - 1) / useful;
  // (4) kxn: one MMA per kernel ROW and K step, N = 3*BN -- column group kx of the accumulator holds the products of tap
  // (ky, kx) with the UNSHIFTED window, and the epilogue adds group kx of lane + kx*d (a warp shuffle: a unit row is one
  // warp).  A third of the MMAs, each above the ~40-cycle floor that N <= 32 instructions pay (note (3)), and a third of
  // the window descriptors / commits / polls per channel block.  The same packed weights serve both forms: the three
  // taps of a kernel row are consecutive [BN][cin_pad] slabs, one {32, BN, 3} box.
  // Measured (tools/bench_win.py, 2x94x311): 544->32 85 -> 64 us, 480->64 104 -> 91, 160->16 35 -> 23, 184->3 37 -> 22; at
  // 1/8 resolution 544->32 49 -> 27, 480->64 47 -> 29.  With one or two channel blocks at BN >= 32 the longer epilogue (three
  // column groups to read and shift) outweighs the shorter K loop (64->32 22.5 -> 24.5, 32->32 at 1/2 res 47 -> 58).
  const bool kxn = g_win_kxn && BN <= 64 && (BN <= 16 || kblocks >= 3);
  const int NB = kxn ? 3 * BN : BN;                          // accumulator columns per unit
  // kxn = 2: a ring item is a whole CHANNEL BLOCK -- the nine weight slabs in one stage, the issuer walks the three kernel rows
  // inside the item: a third of the commits / polls again (a commit idles the tensor pipe for ~280 cycles, and with the unit
  // split both issuers reach it together), where two halo stages and two 9-slab weight stages fit.  tools/bench_win.py at
  // 2x94x311: 544->32 59.3 -> 51.1 us, 480->64 86.0 -> 71.6; at 2x47x156 544->32 26.5 -> 22.4; KITTI forward 2.406 -> 2.369 ms,
  // Sintel b8 12.13 -> 11.90 ms (profiles/r2_ab_kxn2.txt)
  bool kxn2 = kxn && g_win_kxn2 && (BN >= 32 || (kblocks >= g_win_kxn2_min_kb16)) && kblocks >= 3 && 2 * (8 + 2 * dil) * CW_ROW_BYTES + 2 * 9 * BN * 128 <= 224 * 1024;
  const int b_stage_bytes = (kxn2 ? 9 * BN : NB) * 128;      // multiple of 2048
  // two resident CTAs per SM (~108 KB rings each): one CTA's epilogue -- 4-8 us of stores per tile with the tensor pipe idle --
  // runs under the other's MMAs.  Measured on whole forwards (tools/ab_win2.py): KITTI b1 2.813 -> 2.790 ms, Sintel b8
  // 15.10 -> 14.56 ms, HD b2 16.39 -> 15.94 ms.  force_m bit 4 (16): one CTA with 224 KB rings (A/B switch)
  // Only where a CTA still gets >= 4 weight slots: at BN = 64 it gets two, the MMA threads then wait for weights 30 % of the
  // time and the layer is no faster than with one CTA (480->64 at 94x311: 101.4 vs 99.3 us; 544->32: 81.9 vs 93.2 us).
  const bool two_cta = (g_win_force_m & 16) == 0 && (108 * 1024 - 2 * (8 + 2 * dil) * CW_ROW_BYTES) / b_stage_bytes >= 4;
  const int budget = two_cta ? 108 * 1024 : 224 * 1024;
  // units per CTA and issuer split: minimise waves x (time of one ring item), modelled from the microbenchmark --
  // one issuer needs ~(490 + 145 m) cycles per item (4m MMAs + commit + poll), two run in parallel, and the tensor
  // pipe needs 4 m T(N) cycles per item, T = 40 / 49 / 57 / 65 cycles at N <= 32 / 64 / 96 / 128, ~N/2 above.
  // Ties go to the larger tile (fewer weight fetches per pixel).  force_m: +8 forces the unit split.
  const double T = NB <= 32 ? 40. : NB <= 64 ? 49. : NB <= 96 ? 57. : NB <= 128 ? 65. : NB * 0.5;
  int m = 0, tap_split = 1;
  double best = 1e30;
  for (int cand = 4; cand >= 1; cand >>= 1) {
    if ((g_win_force_m & 7) && cand != (g_win_force_m & 7)) continue;
    int split = (2 * cand * NB <= 512) ? 1 : 0;
    if (g_win_force_m & 8) split = 0;
    if (!split && (cand * NB > 512 || cand < 2)) continue;
    const int a_bytes = (4 * cand + 2 * dil) * CW_ROW_BYTES;
    if (2 * a_bytes + ((two_cta || kxn2) ? 2 : 3) * b_stage_bytes > budget) continue;
    if (two_cta && (split ? 2 : 1) * cand * NB > 256) continue;          // both CTAs' accumulators must fit the 512 TMEM columns
    const long long t = (long long)tiles_x * ((H + 4 * cand - 1) / (4 * cand)) * N;
    const double issuer = split ? (490. + 145. * cand) / 2 : (490. + 145. * cand / 2);
    const double tap_time = issuer > 4. * cand * T ? issuer : 4. * cand * T;
    const int slots = two_cta ? 2 * UPF_NUM_SMS : UPF_NUM_SMS;
    const double cost = (double)((t + slots - 1) / slots) * tap_time * (two_cta ? 2 : 1);
    if (cost < best * 0.97) { best = cost; m = cand; tap_split = split; }
  }
  if (m == 0) return 0;
  const int tiles_y = (H + 4 * m - 1) / (4 * m);
  const long long tiles = (long long)tiles_x * tiles_y * N;
  if (tiles * m < 96) return 0;                              // coarse levels: the cluster split-K kernel is the better fit
  const int rows = 4 * m + 2 * dil;

  // pairs of CTAs sharing every weight tile by TMA multicast (as conv_halo.cu does): built, parity-green, and measured SLOWER here
  // (544->32 at 2x94x311 61.4 vs 59.4 us, KITTI forward 2.403 vs 2.386 ms, Sintel b8 12.10 vs 12.00 ms; profiles/r2_ab_mcw.txt):
  // these layers are not bound by weight traffic out of L2.  ncu on 544->32 (profiles/r2_ncu_full_conv_win_544to32_kxn.md): L2
  // 20 % of its peak throughput, tensor-core shared-memory reads 40 %, tensor pipe 43 % active -- no unit is saturated; what is
  // left is the issue side (commit / poll bubbles per ring item), the prologue and the epilogue, and the lock-step of a pair
  // costs more than the halved fetches save.  Off by default; force_m bit 6 (64) switches it on (A/B, tests).
  const bool mc = (g_win_force_m & 64) != 0 && !two_cta && (tiles % 2 == 0) && BN >= 32 && kblocks >= 3;
  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    const cuuint32_t box[4] = {32, CW_POS, (cuuint32_t)rows, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    MapKey key{x, ldx, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)Cin, 200000 + rows, 4};
    int e = encode_cached(key, &mx, 4, const_cast<float*>(x), dims, strides, box, estr);
    if (e) return e;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)BN, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 4, (cuuint64_t)cin_pad * BN * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)(mc ? BN / 2 : BN), mc ? 1u : kxn2 ? 9u : kxn ? 3u : 1u};
    const cuuint32_t estr[3] = {1, 1, 1};
    MapKey key{w_packed, cin_pad, BN, 9, mc ? 6000 + BN : kxn2 ? 9000 + BN : kxn ? 3000 + BN : BN, 3};
    int e = encode_cached(key, &mw, 3, const_cast<float*>(w_packed), dims, strides, box, estr);
    if (e) return e;
  }
  WinParams p;
  p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr; p.bias = bias;
  p.H = H; p.W = W; p.Cout = Cout; p.BN = BN; p.m = m; p.dil = dil; p.kblocks = kblocks;
  p.tiles_x = tiles_x; p.tiles_y = tiles_y;
  p.a_bytes = rows * CW_ROW_BYTES;
  p.b_stage_bytes = b_stage_bytes;
  p.slope = slope;
  p.flags = flags;
  p.probe = g_halo_probe;
  int cols = 32;
  while (cols < (tap_split ? 2 : 1) * m * NB) cols <<= 1;
  p.tmem_cols = cols;
  p.tap_split = tap_split;
  p.kxn = kxn2 ? 2 : kxn ? 1 : 0;
  p.mc = mc ? 1 : 0;
  p.epi_helpers = (g_win_force_m & 32) ? 0 : 1;
  int na = (kblocks >= 3 && 3 * p.a_bytes + 4 * b_stage_bytes <= budget) ? 3 : 2;
  if (na > kblocks) na = kblocks;
  int nb = (budget - na * p.a_bytes) / b_stage_bytes;
  if (nb > 8) nb = 8;
  if (tap_split) nb &= ~1;   // EVEN: each issuer then owns the weight slots of its parity and sees every phase of their barriers
                             // (an odd ring would bring an issuer back to a slot two phases later -- parity waits cannot tell)
  if (nb < 2) return 0;
  p.na = na; p.nb = nb;
  // the windows of the last unit read up to 2*dil positions past their stage: the weight ring follows the halo ring
  const size_t smem = (size_t)na * p.a_bytes + (size_t)nb * b_stage_bytes + (2 * na + 2 * nb + 2) * 8 + 16 + 128 * 4 + 1024;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_win_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_win_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_win smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set.mark();
  }
  if (smem > 227 * 1024) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(CW_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na_ = 0;
  if (g_tc_pdl) {
    attr[na_].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na_].val.programmaticStreamSerializationAllowed = 1;
    ++na_;
  }
  if (mc) {
    attr[na_].id = cudaLaunchAttributeClusterDimension;
    attr[na_].val.clusterDim.x = 2; attr[na_].val.clusterDim.y = 1; attr[na_].val.clusterDim.z = 1;
    ++na_;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na_;
  {
    cudaError_t e = p.probe ? cudaLaunchKernelEx(&cfg, conv_win_kernel<true>, mx, mw, p) : cudaLaunchKernelEx(&cfg, conv_win_kernel<false>, mx, mw, p);
    if (e != cudaSuccess) { set_error("conv_win launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  }
  *taken = 1;
  return check_launch("conv_win");
}

}  // namespace upf

// test / tuning hook (not part of the hot-path ABI): enable / disable the window kernel, set its minimum Cin and
// force the units per CTA (0 = automatic)
extern "C" int upf_debug_conv_win(int enabled, int min_cin, int force_m) {
  upf::g_win_enabled = enabled & 1;
  upf::g_win_max_cout = (enabled & 2) ? 128 : 64;      // bit 1: take every Cout <= 128 (A/B runs)
  upf::g_win_kxn2 = (enabled & 16) ? 1 : 0;            // bit 4: one ring item per channel block instead of per kernel row (see g_win_kxn2)
  upf::g_win_kxn = (enabled & 4) ? 0 : 1;              // bit 2: one MMA per TAP (N = BN) instead of per kernel row (N = 3 BN)
  if (min_cin >= 1000) { upf::g_win_kxn2_min_kb16 = min_cin - 1000; min_cin = 0; }
  if (min_cin >= 0) upf::g_win_min_cin = min_cin;
  upf::g_win_force_m = force_m;
  return 0;
}
