// 3x3 stride-1 convolution, dilation 1/2/4, tensor cores (tcgen05, TF32), with the
// activation tile loaded ONCE per 32-channel block and shared by the nine taps,
// and with the GEMM TRANSPOSED so that the pixels are the wide N dimension.
//
// Two measured facts shape this kernel (B200, tools/bench_halo.py):
//  (1) conv_tc.cu issues one TMA box per (tap, channel block): every activation
//      byte travels L2 -> shared memory nine times;
//  (2) tcgen05.mma kind::tf32 M=128 K=8 costs 46 / 49 / 65 / 129 cycles at
//      N = 16-32 / 64 / 128 / 256 (tools/microbench/tc_probe.cu; the shifted-window
//      operand below costs nothing extra): one N=256 instruction per K step keeps the
//      issue rate low and the accumulator in one TMEM tile.  (For Cout <= 64 the
//      untransposed form -- pixels as M, 46-49 cycles per 128 pixels -- would be up to
//      1.4x cheaper in tensor time; not built yet.)
// Here  D^T[Cout (M=128, zero-padded), pixels (N = 128 or 256)] += Wt[Cout, K] * X[pixels, K]^T :
// the weights tile of a tap is the A operand (128 rows), the activation window is
// the B operand with N = 256 pixels per instruction.  A CTA owns an 8-column x
// (16*MT)-row pixel tile (MT = 1 or 2) and loads, per channel block, ONE halo box
// {32 ch, 16 cols, 16*MT+2d rows}: 128-byte rows, SWIZZLE_128B.  For tap (ky,kx)
// the B operand is a *shifted window of that box*:
//     pixel n = h*8 + w   ->  halo position (h + ky*d)*16 + (w + kx*d)
// i.e. 8-row core-matrix groups 2048 bytes (one halo row) apart, starting
// (ky*d*16 + kx*d) rows into the box -- expressible in the UMMA shared-memory
// descriptor (SBO = 2048 B, shifted start address, base_offset 0: measured, the
// 128-byte swizzle is a function of absolute shared-memory address bits, so TMA's
// write pattern and the MMA's read pattern agree for any 128-byte-aligned start).
// The halo width is fixed at 16 positions.  L2 -> SM activation traffic drops 3-6x.
//   accumulator: TMEM lane = output channel, column = pixel; the epilogue thread of
//   lane c adds bias[c], applies LeakyReLU (+ residual) and stores channel c of 16
//   pixels per tcgen05.ld -- a warp writes 32 consecutive channels (128 B) per pixel.
//   pipeline: activation ring (halo boxes, 2-3 stages) + weight ring (3-8 stages),
//   warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#include "tc_common.cuh"

namespace upf {

constexpr int HC_THREADS = 256;                   // warps: 0 weights TMA, 1 and 7 MMA issuers, 2-5 epilogue, 6 halo TMA
constexpr int HC_PITCH = 132;                     // floats per pixel row of the epilogue staging tile
constexpr int HC_COLS = 16;                       // halo positions per row (2048 B)
constexpr int HC_ROW_BYTES = HC_COLS * 128;

struct HaloParams {
  float* out; int ldo;
  const float* res; int ldr;
  const float* bias;
  int H, W, Cout, BN;
  int MT, dil, kblocks;
  int tiles_x, tiles_y;
  int na, nb;                 // A / B ring depths
  int ra, nba, b_rows, nbb;   // halo rows per TMA box / boxes per halo tile; weight rows per box / boxes per tile
  int a_bytes, b_stage_bytes; // per stage (multiples of 1024)
  int tmem_cols;
  int epi_helpers;            // 1: warps 0, 1, 6, 7 take half of the epilogue
  int two_issuers;            // 1: two MMA-issuing warps, two accumulators (2 * 128 * MT TMEM columns)
  int mc;                     // 1: launched as clusters of two CTAs that share every weight tile (each loads half, TMA multicast)
  int bo_mode;                // 1: descriptor base_offset = (start >> 7) & 7
  long long* probe;           // debug: per-role wait/issue cycle counters of CTA 0 (nullptr = off)
  float slope;
  int flags;
};

// half a weight tile, delivered to the same shared-memory offset (and counted on the same mbarrier offset) of BOTH CTAs of the pair
__device__ __forceinline__ void tma_load_3d_mc2(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  const uint16_t mask = 3;
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask) : "memory");
}
// tcgen05.commit arriving on the mbarrier at this offset in BOTH CTAs of the pair (a shared weight slot is free when both have read it)
__device__ __forceinline__ void umma_commit_mc2(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__global__ void __launch_bounds__(HC_THREADS)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (a uintptr_t round trip would turn every later
  // access into a generic-address load: measured 4x slower epilogue)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = base;
  uint8_t* b_ring = base + (size_t)p.na * p.a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.nb * p.b_stage_bytes);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + p.na;
  uint64_t* fullB = emptyA + p.na;
  uint64_t* emptyB = fullB + p.nb;
  uint64_t* accum_full = emptyB + p.nb;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next kernel's prologue start early (PDL)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int n = tile / p.tiles_y;
  const int x0 = tx * 8, y0 = ty * 16 * p.MT;
  const int d = p.dil;
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.na; ++s) { mbar_init(smem_u32(&fullA[s]), 1); mbar_init(smem_u32(&emptyA[s]), p.two_issuers ? 2 : 1); }
    for (int s = 0; s < p.nb; ++s) { mbar_init(smem_u32(&fullB[s]), 1); mbar_init(smem_u32(&emptyB[s]), p.mc ? 2 : 1); }
    mbar_init(smem_u32(accum_full), p.two_issuers ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (p.mc) cluster_sync_all();        // the peer's barriers exist before anything is multicast into this CTA
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int crank = p.mc ? (int)cluster_ctarank() : 0;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias) touched
  // only this CTA's own state and constant weights, and may overlap the tail of the previous kernel in the stream;
  // from here on we read activations / write outputs, so wait for the upstream grid to complete and flush.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer: weight tiles (one per tap) =====================
    if (elect_one()) {
      long long w_eb = 0, t_start = clock64();
      const int total = p.kblocks * 9;
      for (int ib = 0; ib < total; ++ib) {
        const int kb = ib / 9, tap = ib - kb * 9;
        const int sb = ib % p.nb;
        long long t0 = clock64();
        mbar_wait(smem_u32(&emptyB[sb]), (((uint32_t)(ib / p.nb)) & 1u) ^ 1u);
        w_eb += clock64() - t0;
        const uint32_t fb = smem_u32(&fullB[sb]);
        const uint32_t b_dst = smem_u32(b_ring + (size_t)sb * p.b_stage_bytes);
        if (p.mc) {
          // this CTA fetches weight rows [64 rank, +64) for both CTAs; the peer's half arrives on the same barrier
          mbar_expect_tx(fb, 2u * 64u * 128u);
          tma_load_3d_mc2(b_dst + (uint32_t)(crank * 64 * 128), &map_w, fb, kb * 32, crank * 64, tap);
        } else {
          mbar_expect_tx(fb, b_bytes);
          for (int jb = 0; jb < p.nbb; ++jb)
            tma_load_3d(b_dst + (uint32_t)(jb * p.b_rows * 128), &map_w, fb, kb * 32, jb * p.b_rows, tap);
        }
      }
      if (p.probe && blockIdx.x == 0) { p.probe[1] = w_eb; p.probe[2] = clock64() - t_start; }
    }
    __syncwarp();
  } else if (warp == 6) {
    // ===================== TMA producer: activation halo boxes (one per 32-channel block) =====================
    // (its own warp: a halo box must be requested the moment its ring slot frees, independently of the weight stream)
    if (elect_one()) {
      long long w_ea = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        const int sa = kb % p.na;
        long long t0 = clock64();
        mbar_wait(smem_u32(&emptyA[sa]), (((uint32_t)(kb / p.na)) & 1u) ^ 1u);
        w_ea += clock64() - t0;
        const uint32_t fa = smem_u32(&fullA[sa]);
        mbar_expect_tx(fa, (uint32_t)p.a_bytes);
        const uint32_t a_dst = smem_u32(a_ring + (size_t)sa * p.a_bytes);
        for (int j = 0; j < p.nba; ++j)
          tma_load_4d(a_dst + (uint32_t)(j * p.ra * HC_ROW_BYTES), &map_x, fa, kb * 32, x0 - d, y0 - d + j * p.ra, n);
      }
      if (p.probe && blockIdx.x == 0) p.probe[0] = w_ea;
    }
    __syncwarp();
  } else if (warp == 1 || warp == 7) {
    // ===================== MMA issuers =====================
    // M = 128 (output channels, rows past Cout are never stored), N = 128*MT pixels.  One thread needs ~635 cycles to issue a
    // tap (4 MMAs + commits + the next mbarrier poll, tools/microbench/tc_probe.cu) against 516 cycles of tensor work at
    // N = 256: measured 730-790 cycles per tap with ONE issuer (tools/probe_fine.py).  TWO issuers: issuer w takes the taps
    // with (kb*9 + tap) % 2 == w into its OWN accumulator (fixed order each: bitwise reproducible), the epilogue adds the two.
    const int wi = warp == 1 ? 0 : 1;
    if (wi == 0 || p.two_issuers) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((128 * p.MT) >> 3) << 17) | ((128u >> 4) << 24);
      const int total = p.kblocks * 9;
      const int step = p.two_issuers ? 2 : 1;
      const uint32_t tacc = tmem_base + (uint32_t)(wi * 128 * p.MT);
      long long w_fa = 0, w_fb = 0, t_start = clock64();
      int kb_ready = -1;
      uint32_t acc = 0;
      for (int ib = wi; ib < total; ib += step) {
        const int kb = ib / 9, tap = ib - kb * 9;
        const int sa = kb % p.na, sb = ib % p.nb;
        long long t0 = clock64();
        if (kb != kb_ready) {
          mbar_wait(smem_u32(&fullA[sa]), ((uint32_t)(kb / p.na)) & 1u);
          kb_ready = kb;
        }
        w_fa += clock64() - t0;
        const uint32_t a_base = smem_u32(a_ring + (size_t)sa * p.a_bytes);
        t0 = clock64();
        mbar_wait(smem_u32(&fullB[sb]), ((uint32_t)(ib / p.nb)) & 1u);
        w_fb += clock64() - t0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const int ky = tap / 3, kx = tap - ky * 3;
          const uint64_t dw = umma_desc_sw128(smem_u32(b_ring + (size_t)sb * p.b_stage_bytes));            // weights: A
          const uint32_t x_addr = a_base + (uint32_t)(((ky * d) * HC_COLS + kx * d) * 128);
          const uint64_t dx = umma_desc_sw128_ex(x_addr, HC_ROW_BYTES, (p.bo_mode & 1) ? (x_addr >> 7) : 0u);     // pixels: B
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tacc, dw + (uint64_t)(k * 2), dx + (uint64_t)(k * 2), idesc, acc | (uint32_t)k);
          if (p.mc) umma_commit_mc2(smem_u32(&emptyB[sb]));
          else umma_commit(smem_u32(&emptyB[sb]));
          if ((ib + step) / 9 != kb) umma_commit(smem_u32(&emptyA[sa]));      // this issuer's last tap of the channel block
          if (ib + step >= total) umma_commit(smem_u32(accum_full));
        }
        __syncwarp();
        acc = 1;
      }
      if (p.probe && blockIdx.x == 0 && lane == 0 && wi == 0) { p.probe[3] = w_fa; p.probe[4] = w_fb; p.probe[5] = clock64() - t_start; }
    }
  }
  // ===================== epilogue: warps 2..5, joined by the producer and issuer warps (0, 1, 6, 7 -- one per TMEM lane
  // quarter as well) once their loops are done: both phases are chains of load -> store latencies, two groups of four warps
  // run two such chains at once =====================
  const bool primary = warp >= 2 && warp < 6;
  const int eh = p.epi_helpers ? 1 : 0;                        // 1: eight epilogue warps
  if (primary || eh) {
    const int hw = primary ? 0 : 1;                            // which half of the work this group of four warps takes
    const int q = warp & 3;                                    // TMEM lane quarter = 32 output channels
    const int c = q * 32 + lane;                               // this thread's output channel (phase 1)
    long long t_e0 = clock64();
    mbar_wait(smem_u32(accum_full), 0);
    long long t_e1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // phase 1: TMEM (lane = channel, column = pixel) -> shared memory stage[pixel][channel] (pitch 132 floats:
    // lanes write consecutive channels of one pixel, conflict-free).  The rings are drained (every MMA retired).
    float* stage = reinterpret_cast<float*>(base);
    const int npx = 128 * p.MT;
    if (q * 32 < p.Cout) {                                     // warp-uniform: quarters past Cout hold nothing
      for (int n0 = eh * hw * 16; n0 < npx; n0 += 16 << eh) {  // 16 pixels = 2 tile rows of 8
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)n0, v);
        if (p.two_issuers) {
          uint32_t v2[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(npx + n0), v2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[(n0 + j) * HC_PITCH + c] = __uint_as_float(v[j]);
      }
    }
    // bias goes through shared memory: with ~206 KB of dynamic shared memory the L1 cache is a few KB, and a
    // __ldg per output element costs an L2 round trip (measured: 630 cycles per store iteration)
    float* s_bias = stage + 128 * p.MT * HC_PITCH;
    if (primary) s_bias[threadIdx.x - 64] = ((int)threadIdx.x - 64 < p.Cout) ? __ldg(p.bias + threadIdx.x - 64) : 0.f;   // 128 entries
    if (eh) asm volatile("bar.sync 1, 256;" ::: "memory");    // the eight epilogue warps
    else asm volatile("bar.sync 1, 128;" ::: "memory");       // the four epilogue warps
    const long long t_e2 = clock64();
    // phase 2: pixel-major 16-byte stores with bias + LeakyReLU (+ residual).  A warp serves 32/c4p pixels per
    // step (c4p = channel quads per pixel rounded up to a power of two): a lane keeps the same 4 channels for the
    // whole tile, so its bias quad lives in registers and no division is needed.
    const int c4n = (p.Cout + 3) >> 2;
    int c4p = 1;
    while (c4p < c4n) c4p <<= 1;
    const int ppw = 32 / c4p;                                  // pixels per warp step
    const int sub = lane / c4p, c0 = (lane % c4p) * 4;
    const bool lane_on = c0 < p.Cout;
    const bool vec_out = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const float4 bv = lane_on ? *reinterpret_cast<const float4*>(s_bias + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t img = (size_t)n * p.H * p.W;
    for (int nn0 = q * ppw + sub + eh * hw * 16 * ppw; nn0 < npx; nn0 += (16 * ppw) << eh) {   // 4 pixels per thread per trip: loads first, then stores
      float4 t[4];
      size_t pix[4];
      bool ok[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nn = nn0 + j * 4 * ppw;
        const int py = y0 + (nn >> 3), px = x0 + (nn & 7);
        ok[j] = lane_on && nn < npx && py < p.H && px < p.W;
        pix[j] = img + (size_t)py * p.W + px;
        t[j] = ok[j] ? *reinterpret_cast<const float4*>(stage + nn * HC_PITCH + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!ok[j]) continue;
        float f[4] = {lrelu(t[j].x + bv.x, p.slope), lrelu(t[j].y + bv.y, p.slope), lrelu(t[j].z + bv.z, p.slope),
                      lrelu(t[j].w + bv.w, p.slope)};
        if (p.res) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < p.Cout) f[k] += __ldg(p.res + pix[j] * p.ldr + c0 + k);
        }
        if (p.flags & UPF_FLAG_ROUND_TF32) {
#pragma unroll
          for (int k = 0; k < 4; ++k) f[k] = round_tf32(f[k]);
        }
        float* o = p.out + pix[j] * p.ldo + c0;
        if (vec_out && c0 + 4 <= p.Cout) {
          *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c0 + k < p.Cout) o[k] = f[k];
        }
      }
    }
    if (p.probe && blockIdx.x == 0 && threadIdx.x == 64) { p.probe[6] = t_e2 - t_e1; p.probe[7] = clock64() - t_e2; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }
  if (p.mc) cluster_sync_all();        // nobody leaves while the peer may still multicast into it or arrive on its barriers
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

static int g_halo_bo_mode = 0;   // measured on B200: the 128-byte swizzle is applied to absolute shared-memory address bits,
                                 // so shifted windows of a swizzled box need NO base_offset (bo_mode 1 gives garbage)
extern int g_tc_box_rows;
extern int g_tc_pdl;
static int g_halo_enabled = 1;
static int g_halo_mc = 1;          // pairs of CTAs share every weight tile by TMA multicast (upf_debug_conv_halo enabled bit 2 = OFF): with ONE MMA
                                   // issuer no faster (2.765 vs 2.764 ms per KITTI forward), with two 2.657 vs 2.672 ms
static int g_halo_a_boxes = 1;     // TMA boxes per halo tile (upf_debug_conv_halo enabled bits 4..7; must divide the halo rows)
static int g_halo_epi_helpers = 1; // upf_debug_conv_halo bo_mode bit 4 = four epilogue warps (A/B)
static int g_halo_two_issuers = 1; // upf_debug_conv_halo enabled bit 3 = one issuer (A/B)
static int g_halo_two_cta = 0;     // A/B (upf_debug_conv_halo enabled bit 1): 8x16-pixel tiles with ~108 KB rings, two resident CTAs per SM
static int g_halo_min_cin = 64;    // A/B on the whole KITTI forward (tools/ab_forward.py): off 3.83 ms, >=192 3.72, >=64 3.60, all 3.61
long long* g_halo_probe = nullptr;

// returns 1 when the launch was taken, 0 when the shape is not eligible (caller falls through to conv_tc), <0 / cudaError on failure
int conv2d_fwd_halo(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                    const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                    float slope, int flags, cudaStream_t st, int* taken) {
  *taken = 0;
  if (!g_halo_enabled || ks != 3 || stride != 1 || !(dil == 1 || dil == 2 || dil == 4) || Cout > 128) return 0;
  const int tiles_x = (W + 7) / 8;
  // two M-tiles per CTA when that still gives every SM work; else one
  int MT = 2;
  if ((long long)tiles_x * ((H + 31) / 32) * N < UPF_NUM_SMS) MT = 1;
  // two resident CTAs per SM (small tiles, short rings): one CTA's prologue / epilogue and commit bubbles run under the
  // other's MMAs, and the hardware balances the tail at 128-pixel granularity
  bool two_cta = g_halo_two_cta && MT == 2 && (16 + 2 * dil) * HC_ROW_BYTES * 2 + 2 * 128 * 128 <= 108 * 1024;
  if (two_cta) MT = 1;
  const int tiles_y = (H + 16 * MT - 1) / (16 * MT);
  const long long tiles = (long long)tiles_x * tiles_y * N;
  if (tiles < 96) return 0;                      // coarse levels: the cluster split-K kernel is the better fit
  // measured (tools/bench_halo.py, 1/4-res KITTI): the shared halo wins once the K loop is long (576->128: 152 vs
  // 193 us, 544->32: 138 vs 160) and loses on short-K / high-resolution layers (32->32 at 188x621: 76 vs 61 us)
  if (Cin < g_halo_min_cin) return 0;
  const int BN = (Cout + 15) & ~15;
  const int kblocks = (Cin + 31) / 32;
  const int cin_pad = kblocks * 32;
  const int rows = 16 * MT + 2 * dil;
  // TMA sub-boxes: `ra` halo rows (16 positions each) per box, must divide rows (rows is even)
  int ra = (g_tc_box_rows >= 16 && g_tc_box_rows <= 128) ? g_tc_box_rows / 16 : rows;
  if (g_halo_a_boxes > 1 && rows % g_halo_a_boxes == 0) ra = rows / g_halo_a_boxes;   // the halo as several concurrent boxes
  while (rows % ra) --ra;
  int b_rows = (g_tc_box_rows >= 8 && g_tc_box_rows <= 128 && g_tc_box_rows < BN) ? g_tc_box_rows : BN;
  while (BN % b_rows) b_rows -= 8;
  // weight tiles shared by a pair of CTAs: the kernel's operand traffic (16 KB of weights per tap + the halo) runs at ~47 of the
  // ~58 B/clk an SM can pull out of L2 and its MMA thread waits for weight slots 19 % of the time (tools/probe_fine.py);
  // with each CTA fetching half of every tile for both, the weight traffic per SM halves
  const bool mc = g_halo_mc && !two_cta && (tiles % 2 == 0) && BN > 64;
  if (mc) b_rows = 64;

  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    const cuuint32_t box[4] = {32, HC_COLS, (cuuint32_t)ra, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    MapKey key{x, ldx, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)Cin, 100000 + ra, 4};
    int e = encode_cached(key, &mx, 4, const_cast<float*>(x), dims, strides, box, estr);
    if (e) return e;
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)BN, 9};
    const cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 4, (cuuint64_t)cin_pad * BN * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)b_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    MapKey key{w_packed, cin_pad, b_rows, 9, BN, mc ? 33 : 3};
    int e = encode_cached(key, &mw, 3, const_cast<float*>(w_packed), dims, strides, box, estr);
    if (e) return e;
  }
  HaloParams p;
  p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr; p.bias = bias;
  p.H = H; p.W = W; p.Cout = Cout; p.BN = BN; p.MT = MT; p.dil = dil; p.kblocks = kblocks;
  p.tiles_x = tiles_x; p.tiles_y = tiles_y;
  p.a_bytes = rows * HC_ROW_BYTES;
  p.b_stage_bytes = 128 * 128;                 // the MMA reads M = 128 weight rows; rows past cout_pad16 are stale, never stored
  p.slope = slope;
  p.flags = flags;
  p.bo_mode = g_halo_bo_mode;
  p.mc = mc ? 1 : 0;
  p.probe = g_halo_probe;
  const bool two_issuers = g_halo_two_issuers && !two_cta;
  p.two_issuers = two_issuers ? 1 : 0;
  p.epi_helpers = g_halo_epi_helpers;
  p.tmem_cols = (two_issuers ? 2 : 1) * 128 * MT;   // lanes = channels, columns = pixels; one accumulator per issuer
  // ring depths within ~212 KB: at least 2 A stages, then as many B stages as fit (3..8)
  const int budget = two_cta ? 108 * 1024 : 212 * 1024;
  int na = (kblocks >= 3 && 3 * p.a_bytes + 4 * p.b_stage_bytes <= budget) ? 3 : 2;
  if (na > kblocks) na = kblocks < 1 ? 1 : kblocks;
  int nb = (budget - na * p.a_bytes) / p.b_stage_bytes;
  if (nb > 8) nb = 8;
  if (two_issuers) nb &= ~1;     // EVEN: each issuer owns the weight slots of its parity and sees every phase of their barriers
  if (nb < 2) return 0;
  p.na = na; p.nb = nb;
  if ((long long)na * p.a_bytes + (long long)nb * p.b_stage_bytes < 128ll * MT * HC_PITCH * 4 + 512) return 0;   // epilogue staging tile + bias
  p.ra = ra; p.nba = rows / ra; p.b_rows = b_rows; p.nbb = BN / b_rows;
  const size_t smem = (size_t)na * p.a_bytes + (size_t)nb * p.b_stage_bytes + (2 * na + 2 * nb + 2) * 8 + 1024;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_halo smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(HC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na_ = 0;
  if (mc) {
    attr[na_].id = cudaLaunchAttributeClusterDimension;
    attr[na_].val.clusterDim.x = 2; attr[na_].val.clusterDim.y = 1; attr[na_].val.clusterDim.z = 1;
    ++na_;
  }
  if (g_tc_pdl) {
    attr[na_].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na_].val.programmaticStreamSerializationAllowed = 1;
    ++na_;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na_;
  {
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv_halo_kernel, mx, mw, p);
    if (e != cudaSuccess) { set_error("conv_halo launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  }
  *taken = 1;
  return check_launch("conv_halo");
}

}  // namespace upf

// test / tuning hooks (not part of the hot-path ABI): select the descriptor base-offset convention and
// enable / disable the halo kernel so the two tensor-core kernels can be compared on the same shapes
extern "C" int upf_debug_probe(void* device_buffer_8x_int64) {
  upf::g_halo_probe = reinterpret_cast<long long*>(device_buffer_8x_int64);
  return 0;
}

extern "C" int upf_debug_conv_halo(int enabled, int bo_mode) {
  upf::g_halo_enabled = enabled & 1;
  upf::g_halo_two_cta = (enabled >> 1) & 1;
  upf::g_halo_mc = ((enabled >> 2) & 1) ? 0 : 1;
  upf::g_halo_two_issuers = ((enabled >> 3) & 1) ? 0 : 1;
  upf::g_halo_a_boxes = ((enabled >> 4) & 15) ? ((enabled >> 4) & 15) : 1;
  upf::g_halo_bo_mode = bo_mode & 7;
  upf::g_tc_pdl = (bo_mode & 8) ? 0 : 1;
  upf::g_halo_epi_helpers = (bo_mode & 16) ? 0 : 1;
  if ((bo_mode >> 8) & 0xff) upf::g_tc_box_rows = (bo_mode >> 8) & 0xff;     // tuning: rows per TMA box in bits 8..15 (0 = keep)
  if (bo_mode >> 16) upf::g_halo_min_cin = (bo_mode >> 16) - 1;              // tuning: min Cin + 1 in bits 16.. (0 = keep)
  return 0;
}
