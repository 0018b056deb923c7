// Correlation + normalisation + LeakyReLU, pipelined variant for large images
// (C % 4 == 0): persistent CTAs (one per SM), warp-specialised, mbarrier ring.
//
// corr.cu's tiled kernel alternates "stage a tile" and "compute a tile" inside
// every CTA and relies on a second resident CTA for overlap; ncu shows the two
// phases serialised most of the time (35 % of stall samples wait on the staging
// loads, 25 % FMA-pipe utilisation).  Here the roles run concurrently:
//   * (2d+1) COMPUTE warps run the register-tiled inner product (warp =
//     horizontal displacement, lane = tile column, TY rows x (2d+1) vertical
//     displacements per thread) on 16-channel chunks;
//   * one ISSUER thread keeps the ring of chunks in flight by TMA: two box loads
//     per chunk (f2 search window {16 ch, TX+2d, TY+2d}, f1 tile {16 ch, TX, TY}),
//     SWIZZLE_64B, out-of-bounds zero fill = the zero padding of the correlation,
//     waiting only for the compute warps' "slot empty" arrivals;
//   * with fused normalisation two NORMALISER warps apply (x-mean)*(1/std) in
//     place to the in-image positions of a landed chunk and hand it on.
// Ring geometry is a compile-time choice (UPF_CP_CC channels per chunk x
// UPF_CP_STAGES slots, 114 KB either way): 16 x 2 measured 59.4 us on the HD
// shape, 8 x 4 (32-byte box rows, deeper prefetch) 61.4 us.
// Handshakes are mbarriers: full[s] (TMA bytes), ready[s] (normalised), empty[s]
// (one arrival per compute warp).  The finished tile goes through a transpose
// buffer (pitch = (2d+1)^2 floats, odd: conflict-free).  A contiguous output
// ([N,H,W,(2d+1)^2], W % 4 == 0) leaves it as one shared->global BULK COPY per
// tile row (cp.async.bulk, one thread, drained while the warps compute on); a
// channel slice of a wider buffer is written by the compute warps of schedulers
// 1..3, a warp per pixel.  (first version: ONE storer warp copying 83 KB per tile
// was busy 93 % of the time and the compute warps spent 14 % of theirs waiting
// for it.  Also tried and measured slower: raw sums parked column-minor and six
// finisher warps converting and storing them, 67.6 vs 61.4 us; the operands staged
// by three service warps with 16-byte cp.async instead of TMA, 79.9 us -- the
// copies share the LSU path with the compute warps' shared-memory loads.)
// Where the ~59 us go (clock64 around the phases of every compute warp, CTA 0, HD
// shape, d=4, 7 tiles, 101k cycles): waiting for operands 6.6k in total -- the TMA
// ring keeps up; FMA loop 10.3k cycles per tile for the three warps that share
// scheduler 0 (issue-bound, IPC 0.73) and 7.2k for the two-warp schedulers, which
// is the SHARED-MEMORY bound: 216 LDS.128 per channel quad = 6.9k wavefront cycles
// per tile, the pipe is 96 % busy; epilogue (convert, transpose, two CTA barriers,
// bulk store) ~4k per tile.  A variant that splits the last displacement column by
// rows over four helper warps (one per scheduler) and replaces the CTA barriers by
// per-half mbarriers measured 57.6 us: balanced, but the helpers' 2-row tiles read
// 1 word per 1.5 FMA and every warp then runs at the shared-memory bound (10.2k).
// The lever left is the FMA : shared-load ratio (3 with 8x9 blocking), not the pipeline.
// Shared-memory rows are 64 (32) bytes; the 16-byte chunk index is XOR-swizzled
// with (position >> 1) & 3 (TMA's SWIZZLE_64B; (position >> 2) & 1 = SWIZZLE_32B) so that 8
// consecutive lanes reading the same chunk of 8 consecutive positions hit 8
// distinct 16-byte bank groups.
#include "tc_common.cuh"

namespace upf {

#ifndef UPF_CP_CC
#define UPF_CP_CC 16
#define UPF_CP_STAGES 2
#endif
constexpr int CP_TX = 32, CP_CC = UPF_CP_CC, CP_STAGES = UPF_CP_STAGES;   // tile width, channels per chunk, ring depth
// XOR term of the 16-byte chunk index for position pos (= TMA SWIZZLE_32B for 8-channel rows, SWIZZLE_64B for 16)
__device__ __forceinline__ int cp_sw(int pos) { return CP_CC == 8 ? ((pos >> 2) & 1) : ((pos >> 1) & 3); }

template <int D>
struct CPCfg {
  static constexpr int WIN = 2 * D + 1;
  static constexpr int TY = D <= 4 ? 8 : 4;                  // tile rows (the transpose buffer bounds it)
  static constexpr int NCW = WIN;                             // compute warps 0 .. NCW-1
  static constexpr int ISS = NCW;                             // issuer warp (one thread)
  static constexpr int NNORM = 2;                             // normaliser warps (exit at once without statistics)
  static constexpr int NT = (NCW + 1 + NNORM) * 32;
  static constexpr int HROWS = TY + 2 * D;
  static constexpr int HCOLS = (CP_TX + 2 * D + 7) & ~7;
  static constexpr int HPOS = HROWS * HCOLS;
  static constexpr int F1POS = CP_TX * TY;
  static constexpr int NOUT = WIN * WIN;
  static constexpr int STAGE_FLOATS = (HPOS + F1POS) * CP_CC;  // multiple of 64 floats (256 B)
  static constexpr int OUT_FLOATS = F1POS * NOUT;              // [pixel][displacement]: odd pitch, conflict-free
  static constexpr int SMEM_BYTES = (CP_STAGES * STAGE_FLOATS + OUT_FLOATS) * 4 + 1024;
};

__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int D>
__global__ void __launch_bounds__(CPCfg<D>::NT, 1)
corr_pipe_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                 float* __restrict__ out, int ldo, int H, int W, int C,
                 const double* __restrict__ stats1, const double* __restrict__ stats2,
                 float slope, int flags, int tiles_x, int tiles_y, int n2_shift, int N, int total_tiles, int bulk_out) {
  pdl_prologue();
  using K = CPCfg<D>;
  constexpr int TY = K::TY;
  extern __shared__ __align__(1024) float smem_raw_f[];
  // (pointer arithmetic on the __shared__ array keeps every access an LDS/STS)
  float* smem = smem_raw_f + (((1024u - (smem_u32(smem_raw_f) & 1023u)) & 1023u) >> 2);
  float* s_out = smem + CP_STAGES * K::STAGE_FLOATS;
  __shared__ __align__(16) float s_stat[4][256];       // mean1, rstd1, mean2, rstd2 of the image being staged
  __shared__ uint64_t s_full[CP_STAGES], s_ready[CP_STAGES], s_empty[CP_STAGES];
  const bool norm = stats1 != nullptr;
  if (threadIdx.x == 0) {
    for (int s = 0; s < CP_STAGES; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_ready[s]), K::NNORM * 32);
      mbar_init(smem_u32(&s_empty[s]), K::NCW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunks = (C + CP_CC - 1) / CP_CC;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == K::ISS) {
    // ============================== ISSUER (one thread) ==============================
    if (lane == 0) {
      long long it = 0;
      for (int tloc = 0; tloc < my_tiles; ++tloc) {
        int tile = blockIdx.x + tloc * gridDim.x;
        const int tx = tile % tiles_x; tile /= tiles_x;
        const int ty = tile % tiles_y;
        const int n = tile / tiles_y, n2 = (n + n2_shift) % N;
        const int x0 = tx * CP_TX, y0 = ty * TY;
        for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
          const int s = (int)(it % CP_STAGES);
          const uint32_t ph = (uint32_t)((it / CP_STAGES) & 1);
          if (it >= CP_STAGES) mbar_wait(smem_u32(&s_empty[s]), ph ^ 1u);   // the compute warps have drained this slot
          // generic-proxy accesses of this slot (compute warps, normalisation) precede the async-proxy refill
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          float* st2 = smem + s * K::STAGE_FLOATS;
          float* st1 = st2 + K::HPOS * CP_CC;
          const uint32_t mb = smem_u32(&s_full[s]);
          mbar_expect_tx(mb, (uint32_t)(K::STAGE_FLOATS * 4));
          tma_load_4d(smem_u32(st2), &map2, mb, chunk * CP_CC, x0 - D, y0 - D, n2);
          tma_load_4d(smem_u32(st1), &map1, mb, chunk * CP_CC, x0, y0, n);
        }
      }
    }
  } else if (warp >= K::NCW) {
    const int si = warp - K::NCW - 1;
    if (norm) {
      // ============================== NORMALISER WARPS ==============================
      const int lt = si * 32 + lane;
      constexpr int NLT = K::NNORM * 32, PS = NLT / (CP_CC / 4);  // threads, position slots per pass
      const double npix = (double)H * (double)W;
      const int ch = lt & (CP_CC / 4 - 1), q0 = lt / (CP_CC / 4);   // a lane keeps one 16-byte channel quad
      int stat_n = -1;
      long long it = 0;
      for (int tloc = 0; tloc < my_tiles; ++tloc) {
        int tile = blockIdx.x + tloc * gridDim.x;
        const int tx = tile % tiles_x; tile /= tiles_x;
        const int ty = tile % tiles_y;
        const int n = tile / tiles_y, n2 = (n + n2_shift) % N;
        const int x0 = tx * CP_TX, y0 = ty * TY;
        if (n != stat_n) {                                // per-image statistics table (these warps only)
          named_sync(7, NLT);
          for (int c = lt; c < C; c += NLT) {
            float m, sd;
            stats_to_mean_std(stats1 + ((size_t)n * C + c) * 2, npix, m, sd);
            s_stat[0][c] = m; s_stat[1][c] = __fdiv_rn(1.0f, sd);
            stats_to_mean_std(stats2 + ((size_t)n2 * C + c) * 2, npix, m, sd);
            s_stat[2][c] = m; s_stat[3][c] = __fdiv_rn(1.0f, sd);
          }
          named_sync(7, NLT);
          stat_n = n;
        }
        for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
          const int s = (int)(it % CP_STAGES);
          const uint32_t ph = (uint32_t)((it / CP_STAGES) & 1);
          float* st2 = smem + s * K::STAGE_FLOATS;
          float* st1 = st2 + K::HPOS * CP_CC;
          mbar_wait(smem_u32(&s_full[s]), ph);
          const int c = chunk * CP_CC + ch * 4;
          if (c < C) {
            // in-place (x - mean) * (1/std) on the in-image positions; TMA's zero fill stays zero
            const float4 m2 = *reinterpret_cast<const float4*>(&s_stat[2][c]), r2 = *reinterpret_cast<const float4*>(&s_stat[3][c]);
            constexpr int U = 4;
            for (int p0 = q0; p0 < K::HPOS; p0 += PS * U) {
              float4 v[U];
              float4* qp[U];
#pragma unroll
              for (int i = 0; i < U; ++i) {                 // loads first (independent), then the arithmetic and the stores
                const int pos = p0 + PS * i;
                const int r = pos / K::HCOLS, cc = pos - r * K::HCOLS;
                const int y = y0 - D + r, x = x0 - D + cc;
                qp[i] = (pos < K::HPOS && y >= 0 && y < H && x >= 0 && x < W)
                            ? reinterpret_cast<float4*>(st2 + pos * CP_CC + (((ch ^ cp_sw(pos)) & (CP_CC / 4 - 1)) << 2)) : nullptr;
                if (qp[i]) v[i] = *qp[i];
              }
#pragma unroll
              for (int i = 0; i < U; ++i)
                if (qp[i]) {
                  v[i].x = __fmul_rn(__fsub_rn(v[i].x, m2.x), r2.x); v[i].y = __fmul_rn(__fsub_rn(v[i].y, m2.y), r2.y);
                  v[i].z = __fmul_rn(__fsub_rn(v[i].z, m2.z), r2.z); v[i].w = __fmul_rn(__fsub_rn(v[i].w, m2.w), r2.w);
                  *qp[i] = v[i];
                }
            }
            const float4 m1 = *reinterpret_cast<const float4*>(&s_stat[0][c]), r1 = *reinterpret_cast<const float4*>(&s_stat[1][c]);
            for (int p0 = q0; p0 < K::F1POS; p0 += PS * U) {
              float4 v[U];
              float4* qp[U];
#pragma unroll
              for (int i = 0; i < U; ++i) {
                const int pos = p0 + PS * i;                // CP_TX == 32: row = pos >> 5, column = pos & 31
                qp[i] = (pos < K::F1POS && y0 + (pos >> 5) < H && x0 + (pos & 31) < W)
                            ? reinterpret_cast<float4*>(st1 + pos * CP_CC + (((ch ^ cp_sw(pos)) & (CP_CC / 4 - 1)) << 2)) : nullptr;
                if (qp[i]) v[i] = *qp[i];
              }
#pragma unroll
              for (int i = 0; i < U; ++i)
                if (qp[i]) {
                  v[i].x = __fmul_rn(__fsub_rn(v[i].x, m1.x), r1.x); v[i].y = __fmul_rn(__fsub_rn(v[i].y, m1.y), r1.y);
                  v[i].z = __fmul_rn(__fsub_rn(v[i].z, m1.z), r1.z); v[i].w = __fmul_rn(__fsub_rn(v[i].w, m1.w), r1.w);
                  *qp[i] = v[i];
                }
            }
          }
          mbar_arrive(smem_u32(&s_ready[s]));   // every thread releases its own in-place writes to the waiters
                                                // (an elected lane after __syncwarp is equivalent, but racecheck cannot see it)
        }
      }
    }
  } else {
    // ============================== COMPUTE WARPS ==============================
    const int dxi = warp;
    const int col2 = lane + dxi;
    constexpr int NCT = K::NCW * 32;                    // compute threads
    const int ct = threadIdx.x;                         // compute warps are warps 0 .. NCW-1
    const int swa = cp_sw(lane), swb = cp_sw(col2);
    float acc[TY][K::WIN];
    long long it = 0;
    for (int tloc = 0; tloc < my_tiles; ++tloc) {
#pragma unroll
      for (int p = 0; p < TY; ++p)
#pragma unroll
        for (int q = 0; q < K::WIN; ++q) acc[p][q] = 0.f;
      for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
        const int s = (int)(it % CP_STAGES);
        const uint32_t ph = (uint32_t)((it / CP_STAGES) & 1);
        mbar_wait(smem_u32(norm ? &s_ready[s] : &s_full[s]), ph);
        const float* stage = smem + s * K::STAGE_FLOATS;
        const float* a_base = stage + K::HPOS * CP_CC + lane * CP_CC;
        const float* b_base = stage + col2 * CP_CC;
        const int cend = (C - chunk * CP_CC < CP_CC ? C - chunk * CP_CC : CP_CC);
#pragma unroll
        for (int ch = 0; ch < CP_CC / 4; ++ch) {
          if (ch * 4 < cend) {
            const float* ap = a_base + ((ch ^ swa) << 2);
            const float* bp = b_base + ((ch ^ swb) << 2);
            float4 a[TY];
#pragma unroll
            for (int p = 0; p < TY; ++p) a[p] = *reinterpret_cast<const float4*>(ap + p * CP_TX * CP_CC);
#pragma unroll
            for (int j = 0; j < K::HROWS; ++j) {
              const float4 b = *reinterpret_cast<const float4*>(bp + j * K::HCOLS * CP_CC);
#pragma unroll
              for (int p = 0; p < TY; ++p) {
                const int dyi = j - p;
                if (dyi >= 0 && dyi < K::WIN) {
                  float sacc = acc[p][dyi];
                  sacc = fmaf(a[p].x, b.x, sacc);
                  sacc = fmaf(a[p].y, b.y, sacc);
                  sacc = fmaf(a[p].z, b.z, sacc);
                  sacc = fmaf(a[p].w, b.w, sacc);
                  acc[p][dyi] = sacc;
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));   // slot s may be refilled
      }
      // ---- tile done: mean over channels, LeakyReLU, transpose through s_out, copy-out
      int tile = blockIdx.x + tloc * gridDim.x;
      const int tx = tile % tiles_x; tile /= tiles_x;
      const int ty = tile % tiles_y;
      const int n = tile / tiles_y;
      const int x0 = tx * CP_TX, y0 = ty * TY;
      if (bulk_out && ct == 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous tile's rows were read
      named_sync(1, NCT);                               // the previous tile's rows have left s_out
      const float fC = (float)C, inv = __fdiv_rn(1.0f, fC);
      if ((C & (C - 1)) == 0) {                         // power of two: sum * (1/C) is the correctly rounded mean
#pragma unroll
        for (int p = 0; p < TY; ++p)
#pragma unroll
          for (int q = 0; q < K::WIN; ++q)
            s_out[(p * CP_TX + lane) * K::NOUT + q * K::WIN + dxi] = maybe_round(lrelu(__fmul_rn(acc[p][q], inv), slope), flags);
      } else {
#pragma unroll
        for (int p = 0; p < TY; ++p)
#pragma unroll
          for (int q = 0; q < K::WIN; ++q) {
            const float sacc = acc[p][q];
            float v = __fmul_rn(sacc, inv);
            v = __fmaf_rn(__fmaf_rn(-v, fC, sacc), inv, v);     // correctly rounded sum / C (torch.mean)
            s_out[(p * CP_TX + lane) * K::NOUT + q * K::WIN + dxi] = maybe_round(lrelu(v, slope), flags);
          }
      }
      if (bulk_out) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> bulk-copy reads
      named_sync(2, NCT);
      const int wvalid = (W - x0 < CP_TX ? W - x0 : CP_TX);
      const int hvalid = (H - y0 < TY ? H - y0 : TY);
      if (bulk_out) {
        // every tile row is ONE contiguous, 16-byte aligned run of wvalid * NOUT floats in shared and in global
        // memory: one bulk copy per row, issued by one thread, drained by the copy engine while the warps compute on
        if (ct == 32) {
          const uint32_t nbytes = (uint32_t)(wvalid * K::NOUT * 4);
          for (int r = 0; r < hvalid; ++r) {
            float* orow = out + ((size_t)((size_t)n * H + y0 + r) * W + x0) * K::NOUT;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         ::"l"(orow), "r"(smem_u32(s_out + r * CP_TX * K::NOUT)), "r"(nbytes) : "memory");
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      } else if (warp & 3) {
        // channel slice of a wider buffer: a warp per pixel, NOUT consecutive floats.  Scheduler 0 hosts three compute
        // warps, the others two: the copy is done by the warps of schedulers 1..3 only (ncu: the warps of the
        // lighter schedulers waited 17 % of their time at the tile barrier)
        constexpr int NCOPY = K::NCW - ((K::NCW + 3) >> 2);
        const int npx = hvalid * CP_TX;
        float* obase = out + ((size_t)((size_t)n * H + y0) * W + x0) * ldo;
        for (int px = warp - ((warp + 3) >> 2); px < npx; px += NCOPY) {
          const int r = px >> 5, cx = px & 31;
          if (cx < wvalid) {
            float* o = obase + (r * W + cx) * ldo;
            const float* sp = s_out + px * K::NOUT;
#pragma unroll
            for (int k = 0; k < (K::NOUT + 31) / 32; ++k)
              if (k * 32 + lane < K::NOUT) o[k * 32 + lane] = sp[k * 32 + lane];
          }
        }
      }
    }
    if (bulk_out && ct == 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

template <int D>
static int launch_corr_pipe_t(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                              int N, int H, int W, int C, const double* s1, const double* s2, int shift,
                              float slope, int flags, cudaStream_t st) {
  using K = CPCfg<D>;
  const int tiles_x = (W + CP_TX - 1) / CP_TX, tiles_y = (H + K::TY - 1) / K::TY;
  const long long tiles = (long long)tiles_x * tiles_y * N;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(corr_pipe_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("corr_pipe smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.mark();
  }
  CUtensorMap m1, m2;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t str1[3] = {(cuuint64_t)ld1 * 4, (cuuint64_t)W * ld1 * 4, (cuuint64_t)H * W * ld1 * 4};
    const cuuint32_t box1[4] = {CP_CC, CP_TX, (cuuint32_t)K::TY, 1};
    MapKey k1{f1, ld1, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)C, 7000 + D, CP_CC * 4};
    int e = encode_cached(k1, &m1, 4, const_cast<float*>(f1), dims, str1, box1, estr, CP_CC * 4);
    if (e) return e;
    const cuuint64_t str2[3] = {(cuuint64_t)ld2 * 4, (cuuint64_t)W * ld2 * 4, (cuuint64_t)H * W * ld2 * 4};
    const cuuint32_t box2[4] = {CP_CC, (cuuint32_t)K::HCOLS, (cuuint32_t)K::HROWS, 1};
    MapKey k2{f2, ld2, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)C, 8000 + D, CP_CC * 4};
    e = encode_cached(k2, &m2, 4, const_cast<float*>(f2), dims, str2, box2, estr, CP_CC * 4);
    if (e) return e;
  }
  // contiguous output whose tile rows are 16-byte aligned runs: shared -> global bulk copies instead of ld/st
  const int bulk_out = (ldo == K::NOUT && W % 4 == 0 && aligned16(out)) ? 1 : 0;
  const unsigned grid = (unsigned)(tiles < UPF_NUM_SMS ? tiles : UPF_NUM_SMS);
  UPF_LAUNCH((corr_pipe_kernel<D>), grid, K::NT, K::SMEM_BYTES, st, m1, m2, out, ldo, H, W, C, s1, s2, slope, flags, tiles_x, tiles_y, shift,
                                                          N, (int)tiles, bulk_out);
  return check_launch("corr_pipe");
}

static int g_corr_pipe_enabled = 1;
// tiles PER IMAGE from which the pipelined kernel is used (never a function of N: an image must give the same bits
// in any batch).  Measured (tools/time_corr.py, d=4, fused normalisation): 2x94x311 C=32 32.5 us vs 45.4 us tiled,
// 2x47x156 C=64 (30 tiles per image) 29.3 vs 40.7
static int g_corr_pipe_min_tiles = 25;

// returns 1 in *taken when this kernel handled the call
int launch_corr_pipe(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                     int N, int H, int W, int C, int D, const double* s1, const double* s2, int shift,
                     float slope, int flags, cudaStream_t st, int* taken) {
  *taken = 0;
  const bool vec = (C % 4 == 0) && (ld1 % 4 == 0) && (ld2 % 4 == 0) && aligned16(f1) && aligned16(f2);
  const int ty = D <= 4 ? 8 : 4;
  const long long tiles_img = (long long)((W + CP_TX - 1) / CP_TX) * ((H + ty - 1) / ty);
  if (!g_corr_pipe_enabled || !vec || D > 6 || C > 256 || tiles_img * N >= (1ll << 30) || tiles_img < g_corr_pipe_min_tiles) return 0;
  *taken = 1;
  switch (D) {
    case 1: return launch_corr_pipe_t<1>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
    case 2: return launch_corr_pipe_t<2>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
    case 3: return launch_corr_pipe_t<3>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
    case 4: return launch_corr_pipe_t<4>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
    case 5: return launch_corr_pipe_t<5>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
    default: return launch_corr_pipe_t<6>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, flags, st);
  }
}

}  // namespace upf

// enabled: 0 = never, 1 = default threshold, n > 1 = use the pipelined kernel from n tiles per image on (A/B runs)
extern "C" int upf_debug_corr_pipe(int enabled) {
  upf::g_corr_pipe_enabled = enabled != 0;
  upf::g_corr_pipe_min_tiles = enabled > 1 ? enabled : 25;
  return 0;
}
