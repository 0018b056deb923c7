// Correlation + normalisation + LeakyReLU, pipelined variant for large images
// (max_disp <= 4, C % 4 == 0): persistent CTAs (one per SM), warp-specialised.
//
// corr.cu's tiled kernel alternates "stage a tile" and "compute a tile" inside
// every CTA and relies on a second resident CTA for overlap; ncu shows the two
// phases serialised most of the time (35 % of stall samples wait on the staging
// loads, 25 % FMA-pipe utilisation).  Here three warp roles run concurrently:
//   * (2d+1) COMPUTE warps run the same register-tiled inner product as corr.cu
//     (warp = horizontal displacement, lane = tile column, 8 rows x (2d+1)
//     vertical displacements per thread) on 16-channel chunks;
//   * two LOADER warps fetch the NEXT chunk (of this tile or of the CTA's next
//     tile) by TMA into the other half of a two-stage shared-memory ring: two box
//     loads per chunk (f2 search window {16 ch, 40, 16}, f1 tile {16 ch, 32, 8}),
//     SWIZZLE_64B, out-of-bounds zero fill = the zero padding of the correlation;
//     it waits for the bytes (mbarrier) and applies the normalisation
//     (x-mean)*(1/std) in place to the in-image positions;
//   * one STORER warp copies the finished (2d+1)^2-float result rows from the
//     transpose buffer to global memory while the compute warps are already in
//     the next tile.
// Named barriers (bar.sync / bar.arrive) hand the stages and the transpose
// buffer back and forth; there is no __syncthreads in the steady state.
// Shared-memory rows are 64 bytes (16 channels); the 16-byte chunk index is
// XOR-swizzled with (position >> 1) & 3 (= TMA's SWIZZLE_64B) so that 8
// consecutive lanes reading the same chunk of 8 consecutive positions hit 8
// distinct 16-byte bank groups.
#include "tc_common.cuh"

namespace upf {

constexpr int CP_TX = 32, CP_TY = 8, CP_CC = 16;   // tile, channels per chunk
constexpr int CP_NLOAD = 2, CP_NSTORE = 1;         // loader warps (TMA wait + in-place normalisation), storer warps

template <int D>
struct CPCfg {
  static constexpr int WIN = 2 * D + 1;
  static constexpr int NCW = WIN;                             // compute warps
  static constexpr int NT = (NCW + CP_NLOAD + CP_NSTORE) * 32;
  static constexpr int HROWS = CP_TY + 2 * D;
  static constexpr int HCOLS = (CP_TX + 2 * D + 7) & ~7;
  static constexpr int HPOS = HROWS * HCOLS;
  static constexpr int F1POS = CP_TX * CP_TY;
  static constexpr int NOUT = WIN * WIN;
  static constexpr int STAGE_FLOATS = (HPOS + F1POS) * CP_CC;
  static constexpr int OUT_FLOATS = F1POS * NOUT;
  static constexpr int SMEM_BYTES = (2 * STAGE_FLOATS + OUT_FLOATS) * 4 + 1024;
};

__device__ __forceinline__ int cp_swz(int pos, int chunk) { return pos * CP_CC + (((chunk ^ (pos >> 1)) & 3) << 2); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// barrier ids: 1,2 = FULL[stage] (loaders -> compute); 3,4 = EMPTY[stage] (compute -> loaders);
//              5 = OUT_FULL (compute -> storer); 6 = OUT_EMPTY (storer -> compute); 7 = loaders only
template <int D>
__global__ void __launch_bounds__(CPCfg<D>::NT, 1)
corr_pipe_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                 float* __restrict__ out, int ldo, int H, int W, int C,
                 const double* __restrict__ stats1, const double* __restrict__ stats2,
                 float slope, int tiles_x, int tiles_y, int n2_shift, int N, int total_tiles) {
  pdl_prologue();
  using K = CPCfg<D>;
  extern __shared__ __align__(1024) float smem_raw_f[];
  // (pointer arithmetic on the __shared__ array keeps every access an LDS/STS)
  float* smem = smem_raw_f + (((1024u - (smem_u32(smem_raw_f) & 1023u)) & 1023u) >> 2);
  float* s_out = smem + 2 * K::STAGE_FLOATS;
  __shared__ __align__(16) float s_stat[4][256];       // mean1, rstd1, mean2, rstd2 of the image being staged
  __shared__ uint64_t s_mbar[2];
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&s_mbar[0]), 1);
    mbar_init(smem_u32(&s_mbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool norm = stats1 != nullptr;
  const int nchunks = (C + CP_CC - 1) / CP_CC;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  constexpr int N_FE = (K::NCW + CP_NLOAD) * 32;        // participants of FULL / EMPTY
  constexpr int N_OUT = (K::NCW + CP_NSTORE) * 32;      // participants of OUT_FULL / OUT_EMPTY

  if (warp >= K::NCW && warp < K::NCW + CP_NLOAD) {
    // ============================== LOADER WARPS ==============================
    const int lt = threadIdx.x - K::NCW * 32;           // 0 .. 63
    constexpr int NLT = CP_NLOAD * 32;
    const double npix = (double)H * (double)W;
    int stat_n = -1;
    long long it = 0;
    for (int tloc = 0; tloc < my_tiles; ++tloc) {
      int tile = blockIdx.x + tloc * gridDim.x;
      const int tx = tile % tiles_x; tile /= tiles_x;
      const int ty = tile % tiles_y;
      const int n = tile / tiles_y, n2 = (n + n2_shift) % N;
      const int x0 = tx * CP_TX, y0 = ty * CP_TY;
      if (norm && n != stat_n) {                        // per-image statistics table (this warp only)
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
        const int s = (int)(it & 1);
        const int c0 = chunk * CP_CC;
        if (it >= 2) named_sync(3 + s, N_FE);           // the compute warps have drained this stage
        float* st2 = smem + s * K::STAGE_FLOATS;
        float* st1 = st2 + K::HPOS * CP_CC;
        if (lt == 0) {
          // generic-proxy reads/writes of this stage (compute warps, normalisation) precede the async-proxy refill
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          const uint32_t mb = smem_u32(&s_mbar[s]);
          mbar_expect_tx(mb, (uint32_t)(K::STAGE_FLOATS * 4));
          tma_load_4d(smem_u32(st2), &map2, mb, c0, x0 - D, y0 - D, n2);
          tma_load_4d(smem_u32(st1), &map1, mb, c0, x0, y0, n);
        }
        mbar_wait(smem_u32(&s_mbar[s]), (uint32_t)((it >> 1) & 1));
        if (norm) {
          // in-place (x - mean) * (1/std) on the in-image positions; TMA's zero fill stays zero.  A lane keeps one
          // 16-byte channel quad (its mean / rstd in registers) and walks positions, 8 per step.
          const int ch = lt & 3, q0 = lt >> 2;            // 16 position slots per pass
          const int c = c0 + ch * 4;
          if (c < C) {
            const float4 m2 = *reinterpret_cast<const float4*>(&s_stat[2][c]), r2 = *reinterpret_cast<const float4*>(&s_stat[3][c]);
            constexpr int P2 = (K::HCOLS + 15) / 16;
            for (int r = 0; r < K::HROWS; ++r) {
              const int y = y0 - D + r;
              if (y < 0 || y >= H) continue;
              float4 v[P2];
              float4* qp[P2];
#pragma unroll
              for (int i = 0; i < P2; ++i) {                 // loads first (independent), then the arithmetic and the stores
                const int cc = q0 + 16 * i, x = x0 - D + cc;
                qp[i] = (cc < K::HCOLS && x >= 0 && x < W) ? reinterpret_cast<float4*>(st2 + cp_swz(r * K::HCOLS + cc, ch)) : nullptr;
                if (qp[i]) v[i] = *qp[i];
              }
#pragma unroll
              for (int i = 0; i < P2; ++i)
                if (qp[i]) {
                  v[i].x = __fmul_rn(__fsub_rn(v[i].x, m2.x), r2.x); v[i].y = __fmul_rn(__fsub_rn(v[i].y, m2.y), r2.y);
                  v[i].z = __fmul_rn(__fsub_rn(v[i].z, m2.z), r2.z); v[i].w = __fmul_rn(__fsub_rn(v[i].w, m2.w), r2.w);
                  *qp[i] = v[i];
                }
            }
            const float4 m1 = *reinterpret_cast<const float4*>(&s_stat[0][c]), r1 = *reinterpret_cast<const float4*>(&s_stat[1][c]);
            for (int r = 0; r < CP_TY; r += 2) {
              float4 v[4];
              float4* qp[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int rr = r + (i >> 1), cc = q0 + 16 * (i & 1);
                qp[i] = (y0 + rr < H && x0 + cc < W) ? reinterpret_cast<float4*>(st1 + cp_swz(rr * CP_TX + cc, ch)) : nullptr;
                if (qp[i]) v[i] = *qp[i];
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (qp[i]) {
                  v[i].x = __fmul_rn(__fsub_rn(v[i].x, m1.x), r1.x); v[i].y = __fmul_rn(__fsub_rn(v[i].y, m1.y), r1.y);
                  v[i].z = __fmul_rn(__fsub_rn(v[i].z, m1.z), r1.z); v[i].w = __fmul_rn(__fsub_rn(v[i].w, m1.w), r1.w);
                  *qp[i] = v[i];
                }
            }
          }
        }
        named_arrive(1 + s, N_FE);                      // stage s is full
      }
    }
  } else if (warp >= K::NCW + CP_NLOAD) {
    // ============================== STORER WARP(S) ==============================
    const int st = threadIdx.x - (K::NCW + CP_NLOAD) * 32;
    constexpr int NST = CP_NSTORE * 32;
    for (int tloc = 0; tloc < my_tiles; ++tloc) {
      int tile = blockIdx.x + tloc * gridDim.x;
      const int tx = tile % tiles_x; tile /= tiles_x;
      const int ty = tile % tiles_y;
      const int n = tile / tiles_y;
      const int x0 = tx * CP_TX, y0 = ty * CP_TY;
      named_sync(5, N_OUT);                             // the tile's results are in s_out
      const int wvalid = (W - x0 < CP_TX ? W - x0 : CP_TX);
      for (int r = 0; r < CP_TY; ++r) {
        const int y = y0 + r;
        if (y >= H) break;
        float* orow = out + ((size_t)((size_t)n * H + y) * W + x0) * ldo;
        const float* srow = s_out + r * CP_TX * K::NOUT;
        if (ldo == K::NOUT) {
          const int total = wvalid * K::NOUT;           // one contiguous run
          if ((reinterpret_cast<uintptr_t>(orow) & 15) == 0) {
            const int t4 = total >> 2;
            for (int e = st; e < t4; e += NST) reinterpret_cast<float4*>(orow)[e] = reinterpret_cast<const float4*>(srow)[e];
            for (int e = (t4 << 2) + st; e < total; e += NST) orow[e] = srow[e];
          } else {
            for (int e = st; e < total; e += NST) orow[e] = srow[e];
          }
        } else {
          // one warp per pixel: NOUT consecutive floats, lanes stride the run
          for (int px = warp - K::NCW - CP_NLOAD; px < wvalid; px += CP_NSTORE) {
            float* o = orow + (size_t)px * ldo;
            const float* sp = srow + px * K::NOUT;
            for (int k = lane; k < K::NOUT; k += 32) o[k] = sp[k];
          }
        }
      }
      named_arrive(6, N_OUT);                           // s_out may be overwritten
    }
  } else {
    // ============================== COMPUTE WARPS ==============================
    const int dxi = warp;
    const int col2 = lane + dxi;
    float acc[CP_TY][K::WIN];
    long long it = 0;
    for (int tloc = 0; tloc < my_tiles; ++tloc) {
#pragma unroll
      for (int p = 0; p < CP_TY; ++p)
#pragma unroll
        for (int q = 0; q < K::WIN; ++q) acc[p][q] = 0.f;
      for (int chunk = 0; chunk < nchunks; ++chunk, ++it) {
        const int s = (int)(it & 1);
        named_sync(1 + s, N_FE);                        // wait until the loader filled stage s
        const float* stage = smem + s * K::STAGE_FLOATS;
        const float* a_base = stage + K::HPOS * CP_CC + lane * CP_CC;
        const float* b_base = stage + col2 * CP_CC;
        const int cend = (C - chunk * CP_CC < CP_CC ? C - chunk * CP_CC : CP_CC);
        const int nq = (cend + 3) >> 2;
#pragma unroll 1
        for (int ch = 0; ch < nq; ++ch) {
          const float* ap = a_base + (((ch ^ (lane >> 1)) & 3) << 2);
          const float* bp = b_base + (((ch ^ (col2 >> 1)) & 3) << 2);
          float4 a[CP_TY];
#pragma unroll
          for (int p = 0; p < CP_TY; ++p) a[p] = *reinterpret_cast<const float4*>(ap + p * CP_TX * CP_CC);
#pragma unroll
          for (int j = 0; j < K::HROWS; ++j) {
            const float4 b = *reinterpret_cast<const float4*>(bp + j * K::HCOLS * CP_CC);
#pragma unroll
            for (int p = 0; p < CP_TY; ++p) {
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
        named_arrive(3 + s, N_FE);                      // stage s may be refilled
      }
      // ---- tile done: mean over channels, LeakyReLU, transpose into s_out for the storer warps
      if (tloc > 0) named_sync(6, N_OUT);               // previous tile's rows have left s_out
      const float fC = (float)C, inv = __fdiv_rn(1.0f, fC);
#pragma unroll
      for (int p = 0; p < CP_TY; ++p)
#pragma unroll
        for (int q = 0; q < K::WIN; ++q) {
          const float sacc = acc[p][q];
          float v = __fmul_rn(sacc, inv);
          v = __fmaf_rn(__fmaf_rn(-v, fC, sacc), inv, v);       // correctly rounded sum / C (torch.mean)
          s_out[(p * CP_TX + lane) * K::NOUT + q * K::WIN + dxi] = lrelu(v, slope);
        }
      named_arrive(5, N_OUT);
    }
  }
}

template <int D>
static int launch_corr_pipe_t(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                              int N, int H, int W, int C, const double* s1, const double* s2, int shift,
                              float slope, cudaStream_t st) {
  using K = CPCfg<D>;
  const int tiles_x = (W + CP_TX - 1) / CP_TX, tiles_y = (H + CP_TY - 1) / CP_TY;
  const long long tiles = (long long)tiles_x * tiles_y * N;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(corr_pipe_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("corr_pipe smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done = true;
  }
  CUtensorMap m1, m2;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t str1[3] = {(cuuint64_t)ld1 * 4, (cuuint64_t)W * ld1 * 4, (cuuint64_t)H * W * ld1 * 4};
    const cuuint32_t box1[4] = {CP_CC, CP_TX, CP_TY, 1};
    MapKey k1{f1, ld1, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)C, 7000 + D, 64};
    int e = encode_cached(k1, &m1, 4, const_cast<float*>(f1), dims, str1, box1, estr, 64);
    if (e) return e;
    const cuuint64_t str2[3] = {(cuuint64_t)ld2 * 4, (cuuint64_t)W * ld2 * 4, (cuuint64_t)H * W * ld2 * 4};
    const cuuint32_t box2[4] = {CP_CC, (cuuint32_t)K::HCOLS, (cuuint32_t)K::HROWS, 1};
    MapKey k2{f2, ld2, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)C, 8000 + D, 64};
    e = encode_cached(k2, &m2, 4, const_cast<float*>(f2), dims, str2, box2, estr, 64);
    if (e) return e;
  }
  const unsigned grid = (unsigned)(tiles < UPF_NUM_SMS ? tiles : UPF_NUM_SMS);
  UPF_LAUNCH((corr_pipe_kernel<D>), grid, K::NT, K::SMEM_BYTES, st, m1, m2, out, ldo, H, W, C, s1, s2, slope, tiles_x, tiles_y, shift,
                                                          N, (int)tiles);
  return check_launch("corr_pipe");
}

static int g_corr_pipe_enabled = 1;

// returns 1 in *taken when this kernel handled the call
int launch_corr_pipe(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                     int N, int H, int W, int C, int D, const double* s1, const double* s2, int shift,
                     float slope, cudaStream_t st, int* taken) {
  *taken = 0;
  const bool vec = (C % 4 == 0) && (ld1 % 4 == 0) && (ld2 % 4 == 0) && aligned16(f1) && aligned16(f2);
  const long long tiles = (long long)((W + CP_TX - 1) / CP_TX) * ((H + CP_TY - 1) / CP_TY) * N;
  // persistent CTAs need several tiles each to amortise the pipeline fill and to balance the SMs: measured, at 240
  // tiles (1/4-res KITTI, both directions) the two-CTA-per-SM tiled kernel of corr.cu is still ~10 % faster
  if (!g_corr_pipe_enabled || !vec || D > 4 || C > 256 || tiles >= (1ll << 30) || tiles < 3 * UPF_NUM_SMS) return 0;
  *taken = 1;
  switch (D) {
    case 1: return launch_corr_pipe_t<1>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, st);
    case 2: return launch_corr_pipe_t<2>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, st);
    case 3: return launch_corr_pipe_t<3>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, st);
    default: return launch_corr_pipe_t<4>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, s1, s2, shift, slope, st);
  }
}

}  // namespace upf

extern "C" int upf_debug_corr_pipe(int enabled) {
  upf::g_corr_pipe_enabled = enabled;
  return 0;
}
