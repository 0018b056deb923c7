// Fused cost volume + (optional) feature normalisation + LeakyReLU, forward and
// backward, for sm_100a.
//
// Replaces correlation_cuda_kernel.cu:15-114 (channels_first x2 + one block of
// 32 threads per output pixel, 81 serial shared-memory reductions) and the
// F.unfold formulation of utils/pytorch_correlation.py:27-50.
//
// Forward design (HBM-bound at C=32: 4*(2C+81) B/pixel vs 2*81*C flop/pixel):
//   * one CTA owns an 8-row x 32-column pixel tile; the f2 search window
//     ((8+2d) x (32+2d) positions) and the f1 tile are staged ONCE in shared
//     memory, 32 channels (one 128-byte row) per position, 16-byte chunks
//     XOR-swizzled by (position & 7) -- the layout TMA's SWIZZLE_128B produces
//     -- so that 8 consecutive lanes reading the same chunk of 8 consecutive
//     positions hit 8 different bank groups;
//   * zero padding is produced while staging (out-of-image positions are
//     written as zeros): no padded copy, no transpose, no memset;
//   * the normalisation (x-mean)/std of model/upflow.py:126-135 is applied in
//     registers between the global load and the shared store, so normalised
//     tensors never exist in HBM;
//   * warp w handles horizontal displacement dx = w - d, lane l handles column
//     l of the tile and walks the 8 rows: every f2 value loaded from shared
//     memory feeds up to 8 pixels x 4 channels = 32 FMAs and every f1 value 9
//     (register tile 8 pixels x (2d+1) vertical displacements, 3 FMA per float
//     read from shared memory);
//   * results are transposed through shared memory and stored as contiguous
//     (2d+1)^2-float runs per pixel straight into the channel slice of the
//     decoder's feature buffer.
#include "upf_common.cuh"

namespace upf {

constexpr int CORR_TX = 32;   // tile columns == lanes
constexpr int CORR_TY = 8;    // tile rows == pixels per thread
constexpr int CORR_CC = 32;   // channels staged per pass (128 B per position)

template <int D>
struct CorrCfg {
  static constexpr int WIN = 2 * D + 1;
  static constexpr int NWARPS = WIN;
  static constexpr int NT = NWARPS * 32;
  static constexpr int HROWS = CORR_TY + 2 * D;
  static constexpr int HCOLS = (CORR_TX + 2 * D + 7) & ~7;   // multiple of 8: the swizzle phase of a column is row-independent
  static constexpr int HPOS = HROWS * HCOLS;
  static constexpr int F1POS = CORR_TX * CORR_TY;
  static constexpr int NOUT = WIN * WIN;
  static constexpr int STAGE_BYTES = (HPOS + F1POS) * CORR_CC * 4;
  static constexpr int OUT_BYTES = F1POS * NOUT * 4;
  static constexpr int SMEM_BYTES = (STAGE_BYTES > OUT_BYTES ? STAGE_BYTES : OUT_BYTES) + 4 * CORR_CC * 4;
  static constexpr int MIN_CTAS = (2 * (SMEM_BYTES + 1024) <= 233472 && 2 * NT <= 1024) ? 2 : 1;
};

__device__ __forceinline__ int swz(int pos, int chunk) { return pos * 32 + (((chunk ^ pos) & 7) << 2); }  // float index

// stage `npos` positions x 32 channels of one operand into swizzled smem.  Loads are issued in batches of
// CORR_LB per thread before any is consumed (memory-level parallelism); the normalisation (x-mean)*rstd
// is applied in registers; out-of-image positions and channels >= C are written as zeros (= the zero
// padding of the correlation AFTER normalisation, as in the reference).
constexpr int CORR_LB = 4;
template <int NT, int COLS, bool VEC, typename T>
__device__ __forceinline__ void corr_stage(float* __restrict__ dst, const T* __restrict__ src, int ld,
                                           int n, int H, int W, int C, int c0, int y_org, int x_org,
                                           int npos, const float* __restrict__ mean, const float* __restrict__ rstd,
                                           bool norm) {
  const int units = npos * 8;
  const T* img = src + (size_t)n * H * W * ld;
  for (int u0 = threadIdx.x; u0 < units; u0 += NT * CORR_LB) {
    float4 v[CORR_LB];
    bool ok[CORR_LB];
#pragma unroll
    for (int i = 0; i < CORR_LB; ++i) {
      const int u = u0 + i * NT;
      const int pos = u >> 3, chunk = u & 7;
      const int r = pos / COLS, cidx = pos - r * COLS;
      const int y = y_org + r, x = x_org + cidx;
      const int c = c0 + chunk * 4;
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      ok[i] = u < units && y >= 0 && y < H && x >= 0 && x < W && c < C;
      if (ok[i]) {
        const T* p = img + ((size_t)y * W + x) * ld + c;
        if (VEC) {
          v[i] = ld4(p);
        } else {
          v[i].x = ld1(p);
          if (c + 1 < C) v[i].y = ld1(p + 1);
          if (c + 2 < C) v[i].z = ld1(p + 2);
          if (c + 3 < C) v[i].w = ld1(p + 3);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < CORR_LB; ++i) {
      const int u = u0 + i * NT;
      const int pos = u >> 3, chunk = u & 7;
      if (norm && ok[i]) {
        const int k = chunk * 4;
        const int c = c0 + k;
        v[i].x = __fmul_rn(__fsub_rn(v[i].x, mean[k + 0]), rstd[k + 0]);
        v[i].y = (VEC || c + 1 < C) ? __fmul_rn(__fsub_rn(v[i].y, mean[k + 1]), rstd[k + 1]) : 0.f;
        v[i].z = (VEC || c + 2 < C) ? __fmul_rn(__fsub_rn(v[i].z, mean[k + 2]), rstd[k + 2]) : 0.f;
        v[i].w = (VEC || c + 3 < C) ? __fmul_rn(__fsub_rn(v[i].w, mean[k + 3]), rstd[k + 3]) : 0.f;
      }
      if (u < units) *reinterpret_cast<float4*>(dst + swz(pos, chunk)) = v[i];
    }
  }
}

template <int D, bool VEC, typename T>
// (__maxnreg__ instead of a min-blocks hint: ptxas otherwise settles on 96 registers and spills accumulators
// inside the channel loop; 112 x 288 threads x 2 CTAs = 64512 registers still fits the SM; registers are granted per warp in units of 512,
// so the single-CTA 13-warp variant (d = 6) gets 144, not 152)
// (measured: the 13-warp d = 6 variant launches with 128 registers and fails with 144)
__global__ void __launch_bounds__(CorrCfg<D>::NT) __maxnreg__(CorrCfg<D>::MIN_CTAS == 2 ? 112 : 128)
corr_fwd_kernel(const T* __restrict__ f1, int ld1, const T* __restrict__ f2, int ld2,
                T* __restrict__ out, int ldo, int H, int W, int C,
                const double* __restrict__ stats1, const double* __restrict__ stats2,
                float slope, int flags, int tiles_x, int tiles_y, int n2_shift, int N) {
  pdl_prologue();
  using K = CorrCfg<D>;
  extern __shared__ __align__(128) float smem[];
  float* s_f2 = smem;
  float* s_f1 = smem + K::HPOS * CORR_CC;
  float* s_stat = smem + (K::SMEM_BYTES / 4 - 4 * CORR_CC);   // mean1,rstd1,mean2,rstd2 of the current chunk
  float* s_out = smem;                                        // overlays the staging area

  int tile = blockIdx.x;
  const int tx = tile % tiles_x; tile /= tiles_x;
  const int ty = tile % tiles_y;
  const int n = tile / tiles_y;
  const int n2 = (n + n2_shift) % N;
  const int x0 = tx * CORR_TX, y0 = ty * CORR_TY;
  const int lane = threadIdx.x & 31, dxi = threadIdx.x >> 5;
  const bool norm = stats1 != nullptr;
  const double npix = (double)H * (double)W;

  float acc[CORR_TY][K::WIN];
#pragma unroll
  for (int p = 0; p < CORR_TY; ++p)
#pragma unroll
    for (int q = 0; q < K::WIN; ++q) acc[p][q] = 0.f;

  // shared-memory read addresses: position = row*COLS + column with COLS % 8 == 0, so the XOR phase
  // depends on the column only and every row is a compile-time offset
  const int col2 = lane + dxi;
  const float* a_base = s_f1 + lane * 32;
  const float* b_base = s_f2 + col2 * 32;

  for (int c0 = 0; c0 < C; c0 += CORR_CC) {
    if (c0 > 0) __syncthreads();          // previous pass finished reading the stage
    if (norm) {
      if (threadIdx.x < 2 * CORR_CC) {
        const int which = threadIdx.x >> 5, k = threadIdx.x & 31, c = c0 + k;
        float m = 0.f, sd = 1.f;
        if (c < C) stats_to_mean_std((which ? stats2 + ((size_t)n2 * C + c) * 2 : stats1 + ((size_t)n * C + c) * 2), npix, m, sd);
        s_stat[which * 2 * CORR_CC + k] = m;
        // (x-mean)*(1/std): within 1 ulp of the reference's (x-mean)/std; 1/std itself is a correctly rounded division
        s_stat[which * 2 * CORR_CC + CORR_CC + k] = __fdiv_rn(1.0f, sd);
      }
      __syncthreads();
    }
    corr_stage<K::NT, K::HCOLS, VEC, T>(s_f2, f2, ld2, n2, H, W, C, c0, y0 - D, x0 - D, K::HPOS,
                                     s_stat + 2 * CORR_CC, s_stat + 3 * CORR_CC, norm);
    corr_stage<K::NT, CORR_TX, VEC, T>(s_f1, f1, ld1, n, H, W, C, c0, y0, x0, K::F1POS, s_stat, s_stat + CORR_CC, norm);
    __syncthreads();

    const int cend = (C - c0 < CORR_CC ? C - c0 : CORR_CC);
    const int nchunk = (cend + 3) >> 2;
#pragma unroll 1
    for (int ch = 0; ch < nchunk; ++ch) {
      const float* ap = a_base + (((ch ^ lane) & 7) << 2);
      const float* bp = b_base + (((ch ^ col2) & 7) << 2);
      float4 a[CORR_TY];
#pragma unroll
      for (int p = 0; p < CORR_TY; ++p) a[p] = *reinterpret_cast<const float4*>(ap + p * CORR_TX * 32);
#pragma unroll
      for (int j = 0; j < K::HROWS; ++j) {
        const float4 b = *reinterpret_cast<const float4*>(bp + j * K::HCOLS * 32);
#pragma unroll
        for (int p = 0; p < CORR_TY; ++p) {
          const int dyi = j - p;
          if (dyi >= 0 && dyi < K::WIN) {
            float s = acc[p][dyi];
            s = fmaf(a[p].x, b.x, s);
            s = fmaf(a[p].y, b.y, s);
            s = fmaf(a[p].z, b.z, s);
            s = fmaf(a[p].w, b.w, s);
            acc[p][dyi] = s;
          }
        }
      }
    }
  }
  __syncthreads();   // staging area is dead: reuse it for the output transpose

  // mean over channels (torch.mean = sum / C, utils/pytorch_correlation.py:47): q = s*(1/C) followed by one
  // Newton correction is the correctly rounded quotient (exact product when C is a power of two)
  const float fC = (float)C, inv = __fdiv_rn(1.0f, fC);
#pragma unroll
  for (int p = 0; p < CORR_TY; ++p)
#pragma unroll
    for (int q = 0; q < K::WIN; ++q) {
      const float s = acc[p][q];
      float v = __fmul_rn(s, inv);
      v = __fmaf_rn(__fmaf_rn(-v, fC, s), inv, v);
      s_out[(p * CORR_TX + lane) * K::NOUT + q * K::WIN + dxi] = maybe_round(lrelu(v, slope), flags);
    }
  __syncthreads();

  // coalesced store: each tile row is 32 pixels x NOUT contiguous floats (pitch ldo)
  const int wvalid = (W - x0 < CORR_TX ? W - x0 : CORR_TX);
  const int total = wvalid * K::NOUT;
  for (int r = 0; r < CORR_TY; ++r) {
    const int y = y0 + r;
    if (y >= H) break;
    T* orow = out + ((size_t)((size_t)n * H + y) * W + x0) * ldo;
    const float* srow = s_out + r * CORR_TX * K::NOUT;
    if (ldo == K::NOUT && sizeof(T) == 2 && ((total | (int)(((size_t)(orow - out)) & 1)) & 1) == 0) {
      // 2-byte storage: two outputs per 4-byte store (the run starts on an even element and has an even length)
      for (int e = 2 * threadIdx.x; e < total; e += 2 * K::NT) st2(orow + e, srow[e], srow[e + 1]);
    } else if (ldo == K::NOUT) {
      for (int e = threadIdx.x; e < total; e += K::NT) st1(orow + e, srow[e]);       // one contiguous run
    } else {
      for (int e = threadIdx.x; e < total; e += K::NT) {
        const int px = e / K::NOUT, k = e - px * K::NOUT;
        st1(orow + (size_t)px * ldo + k, srow[e]);
      }
    }
  }
}

template <int D, typename T = float>
static int launch_corr_fwd(const T* f1, int ld1, const T* f2, int ld2, T* out, int ldo,
                           int N, int H, int W, int C, const double* s1, const double* s2, int shift,
                           float slope, int flags, cudaStream_t st) {
  using K = CorrCfg<D>;
  const int tiles_x = (W + CORR_TX - 1) / CORR_TX, tiles_y = (H + CORR_TY - 1) / CORR_TY;
  const long long tiles = (long long)tiles_x * tiles_y * N;
  UPF_REQUIRE(tiles > 0 && tiles < (1ll << 31), "corr: bad tile count %lld", tiles);
  const bool vec = (C % 4 == 0) && (ld1 % 4 == 0) && (ld2 % 4 == 0) && aligned_vec4<T>(f1) && aligned_vec4<T>(f2);
  // opt in to >48 KB dynamic shared memory once per instantiation (per device would need a table: one
  // process drives one GPU here)
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(corr_fwd_kernel<D, true, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(corr_fwd_kernel<D, false, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("corr smem attr: %s", cudaGetErrorString(e));
      return (int)e;
    }
    attr_done.mark();
  }
  if (vec) {
    UPF_LAUNCH((corr_fwd_kernel<D, true, T>), (unsigned)tiles, K::NT, K::SMEM_BYTES, st, f1, ld1, f2, ld2, out, ldo, H, W, C, s1, s2,
                                                                          slope, flags, tiles_x, tiles_y, shift, N);
  } else {
    UPF_LAUNCH((corr_fwd_kernel<D, false, T>), (unsigned)tiles, K::NT, K::SMEM_BYTES, st, f1, ld1, f2, ld2, out, ldo, H, W, C, s1,
                                                                           s2, slope, flags, tiles_x, tiles_y, shift, N);
  }
  return check_launch("corr_fwd");
}

// --------------------------------------------------------------------------
// backward (correlation_cuda_kernel.cu:116-300).  One thread per (pixel,
// 4-channel group): gO (pre-multiplied by the LeakyReLU derivative and 1/C) is
// read 81 times per pixel from L1/L2, f1/f2 rows are 128-byte coalesced.
//   g1[n,y,x,c] = sum_k g[n,y,x,k]        * f2[n,y+dy,x+dx,c]
//   g2[n,y,x,c] = sum_k g[n,y-dy,x-dx,k]  * f1[n,y-dy,x-dx,c]
// --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
corr_bwd_kernel(const float* __restrict__ f1, int ld1, const float* __restrict__ f2, int ld2,
                const float* __restrict__ outv, int ldo, const float* __restrict__ go, int ldg,
                float* __restrict__ g1, int ldg1, float* __restrict__ g2, int ldg2,
                int N, int H, int W, int C, int D, float slope) {
  pdl_prologue();
  const int cg = (C + 3) >> 2;
  const long long total = (long long)N * H * W * cg;
  const int WIN = 2 * D + 1;
  const float invC = 1.0f / (float)C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg) * 4;
    long long pix = i / cg;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int n = (int)(pix / ((long long)W * H));
    float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
    const float* go_p = go + (size_t)pix * ldg;
    const float* ov_p = outv ? outv + (size_t)pix * ldo : nullptr;
    for (int dy = -D; dy <= D; ++dy)
      for (int dx = -D; dx <= D; ++dx) {
        const int k = (dy + D) * WIN + dx + D;
        // gradient wrt f1 at (y,x): partner f2 at (y+dy,x+dx)
        int yy = y + dy, xx = x + dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          float g = __ldg(go_p + k) * invC;
          if (ov_p && __ldg(ov_p + k) < 0.f) g *= slope;
          const float* q = f2 + ((size_t)((size_t)n * H + yy) * W + xx) * ld2 + c;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (c + t < C) a1[t] = fmaf(g, __ldg(q + t), a1[t]);
        }
        // gradient wrt f2 at (y,x): partner f1 at (y-dy,x-dx), whose output channel k looked at us
        yy = y - dy; xx = x - dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const size_t ppix = (size_t)((size_t)n * H + yy) * W + xx;
          float g = __ldg(go + ppix * ldg + k) * invC;
          if (outv && __ldg(outv + ppix * ldo + k) < 0.f) g *= slope;
          const float* q = f1 + ppix * ld1 + c;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            if (c + t < C) a2[t] = fmaf(g, __ldg(q + t), a2[t]);
        }
      }
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (c + t < C) {
        if (g1) g1[(size_t)pix * ldg1 + c + t] = a1[t];
        if (g2) g2[(size_t)pix * ldg2 + c + t] = a2[t];
      }
  }
}

// --------------------------------------------------------------------------
// small images (the 1/64 .. 1/8 pyramid levels: a few hundred to a few thousand
// pixels, up to 196 channels): the tiled kernel above would run on a handful of
// CTAs looping over 7 channel passes.  Here ONE CTA owns ONE pixel: its
// (normalised) f1 vector sits in shared memory, 4 lanes share a displacement
// (each a quarter of the channels, 16-byte loads that hit L1 because neighbouring
// CTAs read the same f2 rows), and a 2-step shuffle finishes the dot product.
// --------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(704)
corr_small_kernel(const float* __restrict__ f1, int ld1, const float* __restrict__ f2, int ld2,
                  float* __restrict__ out, int ldo, int H, int W, int C, int D,
                  const double* __restrict__ stats1, const double* __restrict__ stats2,
                  float slope, int flags, int n2_shift, int N) {
  pdl_prologue();
  extern __shared__ __align__(16) float sm[];        // a[Cp] | mean2[Cp] | rstd2[Cp]
  const int Cp = (C + 3) & ~3;
  float* s_a = sm;
  float* s_m2 = sm + Cp;
  float* s_r2 = sm + 2 * Cp;
  const int pix = blockIdx.x;                         // linear (n, y, x)
  const int x = pix % W, y = (pix / W) % H, n = pix / (W * H);
  const int n2 = (n + n2_shift) % N;
  const bool norm = stats1 != nullptr;
  const double npix = (double)H * (double)W;
  const float* a_src = f1 + (size_t)pix * ld1;
  for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
    float a = 0.f, m2 = 0.f, r2 = 1.f;
    if (c < C) {
      a = __ldg(a_src + c);
      if (norm) {
        float m1, sd1, sd2;
        stats_to_mean_std(stats1 + ((size_t)n * C + c) * 2, npix, m1, sd1);
        stats_to_mean_std(stats2 + ((size_t)n2 * C + c) * 2, npix, m2, sd2);
        a = __fmul_rn(__fsub_rn(a, m1), __fdiv_rn(1.0f, sd1));
        r2 = __fdiv_rn(1.0f, sd2);
      }
    }
    s_a[c] = a; s_m2[c] = m2; s_r2[c] = r2;
  }
  __syncthreads();

  const int WIN = 2 * D + 1, NOUT = WIN * WIN;
  const int k = threadIdx.x >> 2, part = threadIdx.x & 3;       // displacement, channel quarter
  float acc = 0.f;
  if (k < NOUT) {
    const int dy = k / WIN - D, dx = k % WIN - D;
    const int yy = y + dy, xx = x + dx;
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
      const float* b_src = f2 + ((size_t)((size_t)n2 * H + yy) * W + xx) * ld2;
      // channel groups of 4 are dealt round-robin to the 4 lanes of the displacement
      for (int c = part * 4; c < C; c += 16) {
        float4 b;
        if (VEC) {
          b = ldg4(b_src + c);
        } else {
          b.x = __ldg(b_src + c);
          b.y = c + 1 < C ? __ldg(b_src + c + 1) : 0.f;
          b.z = c + 2 < C ? __ldg(b_src + c + 2) : 0.f;
          b.w = c + 3 < C ? __ldg(b_src + c + 3) : 0.f;
        }
        const float4 a = *reinterpret_cast<const float4*>(s_a + c);
        if (norm) {
          const float4 m = *reinterpret_cast<const float4*>(s_m2 + c);
          const float4 r = *reinterpret_cast<const float4*>(s_r2 + c);
          b.x = __fmul_rn(__fsub_rn(b.x, m.x), r.x);
          b.y = __fmul_rn(__fsub_rn(b.y, m.y), r.y);
          b.z = __fmul_rn(__fsub_rn(b.z, m.z), r.z);
          b.w = __fmul_rn(__fsub_rn(b.w, m.w), r.w);
          if (!VEC) {                                       // padded channels carry a = 0 already; keep b finite
            if (c + 1 >= C) b.y = 0.f;
            if (c + 2 >= C) b.z = 0.f;
            if (c + 3 >= C) b.w = 0.f;
          }
        }
        acc = fmaf(a.x, b.x, acc);
        acc = fmaf(a.y, b.y, acc);
        acc = fmaf(a.z, b.z, acc);
        acc = fmaf(a.w, b.w, acc);
      }
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (k < NOUT && part == 0) {
    const float fC = (float)C, inv = __fdiv_rn(1.0f, fC);
    float v = __fmul_rn(acc, inv);
    v = __fmaf_rn(__fmaf_rn(-v, fC, acc), inv, v);          // correctly rounded acc / C
    out[(size_t)pix * ldo + k] = maybe_round(lrelu(v, slope), flags);
  }
}

int launch_corr_pipe(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                     int N, int H, int W, int C, int D, const double* s1, const double* s2, int shift,
                     float slope, int flags, cudaStream_t st, int* taken);

constexpr long long CORR_SMALL_MAX_PIX = 4096;   // measured: at 2x47x156 (C=64) the tiled kernel is already 2x faster

static int launch_corr_small(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                             int N, int H, int W, int C, int D, const double* s1, const double* s2, int shift,
                             float slope, int flags, cudaStream_t st) {
  const int nout = (2 * D + 1) * (2 * D + 1);
  const int threads = ((nout * 4 + 31) / 32) * 32;          // 4 lanes per displacement
  UPF_REQUIRE(threads <= 704, "corr_small: window too large");
  const bool vec = (C % 4 == 0) && (ld2 % 4 == 0) && aligned16(f2);
  const size_t smem = (size_t)3 * ((C + 3) & ~3) * sizeof(float);
  const unsigned grid = (unsigned)((long long)N * H * W);
  if (vec)
    UPF_LAUNCH((corr_small_kernel<true>), grid, threads, smem, st, f1, ld1, f2, ld2, out, ldo, H, W, C, D, s1, s2, slope, flags, shift, N);
  else
    UPF_LAUNCH((corr_small_kernel<false>), grid, threads, smem, st, f1, ld1, f2, ld2, out, ldo, H, W, C, D, s1, s2, slope, flags, shift, N);
  return check_launch("corr_small");
}

}  // namespace upf

extern "C" int upf_corr_lrelu_fwd(const float* f1, int ld1, const float* f2, int ld2, float* out, int ldo,
                                  int N, int H, int W, int C, int max_disp,
                                  const double* stats1, const double* stats2, int f2_batch_shift,
                                  float slope, int flags, void* stream) {
  using namespace upf;
  UPF_REQUIRE(f1 && f2 && out, "corr: null tensor");
  UPF_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < N, "corr: batch shift out of range");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "corr: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  UPF_REQUIRE(ld1 >= C && ld2 >= C, "corr: pitch smaller than C");
  const int nout = (2 * max_disp + 1) * (2 * max_disp + 1);
  UPF_REQUIRE(ldo >= nout, "corr: output pitch %d < %d", ldo, nout);
  UPF_REQUIRE((stats1 == nullptr) == (stats2 == nullptr), "corr: give both stats or neither");
  cudaStream_t st = (cudaStream_t)stream;
  UPF_REQUIRE(max_disp >= 1 && max_disp <= 6, "corr: max_disp %d not in 1..6", max_disp);
  {
    // enough 32x8 tiles per image, d <= 6: persistent warp-specialised kernel (loads overlap the FMA loop), corr_pipe.cu
    // (the choice depends on the IMAGE size only, never on N: an image must give the same bits in any batch)
    int taken = 0;
    const int e = launch_corr_pipe(f1, ld1, f2, ld2, out, ldo, N, H, W, C, max_disp, stats1, stats2, f2_batch_shift, slope, flags,
                                   st, &taken);
    if (e != 0 || taken) return e;
  }
  // coarse pyramid levels: too few tiles to occupy the chip (and up to 7 channel passes each)
  if ((long long)H * W <= CORR_SMALL_MAX_PIX / 2 && C <= 1024)
    return launch_corr_small(f1, ld1, f2, ld2, out, ldo, N, H, W, C, max_disp, stats1, stats2, f2_batch_shift, slope, flags, st);
  switch (max_disp) {
    case 1: return launch_corr_fwd<1>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    case 2: return launch_corr_fwd<2>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    case 3: return launch_corr_fwd<3>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    case 4: return launch_corr_fwd<4>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    case 5: return launch_corr_fwd<5>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    case 6: return launch_corr_fwd<6>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, stats1, stats2, f2_batch_shift, slope, flags, st);
    default: set_error("corr: max_disp %d not in 1..6", max_disp); return UPF_ENOTSUP;
  }
}

// fp16 / bf16 STORAGE: the tiled kernel with T -> fp32 conversion while staging and fp32 -> T at the
// store; every product and sum is fp32, as in the reference's Half dispatch (accumulation in float,
// correlation_cuda_kernel.cu:15-114 with scalar_t = Half)
template <typename T>
static int corr_fwd_lp(const void* f1, int ld1, const void* f2, int ld2, void* out, int ldo, int N, int H, int W, int C, int max_disp,
                       const double* s1, const double* s2, int shift, float slope, cudaStream_t st) {
  using namespace upf;
  const T *a = static_cast<const T*>(f1), *b = static_cast<const T*>(f2);
  T* o = static_cast<T*>(out);
  switch (max_disp) {
    case 1: return launch_corr_fwd<1, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
    case 2: return launch_corr_fwd<2, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
    case 3: return launch_corr_fwd<3, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
    case 4: return launch_corr_fwd<4, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
    case 5: return launch_corr_fwd<5, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
    default: return launch_corr_fwd<6, T>(a, ld1, b, ld2, o, ldo, N, H, W, C, s1, s2, shift, slope, 0, st);
  }
}

extern "C" int upf_corr_lrelu_fwd_lp(const void* f1, int ld1, const void* f2, int ld2, void* out, int ldo, int dtype,
                                     int N, int H, int W, int C, int max_disp, const double* stats1, const double* stats2,
                                     int f2_batch_shift, float slope, void* stream) {
  using namespace upf;
  UPF_REQUIRE(f1 && f2 && out, "corr_lp: null tensor");
  UPF_REQUIRE(dtype == UPF_DTYPE_F16 || dtype == UPF_DTYPE_BF16, "corr_lp: dtype must be UPF_DTYPE_F16 or UPF_DTYPE_BF16");
  UPF_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < N, "corr_lp: batch shift out of range");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "corr_lp: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  UPF_REQUIRE(max_disp >= 1 && max_disp <= 6, "corr_lp: max_disp %d not in 1..6", max_disp);
  UPF_REQUIRE(ld1 >= C && ld2 >= C && ldo >= (2 * max_disp + 1) * (2 * max_disp + 1), "corr_lp: pitch too small");
  UPF_REQUIRE((stats1 == nullptr) == (stats2 == nullptr), "corr_lp: give both stats or neither");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == UPF_DTYPE_F16)
    return corr_fwd_lp<__half>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, max_disp, stats1, stats2, f2_batch_shift, slope, st);
  return corr_fwd_lp<__nv_bfloat16>(f1, ld1, f2, ld2, out, ldo, N, H, W, C, max_disp, stats1, stats2, f2_batch_shift, slope, st);
}

extern "C" int upf_corr_lrelu_bwd(const float* f1, int ld1, const float* f2, int ld2, const float* out, int ldo,
                                  const float* grad_out, int ldg, float* grad_f1, int ldg1, float* grad_f2, int ldg2,
                                  int N, int H, int W, int C, int max_disp, float slope, void* stream) {
  using namespace upf;
  UPF_REQUIRE(f1 && f2 && grad_out, "corr_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && max_disp >= 1 && max_disp <= 6, "corr_bwd: bad shape");
  UPF_REQUIRE(slope == 1.0f || out != nullptr, "corr_bwd: LeakyReLU derivative needs the forward output");
  const long long total = (long long)N * H * W * ((C + 3) / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > UPF_NUM_SMS * 32) blocks = UPF_NUM_SMS * 32;
  UPF_LAUNCH((corr_bwd_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, f1, ld1, f2, ld2, slope == 1.0f ? nullptr : out,
                                                                     ldo, grad_out, ldg, grad_f1, ldg1, grad_f2, ldg2,
                                                                     N, H, W, C, max_disp, slope);
  return check_launch("corr_bwd");
}
