// Backward kernels of the decoder path that are not a re-use of a forward kernel:
//   * weight / bias gradient of conv() (replaces cuDNN wgrad behind nn.Conv2d, model/pwc_modules.py:10-31);
//   * LeakyReLU backward taken from the saved OUTPUT (nn.LeakyReLU(0.1, inplace=True), :26-30);
//   * normalize_features backward THROUGH the moments (model/upflow.py:108-135: torch.mean / torch.var are part
//     of the graph);
//   * upsample2d_flow_as backward (model/pwc_modules.py:77-90: bilinear, align_corners=True, per-channel scale);
//   * the pointwise pieces of sgu_model.forward (model/upflow.py:79-88): sigmoid and the blend, both directions.
// The input gradient of a convolution (cuDNN dgrad) is the FORWARD kernel run on the flipped, transposed weights
// (ops.py: conv_dgrad), and the correlation / warp gradients live next to their forward kernels.
//
// Everything here is deterministic: reductions over pixels are split into fixed ranges whose partial results are
// summed in a fixed order (no floating-point atomics).
#include "upf_common.cuh"

namespace upf {

// ------------------------------------------------------------------------------------------------ wgrad
// dW[tap][ci][co] = sum over output pixels p of X[p (+) tap][ci] * G[p][co]:  per tap a GEMM with the pixels as
// the K dimension.  CTA tile 64 ci x 64 co, 16 pixels per shared-memory step, 4x4 register tile per thread.
// grid = (ci tiles * co tiles, taps, pixel splits); split s writes part[s][tap][ci][co].
constexpr int WG_KP = 16, WG_NT = 256;

// BM x BN CTA tile (ci x co), TM x TN register tile per thread, 256 threads = (BM/TM) x (BN/TN).
// 128x128 / 8x8 reads one shared word per 4 FMAs (the first version, 64x64 / 4x4, one per 2: shared-memory bound).
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(WG_NT, 2)
conv_wgrad_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g, int ldg, float* __restrict__ part,
                  int N, int H, int W, int Ho, int Wo, int Cin, int Cout, int ks, int stride, int dil, int pad,
                  int co_tiles, long long pix_per_split, long long npix) {
  static_assert((BM / TM) * (BN / TN) == WG_NT, "thread layout");
  static_assert(TM % 4 == 0 && TN % 2 == 0, "vector widths");
  __shared__ __align__(16) float Xs[WG_KP][BM];
  __shared__ __align__(16) float Gs[WG_KP][BN];
  const int tile = blockIdx.x, tap = blockIdx.y, split = blockIdx.z;
  const int ci0 = (tile / co_tiles) * BM, co0 = (tile % co_tiles) * BN;
  const int ky = tap / ks, kx = tap % ks;
  const int tid = threadIdx.x;
  constexpr int TXN = BN / TN;                               // threads along co
  const int tx = tid % TXN, ty = tid / TXN;
  const bool vx = (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const bool vg = (ldg % 4 == 0) && ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  const long long p_begin = (long long)split * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > npix) p_end = npix;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  constexpr int XV = WG_KP * BM / 4, GV = WG_KP * BN / 4;    // float4 loads per step for each operand

  float4 xv[(XV + WG_NT - 1) / WG_NT], gv[(GV + WG_NT - 1) / WG_NT];
  // global -> registers for the step starting at pixel p0 (issued one step ahead: the loads fly during the FMAs)
  auto fetch = [&](long long p0) {
#pragma unroll
    for (int r = 0; r < (XV + WG_NT - 1) / WG_NT; ++r) {
      const int v = tid + r * WG_NT;
      xv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int lrow = v / (BM / 4), lcol = (v % (BM / 4)) * 4;
      const long long p = p0 + lrow;
      if (v < XV && p < p_end) {
        const int ox = (int)(p % Wo);
        const long long t = p / Wo;
        const int oy = (int)(t % Ho), n = (int)(t / Ho);
        const int iy = oy * stride - pad + ky * dil, ix = ox * stride - pad + kx * dil;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const float* xp = x + ((size_t)((size_t)n * H + iy) * W + ix) * ldx + ci0 + lcol;
          if (vx && ci0 + lcol + 3 < Cin) xv[r] = ldg4(xp);
          else {
            if (ci0 + lcol < Cin) xv[r].x = __ldg(xp);
            if (ci0 + lcol + 1 < Cin) xv[r].y = __ldg(xp + 1);
            if (ci0 + lcol + 2 < Cin) xv[r].z = __ldg(xp + 2);
            if (ci0 + lcol + 3 < Cin) xv[r].w = __ldg(xp + 3);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < (GV + WG_NT - 1) / WG_NT; ++r) {
      const int v = tid + r * WG_NT;
      gv[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      const int lrow = v / (BN / 4), lcol = (v % (BN / 4)) * 4;
      const long long p = p0 + lrow;
      if (v < GV && p < p_end) {
        const float* gp = g + (size_t)p * ldg + co0 + lcol;
        if (vg && co0 + lcol + 3 < Cout) gv[r] = ldg4(gp);
        else {
          if (co0 + lcol < Cout) gv[r].x = __ldg(gp);
          if (co0 + lcol + 1 < Cout) gv[r].y = __ldg(gp + 1);
          if (co0 + lcol + 2 < Cout) gv[r].z = __ldg(gp + 2);
          if (co0 + lcol + 3 < Cout) gv[r].w = __ldg(gp + 3);
        }
      }
    }
  };
  fetch(p_begin);
  for (long long p0 = p_begin; p0 < p_end; p0 += WG_KP) {
    __syncthreads();                                      // the previous step's tiles have been consumed
#pragma unroll
    for (int r = 0; r < (XV + WG_NT - 1) / WG_NT; ++r) {
      const int v = tid + r * WG_NT;
      if (v < XV) *reinterpret_cast<float4*>(&Xs[v / (BM / 4)][(v % (BM / 4)) * 4]) = xv[r];
    }
#pragma unroll
    for (int r = 0; r < (GV + WG_NT - 1) / WG_NT; ++r) {
      const int v = tid + r * WG_NT;
      if (v < GV) *reinterpret_cast<float4*>(&Gs[v / (BN / 4)][(v % (BN / 4)) * 4]) = gv[r];
    }
    __syncthreads();
    if (p0 + WG_KP < p_end) fetch(p0 + WG_KP);
#pragma unroll
    for (int k = 0; k < WG_KP; ++k) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * TM + i]);
        av[i] = a.x; av[i + 1] = a.y; av[i + 2] = a.z; av[i + 3] = a.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        // a thread's TN output channels are pairs 2*TXN apart: consecutive lanes read consecutive 8-byte words
        // (tx * TN + j made 16 lanes hit 4 bank groups: ncu, 38 % of the shared wavefronts were conflicts)
        const float2 b = *reinterpret_cast<const float2*>(&Gs[k][(j / 2) * (2 * TXN) + tx * 2]);
        bv[j] = b.x; bv[j + 1] = b.y;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
  const int taps = ks * ks;
  float* dst = part + ((size_t)split * taps + tap) * Cin * Cout;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int ci = ci0 + ty * TM + i;
    if (ci >= Cin) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = co0 + (j / 2) * (2 * TXN) + tx * 2 + (j & 1);
      if (co < Cout) dst[(size_t)ci * Cout + co] = acc[i][j];
    }
  }
}

// per-channel sum of G over a pixel range (bias gradient partials): part[split][co].  256 threads = `cw` channel lanes
// x 256/cw pixel slots; the slots are summed through shared memory in slot order (deterministic).  (First version:
// one thread per channel walking its pixel range alone -- 14 % of a training step.)
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ g, int ldg, float* __restrict__ part, int Cout, int cw, long long pix_per_split,
              long long npix) {
  __shared__ float s_part[256];
  const int split = blockIdx.x;
  const long long p_begin = (long long)split * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > npix) p_end = npix;
  const int slots = 256 / cw, lane_c = threadIdx.x % cw, slot = threadIdx.x / cw;
  for (int c0 = 0; c0 < Cout; c0 += cw) {
    const int co = c0 + lane_c;
    float s = 0.f;
    if (co < Cout) {
      float s1 = 0.f, s2 = 0.f, s3 = 0.f;
      long long p = p_begin + slot;
      for (; p + 3 * slots < p_end; p += 4 * slots) {       // four independent loads in flight
        s += __ldg(g + (size_t)p * ldg + co);
        s1 += __ldg(g + (size_t)(p + slots) * ldg + co);
        s2 += __ldg(g + (size_t)(p + 2 * slots) * ldg + co);
        s3 += __ldg(g + (size_t)(p + 3 * slots) * ldg + co);
      }
      for (; p < p_end; p += slots) s += __ldg(g + (size_t)p * ldg + co);
      s = (s + s1) + (s2 + s3);
    }
    s_part[threadIdx.x] = s;
    __syncthreads();
    if (slot == 0 && co < Cout) {
      float t = 0.f;
      for (int k = 0; k < slots; ++k) t += s_part[k * cw + lane_c];
      part[(size_t)split * Cout + co] = t;
    }
    __syncthreads();
  }
}

// out[i] = sum over splits (ascending) of part[s][i]
__global__ void __launch_bounds__(256)
reduce_splits_kernel(const float* __restrict__ part, float* __restrict__ out, long long n, int splits) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * n + i];
    out[i] = s;
  }
}

static unsigned grid_for(long long n);
int launch_reduce_splits(const float* part, float* out, long long n, int splits, cudaStream_t st) {
  reduce_splits_kernel<<<grid_for(n), 256, 0, st>>>(part, out, n, splits);
  return check_launch("reduce_splits");
}

// ------------------------------------------------------------------------------------------------ pointwise
enum { PW_LRELU_BWD = 0, PW_SIGMOID = 1, PW_SIGMOID_BWD = 2 };

// a, b, out: [npix][C] with pitches; LRELU_BWD: out = b * (a > 0 ? 1 : slope) (a = saved output, b = grad);
// SIGMOID: out = 1/(1+exp(-a)); SIGMOID_BWD: out = b * a * (1 - a) (a = saved output)
__global__ void __launch_bounds__(256)
pointwise_kernel(int op, const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                 float* __restrict__ out, int ldo, long long npix, int C, float slope) {
  const long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i - p * C);
    const float av = a[(size_t)p * lda + c];
    float r;
    if (op == PW_LRELU_BWD) r = b[(size_t)p * ldb + c] * (av > 0.f ? 1.0f : slope);
    else if (op == PW_SIGMOID) r = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-av)));
    else r = b[(size_t)p * ldb + c] * av * (1.0f - av);
    out[(size_t)p * ldo + c] = r;
  }
}

// blend of sgu_model.forward (model/upflow.py:88): out_c = w_c * (1 - m) + f_c * m, c in {0,1}
__global__ void __launch_bounds__(256)
blend_fwd_kernel(const float* __restrict__ w, int ldw, const float* __restrict__ f, int ldf, const float* __restrict__ m,
                 int ldm, float* __restrict__ out, int ldo, long long npix) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const float mv = m[(size_t)p * ldm];
#pragma unroll
    for (int c = 0; c < 2; ++c)
      out[(size_t)p * ldo + c] = __fadd_rn(__fmul_rn(w[(size_t)p * ldw + c], __fsub_rn(1.0f, mv)), __fmul_rn(f[(size_t)p * ldf + c], mv));
  }
}
// gw_c = g_c (1 - m); gf_c = g_c m; gm = sum_c g_c (f_c - w_c)
__global__ void __launch_bounds__(256)
blend_bwd_kernel(const float* __restrict__ w, int ldw, const float* __restrict__ f, int ldf, const float* __restrict__ m,
                 int ldm, const float* __restrict__ g, int ldg, float* __restrict__ gw, int ldgw, float* __restrict__ gf,
                 int ldgf, float* __restrict__ gm, int ldgm, long long npix) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
    const float mv = m[(size_t)p * ldm];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float gv = g[(size_t)p * ldg + c];
      gw[(size_t)p * ldgw + c] = gv * (1.0f - mv);
      gf[(size_t)p * ldgf + c] = gv * mv;
      s = fmaf(gv, f[(size_t)p * ldf + c] - w[(size_t)p * ldw + c], s);
    }
    gm[(size_t)p * ldgm] = s;
  }
}

// ------------------------------------------------------------------------------------------------ normalisation
// y = (x - mean) / std with mean, std functions of x (unbiased variance, std = sqrt(var + 1e-16)):
//   dx = ( g - mean(g) - y * sum(g*y) / (n-1) ) / std
// pass 1: per split, per (image, channel): sum g and sum g*y over a pixel range -> part[split][n][c][2] (double)
// Like colsum_kernel: a CTA is cw channel lanes x 256/cw pixel slots (cw = C rounded up to a power of two, <= 256), every
// thread walks its slot's pixels of the split, the slots are summed through shared memory in slot order (deterministic).
// (First version: one thread per channel walking the split alone -- 32 of 256 threads at C = 32, 65 us per call.)
__global__ void __launch_bounds__(256)
featnorm_bwd_sums_kernel(const float* __restrict__ x, int ldx, const double* __restrict__ stats, const float* __restrict__ g,
                         int ldg, double* __restrict__ part, int N, int HW, int C, int splits, int cw) {
  __shared__ double s_sg[256], s_sgy[256];
  const int n = blockIdx.x / splits, split = blockIdx.x % splits;
  const int per = (HW + splits - 1) / splits;
  const int p_begin = split * per, p_end = min(HW, p_begin + per);
  const int slots = 256 / cw, lane_c = threadIdx.x % cw, slot = threadIdx.x / cw;
  for (int c0 = 0; c0 < C; c0 += cw) {
    const int c = c0 + lane_c;
    double sg = 0.0, sgy = 0.0;
    if (c < C) {
      float m, s;
      stats_to_mean_std(stats + ((size_t)n * C + c) * 2, (double)HW, m, s);
      for (int p = p_begin + slot; p < p_end; p += slots) {
        const size_t pix = (size_t)n * HW + p;
        const float gv = __ldg(g + pix * ldg + c);
        const float y = __fdiv_rn(__fsub_rn(__ldg(x + pix * ldx + c), m), s);
        sg += (double)gv;
        sgy += (double)gv * (double)y;
      }
    }
    s_sg[threadIdx.x] = sg; s_sgy[threadIdx.x] = sgy;
    __syncthreads();
    if (slot == 0 && c < C) {
      double a = 0.0, b = 0.0;
      for (int k = 0; k < slots; ++k) { a += s_sg[k * cw + lane_c]; b += s_sgy[k * cw + lane_c]; }
      double* d = part + (((size_t)split * N + n) * C + c) * 2;
      d[0] = a; d[1] = b;
    }
    __syncthreads();
  }
}
__global__ void __launch_bounds__(256)
featnorm_bwd_apply_kernel(const float* __restrict__ x, int ldx, const double* __restrict__ stats, const float* __restrict__ g,
                          int ldg, const double* __restrict__ part, float* __restrict__ gx, int ldgx, int N, int HW, int C,
                          int splits) {
  const long long total = (long long)N * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int n = (int)(pix / HW);
    float m, s;
    stats_to_mean_std(stats + ((size_t)n * C + c) * 2, (double)HW, m, s);
    double sg = 0.0, sgy = 0.0;
    for (int k = 0; k < splits; ++k) {
      const double* d = part + (((size_t)k * N + n) * C + c) * 2;
      sg += d[0]; sgy += d[1];
    }
    const float y = __fdiv_rn(__fsub_rn(x[(size_t)pix * ldx + c], m), s);
    const double v = ((double)g[(size_t)pix * ldg + c] - sg / (double)HW - (double)y * sgy / (double)(HW - 1)) / (double)s;
    gx[(size_t)pix * ldgx + c] = (float)v;
  }
}

// ------------------------------------------------------------------------------------------------ resize backward
// adjoint of resize_bilinear_kernel, separable and in gather form (exact tap test with the forward's own tap
// computation, deterministic):  T[n,oy,ix] = sum_ox wx(ox,ix) g[n,oy,ox];  gin[n,iy,ix] = scale * sum_oy wy(oy,iy) T[n,oy,ix].
// (First version: one thread per INPUT pixel scanning its whole 2-D footprint -- at the 64x upsampling of the
// multi-scale distillation loss that is 416 threads walking 17k output pixels each, 0.86 ms per call.)
__device__ __forceinline__ void resize_bwd_range(int i, float sc, int n_out, int& lo, int& hi) {
  lo = 0; hi = n_out - 1;
  if (sc > 0.f) { lo = max(0, (int)floorf((i - 1) / sc) - 1); hi = min(n_out - 1, (int)ceilf((i + 1) / sc) + 1); }
}
__global__ void __launch_bounds__(256)
resize_bwd_x_kernel(const float* __restrict__ gout, int ldgo, int H, int W, float* __restrict__ tmp, int w, int N, int C, float sx) {
  const long long total = (long long)N * H * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(i % w);
    const long long row = i / w;                       // n * H + oy
    int lo, hi;
    resize_bwd_range(ix, sx, W, lo, hi);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ox = lo; ox <= hi; ++ox) {
      const AxisTap tx = axis_tap(ox, w, sx);
      const float wx = (tx.i0 == ix ? tx.l0 : 0.f) + (tx.i1 == ix ? tx.l1 : 0.f);
      if (wx == 0.f) continue;
      const float* gp = gout + ((size_t)row * W + ox) * ldgo;
      for (int c = 0; c < C; ++c) acc[c] = fmaf(gp[c], wx, acc[c]);
    }
    float* o = tmp + (size_t)i * C;
    for (int c = 0; c < C; ++c) o[c] = acc[c];
  }
}
__global__ void __launch_bounds__(256)
resize_bwd_y_kernel(const float* __restrict__ tmp, int H, float* __restrict__ gin, int ldgi, int h, int w, int N, int C, float sy,
                    float4 scale) {
  const long long total = (long long)N * h * w;
  const float sc[4] = {scale.x, scale.y, scale.z, scale.w};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ix = (int)(i % w);
    const long long t = i / w;
    const int iy = (int)(t % h), n = (int)(t / h);
    int lo, hi;
    resize_bwd_range(iy, sy, H, lo, hi);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int oy = lo; oy <= hi; ++oy) {
      const AxisTap ty = axis_tap(oy, h, sy);
      const float wy = (ty.i0 == iy ? ty.l0 : 0.f) + (ty.i1 == iy ? ty.l1 : 0.f);
      if (wy == 0.f) continue;
      const float* tp = tmp + ((size_t)((size_t)n * H + oy) * w + ix) * C;
      for (int c = 0; c < C; ++c) acc[c] = fmaf(tp[c], wy, acc[c]);
    }
    float* o = gin + (size_t)i * ldgi;
    for (int c = 0; c < C; ++c) o[c] = acc[c] * sc[c];
  }
}


// ------------------------------------------------------------------------------------------------ tensor-core wgrad
// Both operands of dW[tap][ci][co] = sum_p X[p (+) tap][ci] * G[p][co] are pixel-major, i.e. their K index (the pixel)
// is the slow one.  They are transposed once per convolution into PLANAR, ZERO-PADDED form
//     XT[ci][k], GT[co][k],  k = (n*(H+2pad) + y+pad)*(W+2pad) + x+pad,
// where a tap is a constant shift of k (the padding absorbs it: GT is zero there), and the nine tap GEMMs run as ONE
// launch of the tcgen05 kernel of conv_tc.cu with the taps as its "images" (M = Cin, N = Cout, K = k, cluster split-K).
// A TMA box row must start 16-byte aligned, so only the VERTICAL part of the shift ((ky-1)*dil rows of W+2pad rounded
// up to 4 elements) is applied to XT's coordinate; the horizontal part is baked into three copies of GT written at
// k + (kx-1)*dil.
__global__ void __launch_bounds__(256)
nhwc_to_planar_padded_kernel(const float* __restrict__ src, int ld, int C, float* __restrict__ dst, long long Kp, long long P,
                             int H, int W, int pad, int Wp, int shift, int ncopies, long long copy_stride, int copy_shift,
                             int rows) {
  __shared__ float tile[32][33];
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const long long p = p0 + i;
    const int c = c0 + tx;
    tile[i][tx] = (p < P && c < C) ? __ldg(src + (size_t)p * ld + c) : 0.f;
  }
  __syncthreads();
  const long long p = p0 + tx;
  if (p < P) {
    const int Hp = H + 2 * pad;
    const long long hw = (long long)H * W;
    const long long n = p / hw;
    const int r = (int)(p - n * hw);
    const int y = r / W, x = r - y * W;
    const long long k = (n * Hp + y + pad) * Wp + x + pad + shift;
    for (int i = ty; i < 32; i += 8) {
      const int c = c0 + i;
      if (c < C) {
        const float v = tile[tx][i];
        for (int q = 0; q < ncopies; ++q) {                // copy q lives at dst + q*copy_stride, shifted by q*copy_shift
          // BLOCKED planar [k / 32][row][k % 32], `rows` rows per k block, with the 16-byte chunk index XOR-ed by
          // (row & 7): the image of TMA's SWIZZLE_128B, so that a plain bulk copy of rows x 128 bytes lands the K-major
          // operand tile the tcgen05 descriptors expect (tile rows start at multiples of 8)
          const long long kk = k + q * copy_shift;
          const int j = (int)(kk & 31);
          dst[(size_t)q * copy_stride + ((size_t)(kk >> 5) * rows + c) * 32 + ((((j >> 2) ^ (c & 7)) << 2) | (j & 3))] = v;
        }
      }
    }
  }
}

int conv_tc_wgrad_gemm(const float* xt, int ldk, const float* gt_packed, const float* zero_bias, float* gw, int taps,
                       int Cin, int Cout, int K, const int* koffs, const int* wsel, int xt_rows, int kpad, long long gcopy, cudaStream_t st);
int wgrad_taps_gemm(const float* xt, int xt_rows, const float* gt, long long gcopy, int cout_pad, float* part, float* gw,
                    int Cin, int Cout, int kblocks, const int* koff3, cudaStream_t st, int* taken);
long long wgrad_taps_part_elems(int Cin, int Cout);


// ---- 3xTF32 support (engine precision "tf32x3"): an fp32 value x is hi + lo with hi = x truncated to TF32 (what
// tcgen05 kind::tf32 reads) and lo = x - hi (13 significant bits).  conv(x,w) ~ hi*whi + lo*whi + hi*wlo, three
// tensor-core passes accumulated before the activation.
// out = lrelu(t) (+ res);  out_lo = out - trunc_tf32(out)   (t == nullptr: only the split of `out`)
__global__ void __launch_bounds__(256)
act_split_kernel(const float* __restrict__ t, int ldt, const float* __restrict__ res, int ldr, float* __restrict__ out, int ldo,
                 float* __restrict__ lo, int ldlo, long long npix, int C, float slope) {
  const long long total = npix * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / C;
    const int c = (int)(i - p * C);
    float v;
    if (t) {
      v = lrelu(t[(size_t)p * ldt + c], slope);
      if (res) v += res[(size_t)p * ldr + c];
      out[(size_t)p * ldo + c] = v;
    } else {
      v = out[(size_t)p * ldo + c];
    }
    if (lo) lo[(size_t)p * ldlo + c] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  }
}

static unsigned grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > UPF_NUM_SMS * 16) b = UPF_NUM_SMS * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace upf

// number of pixel splits the weight gradient uses for this shape (the caller sizes the workspace with it)
// CTA tile (ci x co) by shape: 32 x 64 for the few-channel convolutions of the image pyramid and output_conv, whose
// pixel count is what is large (a 128 x 64 tile spent 5.9 ms of a 71 ms training step on ten such launches, 1/32 of
// its FMAs useful)
static int wgrad_bm(int Cin, int Cout) { return (Cin <= 32 && Cout <= 64) ? 32 : 128; }
static int wgrad_splits(int Cin, int Cout, int taps, long long npix) {
  const int bn = Cout > 64 ? 128 : 64, bm = wgrad_bm(Cin, Cout);
  const long long tiles = (long long)((Cin + bm - 1) / bm) * ((Cout + bn - 1) / bn) * taps;
  long long s = (4 * UPF_NUM_SMS + tiles - 1) / tiles;
  const long long max_by_pix = (npix + 255) / 256;
  if (s > max_by_pix) s = max_by_pix;
  if (s > 128) s = 128;
  if (s < 1) s = 1;
  return (int)s;
}

#define UPF_BIAS_SPLITS 1024
extern "C" long long upf_conv2d_wgrad_workspace_elems(int N, int H, int W, int Cin, int Cout, int ksize, int stride, int dilation) {
  const int pad = ((ksize - 1) * dilation) / 2;
  const int Ho = (H + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1;
  const int taps = ksize * ksize;
  const int s = wgrad_splits(Cin, Cout, taps, (long long)N * Ho * Wo);
  return (long long)s * ((long long)taps * Cin * Cout) + (long long)UPF_BIAS_SPLITS * Cout;
}

extern "C" int upf_conv2d_wgrad(const float* x, int ldx, const float* grad_out, int ldg, float* grad_w, float* grad_bias,
                                float* workspace, int N, int H, int W, int Cin, int Cout, int ksize, int stride,
                                int dilation, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && grad_out && grad_w && workspace, "wgrad: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "wgrad: bad shape");
  UPF_REQUIRE(ksize == 1 || ksize == 3, "wgrad: kernel size %d not in {1,3}", ksize);
  UPF_REQUIRE(stride >= 1 && dilation >= 1 && ldx >= Cin && ldg >= Cout, "wgrad: bad stride/dilation/pitch");
  const int pad = ((ksize - 1) * dilation) / 2;
  const int Ho = (H + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1, Wo = (W + 2 * pad - dilation * (ksize - 1) - 1) / stride + 1;
  const int taps = ksize * ksize;
  const long long npix = (long long)N * Ho * Wo;
  const int splits = wgrad_splits(Cin, Cout, taps, npix);
  const long long pps = ((npix + splits - 1) / splits + WG_KP - 1) / WG_KP * WG_KP;
  const int bn = Cout > 64 ? 128 : 64, bm = wgrad_bm(Cin, Cout);
  const int ci_tiles = (Cin + bm - 1) / bm, co_tiles = (Cout + bn - 1) / bn;
  cudaStream_t st = (cudaStream_t)stream;
  const long long wn = (long long)taps * Cin * Cout;
  const dim3 grid(ci_tiles * co_tiles, taps, splits);
  if (bn == 128)
    conv_wgrad_kernel<128, 128, 8, 8><<<grid, WG_NT, 0, st>>>(x, ldx, grad_out, ldg, workspace, N, H, W, Ho, Wo, Cin, Cout, ksize,
                                                              stride, dilation, pad, co_tiles, pps, npix);
  else if (bm == 32)
    conv_wgrad_kernel<32, 64, 4, 2><<<grid, WG_NT, 0, st>>>(x, ldx, grad_out, ldg, workspace, N, H, W, Ho, Wo, Cin, Cout, ksize,
                                                            stride, dilation, pad, co_tiles, pps, npix);
  else
    conv_wgrad_kernel<128, 64, 8, 4><<<grid, WG_NT, 0, st>>>(x, ldx, grad_out, ldg, workspace, N, H, W, Ho, Wo, Cin, Cout, ksize,
                                                             stride, dilation, pad, co_tiles, pps, npix);
  int e = check_launch("conv_wgrad");
  if (e) return e;
  reduce_splits_kernel<<<grid_for(wn), 256, 0, st>>>(workspace, grad_w, wn, splits);
  e = check_launch("wgrad_reduce");
  if (e) return e;
  if (grad_bias) {
    float* bpart = workspace + (size_t)splits * wn;
    int cw = 1;
    while (cw < Cout && cw < 256) cw <<= 1;               // channel lanes: power of two <= 256
    int bs = (int)((npix + 255) / 256);                   // its own pixel split: the weight-gradient split can be 1
    if (bs > UPF_BIAS_SPLITS) bs = UPF_BIAS_SPLITS;
    if (bs < 1) bs = 1;
    const long long bpps = (npix + bs - 1) / bs;
    colsum_kernel<<<bs, 256, 0, st>>>(grad_out, ldg, bpart, Cout, cw, bpps, npix);
    e = check_launch("bias_colsum");
    if (e) return e;
    reduce_splits_kernel<<<1, 256, 0, st>>>(bpart, grad_bias, Cout, bs);
    e = check_launch("bias_reduce");
  }
  return e;
}

extern "C" int upf_pointwise(int op, const float* a, int lda, const float* b, int ldb, float* out, int ldo, long long npix,
                             int C, float slope, void* stream) {
  using namespace upf;
  UPF_REQUIRE(a && out && (b || op == PW_SIGMOID), "pointwise: null tensor");
  UPF_REQUIRE(op >= 0 && op <= 2 && npix > 0 && C > 0 && lda >= C && ldo >= C, "pointwise: bad argument");
  pointwise_kernel<<<grid_for(npix * C), 256, 0, (cudaStream_t)stream>>>(op, a, lda, b, ldb, out, ldo, npix, C, slope);
  return check_launch("pointwise");
}

extern "C" int upf_blend_fwd(const float* w, int ldw, const float* f, int ldf, const float* m, int ldm, float* out, int ldo,
                             long long npix, void* stream) {
  using namespace upf;
  UPF_REQUIRE(w && f && m && out && npix > 0, "blend_fwd: bad argument");
  blend_fwd_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(w, ldw, f, ldf, m, ldm, out, ldo, npix);
  return check_launch("blend_fwd");
}
extern "C" int upf_blend_bwd(const float* w, int ldw, const float* f, int ldf, const float* m, int ldm, const float* g, int ldg,
                             float* gw, int ldgw, float* gf, int ldgf, float* gm, int ldgm, long long npix, void* stream) {
  using namespace upf;
  UPF_REQUIRE(w && f && m && g && gw && gf && gm && npix > 0, "blend_bwd: bad argument");
  blend_bwd_kernel<<<grid_for(npix), 256, 0, (cudaStream_t)stream>>>(w, ldw, f, ldf, m, ldm, g, ldg, gw, ldgw, gf, ldgf, gm, ldgm, npix);
  return check_launch("blend_bwd");
}

#define UPF_FEATNORM_BWD_SPLITS 32
extern "C" long long upf_featnorm_bwd_workspace_doubles(int N, int C) { return (long long)UPF_FEATNORM_BWD_SPLITS * N * C * 2; }
extern "C" int upf_featnorm_bwd(const float* x, int ldx, const double* stats, const float* grad_out, int ldg, float* grad_x,
                                int ldgx, double* workspace, int N, int H, int W, int C, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && stats && grad_out && grad_x && workspace, "featnorm_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && ldg >= C && ldgx >= C, "featnorm_bwd: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = UPF_FEATNORM_BWD_SPLITS;
  int cw = 1;
  while (cw < C && cw < 256) cw <<= 1;
  featnorm_bwd_sums_kernel<<<N * splits, 256, 0, st>>>(x, ldx, stats, grad_out, ldg, workspace, N, H * W, C, splits, cw);
  int e = check_launch("featnorm_bwd_sums");
  if (e) return e;
  featnorm_bwd_apply_kernel<<<grid_for((long long)N * H * W * C), 256, 0, st>>>(x, ldx, stats, grad_out, ldg, workspace, grad_x, ldgx,
                                                                               N, H * W, C, splits);
  return check_launch("featnorm_bwd_apply");
}

extern "C" long long upf_resize_bilinear_bwd_workspace_elems(int N, int H, int w, int C) { return (long long)N * H * w * C; }
extern "C" int upf_resize_bilinear_bwd(const float* grad_out, int ldgo, int H, int W, float* grad_in, int ldgi, int h, int w,
                                       int N, int C, const float* scale_host, float* workspace, void* stream) {
  using namespace upf;
  UPF_REQUIRE(grad_out && grad_in && workspace, "resize_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && h > 0 && w > 0 && C > 0 && C <= 4 && ldgo >= C && ldgi >= C, "resize_bwd: bad shape");
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
  if (scale_host) sc = make_float4(scale_host[0], scale_host[1], scale_host[2], scale_host[3]);
  cudaStream_t st = (cudaStream_t)stream;
  resize_bwd_x_kernel<<<grid_for((long long)N * H * w), 256, 0, st>>>(grad_out, ldgo, H, W, workspace, w, N, C, host_ac_scale(w, W));
  int e = check_launch("resize_bwd_x");
  if (e) return e;
  resize_bwd_y_kernel<<<grid_for((long long)N * h * w), 256, 0, st>>>(workspace, H, grad_in, ldgi, h, w, N, C, host_ac_scale(h, H), sc);
  return check_launch("resize_bwd_y");
}

// ---- nn.Conv2d weight [A][B][k][k] -> the library's [k*k][Cin][cout_pad] (cout_pad = Cout rounded up to 4, padding
// zero), one launch per convolution call of the training path (was: fill + permute copy + slice copy, and for the
// input-gradient convolution a flip and a transposing copy before those).
// flip_transpose 0: the forward convolution, Cout = A, Cin = B:          out[t][ci][co] = w[co][ci][t]
// flip_transpose 1: the input-gradient convolution, Cin = A, Cout = B:  out[t][ci][co] = w[ci][co][taps-1-t]
__global__ void __launch_bounds__(256)
repack_weight_kernel(const float* __restrict__ w, float* __restrict__ out, int Cin, int Cout, int cout_pad, int taps, int flip_transpose) {
  const long long total = (long long)taps * Cin * cout_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % cout_pad);
    const long long r = i / cout_pad;
    const int ci = (int)(r % Cin), t = (int)(r / Cin);
    float v = 0.f;
    if (co < Cout)
      v = flip_transpose ? __ldg(w + ((size_t)ci * Cout + co) * taps + (taps - 1 - t)) : __ldg(w + ((size_t)co * Cin + ci) * taps + t);
    out[i] = v;
  }
}

extern "C" int upf_repack_conv_weight(const float* weight, float* out, int A, int B, int ksize, int flip_transpose, void* stream) {
  using namespace upf;
  UPF_REQUIRE(weight && out && A > 0 && B > 0 && ksize > 0, "repack_conv_weight: bad argument");
  const int Cin = flip_transpose ? A : B, Cout = flip_transpose ? B : A, taps = ksize * ksize;
  const int cout_pad = (Cout + 3) / 4 * 4;
  repack_weight_kernel<<<grid_for((long long)taps * Cin * cout_pad), 256, 0, (cudaStream_t)stream>>>(weight, out, Cin, Cout, cout_pad,
                                                                                                   taps, flip_transpose ? 1 : 0);
  return check_launch("repack_weight");
}

// ---- tensor-core weight gradient (stride 1): see nhwc_to_planar_padded_kernel
// (Tried: cutting K into up to 8 grid-level parts on top of the cluster split for the few-channel convolutions at fine
// resolution, 72 -> 576 CTAs, partials summed by reduce_splits_kernel -- parity-green, training step 57.19 vs 57.00 ms:
// those launches are not bound by the GEMM's CTA count; removed.)
// padded row: a multiple of 32, so that a tap's vertical shift (dil * Wp pixels) is a whole number of 32-pixel k blocks
static int wgrad_tc_wp(int W, int ks, int dil) { return ks == 1 ? W : (W + (ks - 1) * dil + 31) / 32 * 32; }
static long long wgrad_tc_kp(int N, int H, int W, int ks, int dil) {
  const int pad = ((ks - 1) * dil) / 2;
  const long long K = (long long)N * (H + 2 * pad) * wgrad_tc_wp(W, ks, dil);
  return (K + 31) / 32 * 32;
}
// zero k blocks in front of / behind the blocked XT buffer: a tap's vertical shift, dil * Wp pixels
static int wgrad_tc_kpad(int W, int ks, int dil) { return ks == 1 ? 0 : dil * wgrad_tc_wp(W, ks, dil) / 32; }
constexpr long long WGRAD_SLACK = 128 * 32;      // a 128-row tile may start at the buffer's last rows: one tile of slack
constexpr int WGRAD_TAIL = 7;                    // zero k blocks behind the last one: a ring slot of the GEMM holds up to 8 blocks
static long long wgrad_tc_xt_elems(int N, int H, int W, int C, int ks, int dil) {
  return (wgrad_tc_kp(N, H, W, ks, dil) / 32 + 2 * wgrad_tc_kpad(W, ks, dil) + WGRAD_TAIL) * (long long)C * 32 + WGRAD_SLACK;
}
// one pre-shifted copy of GT: [kpad zero blocks][K / 32 blocks][kpad + tail zero blocks] x cout_pad rows x 32
static long long wgrad_tc_gt_copy_elems(int N, int H, int W, long long cout_pad, int ks, int dil) {
  return (wgrad_tc_kp(N, H, W, ks, dil) / 32 + 2 * wgrad_tc_kpad(W, ks, dil) + WGRAD_TAIL) * cout_pad * 32;
}
extern "C" long long upf_wgrad_tc_planar_elems(int N, int H, int W, int C, int ksize, int dilation) {
  return wgrad_tc_xt_elems(N, H, W, C, ksize, dilation);
}
extern "C" long long upf_conv2d_wgrad_tc_workspace_elems(int N, int H, int W, int Cin, int Cout, int ksize, int dilation) {
  const long long Kp = wgrad_tc_kp(N, H, W, ksize, dilation);
  const long long cout_pad = (Cout + 15) / 16 * 16;
  return wgrad_tc_xt_elems(N, H, W, Cin, ksize, dilation) + 3 * wgrad_tc_gt_copy_elems(N, H, W, cout_pad, ksize, dilation) +
         WGRAD_SLACK + cout_pad + (long long)UPF_BIAS_SPLITS * Cout + (ksize == 3 ? upf::wgrad_taps_part_elems(Cin, Cout) : 0) + 64;
}
// x != NULL: transpose the input here; xt_pre != NULL: the caller already holds the planar padded input (rows of
// upf_wgrad_tc_transpose_input's output -- a dense block transposes its whole buffer ONCE and every convolution of the
// block reads its own suffix of rows)
static int wgrad_tc_impl(const float* x, int ldx, const float* xt_pre, int xt_rows, int row0, const float* grad_out, int ldg, float* grad_w,
                         float* grad_bias, float* workspace, int N, int H, int W, int Cin, int Cout, int ksize, int dilation,
                         cudaStream_t st) {
  using namespace upf;
  UPF_REQUIRE((x || xt_pre) && grad_out && grad_w && workspace, "wgrad_tc: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (ksize == 1 || ksize == 3) && dilation >= 1, "wgrad_tc: bad shape");
  UPF_REQUIRE((xt_pre || ldx >= Cin) && ldg >= Cout && aligned16(workspace) && (!xt_pre || aligned16(xt_pre)),
              "wgrad_tc: bad pitch / workspace alignment");
  const int pad = ((ksize - 1) * dilation) / 2, taps = ksize * ksize;
  const int Wp = wgrad_tc_wp(W, ksize, dilation);
  const long long Kp = wgrad_tc_kp(N, H, W, ksize, dilation);
  UPF_REQUIRE(Kp < (1ll << 31), "wgrad_tc: padded pixel count too large");
  const long long cout_pad = (Cout + 15) / 16 * 16;
  const long long P = (long long)N * H * W;
  const int ncopies = ksize == 1 ? 1 : 3;
  const int kpad = wgrad_tc_kpad(W, ksize, dilation);
  UPF_REQUIRE(!xt_pre || (row0 % 8) == 0, "wgrad_tc_planar: row0 must be a multiple of 8 (swizzle phase of the operand tiles)");
  float* xt = workspace;
  float* gt = xt + (xt_pre ? 0 : (size_t)wgrad_tc_xt_elems(N, H, W, Cin, ksize, dilation));   // [3][Kp / 32][cout_pad][32]: G written at k + (kx-1)*dil
  const long long gcopy = wgrad_tc_gt_copy_elems(N, H, W, cout_pad, ksize, dilation);
  float* gt0 = gt + (size_t)kpad * cout_pad * 32;           // block 0 of copy 0 (kpad zero blocks in front of every copy)
  float* zb = gt + (size_t)3 * gcopy + WGRAD_SLACK;
  float* bpart = zb + cout_pad;
  float* tpart = bpart + (size_t)UPF_BIAS_SPLITS * Cout;    // partial sums of the taps-along-N kernel (Cout <= 32)
  cudaError_t ce = cudaMemsetAsync(workspace, 0, (size_t)((char*)(zb + cout_pad) - (char*)workspace), st);
  if (ce != cudaSuccess) { set_error("wgrad_tc memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  const unsigned ptiles = (unsigned)((P + 31) / 32);
  int e = 0;
  if (!xt_pre) {
    nhwc_to_planar_padded_kernel<<<dim3(ptiles, (Cin + 31) / 32), 256, 0, st>>>(x, ldx, Cin, xt + (size_t)kpad * Cin * 32, Kp, P, H, W,
                                                                                pad, Wp, 0, 1, 0, 0, Cin);
    e = check_launch("wgrad_tc_transpose_x");
    if (e) return e;
  }
  // the three horizontally shifted copies of G in one pass: one read, three writes
  nhwc_to_planar_padded_kernel<<<dim3(ptiles, (Cout + 31) / 32), 256, 0, st>>>(grad_out, ldg, Cout, gt0, Kp, P, H, W, pad, Wp,
                                                                                ksize == 1 ? 0 : -dilation, ncopies,
                                                                                gcopy, dilation, (int)cout_pad);
  e = check_launch("wgrad_tc_transpose_g");
  if (e) return e;
  int koffs[9], wsel[9];
  for (int t = 0; t < taps; ++t) {
    const int ky = t / ksize, kx = t % ksize;
    koffs[t] = ksize == 1 ? 0 : (ky - 1) * dilation * (Wp / 32);   // in k blocks of 32 (Wp is a multiple of 32)
    wsel[t] = ksize == 1 ? 0 : kx;
  }
  {
    // few output channels: the nine taps along N, the input tile loaded once (wgrad_taps.cu)
    int taken = 0;
    if (ksize == 3) {
      const int koff3[3] = {-dilation * (Wp / 32), 0, dilation * (Wp / 32)};
      const float* xt0 = xt_pre ? xt_pre + (size_t)row0 * 32 + (size_t)kpad * xt_rows * 32 : xt + (size_t)kpad * Cin * 32;
      e = wgrad_taps_gemm(xt0, xt_pre ? xt_rows : Cin, gt0, gcopy, (int)cout_pad, tpart, grad_w, Cin, Cout, (int)(Kp / 32), koff3, st, &taken);
      if (e) return e;
    }
    if (!taken) {
      e = conv_tc_wgrad_gemm(xt_pre ? xt_pre + (size_t)row0 * 32 : xt, (int)Kp, gt0, zb, grad_w, taps, Cin, Cout, (int)Kp, koffs, wsel,
                             xt_pre ? xt_rows : Cin, kpad, gcopy, st);
      if (e) return e;
    }
  }
  if (grad_bias) {
    int cw = 1;
    while (cw < Cout && cw < 256) cw <<= 1;
    int bs = (int)((P + 255) / 256);
    if (bs > UPF_BIAS_SPLITS) bs = UPF_BIAS_SPLITS;
    const long long bpps = (P + bs - 1) / bs;
    colsum_kernel<<<bs, 256, 0, st>>>(grad_out, ldg, bpart, Cout, cw, bpps, P);
    e = check_launch("bias_colsum");
    if (e) return e;
    reduce_splits_kernel<<<1, 256, 0, st>>>(bpart, grad_bias, Cout, bs);
    e = check_launch("bias_reduce");
  }
  return e;
}

extern "C" int upf_conv2d_wgrad_tc(const float* x, int ldx, const float* grad_out, int ldg, float* grad_w, float* grad_bias,
                                   float* workspace, int N, int H, int W, int Cin, int Cout, int ksize, int dilation,
                                   void* stream) {
  UPF_REQUIRE(x, "wgrad_tc: null input");
  return wgrad_tc_impl(x, ldx, nullptr, 0, 0, grad_out, ldg, grad_w, grad_bias, workspace, N, H, W, Cin, Cout, ksize, dilation,
                       (cudaStream_t)stream);
}

extern "C" long long upf_wgrad_tc_planar_pitch(int N, int H, int W, int ksize, int dilation) {
  return wgrad_tc_kp(N, H, W, ksize, dilation);
}

extern "C" int upf_wgrad_tc_transpose_input(const float* x, int ldx, int C, float* xt, int N, int H, int W, int ksize,
                                            int dilation, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && xt && N > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && (ksize == 1 || ksize == 3) && dilation >= 1 &&
                  aligned16(xt), "wgrad_tc_transpose_input: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = ((ksize - 1) * dilation) / 2;
  const long long Kp = wgrad_tc_kp(N, H, W, ksize, dilation), P = (long long)N * H * W;
  cudaError_t ce = cudaMemsetAsync(xt, 0, (size_t)wgrad_tc_xt_elems(N, H, W, C, ksize, dilation) * sizeof(float), st);
  if (ce != cudaSuccess) { set_error("wgrad_tc_transpose_input memset: %s", cudaGetErrorString(ce)); return (int)ce; }
  nhwc_to_planar_padded_kernel<<<dim3((unsigned)((P + 31) / 32), (C + 31) / 32), 256, 0, st>>>(
      x, ldx, C, xt + (size_t)wgrad_tc_kpad(W, ksize, dilation) * C * 32, Kp, P, H, W, pad, wgrad_tc_wp(W, ksize, dilation), 0, 1, 0, 0, C);
  return check_launch("wgrad_tc_transpose_x");
}

extern "C" int upf_conv2d_wgrad_tc_planar(const float* xt, int xt_rows, int row0, const float* grad_out, int ldg, float* grad_w,
                                          float* grad_bias, float* workspace, int N, int H, int W, int Cin, int Cout, int ksize,
                                          int dilation, void* stream) {
  UPF_REQUIRE(xt, "wgrad_tc_planar: null input");
  UPF_REQUIRE(row0 >= 0 && Cin > 0 && row0 + Cin <= xt_rows, "wgrad_tc_planar: rows [%d, %d) outside the %d-row buffer", row0, row0 + Cin, xt_rows);
  return wgrad_tc_impl(nullptr, 0, xt, xt_rows, row0, grad_out, ldg, grad_w, grad_bias, workspace, N, H, W, Cin, Cout, ksize, dilation,
                       (cudaStream_t)stream);
}

extern "C" int upf_act_split(const float* t, int ldt, const float* residual, int ldr, float* out, int ldo, float* out_lo, int ldlo,
                             long long npix, int C, float slope, void* stream) {
  using namespace upf;
  UPF_REQUIRE(out && npix > 0 && C > 0 && ldo >= C && (!t || ldt >= C) && (!out_lo || ldlo >= C) && (!residual || ldr >= C),
              "act_split: bad argument");
  act_split_kernel<<<grid_for(npix * C), 256, 0, (cudaStream_t)stream>>>(t, ldt, residual, ldr, out, ldo, out_lo, ldlo, npix, C, slope);
  return check_launch("act_split");
}
