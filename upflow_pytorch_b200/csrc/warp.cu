// Bilinear warp + validity mask (+ per-channel moments of the result), its
// backward, and the per-image per-channel feature normalisation.
//
// Replaces WarpingLayer_no_div.forward (model/pwc_modules.py:184-207: host-built
// mesh + host-built ones(B,C,H,W) + two grid_sample passes + compare + multiply)
// and tools.torch_warp (utils/tools.py:1274-1304) with ONE pass: the four
// gathers are 16-byte loads of pixel-major rows, the validity mask is computed
// once per pixel (not per channel), and the per-(image,channel) sum / sum of
// squares that normalize_features (model/upflow.py:94-137) needs afterwards
// are reduced on the fly (registers -> warp shuffles -> one shared-memory slot per
// warp -> one double atomic per channel per CTA).
#include "upf_common.cuh"

namespace upf {

constexpr int WARP_NT = 256;

// CTA reduction of the per-thread fp32 partial moments, in double and in a FIXED order: xor-shuffle tree over the
// lanes of a warp that own the same channels, one shared-memory slot per warp (no shared-memory atomics: 4096
// contended double CAS loops per CTA cost more than the warp itself -- 39 vs 16 us at 2x32x94x311), warps summed in
// index order, then one global atomicAdd per channel and CTA (exact in double for fp32-valued partials, so the
// result does not depend on the order the CTAs arrive in).   s_red: [WARP_NT/32][2][4*cgroups] doubles.
__device__ __forceinline__ void cta_reduce_moments(const float (&sm)[2][4], const float (&sq)[2][4], int lpp, int sub, int passes,
                                                   int C, int cgroups, double* s_red, double* stats_n) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nch = cgroups * 4;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int c = (pass * lpp + sub) * 4;
    const bool active = pass < passes && c < C;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double a = active ? (double)sm[pass][k] : 0.0, b = active ? (double)sq[pass][k] : 0.0;
      for (int off = lpp; off < 32; off <<= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        b += __shfl_xor_sync(0xffffffffu, b, off);
      }
      if (active && lane < lpp) {
        s_red[(size_t)(warp * 2 + 0) * nch + c + k] = a;
        s_red[(size_t)(warp * 2 + 1) * nch + c + k] = b;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += WARP_NT) {
    const int which = i / C, c = i - which * C;
    double t = 0.0;
    for (int w = 0; w < WARP_NT / 32; ++w) t += s_red[(size_t)(w * 2 + which) * nch + c];
    atomicAdd(&stats_n[(size_t)c * 2 + which], t);
  }
}

// thread layout: `lpp` lanes per pixel (each lane owns 4 consecutive channels
// per pass), WARP_NT/lpp pixels per CTA step; a CTA strides over the pixels of
// ONE image so the moment reduction stays per image.
template <bool VEC, typename T = float>
__global__ void __launch_bounds__(WARP_NT)
warp_fwd_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ flow, int ldf,
                T* __restrict__ out, int ldo, int H, int W, int C, int align_corners, float mask_thr,
                double* __restrict__ stats, int lpp, int ctas_per_image, int x_shift, int N, int oflags) {
  pdl_prologue();
  extern __shared__ double s_red[];   // [2][cgroups*4] when stats
  const int n = blockIdx.x / ctas_per_image;
  const int cta = blockIdx.x - n * ctas_per_image;
  const int ppc = WARP_NT / lpp;                 // pixels per CTA step
  const int sub = threadIdx.x % lpp;             // which 4-channel group (first pass)
  const int pslot = threadIdx.x / lpp;
  const int cgroups = (C + 3) >> 2;
  const int passes = (cgroups + lpp - 1) / lpp;
  const int npix = H * W;
  const size_t img = (size_t)n * npix;
  const long long ximg = (long long)((n + x_shift) % N) * npix;

  // per-thread moment accumulators for up to 2 passes (C <= 8*lpp <= 256)
  float sm[2][4], sq[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int t = 0; t < 4; ++t) sm[a][t] = sq[a][t] = 0.f;

  // The sampling coordinates, the four weights and the mask are a per-PIXEL computation (two IEEE divisions among
  // ~100 instructions): one lane of the lpp lanes that share a pixel does it and broadcasts seven words.  (Every lane
  // recomputing it made the kernel instruction-bound: 8x redundant at C=32, 32x at C=128.)
  const int lane = threadIdx.x & 31;
  const int leader = lane - (lpp < 32 ? sub : lane);       // lane of this pixel's sub == 0 inside the warp
  for (int base = cta * ppc; base < npix; base += ctas_per_image * ppc) {   // uniform trip count per CTA (shuffles below)
    const int p_raw = base + pslot;
    const bool live = p_raw < npix;
    const int p = live ? p_raw : npix - 1;
    BilinearTaps t;
    int flags = 0;
    if (lane == leader) {
      const int y = p / W, xpix = p - y * W;
      const float* fl = flow + (img + p) * ldf;
      const float u = __ldg(fl), v = __ldg(fl + 1);
      const float ix = sample_coord((float)xpix, u, W, align_corners);
      const float iy = sample_coord((float)y, v, H, align_corners);
      t = bilinear_taps(ix, iy, H, W);
      const bool k = !(mask_thr > 0.f) || t.wsum >= mask_thr;      // mask = (grid_sample(ones) >= 1.0), pwc_modules.py:205-206
      flags = (t.in_nw ? 1 : 0) | (t.in_ne ? 2 : 0) | (t.in_sw ? 4 : 0) | (t.in_se ? 8 : 0) | (k ? 16 : 0);
    }
    t.x0 = __shfl_sync(0xffffffffu, t.x0, leader);
    t.y0 = __shfl_sync(0xffffffffu, t.y0, leader);
    t.w_nw = __shfl_sync(0xffffffffu, t.w_nw, leader);
    t.w_ne = __shfl_sync(0xffffffffu, t.w_ne, leader);
    t.w_sw = __shfl_sync(0xffffffffu, t.w_sw, leader);
    t.w_se = __shfl_sync(0xffffffffu, t.w_se, leader);
    flags = __shfl_sync(0xffffffffu, flags, leader);
    t.in_nw = flags & 1; t.in_ne = flags & 2; t.in_sw = flags & 4; t.in_se = flags & 8;
    const bool keep = (flags & 16) != 0;
    if (!live) continue;
    const T* r_nw = x + (ximg + (long long)t.y0 * W + t.x0) * (long long)ldx;   // may point outside: only dereferenced when in_*
    const T* r_ne = r_nw + ldx;
    const T* r_sw = r_nw + (size_t)W * ldx;
    const T* r_se = r_sw + ldx;
    T* o = out + (img + p) * ldo;
#pragma unroll 2
    for (int pass = 0; pass < passes; ++pass) {
      const int c = (pass * lpp + sub) * 4;
      if (c >= C) break;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (keep) {
        // same accumulation order as ATen's grid_sampler_2d: nw, ne, sw, se
        if (VEC) {
          if (t.in_nw) { float4 q = ld4(r_nw + c); acc.x = fmaf(q.x, t.w_nw, acc.x); acc.y = fmaf(q.y, t.w_nw, acc.y); acc.z = fmaf(q.z, t.w_nw, acc.z); acc.w = fmaf(q.w, t.w_nw, acc.w); }
          if (t.in_ne) { float4 q = ld4(r_ne + c); acc.x = fmaf(q.x, t.w_ne, acc.x); acc.y = fmaf(q.y, t.w_ne, acc.y); acc.z = fmaf(q.z, t.w_ne, acc.z); acc.w = fmaf(q.w, t.w_ne, acc.w); }
          if (t.in_sw) { float4 q = ld4(r_sw + c); acc.x = fmaf(q.x, t.w_sw, acc.x); acc.y = fmaf(q.y, t.w_sw, acc.y); acc.z = fmaf(q.z, t.w_sw, acc.z); acc.w = fmaf(q.w, t.w_sw, acc.w); }
          if (t.in_se) { float4 q = ld4(r_se + c); acc.x = fmaf(q.x, t.w_se, acc.x); acc.y = fmaf(q.y, t.w_se, acc.y); acc.z = fmaf(q.z, t.w_se, acc.z); acc.w = fmaf(q.w, t.w_se, acc.w); }
        } else {
          float* a = &acc.x;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c + k < C) {
              float s = 0.f;
              if (t.in_nw) s = fmaf(ld1(r_nw + c + k), t.w_nw, s);
              if (t.in_ne) s = fmaf(ld1(r_ne + c + k), t.w_ne, s);
              if (t.in_sw) s = fmaf(ld1(r_sw + c + k), t.w_sw, s);
              if (t.in_se) s = fmaf(ld1(r_se + c + k), t.w_se, s);
              a[k] = s;
            }
        }
      }
      if (oflags & UPF_FLAG_ROUND_TF32) { acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w); }
      if (VEC) {
        st4(o + c, acc);
      } else {
        const float* a = &acc.x;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c + k < C) st1(o + c + k, a[k]);
      }
      if (stats && pass < 2) {
        sm[pass][0] += acc.x; sm[pass][1] += acc.y; sm[pass][2] += acc.z; sm[pass][3] += acc.w;
        sq[pass][0] = fmaf(acc.x, acc.x, sq[pass][0]); sq[pass][1] = fmaf(acc.y, acc.y, sq[pass][1]);
        sq[pass][2] = fmaf(acc.z, acc.z, sq[pass][2]); sq[pass][3] = fmaf(acc.w, acc.w, sq[pass][3]);
      }
    }
  }

  if (stats) cta_reduce_moments(sm, sq, lpp, sub, passes, C, cgroups, s_red, stats + (size_t)n * C * 2);
}

// moments of an existing tensor (same reduction skeleton, no sampling)
__global__ void __launch_bounds__(WARP_NT)
featnorm_stats_kernel(const float* __restrict__ x, int ldx, int H, int W, int C, double* __restrict__ stats,
                      int lpp, int ctas_per_image, int vec) {
  pdl_prologue();
  extern __shared__ double s_red[];
  const int n = blockIdx.x / ctas_per_image;
  const int cta = blockIdx.x - n * ctas_per_image;
  const int ppc = WARP_NT / lpp;
  const int sub = threadIdx.x % lpp, pslot = threadIdx.x / lpp;
  const int cgroups = (C + 3) >> 2;
  const int passes = (cgroups + lpp - 1) / lpp;
  const int npix = H * W;
  const size_t img = (size_t)n * npix;
  float sm[2][4], sq[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int t = 0; t < 4; ++t) sm[a][t] = sq[a][t] = 0.f;
  // fp32 partial sums run over at most ~npix/(ctas*ppc) values before they are
  // widened to double, which keeps the moments accurate to ~1e-7 relative
  for (int p = cta * ppc + pslot; p < npix; p += ctas_per_image * ppc) {
    const float* r = x + (img + p) * ldx;
#pragma unroll 2
    for (int pass = 0; pass < passes && pass < 2; ++pass) {
      const int c = (pass * lpp + sub) * 4;
      if (c >= C) break;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) { float4 q = ldg4(r + c); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) if (c + k < C) v[k] = __ldg(r + c + k);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { sm[pass][k] += v[k]; sq[pass][k] = fmaf(v[k], v[k], sq[pass][k]); }
    }
  }
  cta_reduce_moments(sm, sq, lpp, sub, passes, C, cgroups, s_red, stats + (size_t)n * C * 2);
}

__global__ void __launch_bounds__(256)
featnorm_apply_kernel(const float* __restrict__ x, int ldx, const double* __restrict__ stats,
                      float* __restrict__ out, int ldo, int H, int W, int C, long long total) {
  pdl_prologue();
  const double npix = (double)H * (double)W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long pix = i / C;
    const int n = (int)(pix / ((long long)H * W));
    float m, s;
    stats_to_mean_std(stats + ((size_t)n * C + c) * 2, npix, m, s);
    out[(size_t)pix * ldo + c] = __fdiv_rn(__fsub_rn(__ldg(x + (size_t)pix * ldx + c), m), s);
  }
}

// ---- backward of the warp: d/dx (scatter-add) and d/dflow -----------------
// ATen grid_sampler_2d_backward semantics: the mask (a comparison) carries no
// gradient; masked pixels propagate nothing.
__global__ void __launch_bounds__(256)
warp_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ flow, int ldf,
                const float* __restrict__ go, int ldg, float* __restrict__ gx, int ldgx,
                float* __restrict__ gflow, int ldgf, int N, int H, int W, int C, int align_corners, float mask_thr) {
  pdl_prologue();
  // one warp per pixel, lanes stride over channels
  const long long npix = (long long)N * H * W;
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long p = warp0; p < npix; p += nwarps) {
    const int xpix = (int)(p % W);
    const int y = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    const float* fl = flow + (size_t)p * ldf;
    const float ix = sample_coord((float)xpix, __ldg(fl), W, align_corners);
    const float iy = sample_coord((float)y, __ldg(fl + 1), H, align_corners);
    const BilinearTaps t = bilinear_taps(ix, iy, H, W);
    const bool keep = !(mask_thr > 0.f) || t.wsum >= mask_thr;
    float gix = 0.f, giy = 0.f;
    if (keep) {
      const long long base = ((long long)n * H + t.y0) * W + t.x0;   // corners are only touched when in_*
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const float ax1 = (fx0 + 1.f) - ix, ax0 = ix - fx0, ay1 = (fy0 + 1.f) - iy, ay0 = iy - fy0;
      for (int c = lane; c < C; c += 32) {
        const float g = __ldg(go + (size_t)p * ldg + c);
        if (gx) {
          if (t.in_nw) atomicAdd(gx + base * ldgx + c, g * t.w_nw);
          if (t.in_ne) atomicAdd(gx + (base + 1) * ldgx + c, g * t.w_ne);
          if (t.in_sw) atomicAdd(gx + (base + W) * ldgx + c, g * t.w_sw);
          if (t.in_se) atomicAdd(gx + (base + W + 1) * ldgx + c, g * t.w_se);
        }
        if (gflow) {
          const float v_nw = t.in_nw ? __ldg(x + base * ldx + c) : 0.f;
          const float v_ne = t.in_ne ? __ldg(x + (base + 1) * ldx + c) : 0.f;
          const float v_sw = t.in_sw ? __ldg(x + (base + W) * ldx + c) : 0.f;
          const float v_se = t.in_se ? __ldg(x + (base + W + 1) * ldx + c) : 0.f;
          gix += g * ((v_ne - v_nw) * ay1 + (v_se - v_sw) * ay0);
          giy += g * ((v_sw - v_nw) * ax1 + (v_se - v_ne) * ax0);
        }
      }
    }
    if (gflow) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        gix += __shfl_xor_sync(0xffffffffu, gix, o);
        giy += __shfl_xor_sync(0xffffffffu, giy, o);
      }
      if (lane == 0) {
        // d ix / d u: pixel->[-1,1] is 2/(W-1), unnormalise is W/2 (or (W-1)/2)
        const float sx = align_corners ? 1.0f : (float)W / (float)(W > 1 ? W - 1 : 1);
        const float sy = align_corners ? 1.0f : (float)H / (float)(H > 1 ? H - 1 : 1);
        gflow[(size_t)p * ldgf + 0] = gix * sx;
        gflow[(size_t)p * ldgf + 1] = giy * sy;
      }
    }
  }
}


// Vector form (C % 4 == 0, 16-byte aligned rows): lpp lanes share a pixel, a lane owns channel quads -- one 16-byte load of
// the gradient and ONE 16-byte reduction (red.global.add.v4.f32) per corner instead of four scalar atomics, 32/lpp pixels per
// warp; the per-pixel taps are computed by one lane of the group and broadcast, like the forward kernel.
template <bool VEC>
__global__ void __launch_bounds__(256)
warp_bwd_vec_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ flow, int ldf,
                    const float* __restrict__ go, int ldg, float* __restrict__ gx, int ldgx,
                    float* __restrict__ gflow, int ldgf, int N, int H, int W, int C, int align_corners, float mask_thr, int lpp) {
  pdl_prologue();
  const long long npix = (long long)N * H * W;
  const int lane = threadIdx.x & 31;
  const int ppw = 32 / lpp;                                   // pixels per warp
  const int sub = lane % lpp;
  const int leader = lane - sub;
  const int quads = VEC ? C >> 2 : C;                         // work items per pixel: channel quads, or single channels
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pb = warp0 * ppw; pb < npix; pb += nwarps * ppw) {   // warp-uniform trip count (shuffles below)
    const long long p_raw = pb + lane / lpp;
    const bool live = p_raw < npix;
    const long long p = live ? p_raw : npix - 1;
    const int xpix = (int)(p % W);
    const int y = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    BilinearTaps t;
    float ix = 0.f, iy = 0.f;
    int flags = 0;
    if (lane == leader) {
      const float* fl = flow + (size_t)p * ldf;
      ix = sample_coord((float)xpix, __ldg(fl), W, align_corners);
      iy = sample_coord((float)y, __ldg(fl + 1), H, align_corners);
      t = bilinear_taps(ix, iy, H, W);
      const bool k = !(mask_thr > 0.f) || t.wsum >= mask_thr;
      flags = (t.in_nw ? 1 : 0) | (t.in_ne ? 2 : 0) | (t.in_sw ? 4 : 0) | (t.in_se ? 8 : 0) | (k ? 16 : 0);
    }
    t.x0 = __shfl_sync(0xffffffffu, t.x0, leader);
    t.y0 = __shfl_sync(0xffffffffu, t.y0, leader);
    t.w_nw = __shfl_sync(0xffffffffu, t.w_nw, leader);
    t.w_ne = __shfl_sync(0xffffffffu, t.w_ne, leader);
    t.w_sw = __shfl_sync(0xffffffffu, t.w_sw, leader);
    t.w_se = __shfl_sync(0xffffffffu, t.w_se, leader);
    ix = __shfl_sync(0xffffffffu, ix, leader);
    iy = __shfl_sync(0xffffffffu, iy, leader);
    flags = __shfl_sync(0xffffffffu, flags, leader);
    const bool in_nw = flags & 1, in_ne = flags & 2, in_sw = flags & 4, in_se = flags & 8;
    const bool keep = live && (flags & 16) != 0;
    float gix = 0.f, giy = 0.f;
    if (keep) {
      const long long base = ((long long)n * H + t.y0) * W + t.x0;   // corners are only touched when in_*
      const float fx0 = floorf(ix), fy0 = floorf(iy);
      const float ax1 = (fx0 + 1.f) - ix, ax0 = ix - fx0, ay1 = (fy0 + 1.f) - iy, ay0 = iy - fy0;
      for (int q4 = sub; q4 < quads; q4 += lpp) {
        if (!VEC) {
          // few channels (the 3-channel images of the photometric loss), or unaligned rows: one channel per lane
          const int c = q4;
          const float g = __ldg(go + (size_t)p * ldg + c);
          if (gx) {
            if (in_nw) atomicAdd(gx + base * ldgx + c, g * t.w_nw);
            if (in_ne) atomicAdd(gx + (base + 1) * ldgx + c, g * t.w_ne);
            if (in_sw) atomicAdd(gx + (base + W) * ldgx + c, g * t.w_sw);
            if (in_se) atomicAdd(gx + (base + W + 1) * ldgx + c, g * t.w_se);
          }
          if (gflow) {
            const float v_nw = in_nw ? __ldg(x + base * ldx + c) : 0.f;
            const float v_ne = in_ne ? __ldg(x + (base + 1) * ldx + c) : 0.f;
            const float v_sw = in_sw ? __ldg(x + (base + W) * ldx + c) : 0.f;
            const float v_se = in_se ? __ldg(x + (base + W + 1) * ldx + c) : 0.f;
            gix += g * ((v_ne - v_nw) * ay1 + (v_se - v_sw) * ay0);
            giy += g * ((v_sw - v_nw) * ax1 + (v_se - v_ne) * ax0);
          }
          continue;
        }
        const int c = q4 * 4;
        const float4 g = ldg4(go + (size_t)p * ldg + c);
        if (gx) {
          if (in_nw) atomicAdd(reinterpret_cast<float4*>(gx + base * ldgx + c), make_float4(g.x * t.w_nw, g.y * t.w_nw, g.z * t.w_nw, g.w * t.w_nw));
          if (in_ne) atomicAdd(reinterpret_cast<float4*>(gx + (base + 1) * ldgx + c), make_float4(g.x * t.w_ne, g.y * t.w_ne, g.z * t.w_ne, g.w * t.w_ne));
          if (in_sw) atomicAdd(reinterpret_cast<float4*>(gx + (base + W) * ldgx + c), make_float4(g.x * t.w_sw, g.y * t.w_sw, g.z * t.w_sw, g.w * t.w_sw));
          if (in_se) atomicAdd(reinterpret_cast<float4*>(gx + (base + W + 1) * ldgx + c), make_float4(g.x * t.w_se, g.y * t.w_se, g.z * t.w_se, g.w * t.w_se));
        }
        if (gflow) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 v_nw = in_nw ? ldg4(x + base * ldx + c) : z;
          const float4 v_ne = in_ne ? ldg4(x + (base + 1) * ldx + c) : z;
          const float4 v_sw = in_sw ? ldg4(x + (base + W) * ldx + c) : z;
          const float4 v_se = in_se ? ldg4(x + (base + W + 1) * ldx + c) : z;
          gix += g.x * ((v_ne.x - v_nw.x) * ay1 + (v_se.x - v_sw.x) * ay0);
          giy += g.x * ((v_sw.x - v_nw.x) * ax1 + (v_se.x - v_ne.x) * ax0);
          gix += g.y * ((v_ne.y - v_nw.y) * ay1 + (v_se.y - v_sw.y) * ay0);
          giy += g.y * ((v_sw.y - v_nw.y) * ax1 + (v_se.y - v_ne.y) * ax0);
          gix += g.z * ((v_ne.z - v_nw.z) * ay1 + (v_se.z - v_sw.z) * ay0);
          giy += g.z * ((v_sw.z - v_nw.z) * ax1 + (v_se.z - v_ne.z) * ax0);
          gix += g.w * ((v_ne.w - v_nw.w) * ay1 + (v_se.w - v_sw.w) * ay0);
          giy += g.w * ((v_sw.w - v_nw.w) * ax1 + (v_se.w - v_ne.w) * ax0);
        }
      }
    }
    if (gflow) {
      for (int o = lpp >> 1; o > 0; o >>= 1) {                 // within the lpp lanes of a pixel (lpp is a power of two)
        gix += __shfl_xor_sync(0xffffffffu, gix, o);
        giy += __shfl_xor_sync(0xffffffffu, giy, o);
      }
      if (sub == 0 && live) {
        const float sx = align_corners ? 1.0f : (float)W / (float)(W > 1 ? W - 1 : 1);
        const float sy = align_corners ? 1.0f : (float)H / (float)(H > 1 ? H - 1 : 1);
        gflow[(size_t)p * ldgf + 0] = gix * sx;
        gflow[(size_t)p * ldgf + 1] = giy * sy;
      }
    }
  }
}


// ---- normalize_features' other moment modes (model/upflow.py:94-137) ----------------------------------------------
// The correlation kernels normalise with per-(image, channel) statistics given as (sum, sum of squares).  The
// reference can also pool the moments over the channels of an image (moments_across_channels: mean / unbiased var
// over [C,H,W]) and over the two tensors of a pair (moments_across_images: the MEAN of the two means, and -- as the
// reference writes it, :121-124 -- the unbiased VARIANCE of the two variances).  Both are still one (mean, std) per
// image and channel, so this kernel rewrites the raw moments of a pair (A = image n of sa, B = image (n+shift)%N of
// sb) into EQUIVALENT per-channel moments (sum' = mean*npix, sumsq' = var*(npix-1) + sum'*mean) that
// stats_to_mean_std turns back into the pooled mean / std.  One CTA per pair.
__global__ void __launch_bounds__(256)
featnorm_combine_kernel(const double* __restrict__ sa, const double* __restrict__ sb, int shift, double* __restrict__ oa,
                        double* __restrict__ ob, int N, int C, double npix, int across_ch, int across_img) {
  pdl_prologue();
  const int n = blockIdx.x, n2 = (n + shift) % N;
  const double* a = sa + (size_t)n * C * 2;
  const double* b = sb + (size_t)n2 * C * 2;
  __shared__ double red[4];
  if (across_ch) {
    if (threadIdx.x < 4) red[threadIdx.x] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {                 // C <= 256 values per sum: a fixed-order serial sum is deterministic and cheap
      double t[4] = {0.0, 0.0, 0.0, 0.0};
      for (int c = 0; c < C; ++c) { t[0] += a[c * 2]; t[1] += a[c * 2 + 1]; t[2] += b[c * 2]; t[3] += b[c * 2 + 1]; }
      for (int k = 0; k < 4; ++k) red[k] = t[k];
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double cnt = npix, a0 = a[c * 2], a1 = a[c * 2 + 1], b0 = b[c * 2], b1 = b[c * 2 + 1];
    if (across_ch) { cnt = npix * (double)C; a0 = red[0]; a1 = red[1]; b0 = red[2]; b1 = red[3]; }
    double ma = a0 / cnt, mb = b0 / cnt;
    double va = (a1 - a0 * ma) / (cnt - 1.0), vb = (b1 - b0 * mb) / (cnt - 1.0);
    if (va < 0.0) va = 0.0;
    if (vb < 0.0) vb = 0.0;
    if (across_img) {
      // the reference works on the fp32 means / variances of the two tensors from here on
      const double fa = (double)(float)ma, fb = (double)(float)mb, ga = (double)(float)va, gb = (double)(float)vb;
      const double m = (fa + fb) * 0.5, gm = (ga + gb) * 0.5;
      const double v = (ga - gm) * (ga - gm) + (gb - gm) * (gb - gm);     // torch.var of two values, unbiased (n-1 = 1)
      ma = mb = m; va = vb = v;
    }
    double* pa = oa + ((size_t)n * C + c) * 2;
    double* pb = ob + ((size_t)n2 * C + c) * 2;
    pa[0] = ma * npix; pa[1] = va * (npix - 1.0) + ma * npix * ma;
    pb[0] = mb * npix; pb[1] = vb * (npix - 1.0) + mb * npix * mb;
  }
}

static int pick_lpp(int C) {
  int cg = (C + 3) / 4;
  int lpp = 1;
  while (lpp < cg && lpp < 32) lpp <<= 1;   // power of two <= 32 dividing WARP_NT
  return lpp;
}


// ---- forward/backward consistency occlusion masks (tools.occ_check_model, utils/tools.py:501-677, 'for_back_check') ----
// flow [N,H,W,>=2] holds the forward flows in images 0..N/2-1 and the backward flows in N/2..N-1 (the decoder's stacked
// layout); for image n, "other" = image (n + N/2) % N.
//   mag(f) = sqrt(u^2) + sqrt(v^2);  thresh = a1 * (mag(own) + mag(other)) + a2
//   occ = mag(own + torch_warp(other, own)) < thresh          (1 = consistent / visible, 0 = occluded)
// mode 0 ('all'): occ;  1 ('obj'): occ | outgoing, outgoing = the flow leaves the image;  2 ('out'): 1 - outgoing.
// One launch replaces the ~25 elementwise torch kernels + 2 warps the reference runs after every forward.
__global__ void __launch_bounds__(256)
occ_check_kernel(const float* __restrict__ flow, int ldf, float* __restrict__ occ, int ldo, int N, int H, int W,
                 float a1, float a2, int mode, int align_corners) {
  pdl_prologue();
  const long long total = (long long)N * H * W;
  const int half = N / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t0 = i / W;
    const int y = (int)(t0 % H);
    const int n = (int)(t0 / H);
    const int n2 = (n + half) % N;
    const float* own = flow + (size_t)i * ldf;
    const float u = __ldg(own), v = __ldg(own + 1);
    const float px = __fadd_rn((float)x, u), py = __fadd_rn((float)y, v);
    const bool inside = (px <= (float)(W - 1)) && (px >= 0.f) && (py <= (float)(H - 1)) && (py >= 0.f);
    float r;
    if (mode == 2) {
      r = inside ? 1.f : 0.f;
    } else {
      const float* oth = flow + ((size_t)((size_t)n2 * H + y) * W + x) * ldf;
      const float ou = __ldg(oth), ov = __ldg(oth + 1);
      const float ix = sample_coord((float)x, u, W, align_corners), iy = sample_coord((float)y, v, H, align_corners);
      const BilinearTaps t = bilinear_taps(ix, iy, H, W);
      const float* b = flow + ((size_t)((size_t)n2 * H + t.y0) * W + t.x0) * ldf;    // dereferenced only when in_*
      float wu = 0.f, wv = 0.f;
      if (t.in_nw) { wu = fmaf(__ldg(b), t.w_nw, wu); wv = fmaf(__ldg(b + 1), t.w_nw, wv); }
      if (t.in_ne) { wu = fmaf(__ldg(b + ldf), t.w_ne, wu); wv = fmaf(__ldg(b + ldf + 1), t.w_ne, wv); }
      if (t.in_sw) { wu = fmaf(__ldg(b + (size_t)W * ldf), t.w_sw, wu); wv = fmaf(__ldg(b + (size_t)W * ldf + 1), t.w_sw, wv); }
      if (t.in_se) { wu = fmaf(__ldg(b + (size_t)(W + 1) * ldf), t.w_se, wu); wv = fmaf(__ldg(b + (size_t)(W + 1) * ldf + 1), t.w_se, wv); }
      const float mag_sq = __fadd_rn(__fadd_rn(sqrtf(__fmul_rn(u, u)), sqrtf(__fmul_rn(v, v))),
                                     __fadd_rn(sqrtf(__fmul_rn(ou, ou)), sqrtf(__fmul_rn(ov, ov))));
      const float du = __fadd_rn(u, wu), dv = __fadd_rn(v, wv);
      const float diff = __fadd_rn(sqrtf(__fmul_rn(du, du)), sqrtf(__fmul_rn(dv, dv)));
      const float thresh = __fadd_rn(__fmul_rn(a1, mag_sq), a2);
      const bool vis = diff < thresh;
      r = (mode == 1 ? (vis || !inside) : vis) ? 1.f : 0.f;
    }
    occ[(size_t)i * ldo] = r;
  }
}

}  // namespace upf

namespace upf {
template <typename T>
static int warp_fwd_launch(const T* x, int ldx, const float* flow, int ldf, T* out, int ldo,
                           int N, int H, int W, int C, int align_corners, float mask_threshold, int x_batch_shift,
                           double* stats, int flags, void* stream) {
  UPF_REQUIRE(x && flow && out, "warp: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && ldo >= C && ldf >= 2, "warp: bad shape");
  UPF_REQUIRE(stats == nullptr || C <= 256, "warp: fused moments need C <= 256");
  UPF_REQUIRE(x_batch_shift >= 0 && x_batch_shift < N, "warp: batch shift out of range");
  const int lpp = pick_lpp(C);
  const int ppc = WARP_NT / lpp;
  int per_image = (H * W + ppc - 1) / ppc;
  // the CTA count per image fixes how the fp32 partial moments are grouped: it must depend on the image
  // size only, never on N, so that an image gives the same bits in any batch (batch sharding relies on it)
  // (with the fused moments every CTA ends in one double atomic per channel and moment: measured, 914 CTAs per image
  // cost 29 us at 2x32x94x311 where the warp alone takes 13.6 -- fewer, longer CTAs there)
  const int cap = stats ? UPF_NUM_SMS * 2 : UPF_NUM_SMS * 8;
  if (per_image > cap) per_image = cap;
  if (per_image < 1) per_image = 1;
  const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldo % 4 == 0) && aligned_vec4<T>(x) && aligned_vec4<T>(out);
  const size_t smem = stats ? (size_t)(WARP_NT / 32) * 2 * ((C + 3) / 4) * 4 * sizeof(double) : 0;
  if (vec)
    UPF_LAUNCH((warp_fwd_kernel<true, T>), N * per_image, WARP_NT, smem, (cudaStream_t)stream, x, ldx, flow, ldf, out, ldo, H, W, C,
                                                                                 align_corners, mask_threshold, stats, lpp, per_image, x_batch_shift, N, flags);
  else
    UPF_LAUNCH((warp_fwd_kernel<false, T>), N * per_image, WARP_NT, smem, (cudaStream_t)stream, x, ldx, flow, ldf, out, ldo, H, W, C,
                                                                                  align_corners, mask_threshold, stats, lpp, per_image, x_batch_shift, N, flags);
  return check_launch("warp_fwd");
}
}  // namespace upf

extern "C" int upf_warp_fwd(const float* x, int ldx, const float* flow, int ldf, float* out, int ldo,
                            int N, int H, int W, int C, int align_corners, float mask_threshold, int x_batch_shift,
                            double* stats, int flags, void* stream) {
  return upf::warp_fwd_launch<float>(x, ldx, flow, ldf, out, ldo, N, H, W, C, align_corners, mask_threshold, x_batch_shift, stats,
                                     flags, stream);
}

// fp16 / bf16 STORAGE of the warped tensor and its source: taps are converted to fp32, blended in
// fp32 in ATen's order, rounded once at the store; the flow and the mask arithmetic stay fp32 (so the mask is the same
// bit-faithful one); the fused moments are those of the ROUNDED output's fp32 pre-image.
extern "C" int upf_warp_fwd_lp(const void* x, int ldx, const float* flow, int ldf, void* out, int ldo, int dtype,
                               int N, int H, int W, int C, int align_corners, float mask_threshold, int x_batch_shift,
                               double* stats, void* stream) {
  using namespace upf;
  UPF_REQUIRE(dtype == UPF_DTYPE_F16 || dtype == UPF_DTYPE_BF16, "warp_lp: dtype must be UPF_DTYPE_F16 or UPF_DTYPE_BF16");
  if (dtype == UPF_DTYPE_F16)
    return warp_fwd_launch<__half>(static_cast<const __half*>(x), ldx, flow, ldf, static_cast<__half*>(out), ldo, N, H, W, C,
                                   align_corners, mask_threshold, x_batch_shift, stats, 0, stream);
  return warp_fwd_launch<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(x), ldx, flow, ldf, static_cast<__nv_bfloat16*>(out), ldo,
                                        N, H, W, C, align_corners, mask_threshold, x_batch_shift, stats, 0, stream);
}

extern "C" int upf_warp_bwd(const float* x, int ldx, const float* flow, int ldf, const float* grad_out, int ldg,
                            float* grad_x, int ldgx, float* grad_flow, int ldgf,
                            int N, int H, int W, int C, int align_corners, float mask_threshold, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && flow && grad_out, "warp_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "warp_bwd: bad shape");
  const long long npix = (long long)N * H * W;
  const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldg % 4 == 0) && (!grad_x || ldgx % 4 == 0) && aligned16(x) && aligned16(grad_out) &&
                   (!grad_x || aligned16(grad_x));
  if (vec || C <= 16) {
    const int items = vec ? C / 4 : C;
    int lpp = 1;
    while (lpp < items && lpp < 32) lpp <<= 1;
    const int ppw = 32 / lpp;
    long long vblocks = ((npix + ppw - 1) / ppw * 32 + 255) / 256;
    if (vblocks > UPF_NUM_SMS * 16) vblocks = UPF_NUM_SMS * 16;
    if (vec)
      UPF_LAUNCH((warp_bwd_vec_kernel<true>), (unsigned)vblocks, 256, 0, (cudaStream_t)stream, x, ldx, flow, ldf, grad_out, ldg, grad_x, ldgx,
                                                                                    grad_flow, ldgf, N, H, W, C, align_corners, mask_threshold, lpp);
    else
      UPF_LAUNCH((warp_bwd_vec_kernel<false>), (unsigned)vblocks, 256, 0, (cudaStream_t)stream, x, ldx, flow, ldf, grad_out, ldg, grad_x, ldgx,
                                                                                     grad_flow, ldgf, N, H, W, C, align_corners, mask_threshold, lpp);
    return check_launch("warp_bwd");
  }
  long long blocks = (npix * 32 + 255) / 256;
  if (blocks > UPF_NUM_SMS * 16) blocks = UPF_NUM_SMS * 16;
  UPF_LAUNCH((warp_bwd_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, x, ldx, flow, ldf, grad_out, ldg, grad_x, ldgx,
                                                                     grad_flow, ldgf, N, H, W, C, align_corners, mask_threshold);
  return check_launch("warp_bwd");
}

extern "C" int upf_featnorm_stats(const float* x, int ldx, int N, int H, int W, int C, double* stats, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && stats, "featnorm_stats: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 256 && ldx >= C, "featnorm_stats: bad shape (C<=256)");
  const int lpp = pick_lpp(C);
  const int ppc = WARP_NT / lpp;
  int per_image = (H * W + ppc * 8 - 1) / (ppc * 8);
  const int cap = UPF_NUM_SMS * 8;                   // independent of N (see upf_warp_fwd)
  if (per_image > cap) per_image = cap;
  if (per_image < 1) per_image = 1;
  const int vec = (C % 4 == 0) && (ldx % 4 == 0) && aligned16(x);
  const size_t smem = (size_t)(WARP_NT / 32) * 2 * ((C + 3) / 4) * 4 * sizeof(double);
  UPF_LAUNCH((featnorm_stats_kernel), N * per_image, WARP_NT, smem, (cudaStream_t)stream, x, ldx, H, W, C, stats, lpp, per_image, vec);
  return check_launch("featnorm_stats");
}

extern "C" int upf_featnorm_combine(const double* stats_a, const double* stats_b, int b_batch_shift, double* out_a, double* out_b,
                                    int N, int C, long long npix, int across_channels, int across_images, void* stream) {
  using namespace upf;
  UPF_REQUIRE(stats_a && stats_b && out_a && out_b, "featnorm_combine: null tensor");
  UPF_REQUIRE(N > 0 && C > 0 && C <= 256 && npix > 1 && b_batch_shift >= 0 && b_batch_shift < N, "featnorm_combine: bad argument");
  UPF_REQUIRE(out_a != out_b && out_a != stats_a && out_a != stats_b && out_b != stats_a && out_b != stats_b,
              "featnorm_combine: outputs must be distinct buffers");
  UPF_LAUNCH((featnorm_combine_kernel), N, 256, 0, (cudaStream_t)stream, stats_a, stats_b, b_batch_shift, out_a, out_b, N, C,
             (double)npix, across_channels ? 1 : 0, across_images ? 1 : 0);
  return check_launch("featnorm_combine");
}

extern "C" int upf_featnorm_apply(const float* x, int ldx, const double* stats, float* out, int ldo,
                                  int N, int H, int W, int C, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && stats && out, "featnorm_apply: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && ldx >= C && ldo >= C, "featnorm_apply: bad shape");
  const long long total = (long long)N * H * W * C;
  long long blocks = (total + 255) / 256;
  if (blocks > UPF_NUM_SMS * 16) blocks = UPF_NUM_SMS * 16;
  UPF_LAUNCH((featnorm_apply_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, x, ldx, stats, out, ldo, H, W, C, total);
  return check_launch("featnorm_apply");
}

extern "C" int upf_occ_check(const float* flow, int ldf, float* occ, int ldo, int N, int H, int W, float alpha_1,
                             float alpha_2, int mode, int align_corners, void* stream) {
  using namespace upf;
  UPF_REQUIRE(flow && occ, "occ_check: null tensor");
  UPF_REQUIRE(N > 0 && (N % 2) == 0 && H > 0 && W > 0 && ldf >= 2 && ldo >= 1 && mode >= 0 && mode <= 2, "occ_check: bad argument");
  long long blocks = ((long long)N * H * W + 255) / 256;
  if (blocks > UPF_NUM_SMS * 16) blocks = UPF_NUM_SMS * 16;
  UPF_LAUNCH((occ_check_kernel), (unsigned)blocks, 256, 0, (cudaStream_t)stream, flow, ldf, occ, ldo, N, H, W, alpha_1, alpha_2,
             mode, align_corners);
  return check_launch("occ_check");
}
