// Loss terms of the training step as fused kernels: the robust photometric /
// distillation term (model/upflow.py:268-290, used by :436-437 and :461-487) and the first-order edge-aware
// smoothness term (model/upflow.py:198-218).  The reference builds each from 6-10 elementwise ATen kernels forward
// and as many backward; here a term is ONE reduction pass forward (per-CTA partials, summed in a fixed order by a
// one-CTA finisher: deterministic) and ONE elementwise pass backward.  All tensors are pixel-major [npix][ld] like
// the rest of the library; every kernel is HBM-bound (robust: 4*npix*(2C+1) B forward, 4*npix*(3C+1) B backward).
#include "upf_common.cuh"

namespace upf {

constexpr int LOSS_THREADS = 256;
constexpr int LOSS_MAX_BLOCKS = 4 * UPF_NUM_SMS;

__device__ __forceinline__ float robust_val(int kind, float d, float q) {
  if (kind == UPF_LOSS_ABS_ROBUST) return powf(fabsf(d) + 0.01f, q);
  if (kind == UPF_LOSS_CHARBONNIER) return powf(d * d + 1e-6f, q);
  return fabsf(d + 1e-6f);
}
__device__ __forceinline__ float sgn(float v) { return (float)(v > 0.f) - (float)(v < 0.f); }
__device__ __forceinline__ float robust_grad(int kind, float d, float q) {
  if (kind == UPF_LOSS_ABS_ROBUST) return q * powf(fabsf(d) + 0.01f, q - 1.f) * sgn(d);
  if (kind == UPF_LOSS_CHARBONNIER) return q * powf(d * d + 1e-6f, q - 1.f) * 2.f * d;
  return sgn(d + 1e-6f);
}

// two running sums per thread -> part[blockIdx.x], part[gridDim.x + blockIdx.x]; lanes by shuffle, warps in warp order
__device__ __forceinline__ void block_sum2(float a, float b, float* __restrict__ part) {
  __shared__ float s_a[LOSS_THREADS / 32], s_b[LOSS_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a; s_b[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int k = 0; k < LOSS_THREADS / 32; ++k) { ta += s_a[k]; tb += s_b[k]; }
    part[blockIdx.x] = ta;
    part[gridDim.x + blockIdx.x] = tb;
  }
}

__global__ void __launch_bounds__(LOSS_THREADS)
robust_loss_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                       const float* __restrict__ mask, int ldm, float* __restrict__ part, long long npix, int C, int kind,
                       float q) {
  float sd = 0.f, sm = 0.f;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const float m = mask ? __ldg(mask + (size_t)p * ldm) : 1.f;
    const float* xp = x + (size_t)p * ldx;
    const float* yp = y + (size_t)p * ldy;
    float t = 0.f;
    for (int c = 0; c < C; ++c) t += robust_val(kind, __ldg(xp + c) - __ldg(yp + c), q);
    sd += t * m;
    sm += m;
  }
  block_sum2(sd, sm, part);
}

// out[0] = the term, out[1] = 1/denominator (what backward scales by).  masked 0: mean over `count` elements;
// masked 1: sum(d*mask)/(mask_scale*sum(mask)+1e-6) (mask_scale 2: the reference's census normalisation,
// utils/loss.py:44-46).
__global__ void __launch_bounds__(LOSS_THREADS)
robust_loss_finish_kernel(const float* __restrict__ part, int blocks, float* __restrict__ out, double count, int masked,
                          double mask_scale) {
  __shared__ double s_a[LOSS_THREADS], s_b[LOSS_THREADS];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < blocks; k += LOSS_THREADS) { a += part[k]; b += part[blocks + k]; }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < LOSS_THREADS; ++k) { ta += s_a[k]; tb += s_b[k]; }
    const double inv = masked ? 1.0 / (tb * mask_scale + 1e-6) : 1.0 / count;
    out[0] = (float)(ta * inv);
    out[1] = (float)inv;
  }
}

__global__ void __launch_bounds__(LOSS_THREADS)
robust_loss_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ y, int ldy,
                       const float* __restrict__ mask, int ldm, const float* __restrict__ out,
                       const float* __restrict__ gout, float* __restrict__ gx, int ldgx, float* __restrict__ gy, int ldgy,
                       long long npix, int C, int kind, float q) {
  const float s = __ldg(gout) * __ldg(out + 1);
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const float m = (mask ? __ldg(mask + (size_t)p * ldm) : 1.f) * s;
    const float* xp = x + (size_t)p * ldx;
    const float* yp = y + (size_t)p * ldy;
    for (int c = 0; c < C; ++c) {
      const float g = robust_grad(kind, __ldg(xp + c) - __ldg(yp + c), q) * m;
      if (gx) gx[(size_t)p * ldgx + c] = g;
      if (gy) gy[(size_t)p * ldgy + c] = -g;
    }
  }
}

// exp(-mean_c |img[p] - img[p + step]|): the edge weight between a pixel and its neighbour `step` pixels away
__device__ __forceinline__ float edge_weight(const float* __restrict__ ip, long long step_elems, int Ci) {
  float a = 0.f;
  for (int c = 0; c < Ci; ++c) a += fabsf(__ldg(ip + c) - __ldg(ip + step_elems + c));
  return expf(-a / (float)Ci);
}

__global__ void __launch_bounds__(LOSS_THREADS)
edge_smooth1_fwd_kernel(const float* __restrict__ img, int ldi, int Ci, const float* __restrict__ pred, int ldp, int Cp,
                        float* __restrict__ part, int N, int H, int W) {
  const long long npix = (long long)N * H * W;
  float sr = 0.f, sc = 0.f;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const float* ip = img + (size_t)p * ldi;
    const float* pp = pred + (size_t)p * ldp;
    if (h < H - 1) {                                         // the reference's "gradient_x": row differences
      const float wr = edge_weight(ip, (long long)W * ldi, Ci);
      float t = 0.f;
      for (int c = 0; c < Cp; ++c) t += fabsf(__ldg(pp + c) - __ldg(pp + (size_t)W * ldp + c));
      sr += t * wr;
    }
    if (w < W - 1) {
      const float wc = edge_weight(ip, ldi, Ci);
      float t = 0.f;
      for (int c = 0; c < Cp; ++c) t += fabsf(__ldg(pp + c) - __ldg(pp + ldp + c));
      sc += t * wc;
    }
  }
  block_sum2(sr, sc, part);
}

__global__ void __launch_bounds__(LOSS_THREADS)
edge_smooth1_finish_kernel(const float* __restrict__ part, int blocks, float* __restrict__ out, double inv_r, double inv_c) {
  __shared__ double s_a[LOSS_THREADS], s_b[LOSS_THREADS];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < blocks; k += LOSS_THREADS) { a += part[k]; b += part[blocks + k]; }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < LOSS_THREADS; ++k) { ta += s_a[k]; tb += s_b[k]; }
    out[0] = (float)(ta * inv_r + tb * inv_c);
  }
}

// gather form: a pixel collects the (at most four) differences it takes part in
__global__ void __launch_bounds__(LOSS_THREADS)
edge_smooth1_bwd_kernel(const float* __restrict__ img, int ldi, int Ci, const float* __restrict__ pred, int ldp, int Cp,
                        const float* __restrict__ gout, float* __restrict__ gpred, int ldg, int N, int H, int W,
                        float inv_r, float inv_c) {
  const long long npix = (long long)N * H * W;
  const float g0 = __ldg(gout);
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const float* ip = img + (size_t)p * ldi;
    const float* pp = pred + (size_t)p * ldp;
    const long long ri = (long long)W * ldi, rp = (long long)W * ldp;
    const float w_dn = h < H - 1 ? edge_weight(ip, ri, Ci) * inv_r : 0.f;
    const float w_up = h > 0 ? edge_weight(ip - ri, ri, Ci) * inv_r : 0.f;
    const float w_rt = w < W - 1 ? edge_weight(ip, ldi, Ci) * inv_c : 0.f;
    const float w_lf = w > 0 ? edge_weight(ip - ldi, ldi, Ci) * inv_c : 0.f;
    for (int c = 0; c < Cp; ++c) {
      const float v = __ldg(pp + c);
      float g = 0.f;
      if (h < H - 1) g += sgn(v - __ldg(pp + rp + c)) * w_dn;
      if (h > 0) g -= sgn(__ldg(pp - rp + c) - v) * w_up;
      if (w < W - 1) g += sgn(v - __ldg(pp + ldp + c)) * w_rt;
      if (w > 0) g -= sgn(__ldg(pp - ldp + c) - v) * w_lf;
      gpred[(size_t)p * ldg + c] = g * g0;
    }
  }
}

// ---- census term (utils/loss.py:51-91): soft ternary census transform of the grey images over a (2d+1)^2 patch
// (zero padding), soft Hamming distance, robust penalty.  The reference materialises two 49-channel tensors with a
// one-hot convolution; here a pixel walks its patch in registers over a 2-channel grey buffer (L1-resident).
__global__ void __launch_bounds__(LOSS_THREADS)
census_grey_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, float2* __restrict__ grey,
                   long long npix) {
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const float* ap = a + (size_t)p * lda;
    const float* bp = b + (size_t)p * ldb;
    grey[p] = make_float2(0.2989f * __ldg(ap) + 0.5870f * __ldg(ap + 1) + 0.1140f * __ldg(ap + 2),
                          0.2989f * __ldg(bp) + 0.5870f * __ldg(bp + 1) + 0.1140f * __ldg(bp + 2));
  }
}

// h'(u) * tau'(v2) for the pair (neighbour value g, centre value c): u = tau(g.x-c.x) - tau(g.y-c.y),
// tau(v) = v / sqrt(0.81 + v^2), h(u) = u^2 / (0.1 + u^2)
__device__ __forceinline__ float census_pair_grad(float2 g, float2 c) {
  const float v1 = g.x - c.x, v2 = g.y - c.y;
  const float r1 = 1.f / sqrtf(0.81f + v1 * v1), r2 = 1.f / sqrtf(0.81f + v2 * v2);
  const float u = v1 * r1 - v2 * r2;
  const float den = 0.1f + u * u;
  return (0.2f * u / (den * den)) * (0.81f * r2 * r2 * r2);
}

__global__ void __launch_bounds__(LOSS_THREADS)
census_fwd_kernel(const float2* __restrict__ grey, const float* __restrict__ mask, int ldm, float* __restrict__ dist,
                  float* __restrict__ part, int N, int H, int W, int d, float q) {
  const long long npix = (long long)N * H * W;
  float sd = 0.f, sm = 0.f;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    const long long img = p - ((long long)y * W + x);
    const float2 c = grey[p];
    float ds = 0.f;
    for (int dy = -d; dy <= d; ++dy)
      for (int dx = -d; dx <= d; ++dx) {
        const int yy = y + dy, xx = x + dx;
        float2 g = make_float2(0.f, 0.f);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) g = grey[img + (long long)yy * W + xx];
        const float v1 = g.x - c.x, v2 = g.y - c.y;
        const float u = v1 / sqrtf(0.81f + v1 * v1) - v2 / sqrtf(0.81f + v2 * v2);
        const float uu = u * u;
        ds += uu / (0.1f + uu);
      }
    dist[p] = ds;
    float m = 1.f;
    if (mask) m = (y >= d && y < H - d && x >= d && x < W - d) ? __ldg(mask + (size_t)p * ldm) : 0.f;
    sd += powf(fabsf(ds) + 0.01f, q) * m;
    sm += m;
  }
  block_sum2(sd, sm, part);
}

// gradient wrt the SECOND image (the warped one), gather form: pixel p collects its role as the centre of its own patch
// and as a neighbour in the (2d+1)^2 patches around it
__global__ void __launch_bounds__(LOSS_THREADS)
census_bwd_kernel(const float2* __restrict__ grey, const float* __restrict__ mask, int ldm, const float* __restrict__ dist,
                  const float* __restrict__ out, const float* __restrict__ gout, float* __restrict__ gimg, int ldg, int N, int H,
                  int W, int d, float q) {
  const long long npix = (long long)N * H * W;
  const float s = __ldg(gout) * __ldg(out + 1) * q;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const int x = (int)(p % W), y = (int)((p / W) % H);
    const long long img = p - ((long long)y * W + x);
    const float2 c = grey[p];
    // d(term)/d(dist) at pixel (yy, xx)
    auto coef = [&](int yy, int xx) -> float {
      const long long pc = img + (long long)yy * W + xx;
      float m = 1.f;
      if (mask) m = (yy >= d && yy < H - d && xx >= d && xx < W - d) ? __ldg(mask + (size_t)pc * ldm) : 0.f;
      if (m == 0.f) return 0.f;
      const float ds = __ldg(dist + pc);
      return s * m * powf(fabsf(ds) + 0.01f, q - 1.f) * sgn(ds);
    };
    const float a_own = coef(y, x);
    float acc = 0.f;
    for (int dy = -d; dy <= d; ++dy)
      for (int dx = -d; dx <= d; ++dx) {
        if (a_own != 0.f) {                                  // centre of its own patch: neighbour p+o (zero outside)
          const int yy = y + dy, xx = x + dx;
          float2 g = make_float2(0.f, 0.f);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) g = grey[img + (long long)yy * W + xx];
          acc += a_own * census_pair_grad(g, c);
        }
        const int cy = y - dy, cx = x - dx;                  // neighbour in the patch centred at p-o
        if (cy >= 0 && cy < H && cx >= 0 && cx < W) {
          const float a_c = coef(cy, cx);
          if (a_c != 0.f) acc -= a_c * census_pair_grad(c, grey[img + (long long)cy * W + cx]);
        }
      }
    float* gp = gimg + (size_t)p * ldg;
    gp[0] = 0.2989f * acc;
    gp[1] = 0.5870f * acc;
    gp[2] = 0.1140f * acc;
  }
}

// ---- boundary-dilated warp (utils/tools.py:350-499): bilinear lookup of the UN-CROPPED frame I [N,Hf,Wf,C] at
// (x + start_x + u, y + start_y + v); corner indices clamped to the frame, weights taken against the CLAMPED corners
// (so a sample outside the frame extrapolates from the border pixels: that is the reference's arithmetic).
struct BdCorners { float fx, fy, x0c, x1c, y0c, y1c; long long ia, ib, ic, id; };
__device__ __forceinline__ BdCorners bd_corners(const float* __restrict__ flow, int ldf, const float* __restrict__ start, long long p,
                                                int h, int w, int Hf, int Wf) {
  const int x = (int)(p % w), y = (int)((p / w) % h);
  const long long n = p / ((long long)h * w);
  BdCorners c;
  c.fx = ((float)x + __ldg(start + 2 * n)) + __ldg(flow + (size_t)p * ldf);
  c.fy = ((float)y + __ldg(start + 2 * n + 1)) + __ldg(flow + (size_t)p * ldf + 1);
  const float x0 = floorf(c.fx), y0 = floorf(c.fy);
  c.x0c = fminf(fmaxf(x0, 0.f), (float)(Wf - 1));
  c.x1c = fminf(fmaxf(x0 + 1.f, 0.f), (float)(Wf - 1));
  c.y0c = fminf(fmaxf(y0, 0.f), (float)(Hf - 1));
  c.y1c = fminf(fmaxf(y0 + 1.f, 0.f), (float)(Hf - 1));
  const long long base = n * Hf;
  c.ia = (base + (long long)c.y0c) * Wf + (long long)c.x0c;
  c.ib = (base + (long long)c.y1c) * Wf + (long long)c.x0c;
  c.ic = (base + (long long)c.y0c) * Wf + (long long)c.x1c;
  c.id = (base + (long long)c.y1c) * Wf + (long long)c.x1c;
  return c;
}

__global__ void __launch_bounds__(LOSS_THREADS)
bdwarp_fwd_kernel(const float* __restrict__ I, int ldi, int C, int Hf, int Wf, const float* __restrict__ flow, int ldf,
                  const float* __restrict__ start, float* __restrict__ out, int ldo, int N, int h, int w) {
  const long long npix = (long long)N * h * w;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const BdCorners c = bd_corners(flow, ldf, start, p, h, w, Hf, Wf);
    const float wa = (c.x1c - c.fx) * (c.y1c - c.fy), wb = (c.x1c - c.fx) * (c.fy - c.y0c);
    const float wc = (c.fx - c.x0c) * (c.y1c - c.fy), wd = (c.fx - c.x0c) * (c.fy - c.y0c);
    for (int k = 0; k < C; ++k)
      out[(size_t)p * ldo + k] = wa * __ldg(I + (size_t)c.ia * ldi + k) + wb * __ldg(I + (size_t)c.ib * ldi + k) +
                                 wc * __ldg(I + (size_t)c.ic * ldi + k) + wd * __ldg(I + (size_t)c.id * ldi + k);
  }
}

// gradient wrt the flow (floor and clamp are piecewise constant): d/du = sum_k g_k [(y1c-y)(Ic-Ia) + (y-y0c)(Id-Ib)],
// d/dv = sum_k g_k [(x1c-x)(Ib-Ia) + (x-x0c)(Id-Ic)]
__global__ void __launch_bounds__(LOSS_THREADS)
bdwarp_bwd_kernel(const float* __restrict__ I, int ldi, int C, int Hf, int Wf, const float* __restrict__ flow, int ldf,
                  const float* __restrict__ start, const float* __restrict__ g, int ldg, float* __restrict__ gflow, int ldgf,
                  int N, int h, int w) {
  const long long npix = (long long)N * h * w;
  for (long long p = blockIdx.x * (long long)LOSS_THREADS + threadIdx.x; p < npix; p += (long long)gridDim.x * LOSS_THREADS) {
    const BdCorners c = bd_corners(flow, ldf, start, p, h, w, Hf, Wf);
    float gu = 0.f, gv = 0.f;
    for (int k = 0; k < C; ++k) {
      const float a = __ldg(I + (size_t)c.ia * ldi + k), b = __ldg(I + (size_t)c.ib * ldi + k);
      const float cc = __ldg(I + (size_t)c.ic * ldi + k), d = __ldg(I + (size_t)c.id * ldi + k);
      const float gk = __ldg(g + (size_t)p * ldg + k);
      gu += gk * ((c.y1c - c.fy) * (cc - a) + (c.fy - c.y0c) * (d - b));
      gv += gk * ((c.x1c - c.fx) * (b - a) + (c.fx - c.x0c) * (d - cc));
    }
    gflow[(size_t)p * ldgf] = gu;
    gflow[(size_t)p * ldgf + 1] = gv;
  }
}

static int loss_blocks(long long npix) {
  long long b = (npix + LOSS_THREADS - 1) / LOSS_THREADS;
  if (b > LOSS_MAX_BLOCKS) b = LOSS_MAX_BLOCKS;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace upf

extern "C" long long upf_loss_workspace_elems(void) { return 2LL * upf::LOSS_MAX_BLOCKS; }

extern "C" int upf_robust_loss_fwd(const float* x, int ldx, const float* y, int ldy, const float* mask, int ldm,
                                   float* workspace, float* out, long long npix, int C, int kind, float q, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && y && workspace && out, "robust_loss_fwd: null tensor");
  UPF_REQUIRE(npix > 0 && C > 0 && ldx >= C && ldy >= C && (!mask || ldm >= 1), "robust_loss_fwd: bad shape");
  UPF_REQUIRE(kind >= 0 && kind <= 2, "robust_loss_fwd: kind must be UPF_LOSS_ABS_ROBUST, _CHARBONNIER or _L1");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = loss_blocks(npix);
  robust_loss_fwd_kernel<<<blocks, LOSS_THREADS, 0, st>>>(x, ldx, y, ldy, mask, ldm, workspace, npix, C, kind, q);
  int e = check_launch("robust_loss_fwd");
  if (e) return e;
  robust_loss_finish_kernel<<<1, LOSS_THREADS, 0, st>>>(workspace, blocks, out, (double)npix * C, mask != nullptr, 1.0);
  return check_launch("robust_loss_finish");
}

extern "C" int upf_robust_loss_bwd(const float* x, int ldx, const float* y, int ldy, const float* mask, int ldm,
                                   const float* out, const float* grad_out, float* grad_x, int ldgx, float* grad_y,
                                   int ldgy, long long npix, int C, int kind, float q, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && y && out && grad_out && (grad_x || grad_y), "robust_loss_bwd: null tensor");
  UPF_REQUIRE(npix > 0 && C > 0 && ldx >= C && ldy >= C && (!grad_x || ldgx >= C) && (!grad_y || ldgy >= C) &&
                  (!mask || ldm >= 1), "robust_loss_bwd: bad shape");
  UPF_REQUIRE(kind >= 0 && kind <= 2, "robust_loss_bwd: kind must be UPF_LOSS_ABS_ROBUST, _CHARBONNIER or _L1");
  robust_loss_bwd_kernel<<<loss_blocks(npix), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      x, ldx, y, ldy, mask, ldm, out, grad_out, grad_x, ldgx, grad_y, ldgy, npix, C, kind, q);
  return check_launch("robust_loss_bwd");
}

extern "C" int upf_edge_smooth1_fwd(const float* img, int ldi, int Ci, const float* pred, int ldp, int Cp,
                                    float* workspace, float* out, int N, int H, int W, void* stream) {
  using namespace upf;
  UPF_REQUIRE(img && pred && workspace && out, "edge_smooth1_fwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 1 && W > 1 && Ci > 0 && Cp > 0 && ldi >= Ci && ldp >= Cp,
              "edge_smooth1_fwd: needs H, W >= 2 (the reference takes the mean of an empty tensor otherwise)");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = loss_blocks((long long)N * H * W);
  edge_smooth1_fwd_kernel<<<blocks, LOSS_THREADS, 0, st>>>(img, ldi, Ci, pred, ldp, Cp, workspace, N, H, W);
  int e = check_launch("edge_smooth1_fwd");
  if (e) return e;
  edge_smooth1_finish_kernel<<<1, LOSS_THREADS, 0, st>>>(workspace, blocks, out, 1.0 / ((double)N * Cp * (H - 1) * W),
                                                         1.0 / ((double)N * Cp * H * (W - 1)));
  return check_launch("edge_smooth1_finish");
}

extern "C" int upf_edge_smooth1_bwd(const float* img, int ldi, int Ci, const float* pred, int ldp, int Cp,
                                    const float* grad_out, float* grad_pred, int ldg, int N, int H, int W, void* stream) {
  using namespace upf;
  UPF_REQUIRE(img && pred && grad_out && grad_pred, "edge_smooth1_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 1 && W > 1 && Ci > 0 && Cp > 0 && ldi >= Ci && ldp >= Cp && ldg >= Cp,
              "edge_smooth1_bwd: bad shape");
  edge_smooth1_bwd_kernel<<<loss_blocks((long long)N * H * W), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      img, ldi, Ci, pred, ldp, Cp, grad_out, grad_pred, ldg, N, H, W, (float)(1.0 / ((double)N * Cp * (H - 1) * W)),
      (float)(1.0 / ((double)N * Cp * H * (W - 1))));
  return check_launch("edge_smooth1_bwd");
}

extern "C" int upf_census_loss_fwd(const float* img1, int ld1, const float* img2, int ld2, const float* mask, int ldm,
                                   float* grey, float* dist, float* workspace, float* out, int N, int H, int W,
                                   int max_distance, float q, void* stream) {
  using namespace upf;
  UPF_REQUIRE(img1 && img2 && grey && dist && workspace && out, "census_loss_fwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && ld1 >= 3 && ld2 >= 3 && (!mask || ldm >= 1) && max_distance >= 1 && max_distance <= 8,
              "census_loss_fwd: bad shape (3-channel images, patch radius 1..8)");
  UPF_REQUIRE((reinterpret_cast<uintptr_t>(grey) & 7) == 0, "census_loss_fwd: grey buffer must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const long long npix = (long long)N * H * W;
  const int blocks = loss_blocks(npix);
  census_grey_kernel<<<blocks, LOSS_THREADS, 0, st>>>(img1, ld1, img2, ld2, reinterpret_cast<float2*>(grey), npix);
  int e = check_launch("census_grey");
  if (e) return e;
  census_fwd_kernel<<<blocks, LOSS_THREADS, 0, st>>>(reinterpret_cast<const float2*>(grey), mask, ldm, dist, workspace, N, H, W,
                                                     max_distance, q);
  e = check_launch("census_fwd");
  if (e) return e;
  robust_loss_finish_kernel<<<1, LOSS_THREADS, 0, st>>>(workspace, blocks, out, (double)npix, mask != nullptr, 2.0);
  return check_launch("census_finish");
}

extern "C" int upf_census_loss_bwd(const float* grey, const float* dist, const float* mask, int ldm, const float* out,
                                   const float* grad_out, float* grad_img2, int ldg, int N, int H, int W, int max_distance,
                                   float q, void* stream) {
  using namespace upf;
  UPF_REQUIRE(grey && dist && out && grad_out && grad_img2, "census_loss_bwd: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && ldg >= 3 && (!mask || ldm >= 1) && max_distance >= 1 && max_distance <= 8,
              "census_loss_bwd: bad shape");
  census_bwd_kernel<<<loss_blocks((long long)N * H * W), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float2*>(grey), mask, ldm, dist, out, grad_out, grad_img2, ldg, N, H, W, max_distance, q);
  return check_launch("census_bwd");
}

extern "C" int upf_boundary_warp_fwd(const float* image, int ldi, int C, int Hf, int Wf, const float* flow, int ldf,
                                     const float* start, float* out, int ldo, int N, int h, int w, void* stream) {
  using namespace upf;
  UPF_REQUIRE(image && flow && start && out, "boundary_warp_fwd: null tensor");
  UPF_REQUIRE(N > 0 && h > 0 && w > 0 && Hf > 0 && Wf > 0 && C > 0 && ldi >= C && ldo >= C && ldf >= 2, "boundary_warp_fwd: bad shape");
  bdwarp_fwd_kernel<<<loss_blocks((long long)N * h * w), LOSS_THREADS, 0, (cudaStream_t)stream>>>(image, ldi, C, Hf, Wf, flow, ldf,
                                                                                             start, out, ldo, N, h, w);
  return check_launch("boundary_warp_fwd");
}

extern "C" int upf_boundary_warp_bwd(const float* image, int ldi, int C, int Hf, int Wf, const float* flow, int ldf,
                                     const float* start, const float* grad_out, int ldg, float* grad_flow, int ldgf, int N,
                                     int h, int w, void* stream) {
  using namespace upf;
  UPF_REQUIRE(image && flow && start && grad_out && grad_flow, "boundary_warp_bwd: null tensor");
  UPF_REQUIRE(N > 0 && h > 0 && w > 0 && Hf > 0 && Wf > 0 && C > 0 && ldi >= C && ldg >= C && ldf >= 2 && ldgf >= 2,
              "boundary_warp_bwd: bad shape");
  bdwarp_bwd_kernel<<<loss_blocks((long long)N * h * w), LOSS_THREADS, 0, (cudaStream_t)stream>>>(
      image, ldi, C, Hf, Wf, flow, ldf, start, grad_out, ldg, grad_flow, ldgf, N, h, w);
  return check_launch("boundary_warp_bwd");
}
