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

// out[0] = the term, out[1] = 1/denominator (what backward scales by).  mode 0: mean over npix*C; mode 1: masked,
// sum(d*mask)/(sum(mask)+1e-6).
__global__ void __launch_bounds__(LOSS_THREADS)
robust_loss_finish_kernel(const float* __restrict__ part, int blocks, float* __restrict__ out, double count, int masked) {
  __shared__ double s_a[LOSS_THREADS], s_b[LOSS_THREADS];
  double a = 0.0, b = 0.0;
  for (int k = threadIdx.x; k < blocks; k += LOSS_THREADS) { a += part[k]; b += part[blocks + k]; }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int k = 0; k < LOSS_THREADS; ++k) { ta += s_a[k]; tb += s_b[k]; }
    const double inv = masked ? 1.0 / (tb + 1e-6) : 1.0 / count;
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
  robust_loss_finish_kernel<<<1, LOSS_THREADS, 0, st>>>(workspace, blocks, out, (double)npix * C, mask != nullptr);
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
