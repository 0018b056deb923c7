// Flow resize (bilinear, align_corners=True, x rate), the self-guided-upsample
// blend, and layout copies.
//
// upf_resize_bilinear replaces upsample2d_flow_as / upsample2d_as
//   (model/pwc_modules.py:72-90: interpolate + chunk + 2 mul + cat).
// upf_sgu_blend replaces the tail of sgu_model.forward (model/upflow.py:79-88):
//   sigmoid, [two bilinear upsamples + rate], tools.torch_warp (host mesh + H2D
//   copy + grid_sample, utils/tools.py:1274-1304) and the 4-op blend, in one
//   pass that reads 5 low-resolution channels and 2 output-resolution channels
//   per pixel and writes 2.
#include "upf_common.cuh"

namespace upf {

__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const float* __restrict__ in, int ldi, int h, int w, float* __restrict__ out, int ldo,
                       int H, int W, int N, int C, float sh, float sw, float4 mul) {
  pdl_prologue();
  const long long total = (long long)N * H * W;
  const float m[4] = {mul.x, mul.y, mul.z, mul.w};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    const AxisTap ty = axis_tap(y, h, sh), tx = axis_tap(x, w, sw);
    const float* r0 = in + ((size_t)(n * h + ty.i0) * w) * ldi;
    const float* r1 = in + ((size_t)(n * h + ty.i1) * w) * ldi;
    float* o = out + (size_t)i * ldo;
    for (int c = 0; c < C; ++c) {
      const float v00 = __ldg(r0 + (size_t)tx.i0 * ldi + c), v01 = __ldg(r0 + (size_t)tx.i1 * ldi + c);
      const float v10 = __ldg(r1 + (size_t)tx.i0 * ldi + c), v11 = __ldg(r1 + (size_t)tx.i1 * ldi + c);
      // ATen upsample_bilinear2d: h0l*(w0l*v00 + w1l*v01) + h1l*(w0l*v10 + w1l*v11)
      const float top = __fadd_rn(__fmul_rn(tx.l0, v00), __fmul_rn(tx.l1, v01));
      const float bot = __fadd_rn(__fmul_rn(tx.l0, v10), __fmul_rn(tx.l1, v11));
      const float v = __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
      o[c] = __fmul_rn(v, m[c]);
    }
  }
}

__device__ __forceinline__ float sigmoidf_ref(float v) { return 1.0f / (1.0f + expf(-v)); }

// bilinear (align_corners=True) read of channel c of a low-resolution tensor
__device__ __forceinline__ float lowres_tap(const float* r0, const float* r1, const AxisTap& ty, const AxisTap& tx,
                                            int ldi, int c) {
  const float v00 = __ldg(r0 + (size_t)tx.i0 * ldi + c), v01 = __ldg(r0 + (size_t)tx.i1 * ldi + c);
  const float v10 = __ldg(r1 + (size_t)tx.i0 * ldi + c), v11 = __ldg(r1 + (size_t)tx.i1 * ldi + c);
  const float top = __fadd_rn(__fmul_rn(tx.l0, v00), __fmul_rn(tx.l1, v01));
  const float bot = __fadd_rn(__fmul_rn(tx.l0, v10), __fmul_rn(tx.l1, v11));
  return __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
}

__global__ void __launch_bounds__(256)
sgu_blend_kernel(const float* __restrict__ flow_init, int ldf, const float* __restrict__ inter, int ldi, int ih, int iw,
                 float* __restrict__ out, int ldo, float* __restrict__ out2, int ld2, int N, int H, int W, int align_corners,
                 float sh, float sw, float rate_u, float rate_v, int flags) {
  pdl_prologue();
  const long long total = (long long)N * H * W;
  const bool same = (ih == H && iw == W);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const long long n = i / ((long long)W * H);
    float du, dv, m;
    if (same) {
      const float* q = inter + (size_t)i * ldi;
      du = __ldg(q); dv = __ldg(q + 1); m = sigmoidf_ref(__ldg(q + 2));
    } else {
      // inter_flow: upsample then x (W/iw, H/ih); mask: sigmoid at LOW resolution,
      // then upsample (model/upflow.py:82-86)
      const AxisTap ty = axis_tap(y, ih, sh), tx = axis_tap(x, iw, sw);
      const float* r0 = inter + ((size_t)(n * ih + ty.i0) * iw) * ldi;
      const float* r1 = inter + ((size_t)(n * ih + ty.i1) * iw) * ldi;
      du = __fmul_rn(lowres_tap(r0, r1, ty, tx, ldi, 0), rate_u);
      dv = __fmul_rn(lowres_tap(r0, r1, ty, tx, ldi, 1), rate_v);
      const float m00 = sigmoidf_ref(__ldg(r0 + (size_t)tx.i0 * ldi + 2)), m01 = sigmoidf_ref(__ldg(r0 + (size_t)tx.i1 * ldi + 2));
      const float m10 = sigmoidf_ref(__ldg(r1 + (size_t)tx.i0 * ldi + 2)), m11 = sigmoidf_ref(__ldg(r1 + (size_t)tx.i1 * ldi + 2));
      const float top = __fadd_rn(__fmul_rn(tx.l0, m00), __fmul_rn(tx.l1, m01));
      const float bot = __fadd_rn(__fmul_rn(tx.l0, m10), __fmul_rn(tx.l1, m11));
      m = __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
    }
    // torch_warp(flow_init, inter_flow): bilinear, zeros padding, NO validity mask
    const float ix = sample_coord((float)x, du, W, align_corners);
    const float iy = sample_coord((float)y, dv, H, align_corners);
    const BilinearTaps t = bilinear_taps(ix, iy, H, W);
    const long long base = (n * H + t.y0) * (long long)W + t.x0;
    float wu = 0.f, wv = 0.f;
    if (t.in_nw) { const float* q = flow_init + base * ldf; wu = fmaf(__ldg(q), t.w_nw, wu); wv = fmaf(__ldg(q + 1), t.w_nw, wv); }
    if (t.in_ne) { const float* q = flow_init + (base + 1) * ldf; wu = fmaf(__ldg(q), t.w_ne, wu); wv = fmaf(__ldg(q + 1), t.w_ne, wv); }
    if (t.in_sw) { const float* q = flow_init + (base + W) * ldf; wu = fmaf(__ldg(q), t.w_sw, wu); wv = fmaf(__ldg(q + 1), t.w_sw, wv); }
    if (t.in_se) { const float* q = flow_init + (base + W + 1) * ldf; wu = fmaf(__ldg(q), t.w_se, wu); wv = fmaf(__ldg(q + 1), t.w_se, wv); }
    const float* f0 = flow_init + (size_t)i * ldf;
    const float u0 = __ldg(f0), v0 = __ldg(f0 + 1);
    const float om = __fsub_rn(1.0f, m);
    // warp*(1-m) + init*m, two products then one add (model/upflow.py:88)
    const float ru = __fadd_rn(__fmul_rn(wu, om), __fmul_rn(u0, m)), rv = __fadd_rn(__fmul_rn(wv, om), __fmul_rn(v0, m));
    out[(size_t)i * ldo + 0] = ru;
    out[(size_t)i * ldo + 1] = rv;
    if (out2) {                      // the tensor-core consumers' copy: 4-channel slot (u, v, 0, 0)
      float* o2 = out2 + (size_t)i * ld2;
      o2[0] = maybe_round(ru, flags); o2[1] = maybe_round(rv, flags); o2[2] = 0.f; o2[3] = 0.f;
    }
  }
}

// ---- layout copies ---------------------------------------------------------
// [N,C,H,W] planes <-> pixel-major rows through a 32x33 shared tile so both
// sides stay coalesced.
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int ldo, int C, int HW) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    tile[k][tx] = (c < C && p < HW) ? __ldg(in + ((size_t)n * C + c) * HW + p) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    if (p < HW && c < C) out[((size_t)n * HW + p) * ldo + c] = tile[tx][k];
  }
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int C, int HW) {
  pdl_prologue();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int p = p0 + k, c = c0 + tx;
    tile[k][tx] = (p < HW && c < C) ? __ldg(in + ((size_t)n * HW + p) * ldi + c) : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k, p = p0 + tx;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[tx][k];
  }
}

__global__ void __launch_bounds__(256)
copy_channels_kernel(const float* __restrict__ in, int ldi, float* __restrict__ out, int ldo, long long npix, int C,
                     int vec, int flags) {
  pdl_prologue();
  if (vec) {
    const int cg = C >> 2;
    const long long total = npix * cg;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long p = i / cg;
      const int c = (int)(i - p * cg) * 4;
      float4 v = in ? ldg4(in + (size_t)p * ldi + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (flags & UPF_FLAG_ROUND_TF32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
      *reinterpret_cast<float4*>(out + (size_t)p * ldo + c) = v;
    }
  } else {
    const long long total = npix * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
      const long long p = i / C;
      const int c = (int)(i - p * C);
      out[(size_t)p * ldo + c] = in ? maybe_round(__ldg(in + (size_t)p * ldi + c), flags) : 0.f;
    }
  }
}

static unsigned grid_for(long long total, int per_block = 256) {
  long long b = (total + per_block - 1) / per_block;
  if (b > UPF_NUM_SMS * 16) b = UPF_NUM_SMS * 16;
  if (b < 1) b = 1;
  return (unsigned)b;
}


// ---- 3x3 convolution with a handful of output channels as "expand, then combine the taps" ----------------------
// A 3x3 conv with Cout <= 8 wastes the tensor pipe: every tap issues its own MMAs with N padded to 16, nine times
// the instruction count of the products it needs.  Instead the tensor-core kernel runs it as a 1x1 convolution with
// 9*Cout output channels, Y[p][tap*Cout+co] = sum_ci X[p][ci] * W[co][ci][tap]  (one pass over X, no halo), and this
// kernel gathers the nine shifted partial results:
//   out[n,y,x,co] = act( bias[co] + sum_tap Y[n, y+(ky-1)*dil, x+(kx-1)*dil, tap*Cout+co] ) (+ residual)
// with taps whose source pixel lies outside the image contributing nothing (= the conv's zero padding).
__global__ void __launch_bounds__(256)
tap_combine_kernel(const float* __restrict__ y, int ldy, const float* __restrict__ bias, float* __restrict__ out, int ldo,
                   const float* __restrict__ res, int ldr, int N, int H, int W, int Cout, int dil, float slope, int flags) {
  pdl_prologue();
  const long long total = (long long)N * H * W * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long long pix = i / Cout;
    const int x = (int)(pix % W);
    const long long t = pix / W;
    const int yy = (int)(t % H);
    const long long n = t / H;
    float acc = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = yy + (ky - 1) * dil;
      if (sy < 0 || sy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = x + (kx - 1) * dil;
        if (sx < 0 || sx >= W) continue;
        acc += __ldg(y + ((size_t)((size_t)n * H + sy) * W + sx) * ldy + (ky * 3 + kx) * Cout + co);
      }
    }
    float v = lrelu(acc + __ldg(bias + co), slope);
    if (res) v += __ldg(res + (size_t)pix * ldr + co);
    out[(size_t)pix * ldo + co] = maybe_round(v, flags);
  }
}

}  // namespace upf

extern "C" int upf_resize_bilinear(const float* in, int ldi, int h, int w, float* out, int ldo, int H, int W,
                                   int N, int C, const float* scale_host, void* stream) {
  using namespace upf;
  UPF_REQUIRE(in && out, "resize: null tensor");
  UPF_REQUIRE(N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C >= 1 && C <= 4 && ldi >= C && ldo >= C, "resize: bad shape (C<=4)");
  float4 mul = make_float4(1.f, 1.f, 1.f, 1.f);
  if (scale_host) {
    float* m = &mul.x;
    for (int c = 0; c < C; ++c) m[c] = scale_host[c];
  }
  UPF_LAUNCH((resize_bilinear_kernel), grid_for((long long)N * H * W), 256, 0, (cudaStream_t)stream, 
      in, ldi, h, w, out, ldo, H, W, N, C, host_ac_scale(h, H), host_ac_scale(w, W), mul);
  return check_launch("resize_bilinear");
}

extern "C" int upf_sgu_blend(const float* flow_init, int ldf, const float* inter, int ldi, int ih, int iw,
                             float* out, int ldo, float* out_tf32, int ldt, int N, int H, int W, int align_corners, int flags,
                             void* stream) {
  using namespace upf;
  UPF_REQUIRE(flow_init && inter && out, "sgu_blend: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && ih > 0 && iw > 0 && ldf >= 2 && ldi >= 3 && ldo >= 2, "sgu_blend: bad shape");
  UPF_REQUIRE(out != flow_init, "sgu_blend: cannot run in place (neighbouring pixels are gathered)");
  UPF_REQUIRE(out_tf32 == nullptr || ldt >= 4, "sgu_blend: the second output is a 4-channel slot");
  // rate = ratio of SIZES as python floats (model/pwc_modules.py:84-85), rounded to fp32 at the multiply
  const float rate_u = (float)((double)W / (double)iw), rate_v = (float)((double)H / (double)ih);
  UPF_LAUNCH((sgu_blend_kernel), grid_for((long long)N * H * W), 256, 0, (cudaStream_t)stream, 
      flow_init, ldf, inter, ldi, ih, iw, out, ldo, out_tf32, ldt, N, H, W, align_corners, host_ac_scale(ih, H), host_ac_scale(iw, W),
      rate_u, rate_v, flags);
  return check_launch("sgu_blend");
}

extern "C" int upf_nchw_to_nhwc(const float* in, float* out, int ldo, int N, int C, int H, int W, void* stream) {
  using namespace upf;
  UPF_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0 && ldo >= C, "nchw_to_nhwc: bad argument");
  UPF_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "nchw_to_nhwc: N or C too large");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  UPF_LAUNCH((nchw_to_nhwc_kernel), grid, 256, 0, (cudaStream_t)stream, in, out, ldo, C, HW);
  return check_launch("nchw_to_nhwc");
}

extern "C" int upf_nhwc_to_nchw(const float* in, int ldi, float* out, int N, int C, int H, int W, void* stream) {
  using namespace upf;
  UPF_REQUIRE(in && out && N > 0 && C > 0 && H > 0 && W > 0 && ldi >= C, "nhwc_to_nchw: bad argument");
  UPF_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, "nhwc_to_nchw: N or C too large");
  const int HW = H * W;
  dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  UPF_LAUNCH((nhwc_to_nchw_kernel), grid, 256, 0, (cudaStream_t)stream, in, ldi, out, C, HW);
  return check_launch("nhwc_to_nchw");
}

extern "C" int upf_copy_channels(const float* in, int ldi, float* out, int ldo, long long npix, int C, int flags, void* stream) {
  using namespace upf;
  UPF_REQUIRE(out && npix > 0 && C > 0 && (!in || ldi >= C) && ldo >= C, "copy_channels: bad argument");
  const int vec = (C % 4 == 0) && (!in || ((ldi % 4 == 0) && aligned16(in))) && (ldo % 4 == 0) && aligned16(out);
  UPF_LAUNCH((copy_channels_kernel), grid_for(npix * (vec ? C / 4 : C)), 256, 0, (cudaStream_t)stream, in, ldi, out, ldo, npix, C, vec, flags);
  return check_launch("copy_channels");
}

extern "C" int upf_conv3x3_tap_combine(const float* y, int ldy, const float* bias, float* out, int ldo, const float* residual,
                                       int ldr, int N, int H, int W, int Cout, int dilation, float slope, int flags, void* stream) {
  using namespace upf;
  UPF_REQUIRE(y && bias && out, "tap_combine: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && Cout > 0 && dilation >= 1 && ldy >= 9 * Cout && ldo >= Cout && (!residual || ldr >= Cout),
              "tap_combine: bad shape");
  UPF_LAUNCH((tap_combine_kernel), grid_for((long long)N * H * W * Cout), 256, 0, (cudaStream_t)stream, y, ldy, bias, out, ldo,
             residual, ldr, N, H, W, Cout, dilation, slope, flags);
  return check_launch("tap_combine");
}
