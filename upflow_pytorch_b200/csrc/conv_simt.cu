// fp32 SIMT implicit-GEMM convolution (k in {1,3}, any stride / dilation) with
// fused bias + LeakyReLU + residual, reading and writing channel slices of
// pixel-major buffers.
//
// This is the strict-fp32 path (UPF_CONV_FP32): same arithmetic class as the
// reference's nn.Conv2d on CPU (model/pwc_modules.py:10-31), used for parity
// and for the shapes the tensor-core kernel (conv_tc.cu) does not take
// (stride 2).  M = output pixels (flattened n,oy,ox), N = output channels,
// K = taps x input channels.
#include "upf_common.cuh"

namespace upf {

constexpr int CS_BK = 8;
constexpr int CS_NT = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(CS_NT)
conv_simt_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ bias,
                 float* __restrict__ out, int ldo, const float* __restrict__ res, int ldr,
                 int N, int H, int W, int Ho, int Wo, int Cin, int Cout, int cout_pad,
                 int ks, int stride, int dil, float slope, int vec_in, int flags) {
  pdl_prologue();
  static_assert((BM / TM) * (BN / TN) == CS_NT, "thread grid");
  constexpr int A_PER = BM * CS_BK / 4 / CS_NT;      // float4 loads of A per thread per stage
  static_assert(A_PER >= 1, "tile too small");
  __shared__ __align__(16) float As[CS_BK][BM + 4];
  __shared__ __align__(16) float Bs[CS_BK][BN];

  const long long M = (long long)N * Ho * Wo;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int pad = ((ks - 1) * dil) / 2;
  const int t = threadIdx.x;

  // A loader: this thread always stages the same pixel(s)
  int a_m[A_PER], a_kq[A_PER], a_iy0[A_PER], a_ix0[A_PER];
  long long a_img[A_PER];
#pragma unroll
  for (int i = 0; i < A_PER; ++i) {
    const int u = t + i * CS_NT;
    a_m[i] = u % BM;
    a_kq[i] = (u / BM) * 4;
    const long long P = m0 + a_m[i];
    if (P < M) {
      const int ox = (int)(P % Wo);
      const int oy = (int)((P / Wo) % Ho);
      const long long n = P / ((long long)Wo * Ho);
      a_img[i] = n * H;
      a_iy0[i] = oy * stride - pad;
      a_ix0[i] = ox * stride - pad;
    } else {
      a_img[i] = -1;
      a_iy0[i] = a_ix0[i] = 0;
    }
  }
  const int tm = (t % (BM / TM)) * TM;
  const int tn = (t / (BM / TM)) * TN;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int taps = ks * ks;
  for (int tap = 0; tap < taps; ++tap) {
    const int ky = tap / ks, kx = tap - ky * ks;
    for (int c0 = 0; c0 < Cin; c0 += CS_BK) {
      // ---- stage A (activations) transposed to [k][m]
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = c0 + a_kq[i];
        const int iy = a_iy0[i] + ky * dil, ix = a_ix0[i] + kx * dil;
        if (a_img[i] >= 0 && iy >= 0 && iy < H && ix >= 0 && ix < W && c < Cin) {
          const float* p = x + ((size_t)(a_img[i] + iy) * W + ix) * ldx + c;
          if (vec_in && c + 3 < ldx) {
            v = ldg4(p);
            if (c + 1 >= Cin) v.y = 0.f;
            if (c + 2 >= Cin) v.z = 0.f;
            if (c + 3 >= Cin) v.w = 0.f;
          } else {
            v.x = __ldg(p);
            if (c + 1 < Cin) v.y = __ldg(p + 1);
            if (c + 2 < Cin) v.z = __ldg(p + 2);
            if (c + 3 < Cin) v.w = __ldg(p + 3);
          }
        }
        As[a_kq[i] + 0][a_m[i]] = v.x;
        As[a_kq[i] + 1][a_m[i]] = v.y;
        As[a_kq[i] + 2][a_m[i]] = v.z;
        As[a_kq[i] + 3][a_m[i]] = v.w;
      }
      // ---- stage B (weights [tap][cin][cout_pad])
      for (int u = t; u < CS_BK * BN / 4; u += CS_NT) {
        const int k = u / (BN / 4), nq = (u - k * (BN / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = c0 + k, co = n0 + nq;
        if (c < Cin && co < cout_pad) v = ldg4(w + ((size_t)tap * Cin + c) * cout_pad + co);
        *reinterpret_cast<float4*>(&Bs[k][nq]) = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < CS_BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; i += 4) {
          if (TM >= 4) {
            const float4 q = *reinterpret_cast<const float4*>(&As[k][tm + i]);
            a[i] = q.x; a[i + 1] = q.y; a[i + 2] = q.z; a[i + 3] = q.w;
          } else {
            for (int ii = 0; ii < TM; ++ii) a[ii] = As[k][tm + ii];
          }
        }
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 q = *reinterpret_cast<const float4*>(&Bs[k][tn + j]);
          b[j] = q.x; b[j + 1] = q.y; b[j + 2] = q.z; b[j + 3] = q.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // ---- epilogue: bias + LeakyReLU (+ residual), channel-slice store
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long P = m0 + tm + i;
    if (P >= M) continue;
    float* o = out + (size_t)P * ldo;
    const float* r = res ? res + (size_t)P * ldr : nullptr;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tn + j;
      if (co < Cout) {
        float v = lrelu(acc[i][j] + __ldg(bias + co), slope);
        if (r) v += __ldg(r + co);
        o[co] = maybe_round(v, flags);
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
static void launch_simt(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                        const float* res, int ldr, int N, int H, int W, int Ho, int Wo, int Cin, int Cout,
                        int cout_pad, int ks, int stride, int dil, float slope, int vec_in, int flags, cudaStream_t st) {
  const long long M = (long long)N * Ho * Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((cout_pad + BN - 1) / BN));
  UPF_LAUNCH((conv_simt_kernel<BM, BN, TM, TN>), grid, CS_NT, 0, st, x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin,
                                                          Cout, cout_pad, ks, stride, dil, slope, vec_in, flags);
}


// ---------------------------------------------------------------- first layers: 3x3, Cin <= 4 (the RGB image)
// The image convolutions (FeatureExtractor conv 3->16 stride 2, model/pwc_modules.py:122-142, and the full-resolution
// 3->16 of the output-level SGU, model/upflow.py:360-372) have K = 27: the implicit-GEMM tiles above spend their time
// on staging (152 us at 2x375x1242, against 71 MB = 11 us of HBM traffic).  Direct form, persistent CTAs (weights are
// staged in shared memory once): a thread owns C3_PX consecutive output pixels x 4 output channels, so every input
// value it loads (through L1; neighbouring threads share them) and every weight quad is used C3_PX times;
// Cout/4 threads share a pixel group, and a warp's stores are 64-byte runs.
constexpr int C3_PX = 4;

template <int CIN, int STRIDE>
__global__ void __launch_bounds__(256)
conv_c3_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ w, const float* __restrict__ bias,
               float* __restrict__ out, int ldo, const float* __restrict__ res, int ldr,
               int N, int H, int W, int Ho, int Wo, int Cout, int cout_pad, float slope, int flags) {
  pdl_prologue();
  extern __shared__ float s_w[];                         // [9*CIN][cout_pad] then bias[cout_pad]
  for (int i = threadIdx.x; i < 9 * CIN * cout_pad; i += blockDim.x) s_w[i] = __ldg(w + i);
  float* s_b = s_w + 9 * CIN * cout_pad;
  for (int i = threadIdx.x; i < cout_pad; i += blockDim.x) s_b[i] = i < Cout ? __ldg(bias + i) : 0.f;
  __syncthreads();
  constexpr int COLS = (C3_PX - 1) * STRIDE + 3;         // input columns a thread touches
  const int tpp = cout_pad >> 2;                         // threads per pixel group (power of two <= 32)
  const int sh = __ffs(tpp) - 1;
  const bool vec_out = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const int gpb = 256 >> sh;                             // pixel groups per work item
  const int ngx = (Wo + C3_PX - 1) / C3_PX;              // pixel groups per output row
  const int nxb = (ngx + gpb - 1) / gpb;
  const int items = nxb * Ho * N;
  const int c0 = (threadIdx.x & (tpp - 1)) * 4;
  const bool px4 = ldx == 4 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  // (staging the three input rows in shared memory was measured and is SLOWER: 68 vs 56 us at 2x375x1242 -- the two
  //  extra CTA barriers per item cost more than the gathers through L1)
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int xb = item % nxb, row = item / nxb;         // CTA-uniform
    const int oy = row % Ho;
    const long long n = row / Ho;
    const int g = xb * gpb + (threadIdx.x >> sh);
    if (g >= ngx) continue;
    const int ox0 = g * C3_PX;
    float4 acc[C3_PX];
    const float4 b4 = *reinterpret_cast<const float4*>(s_b + c0);
#pragma unroll
    for (int j = 0; j < C3_PX; ++j) acc[j] = b4;
    const int ix0 = ox0 * STRIDE - 1;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * STRIDE + ky - 1;
      if (iy < 0 || iy >= H) continue;                   // warp-uniform (one output row per item)
      const float* prow = x + (((size_t)n * H + iy) * W) * ldx;
      float v[COLS][CIN];
      if (CIN >= 3 && px4) {                             // [N,H,W,4] image buffer: one 16-byte load per pixel instead of three 4-byte ones
#pragma unroll
        for (int cx = 0; cx < COLS; ++cx) {
          const int ix = ix0 + cx;
          const bool in = ix >= 0 && ix < W;
          const float4 q = in ? __ldg(reinterpret_cast<const float4*>(prow + (size_t)ix * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[cx][0] = q.x; v[cx][1] = q.y; v[cx][2] = q.z;
          if (CIN == 4) v[cx][CIN - 1] = q.w;
        }
      } else {
#pragma unroll
        for (int cx = 0; cx < COLS; ++cx) {
          const int ix = ix0 + cx;
          const bool in = ix >= 0 && ix < W;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) v[cx][ci] = in ? __ldg(prow + (size_t)ix * ldx + ci) : 0.f;
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float4 wv = *reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * CIN + ci) * cout_pad + c0);
#pragma unroll
          for (int j = 0; j < C3_PX; ++j) {
            const float a = v[j * STRIDE + kx][ci];
            acc[j].x = fmaf(a, wv.x, acc[j].x); acc[j].y = fmaf(a, wv.y, acc[j].y);
            acc[j].z = fmaf(a, wv.z, acc[j].z); acc[j].w = fmaf(a, wv.w, acc[j].w);
          }
        }
    }
#pragma unroll
    for (int j = 0; j < C3_PX; ++j) {
      const int ox = ox0 + j;
      if (ox >= Wo) break;
      const long long pix = (n * Ho + oy) * Wo + ox;
      float f[4] = {lrelu(acc[j].x, slope), lrelu(acc[j].y, slope), lrelu(acc[j].z, slope), lrelu(acc[j].w, slope)};
      if (res) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < Cout) f[k] += __ldg(res + (size_t)pix * ldr + c0 + k);
      }
      if (flags & UPF_FLAG_ROUND_TF32) {
#pragma unroll
        for (int k = 0; k < 4; ++k) f[k] = round_tf32(f[k]);
      }
      float* o = out + (size_t)pix * ldo + c0;
      if (vec_out && c0 + 4 <= Cout) {
        *reinterpret_cast<float4*>(o) = make_float4(f[0], f[1], f[2], f[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c0 + k < Cout) o[k] = f[k];
      }
    }
  }
}

template <int CIN>
static int launch_c3(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo, const float* res, int ldr,
                     int N, int H, int W, int Ho, int Wo, int Cout, int cout_pad, int stride, float slope, int flags, cudaStream_t st) {
  const int gpb = 256 / (cout_pad >> 2);
  const size_t smem = (size_t)(9 * CIN + 1) * cout_pad * sizeof(float);
  const int ngx = (Wo + C3_PX - 1) / C3_PX;
  const long long items = (long long)((ngx + gpb - 1) / gpb) * Ho * N;
  UPF_REQUIRE(items < (1ll << 31), "conv_c3: too many pixels");
  const long long cap = (long long)UPF_NUM_SMS * 8;
  const unsigned grid = (unsigned)(items < cap ? items : cap);
  if (stride == 1)
    UPF_LAUNCH((conv_c3_kernel<CIN, 1>), grid, 256, smem, st, x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, slope, flags);
  else
    UPF_LAUNCH((conv_c3_kernel<CIN, 2>), grid, 256, smem, st, x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, slope, flags);
  return check_launch("conv_c3");
}

int conv2d_fwd_simt(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                    const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                    float slope, int flags, cudaStream_t st) {
  const int pad = ((ks - 1) * dil) / 2;
  const int Ho = (H + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  UPF_REQUIRE(Ho > 0 && Wo > 0, "conv: empty output");
  const int cout_pad = (Cout + 3) & ~3;
  UPF_REQUIRE(aligned16(w), "conv: weights must be 16-byte aligned");
  const int vec_in = (ldx % 4 == 0) && aligned16(x);
  if (ks == 3 && dil == 1 && (stride == 1 || stride == 2) && Cin <= 4 && (cout_pad == 4 || cout_pad == 8 || cout_pad == 16 || cout_pad == 32)) {
    switch (Cin) {
      case 1: return launch_c3<1>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, stride, slope, flags, st);
      case 2: return launch_c3<2>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, stride, slope, flags, st);
      case 3: return launch_c3<3>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, stride, slope, flags, st);
      default: return launch_c3<4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cout, cout_pad, stride, slope, flags, st);
    }
  }
  if (cout_pad > 64)
    launch_simt<128, 128, 8, 8>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, flags, st);
  else if (cout_pad > 32)
    launch_simt<128, 64, 8, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, flags, st);
  else if (cout_pad > 16)
    launch_simt<128, 32, 4, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, flags, st);
  else if (cout_pad > 4)
    launch_simt<256, 16, 4, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, flags, st);
  else
    launch_simt<256, 4, 1, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, flags, st);
  return check_launch("conv_simt");
}

}  // namespace upf
