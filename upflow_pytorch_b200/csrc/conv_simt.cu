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
                 int ks, int stride, int dil, float slope, int vec_in) {
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
        o[co] = v;
      }
    }
  }
}

template <int BM, int BN, int TM, int TN>
static void launch_simt(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                        const float* res, int ldr, int N, int H, int W, int Ho, int Wo, int Cin, int Cout,
                        int cout_pad, int ks, int stride, int dil, float slope, int vec_in, cudaStream_t st) {
  const long long M = (long long)N * Ho * Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((cout_pad + BN - 1) / BN));
  conv_simt_kernel<BM, BN, TM, TN><<<grid, CS_NT, 0, st>>>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin,
                                                          Cout, cout_pad, ks, stride, dil, slope, vec_in);
}

int conv2d_fwd_simt(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                    const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                    float slope, cudaStream_t st) {
  const int pad = ((ks - 1) * dil) / 2;
  const int Ho = (H + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  UPF_REQUIRE(Ho > 0 && Wo > 0, "conv: empty output");
  const int cout_pad = (Cout + 3) & ~3;
  UPF_REQUIRE(aligned16(w), "conv: weights must be 16-byte aligned");
  const int vec_in = (ldx % 4 == 0) && aligned16(x);
  if (cout_pad > 64)
    launch_simt<128, 128, 8, 8>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, st);
  else if (cout_pad > 32)
    launch_simt<128, 64, 8, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, st);
  else if (cout_pad > 16)
    launch_simt<128, 32, 4, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, st);
  else if (cout_pad > 4)
    launch_simt<256, 16, 4, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, st);
  else
    launch_simt<256, 4, 1, 4>(x, ldx, w, bias, out, ldo, res, ldr, N, H, W, Ho, Wo, Cin, Cout, cout_pad, ks, stride, dil, slope, vec_in, st);
  return check_launch("conv_simt");
}

}  // namespace upf
