// Shared host/device helpers for the upflow_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "../../include/upflow_b200.h"

#define UPF_NUM_SMS 148   // B200: 2 dies x 74 SMs

namespace upf {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void note_kernel(const char* name);   // upf_last_kernel(): which kernel family served the calling thread's last launch

// every launch goes through here so that a bad configuration is reported to
// the caller as a return code (the reference printf's and AT_ERRORs instead:
// correlation_cuda_kernel.cu:383-390, correlation_cuda.cc:81-83)
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  count_launch();
  note_kernel(what);
  return 0;
}

#define UPF_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      upf::set_error(__VA_ARGS__);        \
      return UPF_EINVAL;                  \
    }                                     \
  } while (0)

// Programmatic dependent launch (PDL).  Every forward kernel starts with pdl_prologue(): it lets the NEXT kernel in the
// stream be scheduled while this one runs, and then waits until the PREVIOUS grid has completed and flushed -- so the
// only thing that overlaps is launch latency and CTA scheduling, never data.  UPF_LAUNCH adds the matching launch
// attribute (upf_debug_conv_halo bit 3 turns it off for A/B runs).  Inside a captured graph the edges become
// programmatic dependencies: bit-identical results, a few microseconds less per launch.
extern int g_tc_pdl;
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory");
}
struct PdlLaunch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  PdlLaunch(dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_tc_pdl ? 1 : 0;
  }
};
#define UPF_LAUNCH(kernel, grid, block, smem, st, ...)                      \
  do {                                                                      \
    upf::PdlLaunch _l(dim3(grid), dim3(block), (size_t)(smem), (st));       \
    (void)cudaLaunchKernelEx(&_l.cfg, kernel, __VA_ARGS__);                 \
  } while (0)

// cudaFuncSetAttribute is per DEVICE: a once-per-process flag would leave the second GPU of a process (nn.DataParallel,
// tools.abstract_model.choose_gpu) at the 48 KB default.  One bit per device ordinal, set after the first success.
struct PerDeviceOnce {
  unsigned long long done = 0;
  bool need() const { int d = 0; cudaGetDevice(&d); return d >= 64 || !((done >> d) & 1ull); }
  void mark() { int d = 0; cudaGetDevice(&d); if (d < 64) done |= 1ull << d; }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ float lrelu(float v, float slope) { return v < 0.f ? v * slope : v; }
// nearest TF32 value, ties away from zero (cvt.rna.tf32.f32): what a tensor-core consumer should find in the top 19 bits
__host__ __device__ __forceinline__ float round_tf32(float v) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u);
#else
  uint32_t u; memcpy(&u, &v, 4); u = (u + 0x1000u) & 0xffffe000u; memcpy(&v, &u, 4); return v;
#endif
}
__device__ __forceinline__ float maybe_round(float v, int flags) { return (flags & UPF_FLAG_ROUND_TF32) ? round_tf32(v) : v; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- storage types of the low-precision variants (fp16 / bf16 STORAGE, fp32 arithmetic; the reference
// dispatches its correlation kernels on Half too, correlation_cuda_kernel.cu:352).  ld4 / st4: four consecutive elements
// (16-byte aligned for float, 8-byte for the 2-byte types); ld1 / st1: one element.
__device__ __forceinline__ float4 ld4(const float* p) { return ldg4(p); }
__device__ __forceinline__ float4 ld4(const __half* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));      // bf16 -> fp32 is a 16-bit shift
  return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                     __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st1(float* p, float v) { *p = v; }
__device__ __forceinline__ void st1(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void st1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
// two consecutive elements (4-byte aligned for the 2-byte types)
__device__ __forceinline__ void st2(float* p, float a, float b) { p[0] = a; p[1] = b; }
__device__ __forceinline__ void st2(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }
__device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__half* p, float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  uint2 r; r.x = *reinterpret_cast<const unsigned*>(&a); r.y = *reinterpret_cast<const unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r; r.x = *reinterpret_cast<const unsigned*>(&a); r.y = *reinterpret_cast<const unsigned*>(&b);
  *reinterpret_cast<uint2*>(p) = r;
}
// vector accesses of T need 4 * sizeof(T) alignment
template <typename T> inline bool aligned_vec4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & (4 * sizeof(T) - 1)) == 0; }

// mean / rstd of one (image, channel) from accumulated double sums; the
// reference uses torch.mean and the UNBIASED torch.var (model/upflow.py:113-114)
// and std = sqrt(var + 1e-16) (:126)
__device__ __forceinline__ void stats_to_mean_std(const double* s, double npix, float& mean, float& stdv) {
  double m = s[0] / npix;
  double var = (s[1] - s[0] * m) / (npix - 1.0);
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  // the reference rounds var to fp32, adds 1e-16 in fp32, takes an fp32 sqrt and divides
  float varf = (float)var;
  stdv = sqrtf(varf + 1e-16f);   // callers DIVIDE by it, like the reference (f / std)
}

// ---- bilinear sampling coordinates, bit-faithful to the reference --------
// pixel + flow -> [-1,1] with three separately rounded ops (pwc_modules.py:195-198),
// then ATen's grid_sampler_unnormalize (native/cuda/GridSampler.cuh:23-31); for
// align_corners=False the CPU and CUDA ATen kernels both evaluate
// ((g+1)*size-1)/2 with one rounding (fused multiply-add).
__device__ __forceinline__ float sample_coord(float pix, float disp, int size, int align_corners) {
  float denom = (float)(size > 1 ? size - 1 : 1);
  float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, __fadd_rn(pix, disp)), denom), 1.0f);
  float gp1 = __fadd_rn(g, 1.0f);
  if (align_corners) return __fmul_rn(__fmul_rn(gp1, 0.5f), (float)(size - 1));
  return __fmaf_rn(gp1, 0.5f * (float)size, -0.5f);
}

struct BilinearTaps {
  int x0, y0;            // north-west corner (x1 = x0+1, y1 = y0+1)
  float w_nw, w_ne, w_sw, w_se;
  bool in_nw, in_ne, in_sw, in_se;
  float wsum;            // sum of in-bounds weights, accumulated nw, ne, sw, se from 0
};

__device__ __forceinline__ BilinearTaps bilinear_taps(float ix, float iy, int H, int W) {
  BilinearTaps t;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float fx1 = __fadd_rn(fx0, 1.0f), fy1 = __fadd_rn(fy0, 1.0f);
  t.w_nw = __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(fy1, iy));
  t.w_ne = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(fy1, iy));
  t.w_sw = __fmul_rn(__fsub_rn(fx1, ix), __fsub_rn(iy, fy0));
  t.w_se = __fmul_rn(__fsub_rn(ix, fx0), __fsub_rn(iy, fy0));
  // clamp before the int conversion so wild flows cannot overflow
  float cx = fminf(fmaxf(fx0, -2.0f), (float)W + 1.0f);
  float cy = fminf(fmaxf(fy0, -2.0f), (float)H + 1.0f);
  t.x0 = (int)cx;
  t.y0 = (int)cy;
  bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
  bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
  // NaN coordinates sample nothing
  bool ok = (ix == ix) && (iy == iy);
  t.in_nw = ok && xin0 && yin0;
  t.in_ne = ok && xin1 && yin0;
  t.in_sw = ok && xin0 && yin1;
  t.in_se = ok && xin1 && yin1;
  float s = 0.0f;
  if (t.in_nw) s = __fadd_rn(s, t.w_nw);
  if (t.in_ne) s = __fadd_rn(s, t.w_ne);
  if (t.in_sw) s = __fadd_rn(s, t.w_sw);
  if (t.in_se) s = __fadd_rn(s, t.w_se);
  t.wsum = s;
  return t;
}

// align_corners=True source tap of F.interpolate(bilinear)
// (ATen native/cuda/UpSample.cuh:100-124: scale=(in-1)/(out-1), src=scale*dst)
struct AxisTap { int i0, i1; float l0, l1; };
__device__ __forceinline__ AxisTap axis_tap(int dst, int n_in, float scale) {
  AxisTap t;
  float src = __fmul_rn(scale, (float)dst);
  t.i0 = (int)src;
  if (t.i0 > n_in - 1) t.i0 = n_in - 1;
  t.i1 = t.i0 + (t.i0 < n_in - 1 ? 1 : 0);
  t.l1 = __fsub_rn(src, (float)t.i0);
  t.l0 = __fsub_rn(1.0f, t.l1);
  return t;
}
inline float host_ac_scale(int n_in, int n_out) { return n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f; }

}  // namespace upf
