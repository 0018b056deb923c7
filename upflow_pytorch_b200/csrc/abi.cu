// C-ABI plumbing: error reporting, launch accounting, conv dispatch.
#include "upf_common.cuh"

#include <atomic>

namespace upf {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
static thread_local const char* g_last_kernel = "";
void note_kernel(const char* name) { g_last_kernel = name; }
const char* last_kernel() { return g_last_kernel; }

int conv2d_fwd_simt(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                    const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                    float slope, int flags, cudaStream_t st);
int conv2d_fwd_tc(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                  const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                  float slope, int flags, cudaStream_t st);

}  // namespace upf

extern "C" int upf_abi_version(void) { return UPF_ABI_VERSION; }
extern "C" const char* upf_last_error(void) { return upf::g_err; }
extern "C" const char* upf_last_kernel(void) { return upf::last_kernel(); }
extern "C" long long upf_launch_count(void) { return upf::g_launches.load(std::memory_order_relaxed); }

extern "C" int upf_conv2d_fwd(const float* x, int ldx, const float* w, const float* bias, float* out, int ldo,
                              const float* residual, int ldr, int N, int H, int W, int Cin, int Cout, int ksize,
                              int stride, int dilation, float slope, int precision, void* stream) {
  using namespace upf;
  UPF_REQUIRE(x && w && bias && out, "conv: null tensor");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, "conv: bad shape");
  UPF_REQUIRE(ksize == 1 || ksize == 3, "conv: kernel size %d not in {1,3}", ksize);
  UPF_REQUIRE(stride >= 1 && dilation >= 1, "conv: bad stride/dilation");
  UPF_REQUIRE(ldx >= Cin && ldo >= Cout && (residual == nullptr || ldr >= Cout), "conv: pitch smaller than channels");
  const int flags = (precision & UPF_CONV_ROUND_OUT) ? UPF_FLAG_ROUND_TF32 : 0;
  precision &= ~UPF_CONV_ROUND_OUT;
  if (precision == UPF_CONV_FP32)
    return conv2d_fwd_simt(x, ldx, w, bias, out, ldo, residual, ldr, N, H, W, Cin, Cout, ksize, stride, dilation,
                           slope, flags, (cudaStream_t)stream);
  if (precision == UPF_CONV_TF32) {
    return conv2d_fwd_tc(x, ldx, w, bias, out, ldo, residual, ldr, N, H, W, Cin, Cout, ksize, stride, dilation, slope,
                         flags, (cudaStream_t)stream);
  }
  set_error("conv: unknown precision %d", precision);
  return UPF_EINVAL;
}
