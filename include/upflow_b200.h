/*
 * upflow_b200 -- C ABI of the B200-native UPFlow decoder hot path.
 *
 * Drop-in boundary.  The reference's only native
 * boundary is the pybind module `correlation_cuda`
 * (model/correlation_package/correlation_cuda.cc:169-172: forward/backward);
 * everything else on the hot path is ATen called from Python
 * (model/pwc_modules.py, model/upflow.py, utils/tools.py).  Each entry point
 * below names the reference routine it replaces.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes, no torch types.  Every pointer is a
 *    DEVICE pointer owned by the caller; the library never allocates, never
 *    synchronises and never throws.  Work is enqueued on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream).
 *  - return 0 on success, otherwise a negative UPF_E* code or a positive
 *    cudaError_t; upf_last_error() returns a thread-local message.
 *  - tensors are fp32, PIXEL-MAJOR ("NHWC"): element (n, y, x, c) of a tensor
 *    with row pitch `ld` lives at base[((n*H + y)*W + x)*ld + c].  `ld` may be
 *    larger than the channel count, so a tensor can be a channel slice of a
 *    wider buffer (this is how the dense blocks avoid every torch.cat of
 *    model/pwc_modules.py:280-284).  torch `channels_last` tensors are this
 *    layout with ld == C.
 *  - "stats" buffers are double[N][C][2] = (sum x, sum x*x) over the H*W
 *    pixels of one image and channel; they are ACCUMULATED into (caller zeroes
 *    them, e.g. with cudaMemsetAsync on the same stream).
 */
#ifndef UPFLOW_B200_H
#define UPFLOW_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define UPF_ABI_VERSION 2

#define UPF_EINVAL   (-1)  /* bad argument (shape, alignment, unsupported value) */
#define UPF_ENOTSUP  (-2)  /* valid request this build cannot serve            */
#define UPF_EDRIVER  (-3)  /* CUDA driver entry point / tensor-map failure     */

/* `flags` bit of the producers below: store the result ROUNDED TO THE NEAREST TF32 value (10 mantissa bits, ties away
 * from zero -- cvt.rna.tf32.f32).  tcgen05 kind::tf32 reads the top 19 bits of an fp32 operand, i.e. TRUNCATES; a
 * buffer whose only consumers are tensor-core convolutions is therefore written pre-rounded by whoever produces it
 * (unbiased operand error 2^-12 instead of a one-sided 2^-11: mean EPE vs the fp32 reference 5.6e-4 px instead of
 * 3.8e-3 at KITTI size with the shipped checkpoint, robust-mask diagnostic).  Weights are rounded when packed. */
#define UPF_FLAG_ROUND_TF32 1

int         upf_abi_version(void);
const char* upf_last_error(void);
/* number of kernels this library has launched in the calling process */
long long   upf_launch_count(void);
/* name of the kernel family that served the calling thread's most recent launch ("conv_win", "conv_halo",
 * "conv_tc", "conv_simt", "conv_c3", "corr_pipe", "corr_fwd", "corr_small", ...): entry points such as
 * upf_conv2d_fwd and upf_corr_lrelu_fwd choose a kernel by shape, and measurements attribute time with this */
const char* upf_last_kernel(void);

/* a1+a2+a3 (+ apply half of a5): cost volume + LeakyReLU, optionally with the
 * per-image per-channel normalisation fused into the operand load.
 *   replaces correlation_cuda.forward (correlation_cuda.cc:10-87, kernels
 *   correlation_cuda_kernel.cu:15-114), Corr_pyTorch.forward
 *   (utils/pytorch_correlation.py:27-50), nn.LeakyReLU on the result
 *   (model/upflow.py:563-564) and normalize_features' apply step
 *   (model/upflow.py:126-135).
 * out[n,y,x,(dy+d)*(2d+1)+(dx+d)] = lrelu( (1/C) sum_c a[n,y,x,c]*b[n,y+dy,x+dx,c] ),
 * b zero outside the image; a=(f1-mean1)*rstd1, b=(f2-mean2)*rstd2 when the
 * stats pointers are non-NULL (mean = s/npix, var = unbiased, rstd=1/sqrt(var+1e-16)).
 * max_disp in {1..6}; slope = 1.0f disables the activation.
 * f2_batch_shift: image n of f1 is matched with image (n + shift) % N of f2 /
 * stats2 (the decoder stacks the forward and backward directions in one batch:
 * images [im1.., im2..], partner = shift by N/2; 0 = plain). */
int upf_corr_lrelu_fwd(const float* f1, int ld1, const float* f2, int ld2,
                       float* out, int ldo, int N, int H, int W, int C, int max_disp,
                       const double* stats1, const double* stats2, int f2_batch_shift,
                       float slope, int flags, void* stream);

/* a1+a2+a3 in the reference operator's OWN layout: planar ("NCHW") feature maps in, planar cost volume out
 * (correlation_cuda.forward: input1, input2 [B,C,H,W] -> output [B,(2d+1)^2,H,W], correlation_cuda.cc:10-87;
 * Corr_pyTorch.forward, utils/pytorch_correlation.py:27-50) -- no layout conversion on either side.
 * pitch*[3] = {row pitch, plane (channel) pitch, image pitch} in ELEMENTS; a contiguous [N,C,H,W] tensor has
 * {W, H*W, C*H*W}.  The operands travel by TMA: the pitches of f1 / f2 must be multiples of 4 elements and their base
 * pointers 16-byte aligned, and so must the output's (it leaves by TMA tensor stores); else UPF_ENOTSUP: convert to
 * pixel-major and call upf_corr_lrelu_fwd.
 * out[n,(dy+d)*(2d+1)+(dx+d),y,x] = lrelu( (1/C) sum_c f1[n,c,y,x] * f2[(n+shift)%N,c,y+dy,x+dx] ), f2 zero outside.
 * max_disp <= 4 (5, 6: UPF_ENOTSUP, use the pixel-major entry point).  Bits 8..13 of `flags` are ablation switches of
 * this kernel for tools/dbg_planar_time.py (they skip phases; never set them in production). */
int upf_corr_lrelu_fwd_planar(const float* f1, const long long* pitch1, const float* f2, const long long* pitch2,
                              float* out, const long long* pitch_out, int N, int H, int W, int C, int max_disp,
                              int f2_batch_shift, float slope, int flags, void* stream);

/* fp16 / bf16 STORAGE variants (the reference dispatches its correlation on Half too,
 * correlation_cuda_kernel.cu:352; F.grid_sample takes half tensors).  Tensors are pixel-major 2-byte elements (pitches in
 * ELEMENTS), every product and sum is fp32, one rounding at the store.  dtype: UPF_DTYPE_F16 / UPF_DTYPE_BF16.
 * Same semantics as upf_corr_lrelu_fwd / upf_warp_fwd otherwise (the flow, the mask arithmetic and the moments are fp32). */
#define UPF_DTYPE_F16 1
#define UPF_DTYPE_BF16 2
int upf_corr_lrelu_fwd_lp(const void* f1, int ld1, const void* f2, int ld2, void* out, int ldo, int dtype,
                          int N, int H, int W, int C, int max_disp, const double* stats1, const double* stats2,
                          int f2_batch_shift, float slope, void* stream);
int upf_warp_fwd_lp(const void* x, int ldx, const float* flow, int ldf, void* out, int ldo, int dtype,
                    int N, int H, int W, int C, int align_corners, float mask_threshold, int x_batch_shift,
                    double* stats, void* stream);

/* a11: gradients of the (un-normalised) cost volume wrt f1 and f2.
 *   replaces correlation_cuda.backward (correlation_cuda.cc:89-167, kernels
 *   correlation_cuda_kernel.cu:116-300).  `out` is the saved forward output;
 *   when slope != 1 the LeakyReLU derivative is taken from its sign. */
int upf_corr_lrelu_bwd(const float* f1, int ld1, const float* f2, int ld2,
                       const float* out, int ldo, const float* grad_out, int ldg,
                       float* grad_f1, int ldg1, float* grad_f2, int ldg2,
                       int N, int H, int W, int C, int max_disp, float slope, void* stream);

/* a4 (+ stats half of a5): bilinear warp by a pixel-space flow with the
 * `mask >= 1.0` validity mask, bit-faithful to F.grid_sample as the reference
 * calls it.
 *   replaces WarpingLayer_no_div.forward (model/pwc_modules.py:184-207) and,
 *   with mask_threshold<=0, tools.torch_warp (utils/tools.py:1274-1304).
 * mask_threshold: a pixel is kept when the sum of its in-bounds bilinear weights
 * is >= mask_threshold.  1.0f is the reference; 0.9999f is the DIAGNOSTIC
 * "robust mask" used to show parity without the reference's 1-ulp mask flips
 * (utils/tools.py:1311 has the same idea commented out); <= 0 disables the mask.
 * flow is [N,H,W,>=2] (u,v) with pitch ldf.  If stats != NULL the (sum, sum^2)
 * of the OUTPUT are accumulated per (n,c).  x_batch_shift: output image n
 * samples image (n + shift) % N of x (see upf_corr_lrelu_fwd). */
int upf_warp_fwd(const float* x, int ldx, const float* flow, int ldf, float* out, int ldo,
                 int N, int H, int W, int C, int align_corners, float mask_threshold,
                 int x_batch_shift, double* stats, int flags, void* stream);

/* a11: backward of upf_warp_fwd wrt x (scatter-add; grad_x must be zeroed by the
 * caller) and wrt flow (grad_flow [N,H,W,2], written).  Either may be NULL. */
int upf_warp_bwd(const float* x, int ldx, const float* flow, int ldf, const float* grad_out, int ldg,
                 float* grad_x, int ldgx, float* grad_flow, int ldgf,
                 int N, int H, int W, int C, int align_corners, float mask_threshold, void* stream);

/* Forward/backward consistency masks returned by UPFlow_net.forward (model/upflow.py:386-392): replaces
 * tools.occ_check_model('for_back_check', utils/tools.py:501-677) -- ~25 elementwise ATen kernels and two warps per
 * forward -- by one launch.  flow [N,H,W,>=2]: forward flows in images 0..N/2-1, backward flows in N/2..N-1;
 * occ [N,H,W,>=1] receives 1 = visible, 0 = occluded:  |own + torch_warp(other, own)|_1 < alpha_1*(|own|_1+|other|_1) + alpha_2.
 * mode 0 = 'all', 1 = 'obj' (visible OR flow leaving the image), 2 = 'out' (flow stays inside the image). */
int upf_occ_check(const float* flow, int ldf, float* occ, int ldo, int N, int H, int W, float alpha_1, float alpha_2,
                  int mode, int align_corners, void* stream);

/* a5: normalize_features (model/upflow.py:94-137) with
 * moments_across_channels=False, moments_across_images=False (test.py:24-26). */
int upf_featnorm_stats(const float* x, int ldx, int N, int H, int W, int C, double* stats, void* stream);
int upf_featnorm_apply(const float* x, int ldx, const double* stats, float* out, int ldo,
                       int N, int H, int W, int C, void* stream);
/* normalize_features' other moment modes (model/upflow.py:109-124; the reference's DEFAULT configuration, :311-313):
 * moments pooled over the channels of an image (mean / unbiased var over [C,H,W]) and / or over the two tensors of a
 * (feature, warped feature) pair -- the mean of the two means and, as the reference computes it (:121-124), the unbiased
 * variance of the two variances.  stats_a image n is paired with stats_b image (n + b_batch_shift) % N; out_a[n] and
 * out_b[(n + b_batch_shift) % N] receive EQUIVALENT per-channel moments (sum' = mean*npix, sumsq' = var*(npix-1) +
 * sum'*mean) that upf_corr_lrelu_fwd / upf_featnorm_apply turn back into the pooled mean and std.  C <= 256. */
int upf_featnorm_combine(const double* stats_a, const double* stats_b, int b_batch_shift, double* out_a, double* out_b,
                         int N, int C, long long npix, int across_channels, int across_images, void* stream);

/* a6: upsample2d_flow_as / upsample2d_as (model/pwc_modules.py:72-90):
 * bilinear, align_corners=True, channel c multiplied by scale[c] afterwards
 * (scale = {W/w, H/h} for flows, NULL = no scaling).  C <= 4. */
int upf_resize_bilinear(const float* in, int ldi, int h, int w, float* out, int ldo, int H, int W,
                        int N, int C, const float* scale_host, void* stream);

/* a7 blend: sgu_model.forward's last line (model/upflow.py:79-88).
 *   inter  [N,ih,iw,>=3] = (inter_flow u, v, mask LOGIT) as the dense block emits it
 *   flow_init [N,H,W,2]  = the flow to refine at OUTPUT resolution
 * same resolution (ih==H, iw==W):
 *   out = warp(flow_init, inter_flow)*(1-sigmoid(m)) + flow_init*sigmoid(m)
 * output-level variant (ih<H): inter_flow is bilinearly upsampled (align_corners
 * =True) and scaled by (W/iw, H/ih); sigmoid(m) is upsampled AFTER the sigmoid
 * (model/upflow.py:84-86).
 * out_tf32 (nullable) [N,H,W,>=4] pitch ldt: a second copy of the result for tensor-core consumers, channels
 * (rn_tf32(u), rn_tf32(v), 0, 0) when flags has UPF_FLAG_ROUND_TF32, (u, v, 0, 0) otherwise -- the decoder's estimator
 * buffer keeps flow_up in a 4-channel slot whose second half (flow_up + flow_res) is written later in the level; clearing
 * it here means no value of a previous forward is ever read. */
int upf_sgu_blend(const float* flow_init, int ldf, const float* inter, int ldi, int ih, int iw,
                  float* out, int ldo, float* out_tf32, int ldt, int N, int H, int W, int align_corners, int flags,
                  void* stream);

/* a8/a9 (+ the dense block of a7): Conv2d(k in {1,3}, pad=((k-1)*dil)/2) + bias
 * + LeakyReLU(slope) [+ residual], reading an input channel slice and writing
 * an output channel slice.
 *   replaces conv() (model/pwc_modules.py:10-31) as used by
 *   FlowEstimatorDense_v2 (:279-286), ContextNetwork_v2_ (:401-412) and
 *   sgu_model's dense block (model/upflow.py:52-60); torch.cat disappears.
 * weight layout (prepared once by the host): w[tap][cin][cout_pad] fp32 with
 * tap = ky*k+kx, cout_pad = round_up(Cout, 4) for UPF_CONV_FP32;
 *   residual (nullable) [N,Ho,Wo,>=Cout] pitch ldr is added AFTER the activation.
 * precision: UPF_CONV_FP32 = SIMT fp32 FMA (bit-faithful class, any k/stride);
 *            UPF_CONV_TF32 = tcgen05 tensor cores, TF32 operands, fp32
 *            accumulate in TMEM (3x3/1x1 stride 1; needs the packed weights of
 *            upf_conv_tc_pack_weights).
 *            The tensor-core path takes stride 1 or 2; when the grid cannot fill the 148 SMs it splits the K
 *            loop over a thread-block cluster and reduces through distributed shared memory in a fixed order
 *            (bitwise reproducible, no scratch buffer). */
#define UPF_CONV_FP32 0
#define UPF_CONV_TF32 1
/* OR-ed into `precision`: store the output rounded to the nearest TF32 value (see UPF_FLAG_ROUND_TF32) */
#define UPF_CONV_ROUND_OUT 0x100
int upf_conv2d_fwd(const float* x, int ldx, const float* w, const float* bias,
                   float* out, int ldo, const float* residual, int ldr,
                   int N, int H, int W, int Cin, int Cout, int ksize, int stride, int dilation,
                   float slope, int precision, void* stream);

/* A CHAIN of dependent stride-1 convolutions on ONE small feature map [N,H,W] in one persistent launch (conv_chain.cu):
 * the 19 convolutions a coarse pyramid level runs back to back -- FlowEstimatorDense_v2.forward
 * (model/pwc_modules.py:279-286), ContextNetwork_v2_.forward (:401-412), sgu_model's dense block
 * (model/upflow.py:52-60), as called from UPFlow_net.decode_level_res (model/upflow.py:565-572) -- each of which is
 * 0.1-5 GFLOP and costs 7-18 us as a launch of its own.  Layer l is exactly upf_conv2d_fwd(UPF_CONV_TF32) with
 * stride 1 on the same packed weights: out = lrelu(conv(x) + bias) (+ residual), `flags` as there
 * (UPF_FLAG_ROUND_TF32); layer l may read what layers < l wrote (and nothing a later layer writes).  out2 (nullable):
 * a second copy of the result rounded to TF32 (decode_level_res feeds flow_up + flow_res to the context network, :567-569).
 * Results equal upf_conv2d_fwd's up to the order of the fp32 partial sums over K (bitwise reproducible run to run).
 * The launch occupies most SMs with a co-resident grid (layers are separated by a grid-wide barrier): issue chains
 * of one device from ONE stream at a time.  1 <= n_layers <= 16, Cout <= 1024. */
typedef struct {
  const float* x; int ldx;            /* input [N,H,W,>=Cin], pitch ldx (multiple of 4, 16-byte aligned base) */
  const float* w_packed;              /* upf_conv_tc_pack_weights layout */
  const float* bias;
  float* out; int ldo;
  const float* residual; int ldr;     /* nullable, added after the activation */
  float* out2; int ldo2;              /* nullable */
  int Cin, Cout, ksize, dilation;
  float slope;
  int flags;
} upf_chain_layer;
int upf_conv_chain_fwd(const upf_chain_layer* layers, int n_layers, int N, int H, int W, void* stream);
/* test / tuning hook: cluster size (K split, 1/2/4/8) and number of clusters (0 = as many as are co-resident) of conv_chain.cu;
 * bits 8..15 of cluster_size: rows per TMA box (0 = 128, one box per operand tile) */
int upf_debug_conv_chain(int cluster_size, int n_clusters);
/* debug: device buffer of 16 x 16 int64 receiving CTA 0's per-layer clock64 stamps of conv_chain.cu's roles (NULL = off) */
int upf_debug_conv_chain_probe(void* device_buffer_256x_int64);

/* Second half of a 3x3 convolution with very few output channels run as "expand, then combine the taps" (same
 * reference routine as upf_conv2d_fwd: conv(), model/pwc_modules.py:10-31).  The first half is upf_conv2d_fwd with
 * ksize 1 and 9*Cout output channels, Y[p][tap*Cout+co] = sum_ci X[p][ci]*W[co][ci][tap]; this gathers
 *   out[n,y,x,co] = lrelu( bias[co] + sum_tap Y[n, y+(ky-1)*dil, x+(kx-1)*dil, tap*Cout+co] ) (+ residual),
 * taps falling outside the image contributing nothing.  On the tensor cores a 3x3 conv with Cout <= 8 otherwise
 * issues nine N=16 MMAs per K step for one N<=80 MMA's worth of products. */
int upf_conv3x3_tap_combine(const float* y, int ldy, const float* bias, float* out, int ldo, const float* residual,
                            int ldr, int N, int H, int W, int Cout, int dilation, float slope, int flags, void* stream);

/* TF32 tensor-core path: weights packed as [tap][cout_pad16][cin_pad32] fp32
 * (K-major rows of 32 input channels), done on the device from the SIMT layout; values rounded to the nearest TF32
 * (the MMA would otherwise truncate them). */
long long upf_conv_tc_packed_elems(int Cin, int Cout, int ksize);
int upf_conv_tc_pack_weights(const float* w_simt, float* w_packed, int Cin, int Cout, int ksize, void* stream);

/* test / tuning hook, not part of the hot path: enable the shared-halo tensor-core kernel (conv_halo.cu) and pick
 * its shared-memory-descriptor base-offset convention; returns 0. */
int upf_debug_conv_halo(int enabled, int bo_mode);
/* test / tuning hook: enable the linear-window tensor-core kernel (conv_win.cu), its minimum Cin (<0 = keep) and the
 * 4-row units per CTA (0 = automatic) */
int upf_debug_conv_win(int enabled, int min_cin, int force_m);
/* test / tuning hook: small-grid policy of conv_tc.cu: 0 = default (K split over <= 8 CTAs); 8 / 16 = experimental cost-model
 * policy (narrow N tiles, whole-row pixel tiles, clusters up to that size; measured slower, profiles/r2_ab_conv_tc.txt) */
int upf_debug_conv_tc(int max_cluster);
/* test / tuning hook: largest padded Cout whose 3x3 weight gradient takes the taps-along-N kernel (wgrad_taps.cu); 0 routes
 * every shape to conv_tc.cu's per-tap GEMMs, < 0 restores the default */
int upf_debug_wgrad_taps(int max_cout_pad);
/* test / tuning hook: 0 routes large-image correlations to the non-pipelined tiled kernel (corr.cu) */
int upf_debug_corr_pipe(int enabled);
/* debug: device buffer of 64 int64 receiving CTA 0's per-role wait / busy cycle counters of the halo / window /
 * pipelined-correlation kernels (NULL = off) */
int upf_debug_probe(void* device_buffer_8x_int64);

/* ---- a11: backward of the convolutions and of the small decoder ops (training step, BASELINE config 4) ----
 * The INPUT gradient of conv() is upf_conv2d_fwd itself on flipped / transposed weights (stride 2: on the
 * zero-interleaved output gradient); these entry points are the rest of what autograd asks of
 * model/pwc_modules.py:10-31 (cuDNN wgrad + LeakyReLU backward), model/upflow.py:108-135 (torch.mean / torch.var
 * are differentiated through), model/pwc_modules.py:77-90 and model/upflow.py:79-88.  All reductions are
 * deterministic (fixed pixel ranges, fixed summation order, no floating-point atomics). */

/* weight and bias gradient: grad_w [k*k][Cin][Cout] dense, grad_bias [Cout] (nullable), from the input x
 * [N,H,W,>=Cin] and the gradient wrt the PRE-activation output grad_out [N,Ho,Wo,>=Cout].  workspace: at least
 * upf_conv2d_wgrad_workspace_elems(...) floats. */
long long upf_conv2d_wgrad_workspace_elems(int N, int H, int W, int Cin, int Cout, int ksize, int stride, int dilation);
int upf_conv2d_wgrad(const float* x, int ldx, const float* grad_out, int ldg, float* grad_w, float* grad_bias,
                     float* workspace, int N, int H, int W, int Cin, int Cout, int ksize, int stride, int dilation,
                     void* stream);

/* the same on the tensor cores (stride 1; TF32 operands, fp32 accumulation): both operands are transposed into planar,
 * zero-padded form in `workspace` (upf_conv2d_wgrad_tc_workspace_elems floats, 16-byte aligned), where a tap is a
 * constant shift of the K index (the padded pixel index), and the nine tap GEMMs run as one launch of the convolution
 * kernel.  The planar form is BLOCKED and PRE-SWIZZLED, [K / 32][row][32] with the 16-byte chunk index XOR (row & 7): a
 * GEMM operand tile (rows x 32 k) is then one contiguous run that ONE bulk copy (cp.async.bulk) lands in the K-major
 * SWIZZLE_128B form the tensor-core descriptors expect -- as TMA tensor boxes the same tiles cost ~6 ns per 128-byte row
 * (measured: 1.95 ms for the 576->128 GEMM at 8x64x208).  The padded row length is a multiple of 32 so that a tap's
 * vertical shift is a whole number of blocks. */
long long upf_conv2d_wgrad_tc_workspace_elems(int N, int H, int W, int Cin, int Cout, int ksize, int dilation);
int upf_conv2d_wgrad_tc(const float* x, int ldx, const float* grad_out, int ldg, float* grad_w, float* grad_bias,
                        float* workspace, int N, int H, int W, int Cin, int Cout, int ksize, int dilation, void* stream);
/* The same with the input transposed once for several convolutions that read nested channel ranges of one buffer
 * (the dense blocks): upf_wgrad_tc_transpose_input writes xt (upf_wgrad_tc_planar_elems floats, 16-byte aligned:
 * blocked planar [k blocks][C][32], pre-swizzled, zero-padded in k on both sides); a convolution whose input is
 * channels [row0, row0+Cin) of that buffer (row0 a multiple of 8) passes xt, xt_rows = C and row0.  All of them must
 * share ksize and dilation.  workspace as for upf_conv2d_wgrad_tc. */
long long upf_wgrad_tc_planar_pitch(int N, int H, int W, int ksize, int dilation);   /* padded pixel count K (multiple of 32) */
long long upf_wgrad_tc_planar_elems(int N, int H, int W, int C, int ksize, int dilation);
int upf_wgrad_tc_transpose_input(const float* x, int ldx, int C, float* xt, int N, int H, int W, int ksize, int dilation,
                                 void* stream);
int upf_conv2d_wgrad_tc_planar(const float* xt, int xt_rows, int row0, const float* grad_out, int ldg, float* grad_w,
                               float* grad_bias, float* workspace, int N, int H, int W, int Cin, int Cout, int ksize,
                               int dilation, void* stream);

/* pointwise ops on [npix][C] pitched tensors.  op 0: out = b * (a > 0 ? 1 : slope)  (LeakyReLU backward from the
 * saved output a and the incoming gradient b); op 1: out = sigmoid(a); op 2: out = b * a * (1 - a) (sigmoid
 * backward from the saved output a). */
#define UPF_PW_LRELU_BWD   0
#define UPF_PW_SIGMOID     1
#define UPF_PW_SIGMOID_BWD 2
int upf_pointwise(int op, const float* a, int lda, const float* b, int ldb, float* out, int ldo, long long npix,
                  int C, float slope, void* stream);

/* 3xTF32 convolutions (engine precision "tf32x3": fp32-class results on the tensor cores).  x = hi + lo with hi = x
 * truncated to TF32 (what tcgen05 kind::tf32 reads from an fp32 operand) and lo = x - hi; the engine accumulates
 * conv(lo, w) + conv(x, w_lo) + conv(x, w) + bias with upf_conv2d_fwd (slope 1, residual chaining) and finishes with
 *   out = lrelu(t) (+ residual);  out_lo = out - trunc_tf32(out)        (t == NULL: only the split of `out`). */
int upf_act_split(const float* t, int ldt, const float* residual, int ldr, float* out, int ldo, float* out_lo, int ldlo,
                  long long npix, int C, float slope, void* stream);

/* sgu_model.forward's blend (model/upflow.py:88) as a differentiable piece: out_c = w_c*(1-m) + f_c*m, c in {0,1};
 * backward: gw_c = g_c*(1-m), gf_c = g_c*m, gm = sum_c g_c*(f_c - w_c). */
int upf_blend_fwd(const float* w, int ldw, const float* f, int ldf, const float* m, int ldm, float* out, int ldo,
                  long long npix, void* stream);
int upf_blend_bwd(const float* w, int ldw, const float* f, int ldf, const float* m, int ldm, const float* g, int ldg,
                  float* gw, int ldgw, float* gf, int ldgf, float* gm, int ldgm, long long npix, void* stream);

/* normalize_features backward through the moments: dx = (g - mean(g) - y*sum(g*y)/(n-1)) / std, y = (x-mean)/std.
 * stats = the forward's (sum, sum^2) buffer; workspace: upf_featnorm_bwd_workspace_doubles(N, C) doubles. */
long long upf_featnorm_bwd_workspace_doubles(int N, int C);
int upf_featnorm_bwd(const float* x, int ldx, const double* stats, const float* grad_out, int ldg, float* grad_x,
                     int ldgx, double* workspace, int N, int H, int W, int C, void* stream);

/* adjoint of upf_resize_bilinear: grad_in [N,h,w,C] from grad_out [N,H,W,C] (same scale vector); separable, two
 * passes through workspace (upf_resize_bilinear_bwd_workspace_elems(N, H, w, C) floats). */
long long upf_resize_bilinear_bwd_workspace_elems(int N, int H, int w, int C);
int upf_resize_bilinear_bwd(const float* grad_out, int ldgo, int H, int W, float* grad_in, int ldgi, int h, int w,
                            int N, int C, const float* scale_host, float* workspace, void* stream);

/* nn.Conv2d weight [A][B][k][k] (the reference's parameter layout, model/pwc_modules.py:10-31) -> this library's
 * [k*k][Cin][cout_pad] (cout_pad = Cout rounded up to 4, padding zeroed) in one launch.  flip_transpose 0: the forward
 * convolution (Cout = A, Cin = B).  flip_transpose 1: the weights of the input-gradient convolution -- taps flipped,
 * channel roles exchanged (Cin = A, Cout = B). */
int upf_repack_conv_weight(const float* weight, float* out, int A, int B, int ksize, int flip_transpose, void* stream);
/* the same straight into the tensor-core layout of upf_conv_tc_pack_weights (upf_conv_tc_packed_elems floats) */
int upf_repack_conv_weight_tc(const float* weight, float* w_packed, int A, int B, int ksize, int flip_transpose,
                              void* stream);

/* Loss terms of the training step on pixel-major tensors, each ONE reduction pass forward
 * (deterministic: per-CTA partials summed in a fixed order) and one elementwise pass backward.
 * workspace: upf_loss_workspace_elems() floats.  out: 2 floats, out[0] = the term (out[1] is kept for backward).
 *
 * upf_robust_loss_*: photo_loss_multi_type (model/upflow.py:268-290; the photometric term :436-437 and the multi-scale
 * distillation terms :461-487).  d = x - y over [npix][C]; kind 0 abs_robust (|d|+0.01)^q, 1 charbonnier (d^2+1e-6)^q,
 * 2 L1 |d+1e-6|; mask == NULL: mean over npix*C; mask [npix] (photo_loss_use_occ): sum(v*mask)/(sum(mask)+1e-6).
 * Backward: grad_x = grad_out[0] * d(term)/dx, grad_y = -grad_x (either may be NULL); no gradient reaches the mask. */
#define UPF_LOSS_ABS_ROBUST  0
#define UPF_LOSS_CHARBONNIER 1
#define UPF_LOSS_L1          2
long long upf_loss_workspace_elems(void);
int upf_robust_loss_fwd(const float* x, int ldx, const float* y, int ldy, const float* mask, int ldm, float* workspace,
                        float* out, long long npix, int C, int kind, float q, void* stream);
int upf_robust_loss_bwd(const float* x, int ldx, const float* y, int ldy, const float* mask, int ldm, const float* out,
                        const float* grad_out, float* grad_x, int ldgx, float* grad_y, int ldgy, long long npix, int C,
                        int kind, float q, void* stream);
/* upf_edge_smooth1_*: edge_aware_smoothness_order1 (model/upflow.py:198-218): mean(|d_rows pred| * exp(-mean_c|d_rows
 * img|)) + mean(|d_cols pred| * exp(-mean_c|d_cols img|)); img [N,H,W,Ci], pred [N,H,W,Cp], H, W >= 2.  Backward gives
 * the gradient of pred only (the image is data). */
int upf_edge_smooth1_fwd(const float* img, int ldi, int Ci, const float* pred, int ldp, int Cp, float* workspace,
                         float* out, int N, int H, int W, void* stream);
int upf_edge_smooth1_bwd(const float* img, int ldi, int Ci, const float* pred, int ldp, int Cp, const float* grad_out,
                         float* grad_pred, int ldg, int N, int H, int W, void* stream);

/* upf_census_loss_*: census_loss_torch (utils/loss.py:51-91) with the abs_robust penalty: grey = 0.2989 R + 0.5870 G +
 * 0.1140 B of both 3-channel images, soft ternary transform over the (2*max_distance+1)^2 patch (zero padding), soft
 * Hamming distance `dist`, then (|dist|+0.01)^q: mask == NULL mean over the pixels; mask [npix]: masked by mask * (border
 * of max_distance pixels removed), sum / (2*sum(mask') + 1e-6) -- the reference's factor 2 (utils/loss.py:44-46).
 * grey: 2*npix floats, dist: npix floats (both kept for backward).  Backward: gradient of the SECOND image only. */
int upf_census_loss_fwd(const float* img1, int ld1, const float* img2, int ld2, const float* mask, int ldm, float* grey,
                        float* dist, float* workspace, float* out, int N, int H, int W, int max_distance, float q,
                        void* stream);
int upf_census_loss_bwd(const float* grey, const float* dist, const float* mask, int ldm, const float* out,
                        const float* grad_out, float* grad_img2, int ldg, int N, int H, int W, int max_distance, float q,
                        void* stream);

/* upf_boundary_warp_*: tools.boundary_dilated_warp.warp_im (utils/tools.py:350-499): the photometric loss samples the
 * UN-CROPPED frame image [N,Hf,Wf,C] at (x + start[n][0] + u, y + start[n][1] + v) for every pixel of the crop [N,h,w];
 * corner indices are clamped to the frame and the bilinear weights are taken against the clamped corners, like the
 * reference.  start: [N][2] floats on the device.  Backward: gradient of the flow (the frame is data). */
int upf_boundary_warp_fwd(const float* image, int ldi, int C, int Hf, int Wf, const float* flow, int ldf,
                          const float* start, float* out, int ldo, int N, int h, int w, void* stream);
int upf_boundary_warp_bwd(const float* image, int ldi, int C, int Hf, int Wf, const float* flow, int ldf,
                          const float* start, const float* grad_out, int ldg, float* grad_flow, int ldgf, int N, int h,
                          int w, void* stream);

/* layout helpers for callers holding NCHW-contiguous tensors (the reference's
 * layout): strided copy between [N,C,H,W] planes and pixel-major rows. */
int upf_nchw_to_nhwc(const float* in, float* out, int ldo, int N, int C, int H, int W, void* stream);
int upf_nhwc_to_nchw(const float* in, int ldi, float* out, int N, int C, int H, int W, void* stream);
/* out[n,y,x,0:C] = in[n,y,x,0:C] between two pitched buffers (flags: UPF_FLAG_ROUND_TF32); in == NULL writes zeros */
int upf_copy_channels(const float* in, int ldi, float* out, int ldo, long long npix, int C, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UPFLOW_B200_H */
