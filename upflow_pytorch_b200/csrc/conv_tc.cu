// 3x3 / 1x1 convolution (stride 1 or 2, any dilation) on the 5th-generation
// tensor cores (sm_100a): implicit GEMM, TF32 operands read straight from the
// fp32 feature buffers, fp32 accumulation in tensor memory.
//
// Replaces the cuDNN calls behind conv() (model/pwc_modules.py:10-31) for the
// dense flow estimator (:279-286), the dilated context network (:401-412), the
// SGU dense block (model/upflow.py:52-60), the 1x1 adapters (:349-353), the
// feature pyramid (pwc_modules.py:122-142) and sgu output_conv (upflow.py:66-69).
//
//   GEMM view  D[M=128 pixels, N=Cout] += A[M, K] * B[N, K]^T,  K = taps x Cin
//   A  : im2col-free.  The activation tensor is a 4-D TMA tensor (C, W, H, N);
//        for tap (ky,kx) and channel block kb the producer issues ONE box load
//        {32 ch, TW*s, TH*s, 1} with element strides {1,s,s,1} at
//        (kb*32, x0*s+(kx-1)*dil, y0*s+(ky-1)*dil, n): TMA walks the strided
//        window itself (stride-2 convs cost nothing extra) and its out-of-bounds
//        zero fill IS the convolution's zero padding (and the K remainder when
//        Cin % 32 != 0).  The box lands in shared memory as 128 rows (pixels) x
//        128 bytes with SWIZZLE_128B = the canonical K-major UMMA operand.
//   B  : weights pre-packed [tap][cout_pad16][cin_pad32]; box {32, BN, 1}.
//   D  : TMEM, 128 lanes (pixels) x BN columns (output channels), fp32.
//   pipeline: warp 0 = TMA producer, warps 1 and 6 = MMA issuers (one elected thread each,
//        tcgen05.mma.cta_group::1.kind::tf32, 4 x K=8 per 32-channel block).  TWO issuers because one cannot keep
//        the tensor pipe busy: 4 MMAs + tcgen05.commit + mbarrier poll take 634 cycles of the issuing thread against
//        259 cycles of MMA work at N=128 (tools/microbench/tc_probe.cu).  Issuer w takes the K iterations with
//        it % 2 == w into its OWN accumulator (TMEM columns w*BN..), so each accumulator has one issuer and a fixed
//        order (bitwise reproducible); the epilogue adds the two.
//        warps 2-5 = epilogue (tcgen05.ld 32x32b -> bias + LeakyReLU + residual
//        -> 16-byte stores into the output channel slice).  smem full/empty
//        mbarriers ring over NSTAGE stages; tcgen05.commit releases stages and
//        publishes the accumulator.
//   small grids (the 1/64..1/16 pyramid levels have 2..30 pixel tiles for 148
//        SMs): the K loop is split over a THREAD-BLOCK CLUSTER of S<=8 CTAs
//        (grid z).  Each CTA parks its partial accumulator tile in its own
//        shared memory; after a cluster barrier every CTA reduces 128/S rows by
//        reading its peers' tiles through distributed shared memory in rank
//        order (bitwise reproducible, no global scratch, no atomics) and runs
//        the epilogue for those rows.
#include "tc_common.cuh"

#include <mutex>
#include <unordered_map>

namespace upf {

constexpr int TC_THREADS = 224;         // warps: 0 TMA, 1 and 6 MMA issuers, 2-5 epilogue
constexpr int TC_KC = 32;               // channels per K block (128 B of fp32)
constexpr int TC_A_BYTES = 128 * 128;   // 128 pixels x 128 B

struct TcParams {
  float* out; int ldo;
  const float* res; int ldr;
  const float* bias;
  int Ho, Wo, Cout, BN;        // output size; BN = N tile (multiple of 16, <= 128)
  int TH, TW, tiles_x, tiles_y;
  int ks, dil, stride, kblocks; // kblocks = ceil(Cin/32)
  int nstage, tmem_cols;
  int splits, ips;             // K split over a cluster of `splits` CTAs, `ips` iterations each
  // every operand tile is fetched as several small TMA boxes (bw x bh output pixels / b_rows weight rows each):
  // one box is walked row by row (~10 ns per 128-byte row, measured), independent boxes proceed concurrently
  int bw, bh, nbx, nby, b_rows, nbb;
  int a_bytes;                 // bytes one activation tile really brings in: TH * TW pixels x 128 B (<= TC_A_BYTES)
  float slope;
  int flags;                   // UPF_FLAG_ROUND_TF32: store the output rounded to the nearest TF32 value
  // weight-gradient mode (backward.cu): the "images" are the taps of ONE planar operand -- image n reads the same
  // tensor with its K (channel) coordinate shifted by koffs[n]
  int wgrad;
  int koffs[9];               // in k BLOCKS of 32 (the blocked planar layout of backward.cu)
  int wsel[9];                // which of the (pre-shifted) copies of the other operand tap n multiplies
  // the blocked, PRE-SWIZZLED planar operands of the weight gradient: [k block][row][32 k], `kpad` zero blocks before
  // block 0 of xg (a tap's vertical shift may reach that far); an operand tile is one contiguous run -> one bulk copy
  const float* xg; const float* wg;
  int xrows, wrows, kpad;
  int stage_bytes;            // bytes of one ring slot: [A blocks: 16 KB][B blocks]
  int kbs;                    // k blocks per ring slot in weight-gradient mode (1 otherwise): small operand tiles travel 2..8 at a time
  long long wcopy;            // elements between the pre-shifted copies of wg
};

// contiguous global -> shared bulk copy completing on an mbarrier (no tensor map: no per-row cost)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// bias + LeakyReLU (+ residual) and the store of 4 consecutive output channels of one pixel
__device__ __forceinline__ void store4(const TcParams& p, const float* s_bias, int co0, float* o, const float* r, int co, float4 f, bool vec_out) {
  float v[4] = {f.x, f.y, f.z, f.w};
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (co + j < p.Cout) {
      float a = lrelu(v[j] + s_bias[co + j - co0], p.slope);
      if (r) a += __ldg(r + co + j);
      v[j] = maybe_round(a, p.flags);
    }
  if (vec_out && co + 4 <= p.Cout) {
    *reinterpret_cast<float4*>(o + co) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (co + j < p.Cout) o[co + j] = v[j];
  }
}

__global__ void __launch_bounds__(TC_THREADS)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages: A | B] ... then barriers
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (a uintptr_t round trip would turn every later
  // access into a generic-address load: measured 4x slower epilogue)
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t b_bytes = (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = (uint32_t)p.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)p.nstage * stage_bytes);
  uint64_t* full = bars;                       // [nstage]
  uint64_t* empty = bars + p.nstage;           // [nstage]
  uint64_t* accum_full = bars + 2 * p.nstage;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.nstage + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);      // [BN]: the L1 is tiny next to ~100-200 KB of shared memory

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next kernel's prologue start early (PDL)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int tile = blockIdx.x;
  const int tx = tile % p.tiles_x; tile /= p.tiles_x;
  const int ty = tile % p.tiles_y;
  const int n = tile / p.tiles_y;
  const int x0 = tx * p.TW, y0 = ty * p.TH;      // output-pixel origin of this tile
  const int co0 = blockIdx.y * p.BN;             // first output channel of this CTA's N tile
  const int taps = p.ks * p.ks;
  const int iters_all = taps * p.kblocks;
  const int it_begin = blockIdx.z * p.ips;
  const int iters = (iters_all - it_begin < p.ips ? iters_all - it_begin : p.ips);   // >= 1 by construction

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < p.nstage; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), iters >= 2 ? 2 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 64 && (int)threadIdx.x - 64 < p.BN) {
    const int co = blockIdx.y * p.BN + (int)threadIdx.x - 64;
    s_bias[threadIdx.x - 64] = co < p.Cout ? __ldg(p.bias + co) : 0.f;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch, bias) touched
  // only this CTA's own state and constant weights, and may overlap the tail of the previous kernel in the stream;
  // from here on we read activations / write outputs, so wait for the upstream grid to complete and flush.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const bool vec_out = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  const int pitch = p.BN + 4;                    // floats per row of the parked partial tile (bank-conflict free)
  float* part = reinterpret_cast<float*>(base);  // overlays the (drained) pipeline stages

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      const int half = (p.ks - 1) / 2;
      for (int it = 0; it < iters; ++it) {
        const int g = it_begin + it;
        const int tap = g / p.kblocks, kb = g - tap * p.kblocks;
        const int ky = tap / p.ks, kx = tap - ky * p.ks;
        const int cy = y0 * p.stride + (ky - half) * p.dil, cx = x0 * p.stride + (kx - half) * p.dil;
        const int s = it % p.nstage;
        const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
        mbar_wait(smem_u32(&empty[s]), ph ^ 1u);            // slot free (first pass returns at once)
        const uint32_t a_dst = smem_u32(base + (size_t)s * stage_bytes);
        const uint32_t b_dst = a_dst + TC_A_BYTES;
        const uint32_t fb = smem_u32(&full[s]);
        mbar_expect_tx(fb, (uint32_t)p.kbs * ((uint32_t)p.a_bytes + b_bytes));
        if (p.wgrad) {
          // weight gradient: both operands are BLOCKED planar tensors [k block][row][32 k] (backward.cu), so a box of
          // `rows` x 32 k is one contiguous run of rows x 128 bytes; the tap's vertical shift is a whole number of k blocks
          // (as TMA tensor boxes these tiles cost ~6 ns per 128-byte row whatever the layout: 1.95 ms for the 576->128
          // gradient at 8x64x208; the data is written pre-swizzled by nhwc_to_planar_padded_kernel)
          // a ring slot holds p.kbs consecutive k blocks of both operands (a K loop of small tiles is bound by its ~0.35 us
          // hand-off chain per slot, not by bytes: measured 1.4 ms for the 3->16 full-resolution gradient, one block per slot)
          for (int j = 0; j < p.kbs; ++j) {
            const int kq = kb * p.kbs + j;
            bulk_g2s(a_dst + (uint32_t)(j * p.a_bytes), p.xg + ((size_t)(kq + p.koffs[n] + p.kpad) * p.xrows + cx) * 32, (uint32_t)p.a_bytes, fb);
            bulk_g2s(b_dst + (uint32_t)j * b_bytes, p.wg + (size_t)p.wsel[n] * p.wcopy + ((size_t)kq * p.wrows + co0) * 32, b_bytes, fb);
          }
        } else {
          const int kc = kb * TC_KC;
          for (int jy = 0; jy < p.nby; ++jy)
            for (int jx = 0; jx < p.nbx; ++jx)
              tma_load_4d(a_dst + (uint32_t)((jy * p.bh * p.TW + jx * p.bw) * 128), &map_x, fb, kc,
                          cx + jx * p.bw * p.stride, cy + jy * p.bh * p.stride, n);
          for (int jb = 0; jb < p.nbb; ++jb)
            tma_load_3d(b_dst + (uint32_t)(jb * p.b_rows * 128), &map_w, fb, kb * TC_KC, co0 + jb * p.b_rows, tap);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 6) {
    // ===================== MMA issuers =====================
    // instruction descriptor: D=f32 (bit4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 @17, M>>4 @24
    const int wi = warp == 1 ? 0 : 1;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t tacc = tmem_base + (uint32_t)(wi * p.BN);
    for (int it = wi; it < iters; it += 2) {
      const int s = it % p.nstage;
      const uint32_t ph = (uint32_t)(it / p.nstage) & 1u;
      mbar_wait(smem_u32(&full[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a_addr = smem_u32(base + (size_t)s * stage_bytes);
        for (int j = 0; j < p.kbs; ++j) {
          const uint64_t da = umma_desc_sw128(a_addr + (uint32_t)(j * p.a_bytes));
          const uint64_t db = umma_desc_sw128(a_addr + TC_A_BYTES + (uint32_t)j * b_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 4 x (K = 8 tf32 = 32 bytes): advance the start address inside the swizzle atom
            umma_tf32(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (it > wi || j > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty[s]));                      // frees the stage when these MMAs retire
        if (it + 2 >= iters) umma_commit(smem_u32(accum_full));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                                    // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                             // pixel index inside the tile
    const bool two = iters >= 2;
    mbar_wait(smem_u32(accum_full), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (p.splits == 1) {
      const int py = y0 + row / p.TW, px = x0 + row % p.TW;
      const bool valid = (row < p.TH * p.TW) && (py < p.Ho) && (px < p.Wo);   // a tile may hold fewer than 128 pixels
      const size_t pix = ((size_t)n * p.Ho + py) * p.Wo + px;
      float* o = p.out + pix * p.ldo;
      const float* r = p.res ? p.res + pix * p.ldr : nullptr;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (two) {                                             // second issuer's accumulator (odd K iterations)
          uint32_t v2[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            store4(p, s_bias, co0, o, r, co0 + c0 + j, make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), vec_out);
        }
      }
    } else {
      // park the partial accumulator tile in this CTA's shared memory (the stages are drained: every TMA load
      // was consumed and the last MMA has retired)
      float* mine = part + row * pitch;
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        if (two) {
          uint32_t v2[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), v2);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(mine + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                  __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  if (p.splits > 1) {
    // ---- cluster reduction through distributed shared memory, rank order (deterministic)
    cluster_sync_all();
    const int rank = (int)cluster_ctarank();
    const int rows_per = 128 / p.splits;
    const int c4n = p.BN >> 2;
    const uint32_t part_addr = smem_u32(part);
    for (int u = threadIdx.x; u < rows_per * c4n; u += TC_THREADS) {
      const int rl = u / c4n, c4 = u - rl * c4n;
      const int row = rank * rows_per + rl;
      const uint32_t off = part_addr + (uint32_t)(row * pitch + c4 * 4) * 4u;
      // all peers' loads in flight at once (a dependent chain of distributed-shared-memory round trips costs ~0.3 us each)
      float4 t[8];
#pragma unroll
      for (int sp = 0; sp < 8; ++sp)
        if (sp < p.splits) t[sp] = ld_dsmem_f4(off, (uint32_t)sp);
      float4 acc = t[0];
#pragma unroll
      for (int sp = 1; sp < 8; ++sp)
        if (sp < p.splits) { acc.x += t[sp].x; acc.y += t[sp].y; acc.z += t[sp].z; acc.w += t[sp].w; }
      for (int sp = 8; sp < p.splits; ++sp) {
        const float4 u8 = ld_dsmem_f4(off, (uint32_t)sp);
        acc.x += u8.x; acc.y += u8.y; acc.z += u8.z; acc.w += u8.w;
      }
      const int py = y0 + row / p.TW, px = x0 + row % p.TW;
      if (row < p.TH * p.TW && py < p.Ho && px < p.Wo) {
        const size_t pix = ((size_t)n * p.Ho + py) * p.Wo + px;
        store4(p, s_bias, co0, p.out + pix * p.ldo, p.res ? p.res + pix * p.ldr : nullptr, co0 + c4 * 4, acc, vec_out);
      }
    }
    cluster_sync_all();          // nobody leaves while a peer may still read its tile
  } else {
    __syncthreads();
  }
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// weights: SIMT layout [tap][cin][cout_pad4] -> packed [tap][cout_pad16][cin_pad32]
__global__ void pack_weights_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int taps,
                                    int cout_pad4, int cout_pad16, int cin_pad32) {
  const long long total = (long long)taps * cout_pad16 * cin_pad32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad32);
    const int co = (int)((i / cin_pad32) % cout_pad16);
    const int tap = (int)(i / ((long long)cin_pad32 * cout_pad16));
    float v = 0.f;
    if (ci < Cin && co < Cout) v = w[((size_t)tap * Cin + ci) * cout_pad4 + co];
    wp[i] = round_tf32(v);      // kind::tf32 would truncate: round to nearest once, here
  }
}

// nn.Conv2d weight [A][B][k][k] -> packed [tap][cout_pad16][cin_pad32] in one launch (upf_repack_conv_weight followed
// by pack_weights_kernel, fused: the training path packs every weight twice per step).  flip_transpose as there.
__global__ void repack_weights_tc_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cin, int Cout, int taps,
                                         int cout_pad16, int cin_pad32, int flip_transpose) {
  const long long total = (long long)taps * cout_pad16 * cin_pad32;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin_pad32);
    const int co = (int)((i / cin_pad32) % cout_pad16);
    const int tap = (int)(i / ((long long)cin_pad32 * cout_pad16));
    float v = 0.f;
    if (ci < Cin && co < Cout)
      v = flip_transpose ? __ldg(w + ((size_t)ci * Cout + co) * taps + (taps - 1 - tap)) : __ldg(w + ((size_t)co * Cin + ci) * taps + tap);
    wp[i] = round_tf32(v);
  }
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    for (long long v : {k.a, k.b, k.c, k.d, k.e}) h = h * 1000003u ^ std::hash<long long>()(v);
    return h;
  }
};
static std::mutex g_map_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

int encode_cached(const MapKey& key, CUtensorMap* out, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, int swizzle_bytes) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) { *out = it->second; return 0; }
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("conv_tc: cuTensorMapEncodeTiled not available"); return UPF_EDRIVER; }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("conv_tc: cuTensorMapEncodeTiled failed (%d)", (int)r); return UPF_EDRIVER; }
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps.emplace(key, *out);
  return 0;
}

// Small-grid policy.  0 (default) = K split over a cluster of <= 8 only.  8 / 16 = additionally narrow N tiles and whole-row
// pixel tiles chosen by an operand-traffic cost model, clusters up to that size -- measured SLOWER on the KITTI forward
// (2.855 vs 2.77 ms, tools/ab_conv_tc.py, profiles/r2_ab_conv_tc.txt): the coarse levels are bound by the serial
// launch -> prologue -> first TMA -> commit -> epilogue chain, not by operand bandwidth.  Kept as an A/B switch.
static int g_tc_cluster_cap = 0;
static int g_tc_wgrad_kbs1 = 0;        // 1: one k block per ring slot in the weight-gradient GEMM (A/B switch, upf_debug_conv_tc bit 9)
static int g_tc_wgrad_one_cta = 1;     // weight-gradient GEMM: one resident CTA with a deep ring (0: two CTAs, two slots each)
static void pick_tile(int H, int W, int max_tw, int* TH, int* TW) {
  // <= 128 pixels per tile; pick the shape wasting the fewest pixels
  long long best = -1;
  for (int tw = 8; tw <= max_tw; tw <<= 1) {
    const int th = 128 / tw;
    const long long cover = (long long)((H + th - 1) / th) * ((W + tw - 1) / tw);
    if (best < 0 || cover < best || (cover == best && tw == 16)) { best = cover; *TH = th; *TW = tw; }
  }
  // whole image rows (TW = W, any width): the box is dense, so its pixels are consecutive 128-byte rows of the stage like
  // any other tile's; the MMA reads 128 rows, the rows past TH*TW are never stored.  6x20 (the 1/64 level of a KITTI
  // frame) is ONE tile per image instead of two half-empty ones.
  if (W <= max_tw && W >= 4 && g_tc_cluster_cap != 0) {
    int th = 128 / W;
    if (th > H) th = H;
    const long long cover = (H + th - 1) / th;
    if (cover < best) { *TH = th; *TW = W; }
  }
}

int g_tc_pdl = 1;         // programmatic dependent launch between consecutive conv kernels (upf_debug_conv_halo bit 3 = off)
int g_tc_box_rows = 128;  // pixels (128-byte rows) per TMA box (128 = one box per tile).  Measured: SMALLER boxes are
                          // slower (tools/bench_halo.py: 16-row boxes cost 1.5-2x), so the tile is fetched as one box

int conv2d_fwd_halo(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                    const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                    float slope, int flags, cudaStream_t st, int* taken);

int conv2d_fwd_win(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                   const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                   float slope, int flags, cudaStream_t st, int* taken);

// Largest cluster the K split may use: 16 (non-portable size, opt-in per kernel) when the device schedules it, else 8.
static int tc_max_cluster() {
  static PerDeviceOnce once;
  static int allowed[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) return 8;
  if (once.need()) {
    allowed[dev] = 8;
    if (cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1, 1, 16);
      cfg.blockDim = dim3(TC_THREADS);
      cfg.dynamicSmemBytes = 210 * 1024;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 16;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int nclusters = 0;
      if (cudaOccupancyMaxActiveClusters(&nclusters, conv_tc_kernel, &cfg) == cudaSuccess && nclusters >= 8) allowed[dev] = 16;
    }
    (void)cudaGetLastError();
    once.mark();
  }
  return allowed[dev] < g_tc_cluster_cap ? allowed[dev] : g_tc_cluster_cap;
}

static thread_local const int* g_tc_koffs = nullptr;      // set by conv_tc_wgrad_gemm around its call (per calling thread)
static thread_local const int* g_tc_wsel = nullptr;
static thread_local int g_tc_xrows = 0;                   // rows per k block of the blocked planar XT buffer
static thread_local int g_tc_kpad = 0;                    // zero k blocks in front of XT's block 0
static thread_local long long g_tc_gcopy = 0;             // elements between the pre-shifted copies of GT

int conv2d_fwd_tc(const float* x, int ldx, const float* w_packed, const float* bias, float* out, int ldo,
                  const float* res, int ldr, int N, int H, int W, int Cin, int Cout, int ks, int stride, int dil,
                  float slope, int flags, cudaStream_t st) {
  const int* koffs = g_tc_koffs;
  UPF_REQUIRE((ldx % 4) == 0 && aligned16(x) && aligned16(w_packed), "conv_tc: input pitch/pointer must be 16-byte aligned");
  UPF_REQUIRE(stride == 1 || stride == 2, "conv_tc: stride %d not in {1,2}", stride);
  if (!koffs) {
    // fine pyramid levels, 3x3 / dilation <= 4: the halo kernel loads the activation tile once for all nine taps
    int taken = 0;
    const int e0 = conv2d_fwd_win(x, ldx, w_packed, bias, out, ldo, res, ldr, N, H, W, Cin, Cout, ks, stride, dil, slope, flags, st, &taken);
    if (e0 != 0 || taken) return e0;
    const int e = conv2d_fwd_halo(x, ldx, w_packed, bias, out, ldo, res, ldr, N, H, W, Cin, Cout, ks, stride, dil, slope, flags, st, &taken);
    if (e != 0 || taken) return e;
  }
  const int pad = ((ks - 1) * dil) / 2;
  const int Ho = (H + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  const int Wo = (W + 2 * pad - dil * (ks - 1) - 1) / stride + 1;
  UPF_REQUIRE(Ho > 0 && Wo > 0, "conv_tc: empty output");
  const int cout_pad = (Cout + 15) & ~15;
  const int kblocks = (Cin + TC_KC - 1) / TC_KC;
  const int cin_pad = kblocks * TC_KC;
  const int taps = ks * ks;
  int TH = 8, TW = 16;
  pick_tile(Ho, Wo, 256 / stride > 128 ? 128 : 256 / stride, &TH, &TW);   // TMA box extents are <= 256 elements
  if (koffs) {                 // weight gradient: the "image" is one row of Cin operand rows; one box of TW rows per k block
    TH = 1;
    TW = 8;
    while (TW < 128 && TW < Wo) TW <<= 1;
  }
  const int tiles_x = (Wo + TW - 1) / TW, tiles_y = (Ho + TH - 1) / TH;
  const long long tiles = (long long)tiles_x * tiles_y * N;
  const int a_bytes = TH * TW * 128;
  // weight gradient: k blocks per ring slot (the blocked operands carry 7 zero blocks of slack behind the last one)
  int kbs = 1;
  if (koffs) {
    const int cpad = (Cout + 15) & ~15;
    const int bn0 = ((cpad + (cpad + 127) / 128 - 1) / ((cpad + 127) / 128) + 15) & ~15;
    while (kbs < 8 && 2 * kbs * TW <= 128 && 2 * kbs * (a_bytes + bn0 * 128) <= 32 * 1024) kbs <<= 1;
    if (g_tc_wgrad_kbs1) kbs = 1;
  }
  const int iters_all = taps * ((kblocks + kbs - 1) / kbs);

  // N tiling and K split.  Large grids: equal N tiles <= 128 wide, no split.  Small grids (the coarse pyramid levels: 2..30
  // pixel tiles for 148 SMs) are bound by how fast ONE SM can pull its operands out of L2 (~58 B/clk measured,
  // tools/microbench/tc_probe.cu): a CTA that walks `ips` K iterations receives ips * (activation tile + BN weight rows)
  // bytes.  So the work is cut along N (narrow weight tiles: the activation tile is re-fetched per N tile, but by another
  // SM) and along K (a thread-block cluster of up to 16 CTAs reduces its partial tiles through distributed shared memory
  // in rank order), choosing the (N tiles, splits) pair with the smallest per-CTA cost that still fits one wave.
  int ntiles_n = (cout_pad + 127) / 128;
  int BN = ((cout_pad + ntiles_n - 1) / ntiles_n + 15) & ~15;   // rows past cout_pad: TMA zero fill
  int splits = 1;
  if (koffs) {
    splits = 8;                                                   // weight gradient: K is the pixel index, 10^4..10^5 long
    while (splits > 1 && iters_all / splits < 4) splits >>= 1;
  } else if (g_tc_cluster_cap == 0) {
    while (splits < 8 && tiles * ntiles_n * splits * 2 <= UPF_NUM_SMS && iters_all / (splits * 2) >= 3) splits *= 2;
  } else if (tiles * ntiles_n * 2 <= UPF_NUM_SMS) {
    const int max_cluster = tc_max_cluster();
    double best = 1e30;
    const int base_nt = ntiles_n;
    const int cand_nt[4] = {base_nt, 2, 4, 8};
    for (int ci = 0; ci < 4; ++ci) {
      const int nt = cand_nt[ci];
      if (ci > 0 && nt <= base_nt) continue;
      const int bn = ((cout_pad + nt - 1) / nt + 15) & ~15;
      if (nt > base_nt && (nt - 1) * bn >= cout_pad) break;     // an N tile would be empty
      for (int sp = 1; sp <= max_cluster; sp <<= 1) {
        if (tiles * nt * sp > UPF_NUM_SMS && !(nt == base_nt && sp == 1)) break;
        const int ips_c = (iters_all + sp - 1) / sp;
        if (sp > 1 && (ips_c < 2 || (sp - 1) * ips_c >= iters_all)) break;
        const double load = (double)ips_c * (a_bytes + bn * 128) / 58.0;                 // cycles to receive the operands
        const double reduce = sp > 1 ? 1500.0 + 40.0 * sp * (((128 / sp) * (bn / 4) + TC_THREADS - 1) / TC_THREADS) : 0.0;
        const double cost = load + reduce;
        if (cost < best * 0.97) { best = cost; ntiles_n = nt; BN = bn; splits = sp; }
      }
    }
  }

  // sub-boxes: `box_rows` pixels each, either whole tile rows (TW <= box_rows) or a segment of one row
  // (also for the latency-bound split-K launches: four 32-row boxes per tile measured 2.97 vs 2.80 ms per KITTI forward)
  const int box_rows = (g_tc_box_rows >= 8 && g_tc_box_rows <= 128) ? g_tc_box_rows : 128;
  const int bw = TW < box_rows ? TW : box_rows;
  int bh = box_rows / bw;
  if (bh > TH) bh = TH;
  int b_rows = BN < box_rows ? BN : box_rows;
  while (BN % b_rows) b_rows -= 8;           // BN is a multiple of 16
  CUtensorMap mx, mw;
  if (koffs) {
    // weight gradient: x = XT blocked [Cin / 32 k blocks][g_tc_xrows rows][32], this call's rows start at x; w = GT blocked
    // [3 copies][k blocks][cout_pad rows][32]
    UPF_REQUIRE(H == 1 && stride == 1 && ks == 1 && Cin % 32 == 0 && g_tc_xrows >= W, "conv_tc: bad weight-gradient call");
    const cuuint64_t kblk = (cuuint64_t)(Cin / 32);
    const cuuint64_t dims[4] = {32, (cuuint64_t)W, kblk, 1};
    const cuuint64_t strides[3] = {128, (cuuint64_t)g_tc_xrows * 128, (cuuint64_t)g_tc_xrows * 128 * kblk};
    const cuuint32_t box[4] = {TC_KC, (cuuint32_t)bw, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    MapKey key{x, g_tc_xrows, (long long)W, (long long)kblk, bw, 41};
    int e = encode_cached(key, &mx, 4, const_cast<float*>(x), dims, strides, box, estr);
    if (e) return e;
    const cuuint64_t wdims[4] = {32, (cuuint64_t)cout_pad, kblk, 3};
    const cuuint64_t wstrides[3] = {128, (cuuint64_t)cout_pad * 128, (cuuint64_t)cout_pad * 128 * kblk};
    const cuuint32_t wbox[4] = {TC_KC, (cuuint32_t)b_rows, 1, 1};
    MapKey wkey{w_packed, cout_pad, (long long)kblk, b_rows, 3, 42};
    e = encode_cached(wkey, &mw, 4, const_cast<float*>(w_packed), wdims, wstrides, wbox, estr);
    if (e) return e;
  } else {
  {
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)(koffs ? 1 : N)};
    const cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    const cuuint32_t box[4] = {TC_KC, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), 1};
    const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    MapKey key{x, ldx, ((long long)H << 32) | (unsigned)W, ((long long)(koffs ? 1 : N) << 32) | (unsigned)Cin, (bw * 1000 + bh) * 4 + stride, 4};
    int e = encode_cached(key, &mx, 4, const_cast<float*>(x), dims, strides, box, estr);
    if (e) return e;
  }
  {
    const int wtaps = koffs ? 3 : taps;      // weight-gradient mode: three pre-shifted copies of the second operand
    const cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)wtaps};
    const cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 4, (cuuint64_t)cin_pad * cout_pad * 4};
    const cuuint32_t box[3] = {TC_KC, (cuuint32_t)b_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    MapKey key{w_packed, cin_pad, b_rows, wtaps, cout_pad, 3};
    int e = encode_cached(key, &mw, 3, const_cast<float*>(w_packed), dims, strides, box, estr);
    if (e) return e;
  }
  }

  TcParams p;
  p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr; p.bias = bias;
  p.Ho = Ho; p.Wo = Wo; p.Cout = Cout; p.BN = BN;
  p.TH = TH; p.TW = TW; p.tiles_x = tiles_x; p.tiles_y = tiles_y;
  p.ks = ks; p.dil = dil; p.stride = stride; p.kblocks = (kblocks + kbs - 1) / kbs;   // ring slots per tap
  p.kbs = kbs;
  p.bw = bw; p.bh = bh; p.nbx = TW / bw; p.nby = TH / bh; p.b_rows = b_rows; p.nbb = BN / b_rows;
  p.a_bytes = a_bytes;
  p.slope = slope;
  p.flags = flags;
  p.wgrad = koffs ? 1 : 0;
  p.xg = x; p.wg = w_packed; p.xrows = g_tc_xrows; p.wrows = cout_pad; p.kpad = g_tc_kpad;
  p.wcopy = g_tc_gcopy;
  for (int i = 0; i < 9; ++i) { p.koffs[i] = (koffs && i < N) ? koffs[i] : 0; p.wsel[i] = (koffs && g_tc_wsel && i < N) ? g_tc_wsel[i] : 0; }
  p.tmem_cols = BN <= 16 ? 32 : (BN <= 32 ? 64 : (BN <= 64 ? 128 : 256));   // two accumulators (one per MMA issuer)
  // an MMA reads 128 rows of A from its block's start whatever TW is: the LAST block of a slot must still end inside the slot
  int stage_bytes = TC_A_BYTES + ((kbs * BN * 128 + 1023) & ~1023);
  if (stage_bytes < (kbs - 1) * a_bytes + TC_A_BYTES) stage_bytes = (kbs - 1) * a_bytes + TC_A_BYTES;
  const long long ctas = tiles * ntiles_n;
  int ips = (iters_all + splits - 1) / splits;
  while (splits > 1 && (splits - 1) * ips >= iters_all) { splits >>= 1; ips = (iters_all + splits - 1) / splits; }   // no empty CTA
  // stages: two CTAs per SM when the grid is large (epilogue/main-loop overlap across CTAs); a single
  // resident CTA gets the whole shared memory so that more TMA loads are in flight (latency-bound regime)
  // (weight gradient with 128-wide N tiles: its operands stream from HBM and two resident CTAs get only two 32 KB slots
  // each -- measured 1913 us for 576->128 at 8x64x208 against 505 us with ONE resident CTA and a six-slot ring; for
  // N <= 96 two CTAs stay ahead (257 vs 327 us at 384->96, 333 vs 482 at 480->64).  upf_debug_conv_tc bit 8: A/B switch)
  const bool one_cta = ctas * splits <= UPF_NUM_SMS || (koffs && BN > 96 && g_tc_wgrad_one_cta);
  int nstage = ((one_cta ? 200 : 108) * 1024) / stage_bytes;
  if (nstage > (one_cta ? 12 : 6)) nstage = one_cta ? 12 : 6;
  nstage &= ~1;        // EVEN: issuer w then owns the stages of parity w and sees every phase of their barriers -- with an odd
                       // ring an issuer would revisit a stage two phases later, which a parity wait cannot tell from the stale one
  if (nstage < 2) nstage = 2;
  if (splits > 1 && (long long)nstage * stage_bytes < 128ll * (BN + 4) * 4) { splits = 1; ips = iters_all; }
  p.nstage = nstage;
  p.stage_bytes = stage_bytes;
  p.splits = splits; p.ips = ips;
  const size_t smem = (size_t)nstage * stage_bytes + (2 * nstage + 2) * 8 + 16 + BN * 4 + 1024;
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("conv_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_set.mark();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles, (unsigned)ntiles_n, (unsigned)splits);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)splits;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_tc_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_tc_kernel, mx, mw, p);
  if (e != cudaSuccess) { set_error("conv_tc launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  return check_launch("conv_tc");
}

// dW[tap][ci][co] = sum_k XT[ci][k + koffs[tap]] * GT[wsel[tap]][co][k]  (backward.cu): the tensor-core GEMM of this file with the
// taps as "images" (out is [taps][Cin][Cout]), M = Cin, K = the padded pixel index, cluster split-K
int conv_tc_wgrad_gemm(const float* xt, int ldk, const float* gt_packed, const float* zero_bias, float* gw, int taps,
                       int Cin, int Cout, int K, const int* koffs, const int* wsel, int xt_rows, int kpad, long long gcopy, cudaStream_t st) {
  g_tc_koffs = koffs; g_tc_wsel = wsel; g_tc_xrows = xt_rows; g_tc_kpad = kpad; g_tc_gcopy = gcopy;
  const int e = conv2d_fwd_tc(xt, ldk, gt_packed, zero_bias, gw, Cout, nullptr, 0, taps, 1, Cin, K, Cout, 1, 1, 1, 1.0f, 0, st);
  g_tc_koffs = nullptr; g_tc_wsel = nullptr;
  return e;
}

}  // namespace upf

// test / tuning hook: small-grid policy of conv_tc (0 = default, 8 / 16 = cost-model policy with that cluster cap)
extern "C" int upf_debug_conv_tc(int max_cluster) {
  upf::g_tc_wgrad_one_cta = (max_cluster >= 0 && (max_cluster & 0x100)) ? 0 : 1;
  upf::g_tc_wgrad_kbs1 = (max_cluster >= 0 && (max_cluster & 0x200)) ? 1 : 0;
  upf::g_tc_cluster_cap = (max_cluster >= 0 && (max_cluster & 0xff) <= 16) ? (max_cluster & 0xff) : 0;
  return 0;
}

extern "C" long long upf_conv_tc_packed_elems(int Cin, int Cout, int ksize) {
  const long long cin_pad = (Cin + 31) / 32 * 32, cout_pad = (Cout + 15) / 16 * 16;
  return (long long)ksize * ksize * cout_pad * cin_pad;
}

extern "C" int upf_conv_tc_pack_weights(const float* w_simt, float* w_packed, int Cin, int Cout, int ksize, void* stream) {
  using namespace upf;
  UPF_REQUIRE(w_simt && w_packed && Cin > 0 && Cout > 0 && (ksize == 1 || ksize == 3), "pack_weights: bad argument");
  const int taps = ksize * ksize;
  const int cout_pad4 = (Cout + 3) & ~3, cout_pad16 = (Cout + 15) & ~15, cin_pad32 = (Cin + 31) / 32 * 32;
  const long long total = (long long)taps * cout_pad16 * cin_pad32;
  long long blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(w_simt, w_packed, Cin, Cout, taps, cout_pad4,
                                                                         cout_pad16, cin_pad32);
  return check_launch("pack_weights");
}

extern "C" int upf_repack_conv_weight_tc(const float* weight, float* w_packed, int A, int B, int ksize, int flip_transpose,
                                         void* stream) {
  using namespace upf;
  UPF_REQUIRE(weight && w_packed && A > 0 && B > 0 && (ksize == 1 || ksize == 3), "repack_conv_weight_tc: bad argument");
  const int Cin = flip_transpose ? A : B, Cout = flip_transpose ? B : A, taps = ksize * ksize;
  const int cout_pad16 = (Cout + 15) & ~15, cin_pad32 = (Cin + 31) / 32 * 32;
  const long long total = (long long)taps * cout_pad16 * cin_pad32;
  long long blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  repack_weights_tc_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(weight, w_packed, Cin, Cout, taps, cout_pad16,
                                                                              cin_pad32, flip_transpose ? 1 : 0);
  return check_launch("repack_weights_tc");
}
