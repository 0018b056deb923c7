// Correlation + LeakyReLU for PLANAR ("NCHW") feature maps -- the layout the reference's own operator takes and
// returns (correlation_cuda.forward: input1/input2 [B,C,H,W] -> output [B,(2d+1)^2,H,W], correlation_cuda.cc:10-87;
// Corr_pyTorch.forward, utils/pytorch_correlation.py:27-50).
//
// Why another formulation (profiles/r2_corr_notes.md, r2_corr_planar_notes.md): the pixel-major kernel (corr_pipe.cu) loads
// CHANNEL quads, so the two factors of every product sit in registers of the same bank parity (half of its FFMAs pay a
// register-bank conflict: 75 of 128 FMA/clk/SM with no memory traffic at all) and its 8 rows x 9 vertical displacements
// blocking reads one shared word per 3 FMAs, which is more than the 128 B/clk of shared memory can feed.  Here the staged
// operands are PLANAR, [channel][row][column], so a 16-byte shared load delivers four neighbouring COLUMNS of one channel,
// and the register block is an OUTER PRODUCT centred on the second image:
//
//   thread (lane, row y, group g):  b[r][j] = f2[c][y + g*R + r - d][q + j]   j = 0..3  (R rows of one aligned column quad)
//                                   a[k]    = f1[c][y][q - 4 + k]             k = 0..11 (the row of f1 around it)
//                                   acc[r][dx][j] += a[j - dx + 4] * b[r][j]  for all 2d+1 horizontal displacements dx
//
// i.e. the thread owns the f2 pixels (q+j, y+dy) and produces, for each of them, the outputs of the 2d+1 pixels
// x = q+j-dx that see it at displacement dx: out[(x, y), (dy, dx)] is computed by exactly one thread, completely (no
// partial sums).  d = 4: 108 FMAs per 24 loaded words (4.5 per word), 108 accumulators, and the factors of a product have
// the parity of j - dx and j: the accumulator can always be given a register of the other bank (SASS: 8 % of the FFMAs
// keep three same-bank sources without a .reuse, tools/sass_bank_check.py).  Packed fma.rn.f32x2 was built and measured
// SLOWER (80 vs 68 us at that stage: the pairs that start at an odd column are re-packed with ~50 MOVs per channel).
//
// A warp = one output row x 32 column quads (128 f2 columns) x one group of R vertical displacements; a CTA = 12 warps
// (TR rows x G groups), one CTA per SM (168 registers).  Its outputs are the 120 columns [x0, x0+120) (one quad of f2 halo
// per side: the two edge lanes do half-useful work, 6 %), so that the OUTPUT tile is 16-byte aligned and leaves by TMA:
// for displacement dx a thread's four values are the output columns 4*lane + j - (4+dx), re-aligned to whole quads with
// <= 2 lane shuffles, parked in the warp's staging tile and stored as two TMA tensor stores per tile
// ({120 columns, 1 row, half of the dx, R dy}; a store whose start column is not a multiple of 16 bytes is an
// `illegal instruction`, measured).  No CTA barrier anywhere: warps hand ring slots back through mbarriers.
// Operands arrive by TMA (f2 box {128, TR+2d, CC}, f1 box {136, TR, CC}, out-of-bounds zero fill = the zero padding of the
// correlation) through an mbarrier ring.  Requesting a chunk costs ~600 cycles of serial latency; the role rotates over
// the warps and refills the slot released TWO chunks ago, so that no warp ever waits to request (see the kernel).
//
// Measured on B200 (tools/time_corr_planar.py, L2 flushed, [2,32,270,480]): d=4 57.7 us vs 58.6 us pixel-major -- and no
// layout conversion for NCHW callers (the pixel-major route costs two transposes more, ~85 us).  The ablation
// (tools/dbg_planar_time.py) shows what is left: FMA loop alone 8.2 us per tile where issue-bound is 5.7; ring hand-offs
// 0.9; epilogue 1.5; operand latency not hidden by the 2-chunk ring 1.5; and ~8 us of launch per call.
#include "tc_common.cuh"
#include <type_traits>

namespace upf {

#ifndef UPF_PL_CC4
#define UPF_PL_CC4 4
#endif
// ablation switches in `flags` (tools/dbg_planar_time.py; results are garbage with any of them set)
constexpr int PL_DBG_NO_FMA = 0x100, PL_DBG_NO_LOADS = 0x200, PL_DBG_NO_STORE = 0x800, PL_DBG_NO_EPILOGUE = 0x2000;
constexpr int PL_COLS = 128;   // f2 columns per warp row: 32 lanes x one 16-byte column quad

template <int D>
struct PLCfg {
  static_assert(D >= 1 && D <= 4, "planar kernel: d <= 4");
  static constexpr int WIN = 2 * D + 1;
  static constexpr int R = D == 1 ? 3 : D == 2 ? 5 : D == 3 ? 4 : 3;      // f2 rows (vertical displacements) per thread
  static constexpr int G = (WIN + R - 1) / R;                            // warp groups along dy
  static constexpr int TR = 12 / G;                                      // tile rows: 12 warps = TR rows x G groups
  static constexpr int NW = TR * G, NT = NW * 32;
  static constexpr int OUTW = PL_COLS - 8;                               // output columns per tile (a quad of halo per side)
  static constexpr int F1W = PL_COLS + 8;                                // f1 tile width
  static constexpr int HROWS = TR + G * R - 1;                           // f2 rows staged (>= TR + 2d)
  static constexpr int CC = D == 2 ? 1 : (D == 4 ? UPF_PL_CC4 : 2);      // channels per ring slot
  static constexpr int F2_PLANE = HROWS * PL_COLS, F1_PLANE = TR * F1W;
  static constexpr int STAGE_FLOATS = (F2_PLANE + F1_PLANE) * CC;
  static constexpr int QA = (WIN + 1) / 2;                               // horizontal displacements of the first store (the second: WIN - QA)
  static constexpr int WOUT_FLOATS = ((R * QA * OUTW * 4 + 127) / 128) * 32;   // a warp's staging tile [R dy][QA dx][OUTW], 128-byte multiple
  static constexpr int OUT_FLOATS = NW * WOUT_FLOATS;
  static constexpr int STAGES_MAX = 8;
  static constexpr int stages() {
    int s = (222 * 1024 - OUT_FLOATS * 4) / (STAGE_FLOATS * 4);
    return s > STAGES_MAX ? STAGES_MAX : s;
  }
  static constexpr int S = stages();
  static constexpr int SMEM_BYTES = (S * STAGE_FLOATS + OUT_FLOATS) * 4 + 128;
  static_assert(S >= 3, "ring too shallow");
  static_assert((STAGE_FLOATS * 4) % 128 == 0 && (F2_PLANE * CC * 4) % 128 == 0, "TMA alignment");
};

__device__ __forceinline__ void pl_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void tma_prefetch_l2_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int D>
__global__ void __launch_bounds__(PLCfg<D>::NT, 1)
corr_planar_kernel(const __grid_constant__ CUtensorMap map1, const __grid_constant__ CUtensorMap map2,
                   const __grid_constant__ CUtensorMap mapo, const __grid_constant__ CUtensorMap mapo2, int H, int C, float slope, int flags,
                   int tiles_x, int tiles_y, int n2_shift, int N, int total_tiles) {
  pdl_prologue();
  using K = PLCfg<D>;
  constexpr int WIN = K::WIN, R = K::R, TR = K::TR, S = K::S, CC = K::CC;
  extern __shared__ __align__(128) float pl_smem_raw[];
  float* smem = pl_smem_raw + (((128u - (smem_u32(pl_smem_raw) & 127u)) & 127u) >> 2);
  __shared__ uint64_t s_full[K::STAGES_MAX], s_empty[K::STAGES_MAX];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), K::NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int yy = warp % TR, g = warp / TR;               // this warp's tile row and displacement group
  const int nchunks = (C + CC - 1) / CC;
  const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const uint32_t full0 = smem_u32(&s_full[0]), empty0 = smem_u32(&s_empty[0]);
  const uint32_t smem0 = smem_u32(smem);
  // operand offsets of this thread inside a ring slot
  const int off2 = (yy + g * R) * PL_COLS + lane * 4;
  const int off1 = K::F2_PLANE * CC + yy * K::F1W + lane * 4;

  // ---- operand requests.  Requesting a chunk is a serial chain of ~600 cycles (tile decode, `empty` wait, proxy fence, two
  // TMA instructions): done by one fixed thread inline with its FMA loop it set the pace of the whole CTA (measured: the
  // request time ADDED to the compute time, 0.31 + 0.49 us per chunk).  So the role rotates: chunk i of this CTA's stream is
  // requested by lane 0 of warp i % 12, S-1 chunks ahead, and everything it needs is recomputed from i (no shared state).
  const int total_chunks = my_tiles * nchunks;
  auto issue = [&](int i) {
    const int tloc = i / nchunks, chunk = i - tloc * nchunks;
    int t = blockIdx.x + tloc * gridDim.x;
    const int ty = t % tiles_y; t /= tiles_y;            // consecutive tiles are vertical neighbours (shared halo rows)
    const int tx = t % tiles_x;
    const int n = t / tiles_x;
    int n2 = n + n2_shift; if (n2 >= N) n2 -= N;
    const int x0 = tx * K::OUTW, y0 = ty * TR;
    const int round = i / S, islot = i - round * S;
    if (round > 0) {                                     // every warp has released the chunk that used this slot before
      mbar_wait(empty0 + islot * 8, (uint32_t)((round - 1) & 1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const uint32_t mb = full0 + islot * 8;
    const uint32_t dst = smem0 + islot * (K::STAGE_FLOATS * 4);
    mbar_expect_tx(mb, (uint32_t)(K::STAGE_FLOATS * 4));
    tma_load_4d(dst, &map2, mb, x0 - 4, y0 - D, chunk * CC, n2);                        // f2 columns [x0-4, x0+124)
    tma_load_4d(dst + K::F2_PLANE * CC * 4, &map1, mb, x0 - 8, y0, chunk * CC, n);      // f1 columns [x0-8, x0+128)
  };
  // The slot refilled at the top of iteration `it` is the one chunk it-2 used: every warp left it a whole chunk ago, so the
  // requesting warp does not wait (refilling the slot of chunk it-1 would make it wait for the slowest warp and then be the
  // slowest itself: request + compute in series again).  S-2 chunks are in flight.
  constexpr int AHEAD = S - 2;
  if (!(flags & PL_DBG_NO_LOADS) && lane == 0 && warp < AHEAD && warp < total_chunks) issue(warp);
  int next_issue = AHEAD;                                // chunk requested at the top of the next iteration ...
  int turn = AHEAD % K::NW;                              // ... by this warp

  float acc[R][WIN][4];
  int slot = 0;
  uint32_t phase = 0;
  int tile_id = blockIdx.x;
  for (int tloc = 0; tloc < my_tiles; ++tloc, tile_id += gridDim.x) {
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int q = 0; q < WIN; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][q][j] = 0.f;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      // refill: the chunk S-1 ahead goes to the slot the previous chunk has just left
      if (turn == warp && lane == 0 && next_issue < total_chunks && !(flags & PL_DBG_NO_LOADS)) issue(next_issue);
      ++next_issue;
      if (++turn == K::NW) turn = 0;
      __syncwarp();
      if (!(flags & PL_DBG_NO_LOADS)) mbar_wait(full0 + slot * 8, phase);
      const float* st = smem + slot * K::STAGE_FLOATS;
      const float* st2 = st + off2;
      const float* st1 = st + off1;
      if (!(flags & PL_DBG_NO_FMA))
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        float a[12];                                     // f1 columns q-4 .. q+7 of this thread's row
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(st1 + c * K::F1_PLANE + 4 * k);
          a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(st2 + c * K::F2_PLANE + r * PL_COLS);
          const float b[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < WIN; ++q)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[r][q][j] = fmaf(a[j - q + D + 4], b[j], acc[r][q][j]);
        }
      }
      __syncwarp();
      if (lane == 0) pl_mbar_arrive(empty0 + slot * 8);   // this warp is done with the slot
      if (++slot == S) { slot = 0; phase ^= 1u; }
    }
    // ---- tile done: mean over channels, LeakyReLU.  Warp-local epilogue (no CTA barrier): per horizontal displacement the
    // warp re-aligns its R rows to the output columns with lane shuffles, parks them in its private staging buffer and
    // hands them to the copy engine (one TMA tensor store of {OUTW columns, 1 row, R vertical displacements})
    int t = tile_id;
    const int ty = t % tiles_y; t /= tiles_y;
    const int tx = t % tiles_x;
    const int n = t / tiles_x;
    const int x0 = tx * K::OUTW, y0 = ty * TR;
    const float fC = (float)C, inv = __fdiv_rn(1.0f, fC);
    const bool pow2 = (C & (C - 1)) == 0;
    int wtid;                                            // (a thread id the compiler cannot hoist above the FMA loop)
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(wtid));
    float* const wst = smem + S * K::STAGE_FLOATS + (wtid >> 5) * K::WOUT_FLOATS;
    const bool row_ok = y0 + yy < H;
    // the store of the previous tile has had a whole tile's FMA loop to read the staging tile
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    // activation variants are chosen per tile, not per value (the epilogue is ALU-bound: ~100 values per thread):
    // FAST = power-of-two C (sum * (1/C) is the exact mean), 0 <= slope <= 1 (lrelu(v) = max(v, slope*v)), no TF32 rounding
    auto epilogue = [&](auto fast_tag, auto phase_tag) {
      constexpr bool FAST = decltype(fast_tag)::value;
      constexpr int QB = decltype(phase_tag)::value ? K::QA : 0, QE = decltype(phase_tag)::value ? WIN : K::QA;
#pragma unroll
      for (int q = QB; q < QE; ++q) {
        // this thread's values are the output columns x0 + 4*lane + j - e, e = 4 + dx = 4L + s
        const int e = 4 + (q - D), L = e >> 2, s = e & 3;
        const int m = (s == 3) ? lane - L - 1 : lane - L;  // the aligned output quad this lane assembles
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float sacc = acc[r][q][j];
            float v = __fmul_rn(sacc, inv);
            if (FAST) {
              o[j] = fmaxf(v, __fmul_rn(v, slope));
            } else {
              if (!pow2) v = __fmaf_rn(__fmaf_rn(-v, fC, sacc), inv, v);     // correctly rounded sum / C (torch.mean)
              o[j] = maybe_round(lrelu(v, slope), flags);
            }
          }
          float4 w;
          if (s == 0) {
            w = make_float4(o[0], o[1], o[2], o[3]);
          } else if (s == 1) {                             // own columns 1..3, the next lane's column 0
            const float n0 = __shfl_down_sync(0xffffffffu, o[0], 1);
            w = make_float4(o[1], o[2], o[3], n0);
          } else if (s == 2) {
            const float n0 = __shfl_down_sync(0xffffffffu, o[0], 1), n1 = __shfl_down_sync(0xffffffffu, o[1], 1);
            w = make_float4(o[2], o[3], n0, n1);
          } else {                                         // the previous lane's column 3, own columns 0..2
            const float p3 = __shfl_up_sync(0xffffffffu, o[3], 1);
            w = make_float4(p3, o[0], o[1], o[2]);
          }
          if (m >= 0 && m < K::OUTW / 4) *reinterpret_cast<float4*>(wst + (r * (QE - QB) + (q - QB)) * K::OUTW + 4 * m) = w;
        }
      }
    };
    const bool fast = pow2 && slope >= 0.f && slope <= 1.f && !(flags & UPF_FLAG_ROUND_TF32);
    // two stores per tile (the staging tile holds half of the horizontal displacements): the first is read by the copy
    // engine while the warp converts the second half, the second under the next tile's FMA loop
    if (flags & PL_DBG_NO_EPILOGUE) continue;
    if (fast) epilogue(std::true_type{}, std::false_type{}); else epilogue(std::false_type{}, std::false_type{});
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // generic writes -> TMA store reads
    __syncwarp();
    if (lane == 0) {
      if (row_ok && !(flags & PL_DBG_NO_STORE)) tma_store_5d(&mapo, smem_u32(wst), x0, y0 + yy, 0, g * R, n);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (fast) epilogue(std::true_type{}, std::true_type{}); else epilogue(std::false_type{}, std::true_type{});
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0 && row_ok && !(flags & PL_DBG_NO_STORE)) {
      tma_store_5d(&mapo2, smem_u32(wst), x0, y0 + yy, K::QA, g * R, n);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int D>
static int launch_corr_planar_t(const float* f1, const long long* p1, const float* f2, const long long* p2, float* out,
                                const long long* po, int N, int H, int W, int C, int shift, float slope, int flags, cudaStream_t st) {
  using K = PLCfg<D>;
  const int tiles_x = (W + K::OUTW - 1) / K::OUTW, tiles_y = (H + K::TR - 1) / K::TR;
  const long long tiles = (long long)tiles_x * tiles_y * N;
  static PerDeviceOnce attr_done;
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(corr_planar_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("corr_planar smem attr: %s", cudaGetErrorString(e)); return (int)e; }
    attr_done.mark();
  }
  CUtensorMap m1, m2, mo, mo2;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const cuuint64_t str1[3] = {(cuuint64_t)p1[0] * 4, (cuuint64_t)p1[1] * 4, (cuuint64_t)p1[2] * 4};
    const cuuint32_t box1[4] = {(cuuint32_t)K::F1W, (cuuint32_t)K::TR, (cuuint32_t)K::CC, 1};
    MapKey k1{f1, p1[0], p1[1], p1[2], ((long long)H << 32) | (unsigned)W, ((long long)N << 40) | ((long long)C << 8) | (17000 + D)};
    int e = encode_cached(k1, &m1, 4, const_cast<float*>(f1), dims, str1, box1, estr, 0);
    if (e) return e;
    const cuuint64_t str2[3] = {(cuuint64_t)p2[0] * 4, (cuuint64_t)p2[1] * 4, (cuuint64_t)p2[2] * 4};
    const cuuint32_t box2[4] = {PL_COLS, (cuuint32_t)K::HROWS, (cuuint32_t)K::CC, 1};
    MapKey k2{f2, p2[0], p2[1], p2[2], ((long long)H << 32) | (unsigned)W, ((long long)N << 40) | ((long long)C << 8) | (18000 + D)};
    e = encode_cached(k2, &m2, 4, const_cast<float*>(f2), dims, str2, box2, estr, 0);
    if (e) return e;
    // output [N][dy][dx][H][W]: a warp stores {OUTW columns, 1 row, half of the dx, R consecutive dy} twice
    const cuuint64_t odims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)K::WIN, (cuuint64_t)K::WIN, (cuuint64_t)N};
    const cuuint64_t ostr[4] = {(cuuint64_t)po[0] * 4, (cuuint64_t)po[1] * 4, (cuuint64_t)po[1] * 4 * K::WIN, (cuuint64_t)po[2] * 4};
    const cuuint32_t obox[5] = {(cuuint32_t)K::OUTW, 1, (cuuint32_t)K::QA, (cuuint32_t)K::R, 1};
    const cuuint32_t oestr[5] = {1, 1, 1, 1, 1};
    MapKey ko{out, po[0], po[1], po[2], ((long long)H << 32) | (unsigned)W, ((long long)N << 40) | (19000 + D)};
    e = encode_cached(ko, &mo, 5, out, odims, ostr, obox, oestr, 0);
    if (e) return e;
    const cuuint32_t obox2[5] = {(cuuint32_t)K::OUTW, 1, (cuuint32_t)(K::WIN - K::QA), (cuuint32_t)K::R, 1};
    MapKey ko2{out, po[0], po[1], po[2], ((long long)H << 32) | (unsigned)W, ((long long)N << 40) | (19500 + D)};
    e = encode_cached(ko2, &mo2, 5, out, odims, ostr, obox2, oestr, 0);
    if (e) return e;
  }
  const unsigned grid = (unsigned)(tiles < UPF_NUM_SMS ? tiles : UPF_NUM_SMS);
  UPF_LAUNCH((corr_planar_kernel<D>), grid, K::NT, K::SMEM_BYTES, st, m1, m2, mo, mo2, H, C, slope, flags, tiles_x, tiles_y, shift, N, (int)tiles);
  return check_launch("corr_planar");
}

}  // namespace upf

extern "C" int upf_corr_lrelu_fwd_planar(const float* f1, const long long* pitch1, const float* f2, const long long* pitch2,
                                         float* out, const long long* pitch_out, int N, int H, int W, int C, int max_disp,
                                         int f2_batch_shift, float slope, int flags, void* stream) {
  using namespace upf;
  UPF_REQUIRE(f1 && f2 && out && pitch1 && pitch2 && pitch_out, "corr_planar: null argument");
  UPF_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0, "corr_planar: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
  UPF_REQUIRE(f2_batch_shift >= 0 && f2_batch_shift < N, "corr_planar: batch shift out of range");
  UPF_REQUIRE(max_disp >= 1 && max_disp <= 6, "corr_planar: max_disp %d not in 1..6", max_disp);
  if (max_disp > 4) { set_error("corr_planar: max_disp > 4 is served by the pixel-major kernel (upf_corr_lrelu_fwd)"); return UPF_ENOTSUP; }
  const int nout = (2 * max_disp + 1) * (2 * max_disp + 1);
  UPF_REQUIRE(pitch1[0] >= W && pitch2[0] >= W && pitch_out[0] >= W, "corr_planar: row pitch smaller than W");
  UPF_REQUIRE(pitch1[1] >= pitch1[0] * H && pitch2[1] >= pitch2[0] * H && pitch_out[1] >= pitch_out[0] * H, "corr_planar: plane pitch too small");
  UPF_REQUIRE(pitch1[2] >= pitch1[1] * C && pitch2[2] >= pitch2[1] * C && pitch_out[2] >= pitch_out[1] * nout, "corr_planar: image pitch too small");
  // TMA: 16-byte aligned bases and pitches
  const long long* ps[3] = {pitch1, pitch2, pitch_out};
  for (int t = 0; t < 3; ++t)
    for (int i = 0; i < 3; ++i)
      if (ps[t][i] % 4 != 0) { set_error("corr_planar: pitches must be multiples of 4 elements (16 bytes)"); return UPF_ENOTSUP; }
  if (!aligned16(f1) || !aligned16(f2) || !aligned16(out)) { set_error("corr_planar: tensors must be 16-byte aligned"); return UPF_ENOTSUP; }
  cudaStream_t st = (cudaStream_t)stream;
  switch (max_disp) {
    case 1: return launch_corr_planar_t<1>(f1, pitch1, f2, pitch2, out, pitch_out, N, H, W, C, f2_batch_shift, slope, flags, st);
    case 2: return launch_corr_planar_t<2>(f1, pitch1, f2, pitch2, out, pitch_out, N, H, W, C, f2_batch_shift, slope, flags, st);
    case 3: return launch_corr_planar_t<3>(f1, pitch1, f2, pitch2, out, pitch_out, N, H, W, C, f2_batch_shift, slope, flags, st);
    default: return launch_corr_planar_t<4>(f1, pitch1, f2, pitch2, out, pitch_out, N, H, W, C, f2_batch_shift, slope, flags, st);
  }
}
