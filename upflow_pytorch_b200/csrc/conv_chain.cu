// A CHAIN of stride-1 convolutions on one small feature map in ONE persistent launch (tcgen05, TF32).
//
// The three coarse pyramid levels of the decoder (6x20, 12x39, 24x78 at KITTI size: 240 .. 3744 pixels for 148 SMs)
// run FlowEstimatorDense_v2 (model/pwc_modules.py:279-286), ContextNetwork_v2_ (:401-412) and the SGU dense block
// (model/upflow.py:52-60) as 19 dependent convolutions per level.  As separate launches of conv_tc_kernel each layer
// costs 7-18 us for 0.1-5 GFLOP of work: launch -> prologue (barriers, TMEM) -> first TMA -> MMA -> commit -> epilogue
// -> grid drain (profiles/r2_launch_table_kitti_events.txt: 51 layers, 618 us, 22 % of the forward).  Here the layers
// of a chain are a PROGRAM in the kernel's parameter space; a co-resident grid of thread-block clusters walks it:
//
//   layer  = work items (pixel tile of 128, N tile of BN output channels); item i goes to cluster i % G
//   item   = K (taps x 32-channel blocks) split over the CTAs of the cluster, each CTA: TMA ring -> two MMA issuers ->
//            TMEM -> partial tile parked in its shared memory -> the cluster reduces through distributed shared memory
//            in rank order (bitwise reproducible) -> bias + LeakyReLU (+ residual) -> global
//   between layers: a grid-wide barrier (one release-add per CTA, acquire-polled) -- but ONLY the TMA-producer thread
//            waits for it, and only for the ACTIVATION loads: the ring / TMEM / cluster hand-offs are mbarriers whose
//            phases run on across items and layers, so the producer has the next layer's WEIGHT tiles in flight while the
//            previous layer drains, and no role ever meets a CTA-wide barrier inside the program.
//   activations written by the epilogue through the generic proxy are read by TMA (async proxy) of OTHER SMs in the next
//            layer: writers fence.proxy.async + __threadfence before the release-add, the producer fence.proxy.async
//            after its acquire.
// Same operand layout, packed weights, tensor maps and arithmetic as conv_tc.cu (only the K partition differs, i.e. the
// order of fp32 partial sums).
#include "tc_common.cuh"

namespace upf {

constexpr int CH_THREADS = 224;            // warps: 0 TMA producer, 1 and 6 MMA issuers, 2-5 epilogue / reducers
constexpr int CH_NSTAGE = 4;               // even: issuer w owns the slots of parity w
constexpr int CH_STAGE_BYTES = 32 * 1024;  // [A: 128 pixels x 128 B][B: <= 128 weight rows x 128 B]
constexpr int CH_A_BYTES = 128 * 128;
constexpr int CH_PART_FLOATS = 128 * 132;  // parked partial tile, pitch BN + 4 <= 132 floats
constexpr int CH_MAX_LAYERS = 16;
constexpr int CH_MAX_CLUSTER = 8;

struct alignas(64) ChLayer {
  CUtensorMap mx, mw;
  float* out; const float* res; const float* bias; float* out2;
  int ldo, ldr, ldo2;
  int Cout, BN, ntiles_n;
  int TH, TW, tiles_x, tiles_y, tiles;   // tiles = tiles_x * tiles_y * N
  int ks, dil, kblocks;
  int splits, ips, iters_all;
  int bw, bh, nbx, nby, b_rows, nbb;     // an operand tile may travel as several smaller TMA boxes (A/B switch; one box per tile is
                                         // faster: 32-row boxes measured 3.20 vs 3.06 ms per KITTI forward)
  float slope;
  int flags;
};

struct ChProgram {
  int n_layers, Ho, Wo, pad_;
  long long* probe;                        // debug: [layer][16] clock64 stamps of CTA 0's roles (nullptr = off)
  unsigned* sync;                          // [0] arrivals of the grid barrier, [1] finished CTAs (the last one clears both)
  ChLayer L[CH_MAX_LAYERS];
};
static_assert(sizeof(ChProgram) < 32000, "the program travels in the kernel's parameter space");

// ---- cluster-scope mbarrier hand-offs
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP_C:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE_C;\n\t"
      "bra WAIT_LOOP_C;\n\t"
      "DONE_C:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

#define CH_STAMP(l, k) do { if (prog.probe && blockIdx.x == 0) prog.probe[(l) * 16 + (k)] = clock64(); } while (0)

// what one CTA does for item `item` of layer Lr: its pixel tile, N tile and K range
struct ChItem { int n, x0, y0, co0, it_begin, iters; };
__device__ __forceinline__ ChItem ch_item(const ChLayer& Lr, int item, int rank) {
  ChItem w;
  int tile = item % Lr.tiles;
  const int nt = item / Lr.tiles;
  const int tx = tile % Lr.tiles_x; tile /= Lr.tiles_x;
  const int ty = tile % Lr.tiles_y;
  w.n = tile / Lr.tiles_y;
  w.x0 = tx * Lr.TW; w.y0 = ty * Lr.TH;
  w.co0 = nt * Lr.BN;
  w.it_begin = rank * Lr.ips;
  int iters = Lr.iters_all - w.it_begin;
  if (iters > Lr.ips) iters = Lr.ips;
  if (rank >= Lr.splits || iters < 0) iters = 0;
  w.iters = iters;
  return w;
}

__device__ __forceinline__ void ch_load_a(const ChLayer& Lr, const ChItem& w, uint32_t a_dst, uint32_t fb, int gi) {
  const int half = (Lr.ks - 1) / 2;
  const int tap = gi / Lr.kblocks, kb = gi - tap * Lr.kblocks;
  const int ky = tap / Lr.ks, kx = tap - ky * Lr.ks;
  const int cx = w.x0 + (kx - half) * Lr.dil, cy = w.y0 + (ky - half) * Lr.dil;
  for (int jy = 0; jy < Lr.nby; ++jy)
    for (int jx = 0; jx < Lr.nbx; ++jx)
      tma_load_4d(a_dst + (uint32_t)((jy * Lr.bh * Lr.TW + jx * Lr.bw) * 128), &Lr.mx, fb, kb * 32, cx + jx * Lr.bw, cy + jy * Lr.bh, w.n);
}
__device__ __forceinline__ void ch_load_b(const ChLayer& Lr, const ChItem& w, uint32_t b_dst, uint32_t fb, int gi) {
  const int tap = gi / Lr.kblocks, kb = gi - tap * Lr.kblocks;
  for (int jb = 0; jb < Lr.nbb; ++jb)
    tma_load_3d(b_dst + (uint32_t)(jb * Lr.b_rows * 128), &Lr.mw, fb, kb * 32, w.co0 + jb * Lr.b_rows, tap);
}

__global__ void __launch_bounds__(CH_THREADS)
conv_chain_kernel(const __grid_constant__ ChProgram prog) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* part = reinterpret_cast<float*>(base + (size_t)CH_NSTAGE * CH_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(part + CH_PART_FLOATS);
  uint64_t* full = bars;                              // [NSTAGE] TMA -> MMA
  uint64_t* empty = bars + CH_NSTAGE;                 // [NSTAGE] MMA -> TMA
  uint64_t* accum_full = bars + 2 * CH_NSTAGE;        // MMA (both issuers) -> epilogue
  uint64_t* acc_empty = accum_full + 1;               // epilogue (4 warps) -> MMA: TMEM drained
  uint64_t* parts_full = accum_full + 2;              // every CTA's 4 epilogue warps: partial tiles of the cluster parked
  uint64_t* parts_empty = accum_full + 3;             // every CTA's 4 epilogue warps: done reading this CTA's partial tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 4);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [128]

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CS = (int)cluster_nctarank();
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x / CS;
  const int G = gridDim.x / CS;
  const unsigned nctas = gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[0].mx) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[0].mw) : "memory");
    for (int s = 0; s < CH_NSTAGE; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), 2);
    mbar_init(smem_u32(acc_empty), 4);
    mbar_init(smem_u32(parts_full), 4 * CS);
    mbar_init(smem_u32(parts_empty), 4 * CS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                  // barriers initialised before any peer's remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (one thread walks the whole program) =====================
    if (elect_one()) {
      uint32_t git = 0;
      for (int l = 0; l < prog.n_layers; ++l) {
        const ChLayer& Lr = prog.L[l];
        const int items = Lr.tiles * Lr.ntiles_n;
        const uint32_t b_bytes = (uint32_t)Lr.BN * 128u;
        bool synced = (l == 0);
        if (l + 1 < prog.n_layers) {
          asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[l + 1].mx) : "memory");
          asm volatile("prefetch.tensormap [%0];" ::"l"(&prog.L[l + 1].mw) : "memory");
        }
        for (int item = cluster_id; item < items; item += G) {
          const ChItem w = ch_item(Lr, item, rank);
          if (w.iters <= 0) continue;
          int it = 0;
          if (!synced) {
            // the first ring slots of a layer: weights now, activations once every CTA has finished the previous layer
            const int pre = w.iters < CH_NSTAGE ? w.iters : CH_NSTAGE;
            for (int j = 0; j < pre; ++j) {
              const uint32_t g = git + (uint32_t)j;
              const uint32_t s = g % CH_NSTAGE, ph = (g / CH_NSTAGE) & 1u;
              mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
              const uint32_t fb = smem_u32(&full[s]);
              mbar_expect_tx(fb, (uint32_t)CH_A_BYTES + b_bytes);
              ch_load_b(Lr, w, smem_u32(base + (size_t)s * CH_STAGE_BYTES) + CH_A_BYTES, fb, w.it_begin + j);
            }
            CH_STAMP(l, 0);
            const unsigned target = (unsigned)l * nctas;
            while (ld_acquire_gpu(prog.sync) < target) { }
            fence_proxy_async();
            CH_STAMP(l, 1);
            for (int j = 0; j < pre; ++j) {
              const uint32_t g = git + (uint32_t)j;
              const uint32_t s = g % CH_NSTAGE;
              ch_load_a(Lr, w, smem_u32(base + (size_t)s * CH_STAGE_BYTES), smem_u32(&full[s]), w.it_begin + j);
            }
            CH_STAMP(l, 2);
            it = pre;
            synced = true;
          }
          for (; it < w.iters; ++it) {
            const uint32_t g = git + (uint32_t)it;
            const uint32_t s = g % CH_NSTAGE, ph = (g / CH_NSTAGE) & 1u;
            mbar_wait(smem_u32(&empty[s]), ph ^ 1u);
            const uint32_t a_dst = smem_u32(base + (size_t)s * CH_STAGE_BYTES);
            const uint32_t fb = smem_u32(&full[s]);
            mbar_expect_tx(fb, (uint32_t)CH_A_BYTES + b_bytes);
            ch_load_a(Lr, w, a_dst, fb, w.it_begin + it);
            ch_load_b(Lr, w, a_dst + CH_A_BYTES, fb, w.it_begin + it);
          }
          git += (uint32_t)w.iters;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == 6) {
    // ===================== MMA issuers: issuer w takes the ring iterations of parity w into its own accumulator =====
    const uint32_t wi = warp == 1 ? 0u : 1u;
    const uint32_t tacc = tmem_base + wi * 128u;
    uint32_t git = 0, nitem = 0;
    for (int l = 0; l < prog.n_layers; ++l) {
      const ChLayer& Lr = prog.L[l];
      const int items = Lr.tiles * Lr.ntiles_n;
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Lr.BN >> 3) << 17) | ((128u >> 4) << 24);
      for (int item = cluster_id; item < items; item += G) {
        const ChItem w = ch_item(Lr, item, rank);
        if (w.iters <= 0) continue;
        mbar_wait(smem_u32(acc_empty), (nitem & 1u) ^ 1u);      // the epilogue has read the previous item's accumulators
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        bool first = true;
        for (int it = 0; it < w.iters; ++it) {
          const uint32_t g = git + (uint32_t)it;
          if ((g & 1u) != wi) continue;
          const uint32_t s = g % CH_NSTAGE, ph = (g / CH_NSTAGE) & 1u;
          mbar_wait(smem_u32(&full[s]), ph);
          if (first && wi == (git & 1u) && item == cluster_id && lane == 0) CH_STAMP(l, 3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(base + (size_t)s * CH_STAGE_BYTES);
            const uint64_t da = umma_desc_sw128(a_addr);
            const uint64_t db = umma_desc_sw128(a_addr + CH_A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_tf32(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (!first || k > 0) ? 1u : 0u);
            umma_commit(smem_u32(&empty[s]));
          }
          __syncwarp();
          first = false;
        }
        if (wi == 0 && item == cluster_id && lane == 0) CH_STAMP(l, 4);
        if (elect_one()) umma_commit(smem_u32(accum_full));     // both issuers, every item (arrives at once if it issued nothing)
        __syncwarp();
        git += (uint32_t)w.iters;
        ++nitem;
      }
    }
  } else {
    // ===================== epilogue + cluster reduction (warps 2..5) =====================
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;                  // pixel index inside the tile
    const int et = (int)threadIdx.x - 64;           // 0..127
    const uint32_t part_addr = smem_u32(part);
    const int rows_per = 128 / CS;
    uint32_t git = 0, nitem = 0, eitem = 0;
    for (int l = 0; l < prog.n_layers; ++l) {
      const ChLayer& Lr = prog.L[l];
      const int items = Lr.tiles * Lr.ntiles_n;
      const int pitch = Lr.BN + 4;
      const bool vec_out = ((Lr.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(Lr.out) & 15) == 0);
      for (int item = cluster_id; item < items; item += G) {
        const ChItem w = ch_item(Lr, item, rank);
        mbar_wait_cluster(smem_u32(parts_empty), (eitem & 1u) ^ 1u);    // every peer has read this CTA's previous partial tile
        if (w.iters > 0) {
          // which issuers contributed: ring iterations git .. git+iters-1, issuer = parity
          const bool has0 = (w.iters >= 2) || ((git & 1u) == 0u);
          const bool has1 = (w.iters >= 2) || ((git & 1u) == 1u);
          mbar_wait(smem_u32(accum_full), nitem & 1u);
          if (et == 0 && item == cluster_id) CH_STAMP(l, 5);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          float* mine = part + row * pitch;
          for (int c0 = 0; c0 < Lr.BN; c0 += 16) {
            uint32_t v[16], v2[16];
            if (has0) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            if (has1) tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(128 + c0), v2);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = has0 ? __uint_as_float(v[j]) : 0.f;
              const float b = has1 ? __uint_as_float(v2[j]) : 0.f;
              v[j] = __float_as_uint(a + b);
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(mine + c0 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_local(smem_u32(acc_empty));
          ++nitem;
        }
        s_bias[et] = (et < Lr.BN && w.co0 + et < Lr.Cout) ? __ldg(Lr.bias + w.co0 + et) : 0.f;
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        __syncwarp();
        if (lane < CS) mbar_arrive_remote(smem_u32(parts_full), (uint32_t)lane);
        asm volatile("bar.sync 1, 128;" ::: "memory");                  // s_bias written
        if (et == 0 && item == cluster_id) CH_STAMP(l, 6);
        mbar_wait_cluster(smem_u32(parts_full), eitem & 1u);
        if (et == 0 && item == cluster_id) CH_STAMP(l, 7);            // every CTA of the cluster has parked its partial tile
        // ---- reduce rows [rank*rows_per, +rows_per) over the K splits in rank order, finish, store
        const int c4n = Lr.BN >> 2;
        for (int u = et; u < rows_per * c4n; u += 128) {
          const int rl = u / c4n, c4 = u - rl * c4n;
          const int r = rank * rows_per + rl;
          const uint32_t off = part_addr + (uint32_t)(r * pitch + c4 * 4) * 4u;
          float4 t[CH_MAX_CLUSTER];            // every peer's load in flight at once, summed in rank order
#pragma unroll
          for (int sp = 0; sp < CH_MAX_CLUSTER; ++sp)
            if (sp < Lr.splits) t[sp] = ld_dsmem_f4(off, (uint32_t)sp);
          float4 acc = t[0];
#pragma unroll
          for (int sp = 1; sp < CH_MAX_CLUSTER; ++sp)
            if (sp < Lr.splits) { acc.x += t[sp].x; acc.y += t[sp].y; acc.z += t[sp].z; acc.w += t[sp].w; }
          const int py = w.y0 + r / Lr.TW, px = w.x0 + r % Lr.TW;
          if (py < prog.Ho && px < prog.Wo) {
            const size_t pix = ((size_t)w.n * prog.Ho + py) * prog.Wo + px;
            const int co = w.co0 + c4 * 4;
            float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (co + j < Lr.Cout) {
                float a = lrelu(v[j] + s_bias[c4 * 4 + j], Lr.slope);
                if (Lr.res) a += __ldcg(Lr.res + pix * Lr.ldr + co + j);
                v[j] = maybe_round(a, Lr.flags);
              }
            float* o = Lr.out + pix * Lr.ldo;
            if (vec_out && co + 4 <= Lr.Cout) {
              *reinterpret_cast<float4*>(o + co) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (co + j < Lr.Cout) o[co + j] = v[j];
            }
            if (Lr.out2) {                         // a second, TF32-rounded copy (the context network's flow input slot)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (co + j < Lr.Cout) Lr.out2[pix * Lr.ldo2 + co + j] = round_tf32(v[j]);
            }
          }
        }
        if (et == 0 && item == cluster_id) CH_STAMP(l, 8);
        __syncwarp();
        if (lane < CS) mbar_arrive_remote(smem_u32(parts_empty), (uint32_t)lane);
        asm volatile("bar.sync 1, 128;" ::: "memory");                  // s_bias may be rewritten
        git += (uint32_t)w.iters;
        ++eitem;
      }
      // ---- end of the layer: this CTA's outputs are published; its arrival follows everyone's arrival for the previous layer
      if (et == 0) CH_STAMP(l, 9);
      fence_proxy_async();
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        CH_STAMP(l, 10);
        const unsigned target = (unsigned)l * nctas;
        while (ld_acquire_gpu(prog.sync) < target) { }
        red_release_gpu(prog.sync, 1u);
        CH_STAMP(l, 11);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  cluster_sync_all();                  // nobody leaves while a peer may still read its partial tile
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
  if (threadIdx.x == 0) {
    // the last CTA to finish clears the barrier words for the next launch (every poller has left by then)
    __threadfence();
    const unsigned done = atomicAdd(prog.sync + 1, 1u);
    if (done == nctas - 1) {
      prog.sync[0] = 0u;
      prog.sync[1] = 0u;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------- host
static long long* g_chain_probe = nullptr;
static int g_chain_cs = 8;             // cluster size (K split), upf_debug_conv_chain
static int g_chain_box_rows = 128;     // rows per TMA box (upf_debug_conv_chain bits 8..15 of cluster_size; 128 = one box per tile)
static int g_chain_clusters = 0;       // clusters in the grid (0 = as many as are co-resident, see below)
static unsigned* g_chain_sync[64] = {nullptr};
static int g_chain_max_clusters[64] = {0};
static int g_chain_cs_of_max[64] = {0};

static size_t chain_smem_bytes() {
  return (size_t)CH_NSTAGE * CH_STAGE_BYTES + (size_t)CH_PART_FLOATS * 4 + (2 * CH_NSTAGE + 4) * 8 + 16 + 128 * 4 + 1024;
}

static void pick_tile_128(int H, int W, int* TH, int* TW) {
  long long best = -1;
  for (int tw = 8; tw <= 128; tw <<= 1) {
    const int th = 128 / tw;
    const long long cover = (long long)((H + th - 1) / th) * ((W + tw - 1) / tw);
    if (best < 0 || cover < best || (cover == best && tw == 16)) { best = cover; *TH = th; *TW = tw; }
  }
}

}  // namespace upf

/* debug: device buffer of 16 x 16 int64 receiving CTA 0's per-layer clock64 stamps (NULL = off): 0 producer reaches the grid
 * barrier, 1 passes it, 2 activation loads out, 3 first operands landed, 4 last MMA issued, 5 accumulators complete, 6 partial
 * tile parked, 7 cluster's partial tiles complete, 8 reduced + stored, 9 layer's items done, 10 fenced, 11 arrived */
extern "C" int upf_debug_conv_chain_probe(void* device_buffer_256x_int64) {
  upf::g_chain_probe = reinterpret_cast<long long*>(device_buffer_256x_int64);
  return 0;
}

extern "C" int upf_debug_conv_chain(int cluster_size, int n_clusters) {
  using namespace upf;
  const int box_rows = (cluster_size >> 8) & 0xff;
  cluster_size &= 0xff;
  UPF_REQUIRE(cluster_size == 1 || cluster_size == 2 || cluster_size == 4 || cluster_size == 8, "conv_chain: cluster size %d not in {1,2,4,8}", cluster_size);
  UPF_REQUIRE(box_rows == 0 || box_rows == 8 || box_rows == 16 || box_rows == 32 || box_rows == 64 || box_rows == 128, "conv_chain: box rows %d", box_rows);
  g_chain_box_rows = box_rows ? box_rows : 128;
  g_chain_cs = cluster_size;
  g_chain_clusters = n_clusters < 0 ? 0 : n_clusters;
  return 0;
}

extern "C" int upf_conv_chain_fwd(const upf_chain_layer* layers, int n_layers, int N, int H, int W, void* stream) {
  using namespace upf;
  UPF_REQUIRE(layers && n_layers >= 1 && n_layers <= CH_MAX_LAYERS, "conv_chain: 1..%d layers", CH_MAX_LAYERS);
  UPF_REQUIRE(N > 0 && H > 0 && W > 0, "conv_chain: empty feature map");
  cudaStream_t st = (cudaStream_t)stream;
  int dev = 0;
  cudaGetDevice(&dev);
  UPF_REQUIRE(dev < 64, "conv_chain: device ordinal %d", dev);
  const int CS = g_chain_cs;
  const size_t smem = chain_smem_bytes();
  if (!g_chain_sync[dev]) {
    cudaError_t e = cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaMalloc(&g_chain_sync[dev], 64);
    if (e == cudaSuccess) e = cudaMemset(g_chain_sync[dev], 0, 64);
    if (e != cudaSuccess) { set_error("conv_chain init: %s", cudaGetErrorString(e)); g_chain_sync[dev] = nullptr; return (int)e; }
  }
  if (g_chain_cs_of_max[dev] != CS) {
    // every CTA of the grid must be resident at once (the layers are separated by a grid-wide barrier)
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3(CS * UPF_NUM_SMS);
    q.blockDim = dim3(CH_THREADS);
    q.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = (unsigned)CS; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    q.attrs = a; q.numAttrs = 1;
    int nc = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, conv_chain_kernel, &q);
    if (e != cudaSuccess || nc < 1) { set_error("conv_chain occupancy: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return e != cudaSuccess ? (int)e : UPF_EDRIVER; }
    g_chain_max_clusters[dev] = nc;
    g_chain_cs_of_max[dev] = CS;
  }
  int G = g_chain_max_clusters[dev];
  if (g_chain_clusters > 0 && g_chain_clusters < G) G = g_chain_clusters;

  ChProgram prog;
  memset(&prog, 0, sizeof(prog));
  prog.n_layers = n_layers; prog.Ho = H; prog.Wo = W;
  prog.sync = g_chain_sync[dev];
  prog.probe = g_chain_probe;
  int TH = 8, TW = 16;
  pick_tile_128(H, W, &TH, &TW);
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int tiles = tiles_x * tiles_y * N;
  for (int l = 0; l < n_layers; ++l) {
    const upf_chain_layer& a = layers[l];
    ChLayer& Lr = prog.L[l];
    UPF_REQUIRE(a.x && a.w_packed && a.bias && a.out, "conv_chain: layer %d: null pointer", l);
    UPF_REQUIRE(a.Cin > 0 && a.Cout > 0 && (a.ksize == 1 || a.ksize == 3) && a.dilation >= 1, "conv_chain: layer %d: bad shape", l);
    UPF_REQUIRE((a.ldx % 4) == 0 && aligned16(a.x) && aligned16(a.w_packed), "conv_chain: layer %d: input pitch/pointer must be 16-byte aligned", l);
    UPF_REQUIRE(a.ldx >= a.Cin && a.ldo >= a.Cout && (!a.residual || a.ldr >= a.Cout) && (!a.out2 || a.ldo2 >= a.Cout), "conv_chain: layer %d: pitch smaller than the channel count", l);
    const int cout_pad = (a.Cout + 15) & ~15;
    const int kblocks = (a.Cin + 31) / 32, cin_pad = kblocks * 32;
    const int taps = a.ksize * a.ksize;
    const int iters_all = taps * kblocks;
    // N tiles x K splits: the layer is a few rounds of (operand fetch at ~58 B/clk/SM + hand-offs + reduction); narrow N tiles
    // put more clusters on a layer at the price of re-fetching the activation tile
    int best_nt = 1, best_bn = ((cout_pad + ((cout_pad + 127) / 128) - 1) / ((cout_pad + 127) / 128) + 15) & ~15, best_sp = 1;
    double best = 1e30;
    for (int nt = (cout_pad + 127) / 128; nt <= 8; nt <<= 1) {
      const int bn = ((cout_pad + nt - 1) / nt + 15) & ~15;
      if (nt > 1 && (nt - 1) * bn >= cout_pad) break;
      const int rounds = (tiles * nt + G - 1) / G;
      for (int sp = 1; sp <= CS; sp <<= 1) {
        const int ips = (iters_all + sp - 1) / sp;
        if (sp > 1 && (ips < 2 || (sp - 1) * ips >= iters_all)) break;
        const double cost = rounds * (ips * (CH_A_BYTES + bn * 128) / 58.0 + 1500.0 + 40.0 * sp);
        if (cost < best * 0.97) { best = cost; best_nt = nt; best_bn = bn; best_sp = sp; }
      }
    }
    const int BN = best_bn;
    const int box_rows = g_chain_box_rows;
    const int bw = TW < box_rows ? TW : box_rows;
    int bh = box_rows / bw;
    if (bh > TH) bh = TH;
    int b_rows = BN < box_rows ? BN : box_rows;
    while (BN % b_rows) b_rows -= 8;
    {
      const cuuint64_t dims[4] = {(cuuint64_t)a.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      const cuuint64_t strides[3] = {(cuuint64_t)a.ldx * 4, (cuuint64_t)W * a.ldx * 4, (cuuint64_t)H * W * a.ldx * 4};
      const cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, 1};
      const cuuint32_t estr[4] = {1, 1, 1, 1};
      MapKey key{a.x, a.ldx, ((long long)H << 32) | (unsigned)W, ((long long)N << 32) | (unsigned)a.Cin, (bw * 1000 + bh) * 4 + 1, 4};
      int e = encode_cached(key, &Lr.mx, 4, const_cast<float*>(a.x), dims, strides, box, estr);
      if (e) return e;
    }
    {
      const cuuint64_t dims[3] = {(cuuint64_t)cin_pad, (cuuint64_t)cout_pad, (cuuint64_t)taps};
      const cuuint64_t strides[2] = {(cuuint64_t)cin_pad * 4, (cuuint64_t)cin_pad * cout_pad * 4};
      const cuuint32_t box[3] = {32, (cuuint32_t)b_rows, 1};
      const cuuint32_t estr[3] = {1, 1, 1};
      MapKey key{a.w_packed, cin_pad, b_rows, taps, cout_pad, 3};
      int e = encode_cached(key, &Lr.mw, 3, const_cast<float*>(a.w_packed), dims, strides, box, estr);
      if (e) return e;
    }
    Lr.out = a.out; Lr.res = a.residual; Lr.bias = a.bias; Lr.out2 = a.out2;
    Lr.ldo = a.ldo; Lr.ldr = a.ldr; Lr.ldo2 = a.ldo2;
    Lr.Cout = a.Cout; Lr.BN = BN; Lr.ntiles_n = best_nt;
    Lr.TH = TH; Lr.TW = TW; Lr.tiles_x = tiles_x; Lr.tiles_y = tiles_y; Lr.tiles = tiles;
    Lr.ks = a.ksize; Lr.dil = a.dilation; Lr.kblocks = kblocks;
    Lr.splits = best_sp; Lr.ips = (iters_all + best_sp - 1) / best_sp; Lr.iters_all = iters_all;
    Lr.bw = bw; Lr.bh = bh; Lr.nbx = TW / bw; Lr.nby = TH / bh; Lr.b_rows = b_rows; Lr.nbb = BN / b_rows;
    Lr.slope = a.slope; Lr.flags = a.flags;
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(G * CS));
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_tc_pdl ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_chain_kernel, prog);
  if (e != cudaSuccess) { set_error("conv_chain launch: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return (int)e; }
  return check_launch("conv_chain");
}
