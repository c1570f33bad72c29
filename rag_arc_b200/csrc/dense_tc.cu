// Dense scoring on the 5th-generation tensor cores with the per-query selection fused into the
// epilogue (sm_100a only).
//
//   scores[q, r] = sum_k Q[q,k] * X[r,k]          bf16/fp16 operands, fp32 accumulate in TMEM
//
// Work item = (query block, contiguous slice of corpus tiles); persistent CTAs (CG=1) or CTA pairs
// (CG=2, one thread-block cluster of two SMs) walk their items.  Per corpus tile (256 rows) and per
// 64-wide k block one elected thread per CTA issues TMA loads into 128B-swizzled shared memory
// (mbarrier ring): its own 128 query rows and - CG=1: all 256 corpus rows / CG=2: its half (128)
// of them.  One elected thread (of the leader CTA) issues tcgen05.mma
//   CG=1: M=128 x N=256 x K=16, cta_group::1      CG=2: M=256 x N=256 x K=16, cta_group::2
// into one of two 256-column TMEM accumulators; with CG=2 the pair shares the corpus tile, which
// halves its L2->SM traffic, and the smaller per-CTA stage allows a 6-deep ring.
// Four epilogue warps per CTA read the finished accumulator with tcgen05.ld (32 lanes x 32 columns:
// one thread == one query row), build a branch-free pass mask against the row's running threshold,
// and append the survivors to the (item,row) candidate list; lists are pruned warp-cooperatively
// (common.cuh).  The nq x n score matrix never leaves the SM.  While the epilogue drains
// accumulator b, the MMA warp fills b^1.
// CL > 1 (CG=2 only): clusters of CL pairs take CL adjacent query blocks over the same corpus tiles;
// every CTA fetches 1/CL of its corpus half-tile and TMA-multicasts it to the CTAs holding that
// half in the other pairs (L2->SM corpus traffic / CL).  Such clusters fit on only part of the SMs
// (GPC boundaries), so launch_dense_tc gives the remaining SMs a concurrent CL=1 launch over the
// last slices of the corpus.
// MODE_STORE is the threshold-seeding variant: instead of selecting, it writes the maximum of every
// 16 consecutive corpus rows (see merge.cu: seed_select_kernel).
#include <cuda.h>
#include <cstdlib>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include "common.cuh"

// experiment counters (appends, prunes, publications, first-tile waits) are compiled in only with
// -DRAGARC_TC_STATS_BUILD; benchmarks/tc_stats.py reads them
#ifdef RAGARC_TC_STATS_BUILD
#define RA_STAT(...) __VA_ARGS__
#else
#define RA_STAT(...)
#endif

namespace ragarc {
static unsigned long long* g_tc_stats = nullptr;   // RAGARC_TC_STATS=1 counters (experiments)
namespace tc {

constexpr int BM = 128;            // query rows per CTA (= TMEM lanes)
constexpr int BN = 256;            // corpus rows per tile (= TMEM columns per accumulator)
constexpr int BK = 64;             // k elements per stage (= one 128-byte swizzle span)
// warp0 TMA, warp1 MMA/TMEM alloc, then one or two SETS of four epilogue warps (one warp per TMEM
// lane quarter).  With two sets, set e drains accumulator buffer e, i.e. every other corpus tile,
// into its own candidate lists: two epilogue warps per SM sub-partition hide each other's latencies
// (a single warp per sub-partition issued only ~20 % of the time and set the pace of the kernel).
#ifndef RAGARC_TC_MAX_SETS
#define RAGARC_TC_MAX_SETS 1      // build with -DRAGARC_TC_MAX_SETS=2 to allow RAGARC_TC_SETS=2 (measured: no gain)
#endif
constexpr int MAX_SETS = RAGARC_TC_MAX_SETS;
constexpr int MAX_THREADS = 64 + 128 * MAX_SETS;
constexpr int TMEM_COLS = 512;
constexpr int A_BYTES = BM * BK * 2;                       // 16 KB
// per epilogue warp 4 KB: staging of one 32x32 chunk; the first 1 KB doubles as the prune histogram
// (a prune only runs after the chunk's survivors have been fetched)
constexpr int MISC_BYTES = 256 /*barriers*/ + MAX_SETS * 4 * 32 * 32 * 4;

template <int CG> struct Cfg {
  static constexpr int BN_CTA = BN / CG;                   // corpus rows this CTA loads per tile
  static constexpr int B_BYTES = BN_CTA * BK * 2;          // 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 48 KB / 32 KB
  static constexpr int STAGES = CG == 1 ? 4 : 6;           // 192 KB either way
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + MISC_BYTES;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
// Spin on an mbarrier phase.  The bound is wall-clock time, not a spin count: a stall of more than
// ~20 s (checked every 2^16 polls against %globaltimer) traps, so that a protocol bug surfaces as an
// error on the host instead of a hung GPU, while a debugger stop, an MPS time slice or a preemption -
// which can hold a CTA for far longer than any spin count allows - do not kill the context.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  unsigned long long t_first = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if ((spin & 0xFFFFu) == 0xFFFFu) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_first == 0) t_first = now;
      else if (now - t_first > 20000000000ull) __trap();
    }
  }
}
// TMA tile load.  CG=2: the completion bytes are signalled on the LEADER CTA's barrier (the
// barrier address with the peer bit cleared, as CUTLASS's SM100_TMA_2SM_LOAD does).
template <int CG>
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
  }
}
// CG=2 multicast variant: the tile lands at the same shared-memory offset in every CTA of `mask`
// and the bytes are signalled on the leader barrier of each destination's own pair.
__device__ __forceinline__ void tma_load_2d_mc2(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
// TMA prefetch of a tile into L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// MMA completion -> mbarrier.  CG=2: arrive on the barrier (same offset) of every CTA in `mask`.
template <int CG>
__device__ __forceinline__ void tc_commit(uint64_t* bar, uint16_t mask) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// K-major operand tile in 128B-swizzled shared memory (rows of 64 bf16 = 128 B, 8-row groups
// 1024 B apart): start>>4 | LBO(ignored for swizzled K-major)=1 | SBO=1024>>4 | version=1 | SW128.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) |
         (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Pass A of a worker's first tile (see the epilogue): an order statistic of this thread's accumulator
// row that at least m of its scores reach.  Exact top-m selection over 256 values costs a divergent
// insertion per element (measured: 20 us per warp); the maxima of the sixteen 16-column groups are
// scores of sixteen distinct rows, so the m-th largest of THEM is just as valid a rung and needs
// sixteen insertions.  Kept out of line: it runs once per warp and must not cost the steady-state
// loop registers.
struct Top8 { float t[PUB_MAX_M]; };
static __device__ __noinline__ Top8 first_tile_top8(uint32_t taddr, int nvalid) {
  Top8 r;
#pragma unroll
  for (int j = 0; j < PUB_MAX_M; ++j) r.t[j] = -INFINITY;
#pragma unroll 1
  for (int c0 = 0; c0 < 256; c0 += 32) {
    if (c0 >= nvalid) break;
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
    float g0 = -INFINITY, g1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      g0 = fmaxf(g0, (c0 + j < nvalid) ? __uint_as_float(v[j]) : -INFINITY);
      g1 = fmaxf(g1, (c0 + 16 + j < nvalid) ? __uint_as_float(v[16 + j]) : -INFINITY);
    }
    if (g0 > r.t[PUB_MAX_M - 1]) top8_insert(r.t, g0);
    if (g1 > r.t[PUB_MAX_M - 1]) top8_insert(r.t, g1);
  }
  return r;
}

struct Params {
  int64_t n;       // corpus rows
  int nq, k, num_kb, MB, S;
  int64_t tiles;   // corpus tiles of THIS launch, split into S slices, starting at tile_base
  int64_t tile_base;
  int slice_base;  // candidate lists are indexed by (slice_base + slice, query block, row, epilogue set)
  int sets;        // epilogue warp sets per CTA (1 or 2) = candidate lists per (item, row)
  int cap, keep;
  uint32_t idesc;
  uint64_t* lists;
  int* counts;
  uint32_t* gthr;
  int kb_per_plane;    // >0: bf16x3 split operands, k blocks per plane (see launch_dense_tc)
  int prefetch;        // RAGARC_TC_PREFETCH: L2 prefetch distance in tiles (0 = off; measured 2 % slower when on)
  float* seed_out;     // MODE_STORE: [nq, seed_ld] maxima of 16-row groups of rows [0, n)
  int seed_ld;
  // published order statistics (common.cuh): pub[query][PUB_LD]; lists of slices < pub_n READ the
  // minimum of the first pub_n entries as a threshold; lists with list slice < pub_publish also WRITE
  // their pub_m-th best score (pub_publish = pub_n in the main launch, 0 in the left-over launch)
  uint32_t* pub;
  int pub_n, pub_m, pub_publish;
  int pub_wait_cycles; // a worker's first item waits at most this long for the first-tile rungs of all lists
  int pub_dbg;         // RAGARC_TC_PUB_DBG (experiments): 1 = no publication after the first tile, 2 = no first-tile exchange, 4 = no rung tracking in the append loop
  unsigned long long* stats;   // RAGARC_TC_STATS=1 (experiments): appends, prunes, rung publications, first-tile
                               // wait cycles (sum over warps), first-tile waits that timed out, warps that waited
};

constexpr int MODE_TOPK = 0, MODE_STORE = 1;

// CL = CTA pairs per cluster (CG=2 only).  With CL > 1 the pairs of a cluster work on the same
// corpus tiles for CL adjacent query blocks: every CTA loads 1/CL of its corpus half-tile and
// multicasts it to the CTAs holding the same half in the other pairs, which divides the corpus
// operand's L2->SM traffic by CL (the kernel is bound by L2 bandwidth, not by the tensor pipe).
template <int MODE, int CG, int CL>
__global__ void __launch_bounds__(MAX_THREADS, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x,
                const Params p) {
  using C = Cfg<CG>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* tiles_smem = smem;                                    // STAGES x (A | B)
  uint64_t* full_bar = (uint64_t*)(smem + STAGES * STAGE_BYTES);  // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                       // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                      // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                          // [2]
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 2);
  float* stage_all = (float*)(smem + STAGES * STAGE_BYTES + 256);   // [epilogue warps][32][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  static_assert(CL == 1 || CG == 2, "multicast clusters are built from CTA pairs");
  constexpr int CSIZE = CG * CL;                                 // CTAs per cluster
  const uint32_t crank = CSIZE == 1 ? 0u : cluster_ctarank();
  const uint32_t rank = crank & (CG - 1);                        // 0 = leader CTA of the pair
  const uint32_t pair = crank / CG;                              // pair inside the cluster
  const int unit = blockIdx.x / CSIZE;                           // persistent worker (cluster) id
  const int nunits = gridDim.x / CSIZE;
  constexpr int ROWS_ITEM = BM * CG;                             // query rows per work item
  const int MBq = p.MB / CL;                                     // query-block groups (CL blocks each)
  const uint16_t mask_all = (uint16_t)((1u << CSIZE) - 1);       // every CTA of the cluster
  const uint16_t mask_pair = (uint16_t)(3u << (crank & ~1u));    // this CTA's pair

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 1) __syncthreads(); else cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch: once every CTA of this grid is resident and has got here, a kernel
  // launched behind it with programmatic stream serialisation may start.  That is how the left-over
  // pairs' launch is ordered AFTER the placement of the multicast clusters (if its CTAs were placed
  // first they could sit on SMs a 4-CTA cluster needs, and those clusters would only start when the
  // left-over launch is done: measured +0.7 ms per search whenever the stream was busy at launch time).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int64_t items = (int64_t)MBq * p.S;       // cluster work items: (query-block group, slice)

  if (warp == 0) {
    // ------------------------------ TMA producer (every CTA) ------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      constexpr int B_ROWS = C::BN_CTA / CL;       // corpus rows this CTA fetches per tile
      const uint16_t mask_half = (uint16_t)((CL == 1 ? 0x1u : CL == 2 ? 0x5u : 0x55u) << rank);
      for (int64_t item = unit; item < items; item += nunits) {
        const int qb = (int)(item % MBq) * CL + (int)pair;
        const int64_t s = item / MBq;
        const int64_t t0 = p.tile_base + s * p.tiles / p.S, t1 = p.tile_base + (s + 1) * p.tiles / p.S;
        const int q0 = qb * ROWS_ITEM + (int)rank * BM;
        for (int64_t t = t0; t < t1; ++t) {
          const int x0 = (int)(t * BN) + (int)rank * C::BN_CTA + (int)pair * B_ROWS;
          // corpus rows one tile ahead are pulled into L2 now, so that the ring's refills (which
          // have ~5 stage-times to land) see L2 latency instead of HBM latency
          const bool pf = p.prefetch > 0 && (t + p.prefetch < t1);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            if (pf) tma_prefetch_l2_2d(&tmap_x, kb * BK, x0 + p.prefetch * BN);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* a = tiles_smem + stage * STAGE_BYTES;
            if (rank == 0) mbar_expect_tx(&full_bar[stage], STAGE_BYTES * CG);   // both CTAs' bytes
            int ka = kb * BK, kx = kb * BK;
            if (p.kb_per_plane) {
              // bf16x3: k block kb belongs to plane pair (query plane, corpus plane), smallest terms first:
              // (1,1) (0,2) (2,0) (0,1) (1,0) (0,0)
              const int pair = kb / p.kb_per_plane, kk = kb - pair * p.kb_per_plane;
              ka = (((0x010201 >> (4 * pair)) & 0xF) * p.kb_per_plane + kk) * BK;
              kx = (((0x001021 >> (4 * pair)) & 0xF) * p.kb_per_plane + kk) * BK;
            }
            tma_load_2d<CG>(a, &tmap_q, &full_bar[stage], ka, q0);
            if (CL == 1) tma_load_2d<CG>(a + A_BYTES, &tmap_x, &full_bar[stage], kx, x0);
            else tma_load_2d_mc2(a + A_BYTES + pair * (B_ROWS * BK * 2), &tmap_x, &full_bar[stage], kx, x0, mask_half);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) --------------
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t tcount = 0;
      for (int64_t item = unit; item < items; item += nunits) {
        const int64_t s = item / MBq;
        const int64_t t0 = p.tile_base + s * p.tiles / p.S, t1 = p.tile_base + (s + 1) * p.tiles / p.S;
        for (int64_t t = t0; t < t1; ++t, ++tcount) {
          const uint32_t buf = tcount & 1, aphase = (tcount >> 1) & 1;
          mbar_wait(&tempty_bar[buf], aphase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + buf * BN;
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a = smem_u32(tiles_smem + stage * STAGE_BYTES);
            const uint64_t adesc = make_smem_desc(a);
            const uint64_t bdesc = make_smem_desc(a + A_BYTES);
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the >>4 address
              tc_mma<CG>(tmem_d, adesc + (uint64_t)(k4 * 2), bdesc + (uint64_t)(k4 * 2), p.idesc,
                         (uint32_t)((kb | k4) != 0));
            }
            tc_commit<CG>(&empty_bar[stage], mask_all);   // frees the smem stage (in every CTA that may write it) when the MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          tc_commit<CG>(&tfull_bar[buf], mask_pair);    // accumulator complete (signalled in both CTAs of the pair)
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..5, every CTA) ----------
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;          // query row inside the CTA == TMEM lane
    const uint32_t set = (uint32_t)(warp - 2) >> 2;   // this warp drains accumulator buffer `set` when there are two sets
    const uint32_t two_sets = p.sets > 1 ? 1u : 0u;
    uint32_t* myhist = (uint32_t*)(stage_all + (warp - 2) * 1024);
    // per-warp staging of one 32x32 chunk: thread `lane` owns row `lane` (128 B); 16-byte chunks are
    // XOR-swizzled by (lane & 7) so that the 128-bit stores of a warp spread over all banks
    float* mystage = stage_all + (warp - 2) * 1024 + lane * 32;
    const int swz = lane & 7;
    uint32_t tcount = 0;
    bool first_tile_done = false;                 // the first-tile rung exchange happens once per warp
    RA_STAT(unsigned long long n_app = 0, n_pub = 0, n_prune = 0;)
    RA_STAT(unsigned long long* tl = p.stats ? p.stats + 16 + ((p.slice_base ? 160 : 0) + blockIdx.x) * 4 : nullptr;)
    RA_STAT(if (tl && warp == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); tl[0] = g; })
    for (int64_t citem = unit; citem < items; citem += nunits) {
      const int qb = (int)(citem % MBq) * CL + (int)pair;
      const int64_t s = citem / MBq;
      const int64_t item = (p.slice_base + s) * p.MB + qb;   // (slice, query block): index of the candidate lists
      const int64_t t0 = p.tile_base + s * p.tiles / p.S, t1 = p.tile_base + (s + 1) * p.tiles / p.S;
      const int irow = (int)rank * BM + row;       // row inside the work item
      const int qrow = qb * ROWS_ITEM + irow;
      RowState st;
      const size_t list_id = ((size_t)item * ROWS_ITEM + irow) * p.sets + set;
      st.list = p.lists + list_id * (size_t)p.cap;
      st.cnt = 0;
      st.ord_local = 0;
      st.ord_global = 0;
      st.thr = (qrow < p.nq) ? -INFINITY : INFINITY;
      uint32_t* grow = (MODE == MODE_TOPK && qrow < p.nq) ? p.gthr + qrow : nullptr;
      // published rungs: this row's slot (when its list is one of the publishing ones) and its row of
      // everybody's rungs
      const uint32_t* pubrow = (MODE == MODE_TOPK && p.pub_n > 0 && grow) ? p.pub + (size_t)qrow * PUB_LD : nullptr;
      const int64_t pub_id = (p.slice_base + s) * p.sets + set;                        // this list among the publishing ones
      const bool publishes = MODE == MODE_TOPK && pub_id < p.pub_publish;              // warp-uniform
      uint32_t* pubslot = (pubrow && publishes) ? p.pub + (size_t)qrow * PUB_LD + pub_id : nullptr;
      float top[PUB_MAX_M];
#pragma unroll
      for (int j = 0; j < PUB_MAX_M; ++j) top[j] = -INFINITY;
      float published = -INFINITY;
      for (int64_t t = t0; t < t1; ++t, ++tcount) {
        const uint32_t buf = tcount & 1, aphase = (tcount >> 1) & 1;
        if (two_sets && buf != set) continue;    // the other set's tile
        if (MODE == MODE_TOPK && grow) {
          uint32_t g = *(volatile uint32_t*)grow;
          if (g > st.ord_global) { st.ord_global = g; st.thr = combine_thr(st.ord_local, g); }
        }
        const int64_t r0 = t * BN;
        const int nvalid = (int)((p.n - r0) < BN ? (p.n - r0) : BN);
        mbar_wait(&tfull_bar[buf], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + buf * BN;
        if (MODE == MODE_TOPK && p.pub_n > 0 && citem == unit && !first_tile_done && !(p.pub_dbg & 2)) {
          first_tile_done = true;
          // First tile (of this epilogue set) of the worker's first item: nothing is known about the score distribution
          // yet.  Pass A reads the accumulator once only to find this row's pub_m best scores and
          // publishes the last of them; then every row waits (bounded: all first-wave workers reach
          // this point within a few microseconds of each other) until the rungs of all publishing
          // lists of its query are in, and pass B - the normal loop below, over the same
          // accumulator - appends only what beats their minimum instead of all 256 rows.
          if (publishes) {
            const Top8 a = first_tile_top8(taddr, nvalid);
            const float pv = top8_get(a.t, p.pub_m - 1);     // pass B re-inserts its survivors into `top`
            if (pubslot && pv > published) {
              published = pv;
              RA_STAT(++n_pub;)
              __stcg(pubslot, f32_to_ord(pv));
            }
          }
          RA_STAT(if (tl && warp == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); tl[1] = g; })
          const long long tstart = clock64();
          uint32_t pg = 0;
          bool pending = pubrow != nullptr;
          for (;;) {                                 // both exits are warp-uniform
            if (pending) { pg = pub_min(pubrow, p.pub_n); pending = pg == 0; }
            const bool expired = clock64() - tstart > (long long)p.pub_wait_cycles;
            if (!__any_sync(FULL, pending) || __any_sync(FULL, expired)) break;
            __nanosleep(100);
          }
          if (pg > st.ord_global) {
            st.ord_global = pg; st.thr = combine_thr(st.ord_local, pg);
            if (pubslot) atomicMax(grow, pg);        // for the lists that start later and only poll the shared threshold
          }
          RA_STAT(if (tl && warp == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); tl[2] = g; })
          RA_STAT(if (p.stats && lane == 0) {
            atomicAdd(p.stats + 3, (unsigned long long)(clock64() - tstart));
            if (pending) atomicAdd(p.stats + 4, 1ull);
            atomicAdd(p.stats + 5, 1ull);
          })
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (c0 >= nvalid) break;                // warp-uniform
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_ld_wait();
          if (MODE == MODE_STORE) {
            // seed pass: keep only the maximum of every 16 consecutive rows ("group maxima")
            float g0 = -INFINITY, g1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              g0 = fmaxf(g0, (c0 + j < nvalid) ? __uint_as_float(v[j]) : -INFINITY);
              g1 = fmaxf(g1, (c0 + 16 + j < nvalid) ? __uint_as_float(v[16 + j]) : -INFINITY);
            }
            if (qrow < p.nq)
              *reinterpret_cast<float2*>(p.seed_out + (size_t)qrow * p.seed_ld + ((r0 + c0) >> 4)) = make_float2(g0, g1);
            continue;
          }
          // (1) park the chunk in shared memory so that survivors can be fetched by column index
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            *reinterpret_cast<uint4*>(mystage + ((c4 ^ swz) << 2)) =
                make_uint4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
          // (2) branch-free pass mask: bit (31-j) = sign(thr - v[j]) = (v[j] > thr); two chains
          uint32_t ma = 0, mb = 0;
          const float thr = st.thr;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            ma = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v[j])), ma, 1);
            mb = __funnelshift_l(__float_as_uint(thr - __uint_as_float(v[j + 16])), mb, 1);
          }
          uint32_t m = (ma << 16) | (mb & 0xFFFFu);
          if (c0 + 32 > nvalid) m &= ~(0xFFFFFFFFu >> (nvalid - c0));   // drop padding columns
          // (3) append the survivors in ascending column order
          while (m) {
            const int j = __clz(m);
            m &= ~(0x80000000u >> j);
            const float f = mystage[(((j >> 2) ^ swz) << 2) | (j & 3)];
            st.list[st.cnt++] = make_key(f, (uint32_t)(r0 + c0 + j));
            RA_STAT(++n_app;)
            if (publishes && !(p.pub_dbg & 4) && f > top[PUB_MAX_M - 1]) top8_insert(top, f);
          }
          RA_STAT(const int before = st.cnt;)
          prune_if_needed(st, p.k, p.cap, p.cap - 32, grow, myhist);
          RA_STAT(if (st.cnt < before) ++n_prune;)
        }
        // Publication schedule: after the 1st, 2nd, 4th, 8th ... tile of the item and after its last
        // one (the bound ~ m / rows seen, so doubling intervals lose at most a factor of two; the
        // schedule is warp-uniform, so a warp goes through the sequence a handful of times per item
        // instead of whenever one of its 32 lists moved).
        const int64_t ti = t - t0 + 1;
        if (MODE == MODE_TOPK && publishes && !(p.pub_dbg & 1) && ((ti & (ti - 1)) == 0 || t + 1 + (two_sets ? 1 : 0) >= t1)) {
          const float pv = top8_get(top, p.pub_m - 1);
          if (pubslot) {
            // store this list's rung if it moved, then fold the minimum over all publishing lists of
            // the query (a score at least pub_n * pub_m >= k rows reach) into the shared threshold
            if (pv > published) { published = pv; RA_STAT(++n_pub;) __stcg(pubslot, f32_to_ord(pv)); }
            const uint32_t g = pub_min(pubrow, p.pub_n);
            if (g > st.ord_global) {
              atomicMax(grow, g);
              st.ord_global = g; st.thr = combine_thr(st.ord_local, g);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          // the accumulator is drained: tell the MMA issuer (in the leader CTA)
          if (rank == 0) mbar_arrive(&tempty_bar[buf]); else mbar_arrive_remote(&tempty_bar[buf], crank & ~1u);
        }
      }
      if (MODE == MODE_TOPK) {
        // A list longer than the merge kernel's per-list budget is first filtered against the best
        // bound known by now (every thread compacts its own list, no cooperation needed: ~1 us for a
        // whole warp); only what is still too long afterwards pays for an exact radix prune, which
        // serialises the lanes of a warp at 3-4 us each.
        if (grow && st.cnt > p.keep) {
          uint32_t g = *(volatile uint32_t*)grow;
          g = g > st.ord_global ? g : st.ord_global;
          if (g > st.ord_local) {
            // sixteen loads in flight per pass: one key at a time is a dependent L2 round trip each (the
            // compiler may not hoist loads over the compacting stores) and cost ~50-100 us per item
            int w = 0;
            for (int i = 0; i < st.cnt; i += 16) {
              uint64_t kk[16];
#pragma unroll
              for (int u = 0; u < 16; ++u) kk[u] = (i + u < st.cnt) ? st.list[i + u] : 0ull;
#pragma unroll
              for (int u = 0; u < 16; ++u)
                if (uint32_t(kk[u] >> 32) >= g && kk[u] != 0ull) st.list[w++] = kk[u];
            }
            st.cnt = w;
          }
        }
        prune_if_needed(st, p.k, p.cap, p.keep, grow, myhist);
        p.counts[list_id] = st.cnt;
        RA_STAT(if (p.stats) {
          atomicAdd(p.stats + 0, n_app); atomicAdd(p.stats + 1, n_prune); atomicAdd(p.stats + 2, n_pub);
          n_app = 0; n_pub = 0; n_prune = 0;
        })
      }
    }
  }

  RA_STAT(if (p.stats && warp == 2 && lane == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); p.stats[16 + ((p.slice_base ? 160 : 0) + blockIdx.x) * 4 + 3] = g; })
  tc_fence_before();
  if (CG == 1) __syncthreads(); else cluster_sync_all();   // cluster: nobody leaves while a peer still uses us
  if (warp == 1) {
    tc_fence_after();
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p)
    return nullptr;
  fn = (EncodeTiledFn)p;
  return fn;
}

static int make_map(CUtensorMap* map, const void* base, int64_t rows, int d, int dtype, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return RAGARC_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == RAGARC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return RAGARC_ERR_CUDA; }
  return RAGARC_OK;
}

// how many clusters of CG*CL CTAs the device keeps resident at once (GPC boundaries decide)
template <int MODE, int CG, int CL>
static int max_clusters() {
  static int cached = 0;
  if (cached) return cached;
  using C = Cfg<CG>;
  int n = sm_count() / (CG * CL);
  if (CG * CL > 1 &&
      cudaFuncSetAttribute(dense_tc_kernel<MODE, CG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sm_count() / (CG * CL) * (CG * CL)));
    cfg.blockDim = dim3(MAX_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG * CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int got = 0;
    if (cudaOccupancyMaxActiveClusters(&got, dense_tc_kernel<MODE, CG, CL>, &cfg) == cudaSuccess && got > 0 && got < n) n = got;
    (void)cudaGetLastError();
  }
  cached = n > 0 ? n : 1;
  return cached;
}

template <int MODE, int CG, int CL>
static int launch_one(const CUtensorMap& mq, const CUtensorMap& mx, const Params& p, cudaStream_t stream,
                      bool programmatic = false) {
  using C = Cfg<CG>;
  RA_CUDA(cudaFuncSetAttribute(dense_tc_kernel<MODE, CG, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  RA_REQUIRE(p.MB % CL == 0, "dense tcgen05: query blocks not divisible by the cluster's pair count");
  const int64_t items = (int64_t)(p.MB / CL) * p.S;
  const int max_units = max_clusters<MODE, CG, CL>();
  const int units = (int)(items < max_units ? items : max_units);
  constexpr int CG_ = CG * CL;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(units * CG_));
  cfg.blockDim = dim3(64 + 128 * (p.sets > 1 ? 2 : 1));
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG_;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (programmatic) {      // may start as soon as the previous kernel in the stream has issued launch_dependents everywhere
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  RA_CUDA(cudaLaunchKernelEx(&cfg, dense_tc_kernel<MODE, CG, CL>, mq, mx, p));
  count_launch();
  return RAGARC_OK;
}

// Side stream (+ fork/join events) for the launch that runs concurrently with the multicast
// clusters: one per (device, caller stream), created on first use and kept - so that a capture in
// progress on one stream never shares its side stream with eager work on another.
struct Side { cudaStream_t stream; cudaEvent_t fork, join; std::mutex mu; };
static Side* get_side(cudaStream_t user) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, Side*> sides;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = sides.find({dev, user});
  if (it != sides.end()) return it->second;
  Side* s = new (std::nothrow) Side();
  if (!s) return nullptr;
  if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->join, cudaEventDisableTiming) != cudaSuccess) {
    delete s;
    return nullptr;
  }
  sides[{dev, user}] = s;
  return s;
}

}  // namespace tc

int dense_tc_max_sets() { return tc::MAX_SETS; }

// experiments only (not part of the public header): read and reset the RAGARC_TC_STATS counters
extern "C" int ragarc_internal_tc_stats(unsigned long long* out2048_host) {
  if (!g_tc_stats) { for (int i = 0; i < 2048; ++i) out2048_host[i] = 0; return RAGARC_OK; }
  RA_CUDA(cudaDeviceSynchronize());
  RA_CUDA(cudaMemcpy(out2048_host, g_tc_stats, 2048 * 8, cudaMemcpyDeviceToHost));
  RA_CUDA(cudaMemset(g_tc_stats, 0, 2048 * 8));
  return RAGARC_OK;
}

int dense_tc_units(int cg, int cl) {
  using namespace tc;
  if (cg == 1) return max_clusters<MODE_TOPK, 1, 1>();
  if (cl == 4) return max_clusters<MODE_TOPK, 2, 4>() * 4;
  if (cl == 2) return max_clusters<MODE_TOPK, 2, 2>() * 2;
  return max_clusters<MODE_TOPK, 2, 1>();
}

bool dense_tc_supported(const void* corpus, int64_t n, int d, int dtype, const void* queries) {
  if (dtype != RAGARC_BF16 && dtype != RAGARC_F16) return false;
  if (d % 8 != 0 || n <= 0 || n >= (int64_t)0x7FFFFF00ll) return false;
  if (((uintptr_t)corpus & 15) || ((uintptr_t)queries & 15)) return false;
  return true;
}

int launch_dense_tc(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                    int k, const DensePlan& pl, uint64_t* lists, int* counts, uint32_t* gthr, uint32_t* pub,
                    float* seed_scores, void* qpad, cudaEvent_t after_seed, cudaStream_t stream) {
  using namespace tc;
  RA_REQUIRE(dense_tc_supported(corpus, n, d, dtype, queries),
             "dense tcgen05: needs bf16/fp16, d %% 8 == 0 and 16-byte aligned base pointers");
  const int cg = pl.rows_per_item / BM;          // 1 or 2 (chosen by the planner)
  const int cl = pl.cl;                          // pairs per cluster: 1, 2 or 4 (cg == 2 only)
  RA_REQUIRE((cg == 1 && cl == 1) || (cg == 2 && (cl == 1 || cl == 2 || cl == 4) && pl.MB % cl == 0),
             "dense tcgen05: bad plan");
  // Queries are staged into a buffer padded with zero rows up to a whole number of work-item rows,
  // so that every query-tile TMA load is fully in bounds (out-of-bounds fill was measured slower for
  // tiny batches: 1 query in a 128-row box).
  const int64_t q_rows = (int64_t)pl.MB * pl.rows_per_item;
  const void* qsrc = queries;
  if (q_rows > nq) {
    const size_t q_bytes = (size_t)nq * d * 2;
    RA_CUDA(cudaMemcpyAsync(qpad, queries, q_bytes, cudaMemcpyDeviceToDevice, stream));
    RA_CUDA(cudaMemsetAsync((char*)qpad + q_bytes, 0, (size_t)(q_rows - nq) * d * 2, stream));
    qsrc = qpad;
  }
  CUtensorMap mq, mx;
  int rc = make_map(&mq, qsrc, q_rows, d, dtype, BM);
  if (rc) return rc;
  rc = make_map(&mx, corpus, n, d, dtype, BN / cg / cl);
  if (rc) return rc;
  Params p;
  p.n = n; p.nq = nq; p.k = k; p.MB = pl.MB; p.S = pl.S;
  // bf16x3: an fp32 vector v is stored as three bf16 planes v1+v2+v3 (width d = 3*x3_d); the dot
  // product is the sum of the six largest plane-pair products, i.e. six passes over x3_d
  p.kb_per_plane = pl.x3_d > 0 ? pl.x3_d / BK : 0;
  p.num_kb = pl.x3_d > 0 ? 6 * p.kb_per_plane : (d + BK - 1) / BK;
  p.tiles = pl.tiles; p.tile_base = 0; p.slice_base = 0; p.cap = pl.cap; p.keep = pl.keep; p.lists = lists; p.counts = counts; p.gthr = gthr;
  p.seed_out = nullptr; p.seed_ld = 0;
  p.sets = pl.sets;
  p.pub = pub; p.pub_n = pl.pub_n; p.pub_m = pl.pub_m; p.pub_publish = pl.pub_n;
  {
    static const char* envd = getenv("RAGARC_TC_PUB_DBG");
    p.pub_dbg = envd ? atoi(envd) : 0;
  }
  p.stats = nullptr;
  {
    static const char* envst = getenv("RAGARC_TC_STATS");
    if (envst && envst[0] == '1') {
      if (!g_tc_stats) { RA_CUDA(cudaMalloc(&g_tc_stats, 2048 * sizeof(unsigned long long))); RA_CUDA(cudaMemset(g_tc_stats, 0, 2048 * 8)); }
      p.stats = g_tc_stats;
    }
  }
  {
    static const char* envw = getenv("RAGARC_TC_PUB_WAIT");      // experiments: first-tile wait bound in cycles
    p.pub_wait_cycles = envw ? atoi(envw) : 40000;
  }
  {
    static const char* env = getenv("RAGARC_TC_PREFETCH");
    p.prefetch = env ? atoi(env) : 0;
  }
  const uint32_t fmt = dtype == RAGARC_BF16 ? 1u : 0u;
  // instruction descriptor (kind::f16): D=f32 [4,6), A fmt [7,10), B fmt [10,13), A/B K-major,
  // N>>3 at [17,23), M>>4 at [24,29); M is 128 per CTA, i.e. 256 for a cta_group::2 pair
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t((BM * cg) >> 4) << 24);
  if (pl.seed_rows > 0) {
    // seed pass over the first seed_rows rows -> group maxima -> k-th largest per query -> gthr
    Params ps = p;
    ps.pub_n = 0; ps.pub_publish = 0;
    CUtensorMap mxs;
    rc = make_map(&mxs, corpus, pl.seed_rows, d, dtype, BN / cg);
    if (rc) return rc;
    ps.n = pl.seed_rows;
    ps.tiles = (pl.seed_rows + BN - 1) / BN;
    ps.S = pl.seed_S;
    ps.seed_out = seed_scores;
    ps.seed_ld = pl.seed_rows / 16;
    rc = cg == 1 ? launch_one<MODE_STORE, 1, 1>(mq, mxs, ps, stream) : launch_one<MODE_STORE, 2, 1>(mq, mxs, ps, stream);
    if (rc) return rc;
    rc = launch_seed_select(seed_scores, nq, pl.seed_rows / 16, k, gthr, stream);
    if (rc) return rc;
  }
  if (after_seed) RA_CUDA(cudaEventRecord(after_seed, stream));
  if (cg == 1) return launch_one<MODE_TOPK, 1, 1>(mq, mx, p, stream);
  if (cl == 1) return launch_one<MODE_TOPK, 2, 1>(mq, mx, p, stream);
  // Multicast clusters on the caller's stream and - concurrently, on the SMs no cluster fits on -
  // plain pairs on a side stream (forked and joined with events, so the call stays stream-ordered
  // and capturable).  Programmatic dependent launch in one stream was tried instead and did not
  // overlap the two launches (measured: they ran back to back).
  p.S = pl.S - pl.S_tail;
  p.tiles = pl.tiles_main;
  static const char* envpdl = getenv("RAGARC_TC_PDL");       // 0: left-over pairs on a side stream (fork/join events)
  const bool pdl = !(envpdl && envpdl[0] == '0');
  if (pdl && pl.S_tail > 0) {
    // both launches in the caller's stream: the left-over pairs as a programmatic dependent of the clusters
    rc = cl == 4 ? launch_one<MODE_TOPK, 2, 4>(mq, mx, p, stream) : launch_one<MODE_TOPK, 2, 2>(mq, mx, p, stream);
    if (rc) return rc;
    CUtensorMap mxt;
    rc = make_map(&mxt, corpus, n, d, dtype, BN / cg);
    if (rc) return rc;
    Params pt = p;
    pt.S = pl.S_tail;
    pt.tiles = pl.tiles - pl.tiles_main;
    pt.tile_base = pl.tiles_main;
    pt.slice_base = pl.S - pl.S_tail;
    pt.pub_publish = 0;                          // the left-over pairs read the rungs but publish none
    return launch_one<MODE_TOPK, 2, 1>(mq, mxt, pt, stream, true);
  }
  Side* side = nullptr;
  std::unique_lock<std::mutex> side_lock;          // the fork/join sequence on the shared side stream is atomic
  if (pl.S_tail > 0) {
    side = get_side(stream);
    RA_REQUIRE(side != nullptr, "dense tcgen05: cannot create the side stream");
    side_lock = std::unique_lock<std::mutex>(side->mu);
    RA_CUDA(cudaEventRecord(side->fork, stream));
    RA_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  }
  rc = cl == 4 ? launch_one<MODE_TOPK, 2, 4>(mq, mx, p, stream) : launch_one<MODE_TOPK, 2, 2>(mq, mx, p, stream);
  if (rc) return rc;
  if (side) {
    CUtensorMap mxt;
    rc = make_map(&mxt, corpus, n, d, dtype, BN / cg);
    if (rc) return rc;
    Params pt = p;
    pt.S = pl.S_tail;
    pt.tiles = pl.tiles - pl.tiles_main;
    pt.tile_base = pl.tiles_main;
    pt.slice_base = pl.S - pl.S_tail;
    pt.pub_publish = 0;                          // the left-over pairs read the rungs but publish none
    rc = launch_one<MODE_TOPK, 2, 1>(mq, mxt, pt, side->stream);
    if (rc) return rc;
    RA_CUDA(cudaEventRecord(side->join, side->stream));
    RA_CUDA(cudaStreamWaitEvent(stream, side->join, 0));
  }
  return RAGARC_OK;
}

}  // namespace ragarc
