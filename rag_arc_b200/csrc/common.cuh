// Shared device/host helpers for libragarc_b200: error plumbing, sortable keys, and the
// per-query candidate-list machinery (append / warp-cooperative prune) that both dense scoring
// kernels use so that the nq x n score matrix never reaches HBM.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ragarc_b200.h"

namespace ragarc {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define RA_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::ragarc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                   \
                          cudaGetErrorString(_e));                                        \
      return RAGARC_ERR_CUDA;                                                             \
    }                                                                                     \
  } while (0)

#define RA_REQUIRE(cond, ...)                                                             \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::ragarc::set_error(__VA_ARGS__);                                                   \
      return RAGARC_ERR_INVALID;                                                          \
    }                                                                                     \
  } while (0)

#define RA_LAUNCH_CHECK()                                                                 \
  do {                                                                                    \
    ::ragarc::count_launch();                                                             \
    RA_CUDA(cudaGetLastError());                                                          \
  } while (0)

constexpr unsigned FULL = 0xFFFFFFFFu;

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- sortable keys -------------------------------------------------------------------------
// fp32 -> uint32 that sorts like the float (NaN excluded); -0 is canonicalised to +0.
// ord 0 is never produced by a real score and doubles as "unset".
__host__ __device__ __forceinline__ uint32_t f32_to_ord(float f) {
  f += 0.0f;
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
constexpr uint32_t ORD_NEG_INF = 0x007FFFFFu;

// key: larger = better (higher score, then LOWER row id).  0 = empty slot.
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row) {
  return (uint64_t(f32_to_ord(score)) << 32) | uint64_t(0xFFFFFFFFu - row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xFFFFFFFFu - uint32_t(key); }
__device__ __forceinline__ float key_score(uint64_t key) { return ord_to_f32(uint32_t(key >> 32)); }

// Strict float threshold equivalent to "ord(v) > T" (T==0 / below -inf: everything finite passes).
__device__ __forceinline__ float thr_from_ord(uint32_t T) {
  return (T <= ORD_NEG_INF) ? -INFINITY : ord_to_f32(T);
}
// Threshold for a query given its CTA-local k-th best (strict: later rows with an equal score
// have higher ids and lose the tie) and the cross-CTA shared k-th best (non-strict: the rows
// behind it may have higher ids than ours).
__device__ __forceinline__ float combine_thr(uint32_t ord_local, uint32_t ord_global) {
  uint32_t g = ord_global ? ord_global - 1 : 0;
  return thr_from_ord(ord_local > g ? ord_local : g);
}

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// Histogram increment used by every radix-select pass: a plain shared-memory atomic.  (A
// __match_any_sync warp-aggregated variant was measured SLOWER on B200: merge kernel 84 -> 113 us,
// see profiles/README.md.)
__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t bin, bool valid) {
  if (valid) atomicAdd(&hist[bin], 1u);
}

// Radix passes start at the highest bit in which the keys actually differ (given the OR and AND of
// all keys): the first 8-bit digit then resolves the keys' real range instead of their shared
// leading bits.  Later digits step down by 8, the last one is clamped to shift 0.
__device__ __forceinline__ int first_varying_shift(uint64_t all_or, uint64_t all_and) {
  const uint64_t diff = all_or ^ all_and;
  if (diff == 0) return 0;
  const int top = 63 - __clzll((long long)diff);
  return top >= 7 ? top - 7 : 0;
}
__device__ __forceinline__ uint64_t high_bytes_mask(int shift) {   // bits above the digit at `shift`
  return shift + 8 >= 64 ? 0ull : ~((1ull << (shift + 8)) - 1ull);
}
__device__ __forceinline__ int next_shift(int shift) { return shift >= 8 ? shift - 8 : 0; }

// ---- warp-cooperative prune -------------------------------------------------------------------
// Keeps exactly the k largest keys of list[0..n) (n > k, keys distinct) at list[0..k) and
// returns ord of the k-th largest score.  MSB-first 8-bit radix select over the 64-bit keys with
// early exit as soon as the boundary bucket is entirely inside the top-k (typically 2-3 passes
// for float scores), then an in-place stable compaction.  All 32 lanes must call it with the
// same arguments; `hist` is 256 words of shared memory private to the warp.
static __device__ __noinline__ uint32_t warp_prune(uint64_t* __restrict__ list, int n, int k,
                                            uint32_t* hist) {
  const int lane = threadIdx.x & 31;
  uint64_t prefix = 0, mask = 0;
  uint32_t rem = (uint32_t)k;
  __syncwarp();
#pragma unroll 1
  for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane + 32 * i] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      uint64_t key = list[i];
      if ((key & mask) == prefix) atomicAdd(&hist[(uint32_t)(key >> shift) & 0xFFu], 1u);
    }
    __syncwarp();
    uint32_t h[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; sum += h[j]; }
    uint32_t incl = sum;                       // becomes: sum over lanes >= lane
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_down_sync(FULL, incl, o);
      if (lane + o < 32) incl += v;
    }
    uint32_t excl = incl - sum;                // keys in buckets above this lane's 8 buckets
    bool here = (excl < rem) && (incl >= rem);
    int src = __ffs(__ballot_sync(FULL, here)) - 1;
    uint32_t D = 0, above = 0, hD = 0;
    if (lane == src) {
      uint32_t acc = excl;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (acc + h[j] >= rem) { D = 8 * lane + j; above = acc; hD = h[j]; break; }
        acc += h[j];
      }
    }
    D = __shfl_sync(FULL, D, src);
    above = __shfl_sync(FULL, above, src);
    hD = __shfl_sync(FULL, hD, src);
    rem -= above;
    prefix |= uint64_t(D) << shift;
    mask |= uint64_t(0xFF) << shift;
    if (hD == rem) break;                      // whole boundary bucket is kept
  }
  // keep <=> (key & mask) >= prefix : exactly k keys
  int base = 0;
  uint32_t min_ord = 0xFFFFFFFFu;
  for (int r0 = 0; r0 < n; r0 += 32) {
    int i = r0 + lane;
    uint64_t key = (i < n) ? list[i] : 0ull;
    bool keep = (i < n) && ((key & mask) >= prefix);
    unsigned b = __ballot_sync(FULL, keep);
    if (keep) {
      list[base + __popc(b & ((1u << lane) - 1u))] = key;
      min_ord = min(min_ord, uint32_t(key >> 32));
    }
    base += __popc(b);
  }
  __syncwarp();
  return warp_min_u32(min_ord);
}

// Register-resident variant for lists of at most 32*R keys: every key is loaded once, up front
// (R independent loads in flight per lane instead of one L2 round trip per element and pass), the
// radix passes and the compaction then run out of registers.  Same contract as warp_prune.
template <int R>
static __device__ __noinline__ uint32_t warp_prune_reg(uint64_t* __restrict__ list, int n, int k,
                                                       uint32_t* hist) {
  const int lane = threadIdx.x & 31;
  uint64_t key[R];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int i = j * 32 + lane;
    key[j] = (i < n) ? list[i] : 0ull;       // 0 = empty: never selected (valid keys are > 0)
  }
  // skip the radix passes over leading bytes that all keys share
  uint32_t oh = 0, ol = 0, ah = 0xFFFFFFFFu, al = 0xFFFFFFFFu;
#pragma unroll
  for (int j = 0; j < R; ++j)
    if (key[j] != 0ull) {
      oh |= uint32_t(key[j] >> 32); ol |= uint32_t(key[j]);
      ah &= uint32_t(key[j] >> 32); al &= uint32_t(key[j]);
    }
  oh = __reduce_or_sync(FULL, oh); ol = __reduce_or_sync(FULL, ol);
  ah = __reduce_and_sync(FULL, ah); al = __reduce_and_sync(FULL, al);
  const int shift0 = first_varying_shift((uint64_t(oh) << 32) | ol, (uint64_t(ah) << 32) | al);
  uint64_t mask = high_bytes_mask(shift0);
  uint64_t prefix = ((uint64_t(ah) << 32) | al) & mask;
  uint32_t rem = (uint32_t)k;
#pragma unroll 1
  for (int shift = shift0;; shift = next_shift(shift)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane + 32 * i] = 0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j)
      hist_add(hist, (uint32_t)(key[j] >> shift) & 0xFFu, key[j] != 0ull && (key[j] & mask) == prefix);
    __syncwarp();
    uint32_t h[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { h[j] = hist[8 * lane + j]; sum += h[j]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_down_sync(FULL, incl, o);
      if (lane + o < 32) incl += v;
    }
    const uint32_t excl = incl - sum;
    const bool here = (excl < rem) && (incl >= rem);
    const int src = __ffs(__ballot_sync(FULL, here)) - 1;
    uint32_t D = 0, above = 0, hD = 0;
    if (lane == src) {
      uint32_t acc = excl;
#pragma unroll
      for (int j = 7; j >= 0; --j) {
        if (acc + h[j] >= rem) { D = 8 * lane + j; above = acc; hD = h[j]; break; }
        acc += h[j];
      }
    }
    D = __shfl_sync(FULL, D, src);
    above = __shfl_sync(FULL, above, src);
    hD = __shfl_sync(FULL, hD, src);
    rem -= above;
    prefix |= uint64_t(D) << shift;
    mask |= uint64_t(0xFF) << shift;
    __syncwarp();
    if (hD == rem || shift == 0) break;
  }
  int base = 0;
  uint32_t min_ord = 0xFFFFFFFFu;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const bool keep = key[j] != 0ull && ((key[j] & mask) >= prefix);
    const unsigned b = __ballot_sync(FULL, keep);
    if (keep) {
      list[base + __popc(b & ((1u << lane) - 1u))] = key[j];
      min_ord = min(min_ord, uint32_t(key[j] >> 32));
    }
    base += __popc(b);
  }
  __syncwarp();
  return warp_min_u32(min_ord);
}

static __device__ __forceinline__ uint32_t warp_prune_any(uint64_t* list, int n, int k, int cap,
                                                          uint32_t* hist) {
  if (cap <= 512) return warp_prune_reg<16>(list, n, k, hist);
  if (cap <= 1024) return warp_prune_reg<32>(list, n, k, hist);
  return warp_prune(list, n, k, hist);
}

// Per-thread (= per query row) running state of the fused selection.
struct RowState {
  uint64_t* list;     // this (item,row)'s candidate list, capacity `cap` keys
  int cnt;
  uint32_t ord_local; // ord of the k-th best kept so far in THIS list (0 = list never pruned)
  uint32_t ord_global;// last value read from the cross-CTA shared threshold
  float thr;          // strict pass threshold
};

// Called warp-uniformly after a chunk of at most `chunk` appends per lane: prunes every lane's
// list that could overflow on the next chunk (limit = cap - chunk), or - at item end - every list
// longer than `limit` (the merge kernel's per-list budget).  gthr_row = &shared_threshold[query].
__device__ __forceinline__ void prune_if_needed(RowState& st, int k, int cap, int limit,
                                                uint32_t* gthr_row, uint32_t* hist) {
  const int lane = threadIdx.x & 31;
  unsigned m = __ballot_sync(FULL, st.cnt > limit);
  while (m) {
    int src = __ffs(m) - 1;
    m &= m - 1;
    unsigned long long lp = (unsigned long long)st.list;
    lp = __shfl_sync(FULL, lp, src);
    int n = __shfl_sync(FULL, st.cnt, src);
    uint32_t kth = warp_prune_any((uint64_t*)lp, n, k, cap, hist);
    if (lane == src) {
      st.cnt = k;
      st.ord_local = kth;
      if (gthr_row) {
        uint32_t old = atomicMax(gthr_row, kth);
        st.ord_global = old > kth ? old : kth;
      }
      st.thr = combine_thr(st.ord_local, st.ord_global);
    }
  }
}

// ---- published order statistics ("rungs") ---------------------------------------------------------
// Every candidate list of the first wave keeps the m best scores it has produced so far in
// registers and publishes the m-th of them (as an orderable uint32; 0 = fewer than m rows yet) into
// pub[query][slice].  When all P lists of a query have published, min over them is a score that at
// least P*m >= k distinct corpus rows reach - a valid (non-strict) global threshold that tightens
// after every tile, without a separate seeding pass and without waiting for any list to fill up.
// The publisher itself folds that minimum into the query's shared threshold (atomicMax), so readers
// keep polling one word per tile; every value ever written is a valid bound, so stale reads of
// other lists' slots only make the published minimum lag, never wrong.
constexpr int PUB_LD = 32;      // pub row pitch in words (at most 32 publishing lists per query)
constexpr int PUB_MAX_M = 8;    // order statistic tracked per list (registers)

__device__ __forceinline__ void top8_insert(float (&t)[PUB_MAX_M], float f) {   // precondition: f > t[7]
  t[PUB_MAX_M - 1] = f;
#pragma unroll
  for (int j = PUB_MAX_M - 1; j > 0; --j)
    if (t[j] > t[j - 1]) { const float x = t[j]; t[j] = t[j - 1]; t[j - 1] = x; }
}
__device__ __forceinline__ float top8_get(const float (&t)[PUB_MAX_M], int idx) {
  float v = t[0];
#pragma unroll
  for (int j = 1; j < PUB_MAX_M; ++j) v = (j == idx) ? t[j] : v;
  return v;
}
// min over the first n published entries of a query's row (0 if any list has not published yet);
// the row is one 128-byte line, read with eight independent loads
__device__ __forceinline__ uint32_t pub_min(const uint32_t* row, int n) {
  uint4 v[PUB_LD / 4];
#pragma unroll
  for (int i = 0; i < PUB_LD / 4; ++i) v[i] = __ldcg(reinterpret_cast<const uint4*>(row) + i);
  uint32_t g = 0xFFFFFFFFu;
#pragma unroll
  for (int i = 0; i < PUB_LD / 4; ++i) {
    g = min(g, 4 * i + 0 < n ? v[i].x : 0xFFFFFFFFu);
    g = min(g, 4 * i + 1 < n ? v[i].y : 0xFFFFFFFFu);
    g = min(g, 4 * i + 2 < n ? v[i].z : 0xFFFFFFFFu);
    g = min(g, 4 * i + 3 < n ? v[i].w : 0xFFFFFFFFu);
  }
  return g;
}

// Workspace plan shared by both dense paths (host side).
struct DensePlan {
  int rows_per_item;   // query rows per work item (128 tcgen05, 64 simt)
  int tile_n;          // corpus rows per tile
  int MB;              // query blocks
  int S;               // corpus slices
  int64_t tiles;       // total corpus tiles
  int cap;             // list capacity (keys)
  int keep;            // a list longer than this is pruned to k at item end (S*keep <= 8192)
  int seed_rows;       // >0: thresholds are seeded from exact scores of the first seed_rows rows
  int seed_S;          // corpus slices of the seed pass
  int cl;              // tcgen05 pairs per cluster sharing corpus tiles by TMA multicast (1, 2 or 4)
  int S_tail;          // cl > 1: the last S_tail slices run as plain pairs on the SMs no cluster fits on
  int64_t tiles_main;  // cl > 1: corpus tiles covered by the first S - S_tail slices
  int x3_d;            // >0: rows are three bf16 planes [x1|x2|x3] of a d=x3_d fp32 vector (width 3*x3_d)
  int sets;            // tcgen05: epilogue warp sets per CTA (1 or 2) = candidate lists per (item, query row)
  int pub_n, pub_m;    // tcgen05: pub_n > 0 = the first pub_n slices publish their pub_m-th best score (replaces seeding)
  size_t off_lists, off_counts, off_gthr, off_pub, off_keys, off_qpad, off_seed, off_mscratch, total;
};

int plan_dense(int64_t n, int d, int dtype, int nq, int k, int path, DensePlan* plan, int* path_out);

int launch_dense_simt(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                      int k, const DensePlan& pl, uint64_t* lists, int* counts, uint32_t* gthr,
                      cudaStream_t stream);
int launch_dense_tc(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                    int k, const DensePlan& pl, uint64_t* lists, int* counts, uint32_t* gthr, uint32_t* pub,
                    float* seed_scores, void* qpad, cudaEvent_t after_seed, cudaStream_t stream);
int launch_seed_select(const float* seed_scores, int nq, int seed_rows, int k, uint32_t* gthr,
                       cudaStream_t stream);
bool dense_tc_supported(const void* corpus, int64_t n, int d, int dtype, const void* queries);

// Multi-GPU "push" exchange: instead of a local [nq,k] block the sorted keys of query q go straight
// into the inbox of the rank that owns the query (owner = q / nq_per; inboxes[owner] may be peer
// memory reached over NVLink), at [this rank][q - owner*nq_per][k].
struct MergePush {
  uint64_t* const* inboxes;   // device array: one inbox base pointer per rank
  int rank;                   // this rank (row of the inbox it writes)
  int nq_per;                 // queries owned by each rank (the last one may own fewer)
  int n_ranks;                // > 0: every inbox is followed by nq_per arrival counters (uint32) at
                              // inbox + n_ranks*nq_per*k keys; after a query's row has landed the pusher
                              // bumps the owner's counter of that query (release, system scope), which is
                              // what the owner's merge kernel waits on instead of a barrier between kernels
};

// merge of per-slice candidate lists -> sorted keys [nq,k] (+ optional decoded outputs); gthr (may
// be NULL) = per-query score bound below which candidates are dropped while gathering
int launch_merge_lists(const uint64_t* lists, const int* counts, const DensePlan& pl, int nq, int k,
                       uint64_t id_base, const uint32_t* gthr, uint64_t* scratch, uint64_t* out_keys,
                       float* out_scores, int64_t* out_ids, const MergePush* push, cudaStream_t stream);

int sm_count();
int dense_tc_max_sets();              // epilogue warp sets the tcgen05 kernel was built for
int dense_tc_units(int cg, int cl);   // persistent work-item slots (CTAs or pairs) the tcgen05 kernel keeps resident

}  // namespace ragarc
