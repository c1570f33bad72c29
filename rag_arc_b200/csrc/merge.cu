// Final selection: merge the per-slice candidate lists of one query (or the all-gathered
// per-GPU top-k lists) into one sorted top-k.  One CTA per query: the candidate keys (64-bit,
// larger = better, unique) are gathered into shared memory with one warp per list, the k best are
// picked by an MSB-first 8-bit radix select (early exit, typically 3 passes) and only those are
// bitonic-sorted.  The result is deterministic - score descending, ties by ascending row id -
// independent of how the corpus was tiled, sliced or sharded.
// Also here: the threshold seeding select (k-th largest of a dense block of scores per query).
#include "common.cuh"

namespace ragarc {

constexpr int MERGE_THREADS = 512;        // seed_select_kernel
constexpr int MERGE_LISTS_THREADS = 256;  // merge_lists_kernel: 8 CTAs per SM, all queries of a 1024-batch resident at once
constexpr int MERGE_SMEM_KEYS = 2048;     // filtered candidate keys a merge_lists CTA holds in shared memory; a query
                                          // with more survivors (loose thresholds) continues in a global scratch row

// k-th largest of keys[0..T) (T > k): leaves the winners (exactly k) in win[0..k), unordered.
// hist: 256 words, sel: 3 words, nwin: 1 word of shared memory.
__device__ __forceinline__ void block_select_topk(const uint64_t* keys, int T, int k, uint64_t* win,
                                                  uint32_t* hist, uint32_t* sel, uint32_t* nwin,
                                                  unsigned long long* orand) {
  // leading bytes common to all keys need no pass
  if (threadIdx.x == 0) { orand[0] = 0ull; orand[1] = ~0ull; }
  __syncthreads();
  {
    uint32_t oh = 0, ol = 0, ah = 0xFFFFFFFFu, al = 0xFFFFFFFFu;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
      const uint64_t key = keys[i];
      oh |= uint32_t(key >> 32); ol |= uint32_t(key); ah &= uint32_t(key >> 32); al &= uint32_t(key);
    }
    oh = __reduce_or_sync(FULL, oh); ol = __reduce_or_sync(FULL, ol);
    ah = __reduce_and_sync(FULL, ah); al = __reduce_and_sync(FULL, al);
    if ((threadIdx.x & 31) == 0) {
      atomicOr(&orand[0], (unsigned long long)((uint64_t(oh) << 32) | ol));
      atomicAnd(&orand[1], (unsigned long long)((uint64_t(ah) << 32) | al));
    }
  }
  __syncthreads();
  const int shift0 = first_varying_shift(orand[0], orand[1]);
  uint64_t mask = high_bytes_mask(shift0);
  uint64_t prefix = orand[1] & mask;
  uint32_t rem = (uint32_t)k;
  for (int shift = shift0;; shift = next_shift(shift)) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int b = 0; b < T; b += blockDim.x) {
      const int i = b + threadIdx.x;
      const uint64_t key = i < T ? keys[i] : 0ull;
      hist_add(hist, (uint32_t)(key >> shift) & 0xFFu, i < T && (key & mask) == prefix);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
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
      if (excl < rem && incl >= rem) {
        uint32_t a = excl;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (a + h[j] >= rem) { sel[0] = 8 * lane + j; sel[1] = a; sel[2] = h[j]; break; }
          a += h[j];
        }
      }
    }
    __syncthreads();
    const uint32_t D = sel[0], above = sel[1], hD = sel[2];
    rem -= above;
    prefix |= uint64_t(D) << shift;
    mask |= uint64_t(0xFF) << shift;
    __syncthreads();
    if (hD == rem || shift == 0) break;
  }
  if (threadIdx.x == 0) *nwin = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const uint64_t key = keys[i];
    if ((key & mask) >= prefix) win[atomicAdd(nwin, 1u)] = key;
  }
  __syncthreads();
}

// Bitonic sort (descending) of s[0..P).  Only the first min(blockDim, P/2) threads take part and
// they synchronise on named barrier 1, so a 128-key sort does not make 512 threads spin through
// 28 block-wide barriers.  Ends with a block-wide barrier.
__device__ __forceinline__ void block_bitonic_desc(uint64_t* s, int P) {
  int nact = P >> 1;
  if (nact > (int)blockDim.x) nact = blockDim.x;
  nact = (nact + 31) & ~31;
  __syncthreads();
  if ((int)threadIdx.x < nact) {
    for (int size = 2; size <= P; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int i = threadIdx.x; i < (P >> 1); i += nact) {
          const int lo = 2 * i - (i & (stride - 1));
          const int hi = lo + stride;
          const bool desc = (lo & size) == 0;
          const uint64_t a = s[lo], b = s[hi];
          if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nact) : "memory");
      }
    }
  }
  __syncthreads();
}

// The same sort for P = 32*E <= 256 keys by ONE warp, keys in registers (E per lane, key index = lane*E + e):
// exchanges at a distance below E stay inside a lane, the others are two 32-bit shuffles per key - no
// shared-memory round trips and no barriers between the 15-36 exchange steps, which is what the
// shared-memory version spends its time on when only 2-4 of a CTA's warps have work.  Call with warp 0;
// the other warps wait at the caller's barrier.
template <int E>
__device__ __forceinline__ void warp_bitonic_desc(uint64_t* s) {
  const int lane = threadIdx.x & 31;
  uint64_t v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) v[e] = s[lane * E + e];
#pragma unroll
  for (int size = 2; size <= 32 * E; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= E) {
        const int lstride = stride / E;
        // every key of this lane has the same role: low partner iff the lane's bit is clear
        const bool desc = ((lane * E) & size) == 0, is_lo = (lane & lstride) == 0;
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const uint32_t ph = __shfl_xor_sync(FULL, (uint32_t)(v[e] >> 32), lstride);
          const uint32_t pl = __shfl_xor_sync(FULL, (uint32_t)v[e], lstride);
          const uint64_t o = ((uint64_t)ph << 32) | pl;
          const bool take_max = is_lo == desc;
          v[e] = take_max ? (o > v[e] ? o : v[e]) : (o < v[e] ? o : v[e]);
        }
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if ((e & stride) == 0) {
            const int idx = lane * E + e;
            const bool desc = (idx & size) == 0;
            const uint64_t a = v[e], b = v[e | stride];
            const bool swap = (a < b) == desc;
            v[e] = swap ? b : a; v[e | stride] = swap ? a : b;
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) s[lane * E + e] = v[e];
}

// sorts s[0..P) descending; P a power of two >= 32; the block-wide barrier at the end publishes the result
__device__ __forceinline__ void sort_desc(uint64_t* s, int P) {
  if (P > 256) { block_bitonic_desc(s, P); return; }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (P == 32) warp_bitonic_desc<1>(s);
    else if (P == 64) warp_bitonic_desc<2>(s);
    else if (P == 128) warp_bitonic_desc<4>(s);
    else warp_bitonic_desc<8>(s);
  }
  __syncthreads();
}

// out_keys / out_scores / out_ids point at the [nq,k] blocks; the row of query q is written.
__device__ __forceinline__ void emit_topk(const uint64_t* s, int q, int k, uint64_t id_base,
                                          uint64_t* out_keys, float* out_scores, int64_t* out_ids) {
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const uint64_t key = s[j];
    if (out_keys) {
      uint64_t gk = 0;
      if (key) gk = (key & 0xFFFFFFFF00000000ull) | uint64_t(0xFFFFFFFFu - (uint32_t)(key_row(key) + id_base));
      out_keys[(size_t)q * k + j] = gk;
    }
    if (out_scores) out_scores[(size_t)q * k + j] = key ? key_score(key) : -INFINITY;
    if (out_ids) out_ids[(size_t)q * k + j] = key ? (int64_t)(key_row(key) + id_base) : -1;
  }
}



// Shared tail: keys[0..T) gathered in smem -> top-k sorted in win[0..PK) -> outputs.
__device__ __forceinline__ void select_sort_emit(uint64_t* keys, int T, int k, int PK, uint64_t* win,
                                                 uint32_t* hist, uint32_t* sel, uint32_t* nwin,
                                                 unsigned long long* orand, int q,
                                                 uint64_t id_base, uint64_t* out_keys, float* out_scores,
                                                 int64_t* out_ids) {
  int have;
  if (T > k) {
    block_select_topk(keys, T, k, win, hist, sel, nwin, orand);
    have = k;
  } else {
    for (int i = threadIdx.x; i < T; i += blockDim.x) win[i] = keys[i];
    have = T;
  }
  for (int i = have + threadIdx.x; i < PK; i += blockDim.x) win[i] = 0;
  __syncthreads();
  block_bitonic_desc(win, PK);
  emit_topk(win, q, k, id_base, out_keys, out_scores, out_ids);
}

// lists[(slice*MB + qb) * rows + r][cap], counts[(slice*MB + qb) * rows + r] (<= keep each)
//
// Typical input: S*~160 = ~6000 keys per query of which k=100 are wanted.  Instead of radix-
// selecting over all of them: (A) the gather also finds the smallest and largest score, (B) one
// pass histograms the scores into 256 linear bins over that range and locates the bin holding the
// k-th largest, (C) one compare pass keeps only the keys at or above that bin's lower edge
// (k + one bin's population, typically ~130) and only those are bitonic-sorted.  Falls back to the
// full radix select if more than MERGE_WIN keys survive (heavily tied scores).
constexpr int MERGE_WIN = 512;       // survivors that are sorted directly

__global__ void __launch_bounds__(MERGE_LISTS_THREADS, 8)
merge_lists_kernel(const uint64_t* __restrict__ lists, const int* __restrict__ counts, int MB, int S, int sets,
                   int rows, int cap, int k, int PK, int tmax, int smem_keys, uint64_t* __restrict__ scratch,
                   uint64_t id_base, const uint32_t* __restrict__ gthr, uint64_t* out_keys,
                   float* out_scores, int64_t* out_ids, MergePush push) {
  extern __shared__ uint64_t msm[];
  uint64_t* keys = msm;                       // [smem_keys] (or this query's row of `scratch`, see below)
  uint64_t* win = msm + smem_keys;            // [max(PK, MERGE_WIN)] survivors / winners
  int* offs = (int*)(win + (PK > MERGE_WIN ? PK : MERGE_WIN));   // [S+1]
  int* lrow = offs + S + 1;                   // [S] row of list s in the [.., cap] list matrix
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel[3];
  __shared__ uint32_t nwin, omin, omax;
  __shared__ unsigned long long orand[2];
  const int q = blockIdx.x;
  const int qb = q / rows, r = q % rows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (threadIdx.x < 256) hist[threadIdx.x] = 0;
  if (warp == 0) {
    // exclusive prefix sum of the S list lengths
    int carry = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      int c = 0;
      if (s < S) {
        const int row = ((s / sets * MB + qb) * rows + r) * sets + (s % sets);
        lrow[s] = row;
        c = counts[row];
      }
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      if (s < S) offs[s] = carry + incl - c;
      carry += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) { offs[S] = carry; nwin = 0; omin = 0xFFFFFFFFu; omax = 0u; }
  }
  __syncthreads();
  int total_raw = offs[S];
  if (total_raw > tmax) total_raw = tmax;    // cannot happen: every list is <= keep and S*keep <= tmax
  uint64_t* spill = scratch + (size_t)q * tmax;   // this query's global row: only touched if more than smem_keys candidates survive
  // (A) flattened gather: element e of the concatenated lists belongs to the list found by binary
  // search in the prefix sums, so every load is independent of every other: a thread first issues
  // the loads of all its elements of a pass (one L2 round trip for 1024 candidates), then filters.
  // Candidates below the query's shared threshold - a score that at least k rows are known to
  // reach - are dropped on the way in; what survives goes to shared memory (typically a few hundred).
  {
    constexpr int NPT = 4;                    // elements per thread and pass
    const uint32_t bound = gthr ? gthr[q] : 0u;
    uint32_t lmin = 0xFFFFFFFFu, lmax = 0u;
    for (int b = 0; b < total_raw; b += blockDim.x * NPT) {
      uint64_t kreg[NPT];
#pragma unroll
      for (int u = 0; u < NPT; ++u) {
        const int e = b + u * blockDim.x + threadIdx.x;
        kreg[u] = 0ull;
        if (e < total_raw) {
          int lo = 0, hi = S;                 // largest sl with offs[sl] <= e
          while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (offs[mid] <= e) lo = mid; else hi = mid; }
          kreg[u] = lists[(size_t)lrow[lo] * (size_t)cap + (e - offs[lo])];
        }
      }
#pragma unroll
      for (int u = 0; u < NPT; ++u) {
        const uint64_t key = kreg[u];
        const uint32_t ord = uint32_t(key >> 32);
        const bool keep = key != 0ull && ord >= bound;
        const unsigned bal = __ballot_sync(FULL, keep);
        uint32_t base = 0;
        if (lane == 0 && bal) base = atomicAdd(&nwin, (uint32_t)__popc(bal));
        base = __shfl_sync(FULL, base, 0);
        if (keep) {
          const uint32_t idx = base + __popc(bal & ((1u << lane) - 1u));
          if (idx < (uint32_t)smem_keys) keys[idx] = key; else spill[idx] = key;
          lmin = min(lmin, ord); lmax = max(lmax, ord);
        }
      }
    }
    lmin = __reduce_min_sync(FULL, lmin); lmax = __reduce_max_sync(FULL, lmax);
    if (lane == 0) { atomicMin(&omin, lmin); atomicMax(&omax, lmax); }
  }
  __syncthreads();
  if ((int)nwin > smem_keys) {
    // rare (loose thresholds, heavy ties): continue out of the global row - same code, generic pointer
    for (int i = threadIdx.x; i < smem_keys; i += blockDim.x) spill[i] = keys[i];
    keys = spill;
    __syncthreads();
  }
  const int total = (int)nwin;
  __syncthreads();
  if (threadIdx.x == 0) nwin = 0;
  if (push.inboxes) {
    // redirect the key output of this query to its owner's inbox (see MergePush)
    const int owner = q / push.nq_per;
    out_keys = push.inboxes[owner] + (ptrdiff_t)(push.rank - owner) * (ptrdiff_t)push.nq_per * (ptrdiff_t)k;
  }
  __syncthreads();
  bool fast = total > k;
  uint32_t bound = 0;
  if (fast) {
    // (B) 256 linear bins over [omin, omax]
    const uint32_t lo = omin, range = omax - omin;
    const int shift = range >= 256u ? (32 - __clz(range)) - 8 : 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x)
      atomicAdd(&hist[(uint32_t(keys[i] >> 32) - lo) >> shift], 1u);
    __syncthreads();
    if (warp == 0) {
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
      if (excl < (uint32_t)k && incl >= (uint32_t)k) {
        uint32_t a = excl;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (a + h[j] >= (uint32_t)k) { sel[0] = 8 * lane + j; sel[1] = a + h[j]; break; }
          a += h[j];
        }
      }
    }
    __syncthreads();
    bound = lo + (sel[0] << shift);             // every key with score-ord >= bound survives: sel[1] >= k of them
    fast = sel[1] <= (uint32_t)MERGE_WIN;
  }
  if (fast) {
    // (C) compact the survivors, sort them, emit the first k
    for (int b = 0; b < total; b += blockDim.x) {
      const int i = b + threadIdx.x;
      const uint64_t key = i < total ? keys[i] : 0ull;
      const bool keep = i < total && uint32_t(key >> 32) >= bound;
      const unsigned bal = __ballot_sync(FULL, keep);
      uint32_t base = 0;
      if (lane == 0 && bal) base = atomicAdd(&nwin, (uint32_t)__popc(bal));
      base = __shfl_sync(FULL, base, 0);
      if (keep) win[base + __popc(bal & ((1u << lane) - 1u))] = key;
    }
    __syncthreads();
    const int have = (int)nwin;
    int PW = 32; while (PW < have) PW <<= 1;
    for (int i = have + threadIdx.x; i < PW; i += blockDim.x) win[i] = 0;
    sort_desc(win, PW);
    emit_topk(win, q, k, id_base, out_keys, out_scores, out_ids);
  } else {
    select_sort_emit(keys, total, k, PK, win, hist, sel, &nwin, orand, q, id_base, out_keys, out_scores, out_ids);
  }
  if (push.inboxes && push.n_ranks > 0) {
    // the row is in the owner's inbox: one release-add at system scope by one thread, after a CTA barrier,
    // publishes every thread's stores (release is cumulative over what the barrier ordered before it);
    // fencing in all 256 threads instead cost 15 us per search
    __syncthreads();
    if (threadIdx.x == 0) {
      const int owner = q / push.nq_per;
      uint32_t* counters = reinterpret_cast<uint32_t*>(push.inboxes[owner] + (size_t)push.n_ranks * push.nq_per * k);
      asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(counters + (q - owner * push.nq_per)) : "memory");
    }
  }
}

// G per-shard result lists per query, each sorted descending (the output format of
// ragarc_dense_topk_keys; 0 = empty slot, at the tail) -> global top k_out.  No sort: every key's
// global rank = its position in its own list + the number of larger keys in each other list
// (binary search in shared memory); keys are unique, so ranks are a permutation.
// list g of query q lives at (ptrs ? ptrs[g] : base + g*stride_g) + q*k_in  - with `ptrs` the lists
// may be PEER-GPU memory (NVLink loads), which is how the multi-GPU merge avoids an all-gather.
// counters != NULL: the lists are an inbox other GPUs write into (MergePush); the CTA of query q first
// waits until `expected` ranks have counted themselves in for q (acquire, system scope; bounded: on
// timeout *status is set and the merge proceeds with what is there), resets the counter for the slot's
// next use, and reads the rows past L1.
__global__ void __launch_bounds__(256)
merge_sorted_keys_kernel(const uint64_t* base, size_t stride_g, const uint64_t* const* __restrict__ ptrs,
                         int G, int nq, int k_in, int k_out, float* __restrict__ out_scores,
                         int64_t* __restrict__ out_ids, uint32_t* counters, uint32_t expected,
                         long long timeout_cycles, uint32_t* status) {
  extern __shared__ uint64_t sk[];            // [G][k_in]
  const int q = blockIdx.x, total = G * k_in;
  if (counters) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      uint32_t seen = 0;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(counters + q) : "memory");
        if (seen >= expected) break;
        if (clock64() - t0 > timeout_cycles) { if (status) atomicOr(status, 1u); break; }
        __nanosleep(64);
      }
      counters[q] = 0;                        // next use of this slot is two searches away (see sharded.py)
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const int g = e / k_in, i = e - g * k_in;
    const uint64_t* src = (ptrs ? ptrs[g] : base + (size_t)g * stride_g) + (size_t)q * k_in;
    sk[e] = counters ? __ldcg(src + i) : src[i];
  }
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    out_scores[(size_t)q * k_out + j] = -INFINITY;
    out_ids[(size_t)q * k_out + j] = -1;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    const uint64_t key = sk[e];
    if (!key) continue;
    const int g = e / k_in, i = e - g * k_in;
    int rank = i;
    for (int h = 0; h < G; ++h) {
      if (h == g) continue;
      const uint64_t* lst = sk + h * k_in;
      int lo = 0, hi = k_in;                  // first index whose key is < `key` (descending list)
      while (lo < hi) { const int m = (lo + hi) >> 1; if (lst[m] > key) lo = m + 1; else hi = m; }
      rank += lo;
    }
    if (rank < k_out) {
      out_scores[(size_t)q * k_out + rank] = key_score(key);
      out_ids[(size_t)q * k_out + rank] = (int64_t)key_row(key);
    }
  }
}

// Threshold seeding: vals[q][0..S) (fp32; each the maximum score of a distinct group of corpus
// rows) -> gthr[q] = orderable(k-th largest).  "At least k rows score at least this" is all the
// main pass needs to start with a tight filter instead of -inf.
__global__ void __launch_bounds__(MERGE_THREADS)
seed_select_kernel(const float* __restrict__ scores, int S, int k, uint32_t* __restrict__ gthr) {
  extern __shared__ uint32_t ords[];   // [S]
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel[3];
  const int q = blockIdx.x;
  const float* row = scores + (size_t)q * S;
  for (int i = threadIdx.x; i < S; i += blockDim.x) ords[i] = f32_to_ord(row[i]);
  __syncthreads();
  __shared__ uint32_t orand32[2];
  if (threadIdx.x == 0) { orand32[0] = 0u; orand32[1] = 0xFFFFFFFFu; }
  __syncthreads();
  {
    uint32_t o_ = 0, a_ = 0xFFFFFFFFu;
    for (int i = threadIdx.x; i < S; i += blockDim.x) { o_ |= ords[i]; a_ &= ords[i]; }
    o_ = __reduce_or_sync(FULL, o_); a_ = __reduce_and_sync(FULL, a_);
    if ((threadIdx.x & 31) == 0) { atomicOr(&orand32[0], o_); atomicAnd(&orand32[1], a_); }
  }
  __syncthreads();
  const uint32_t diff = orand32[0] ^ orand32[1];
  const int top = diff ? 31 - __clz((int)diff) : 0;
  const int shift0 = top >= 7 ? top - 7 : 0;
  uint32_t mask = shift0 + 8 >= 32 ? 0u : ~((1u << (shift0 + 8)) - 1u);
  uint32_t prefix = orand32[1] & mask, rem = (uint32_t)k;
  for (int shift = shift0;; shift = next_shift(shift)) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int b = 0; b < S; b += blockDim.x) {
      const int i = b + threadIdx.x;
      const uint32_t o = i < S ? ords[i] : 0u;
      hist_add(hist, (o >> shift) & 0xFFu, i < S && (o & mask) == prefix);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
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
      if (excl < rem && incl >= rem) {
        uint32_t a = excl;
#pragma unroll
        for (int j = 7; j >= 0; --j) {
          if (a + h[j] >= rem) { sel[0] = 8 * lane + j; sel[1] = a; sel[2] = h[j]; break; }
          a += h[j];
        }
      }
    }
    __syncthreads();
    rem -= sel[1];
    prefix |= sel[0] << shift;
    mask |= 0xFFu << shift;
    __syncthreads();
    if (shift == 0) break;
  }
  if (threadIdx.x == 0) gthr[q] = prefix;    // all varying bytes resolved: exact k-th largest ord
}

static int next_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

int launch_merge_lists(const uint64_t* lists, const int* counts, const DensePlan& pl, int nq, int k,
                       uint64_t id_base, const uint32_t* gthr, uint64_t* scratch, uint64_t* out_keys,
                       float* out_scores, int64_t* out_ids, const MergePush* push, cudaStream_t stream) {
  const int VS = pl.S * pl.sets;              // candidate lists per query
  const int tmax = VS * pl.keep;
  const int PK = next_pow2(k);
  RA_REQUIRE(tmax <= 16384 && VS <= 1024 && PK <= 2048, "merge: S*sets*keep=%d too large", tmax);
  const int smem_keys = tmax < MERGE_SMEM_KEYS ? tmax : MERGE_SMEM_KEYS;
  const size_t smem = (size_t)(smem_keys + (PK > MERGE_WIN ? PK : MERGE_WIN)) * 8 + (size_t)(2 * VS + 1) * 4;
  RA_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (MERGE_SMEM_KEYS + 2048) * 8 + 2049 * 4));
  MergePush mp{nullptr, 0, 1, 0};
  if (push) mp = *push;
  merge_lists_kernel<<<nq, MERGE_LISTS_THREADS, smem, stream>>>(lists, counts, pl.MB, VS, pl.sets, pl.rows_per_item,
                                                               pl.cap, k, PK, tmax, smem_keys, scratch, id_base, gthr,
                                                               out_keys, out_scores, out_ids, mp);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int launch_seed_select(const float* seed_scores, int nq, int seed_rows, int k, uint32_t* gthr,
                       cudaStream_t stream) {
  RA_REQUIRE(seed_rows >= k && seed_rows <= 32768, "seed_select: bad group count %d", seed_rows);
  RA_CUDA(cudaFuncSetAttribute(seed_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  seed_select_kernel<<<nq, MERGE_THREADS, (size_t)seed_rows * 4, stream>>>(seed_scores, seed_rows, k, gthr);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

}  // namespace ragarc

using namespace ragarc;

static int merge_sorted_launch(const uint64_t* base, size_t stride_g, const uint64_t* const* ptrs, int nlists,
                               int nq, int k_in, int k_out, float* out_scores, int64_t* out_ids, void* stream,
                               uint32_t* counters = nullptr, uint32_t expected = 0, long long timeout_cycles = 0,
                               uint32_t* status = nullptr) {
  RA_REQUIRE(out_scores && out_ids, "merge_topk_keys: null pointer");
  RA_REQUIRE(nlists > 0 && nq >= 0 && k_in > 0 && k_out > 0 && k_out <= nlists * k_in,
             "merge_topk_keys: bad shape G=%d nq=%d k_in=%d k_out=%d", nlists, nq, k_in, k_out);
  if (nq == 0) return RAGARC_OK;
  const size_t smem = (size_t)nlists * k_in * 8;
  RA_REQUIRE(smem <= 200 * 1024, "merge_topk_keys: nlists*k_in=%d too large", nlists * k_in);
  if (smem > 48 * 1024)
    RA_CUDA(cudaFuncSetAttribute(merge_sorted_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_sorted_keys_kernel<<<nq, 256, smem, (cudaStream_t)stream>>>(base, stride_g, ptrs, nlists, nq, k_in, k_out,
                                                                   out_scores, out_ids, counters, expected,
                                                                   timeout_cycles, status);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

extern "C" int ragarc_merge_topk_keys(const uint64_t* keys, int nlists, int nq, int k_in, int k_out,
                                      float* out_scores, int64_t* out_ids, void* stream) {
  RA_REQUIRE(keys, "merge_topk_keys: null keys");
  return merge_sorted_launch(keys, (size_t)nq * k_in, nullptr, nlists, nq, k_in, k_out, out_scores, out_ids, stream);
}

extern "C" int ragarc_merge_topk_keys_p2p(const uint64_t* const* key_ptrs, int nlists, int nq, int k_in, int k_out,
                                          float* out_scores, int64_t* out_ids, void* stream) {
  RA_REQUIRE(key_ptrs, "merge_topk_keys_p2p: null pointer table");
  return merge_sorted_launch(nullptr, 0, key_ptrs, nlists, nq, k_in, k_out, out_scores, out_ids, stream);
}

extern "C" int ragarc_merge_topk_inbox(uint64_t* inbox, int n_ranks, int nq_per_rank, int nq_own, int k_in, int k_out,
                                       float* out_scores, int64_t* out_ids, double timeout_ms,
                                       uint32_t* status, void* stream) {
  RA_REQUIRE(inbox, "merge_topk_inbox: null inbox");
  RA_REQUIRE(n_ranks > 0 && nq_per_rank > 0 && k_in > 0 && nq_own >= 0 && nq_own <= nq_per_rank,
             "merge_topk_inbox: bad shape");
  if (nq_own == 0) return RAGARC_OK;          // this rank owns no query of the batch: nothing arrives, nothing to merge
  uint32_t* counters = reinterpret_cast<uint32_t*>(inbox + (size_t)n_ranks * nq_per_rank * k_in);
  const long long cycles = (long long)((timeout_ms > 0 ? timeout_ms : 2000.0) * 1.5e6);   // ~1.5 GHz worst case
  return merge_sorted_launch(inbox, (size_t)nq_per_rank * k_in, nullptr, n_ranks, nq_own, k_in, k_out,
                             out_scores, out_ids, stream, counters, (uint32_t)n_ranks, cycles, status);
}
