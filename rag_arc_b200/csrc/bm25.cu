// BM25 (Okapi) scoring over CSR postings + exact top-k, fp64, bit-identical to the numpy
// expression the reference evaluates through rank_bm25 (core/retrieval/bm25.py:306):
//
//   score[d] += idf[t] * ( tf * (k1 + 1) / ( tf + k1 * (1 - b + b * dl[d] / avgdl) ) )
//
// Query terms are processed SEQUENTIALLY in query order (duplicates repeat), the postings of one
// term in parallel: a doc id occurs at most once per posting list, so no two threads touch the same
// accumulator inside a term and there are no atomics - that, plus the explicit round-to-nearest
// intrinsics (no FMA contraction), is what makes the fp64 sums reproduce numpy's bit for bit.
//
// Two schedules:
//  * bm25_range_kernel (the top-k path): a CTA owns (query, range of consecutive docs) and keeps that
//    range's fp64 accumulators in SHARED memory: <= 12 288 docs (96 KB) and 512 threads with two CTAs per
//    SM, or <= 24 576 docs (192 KB) and 1024 threads with one when the corpus is large (bm25_range_geometry).  Posting lists are sorted by
//    doc, so the range's postings of a term are one contiguous span, found by a warp-cooperative
//    32-ary search; they are streamed once (coalesced loads of the doc id and the precomputed
//    per-posting factor, four postings per thread in flight), the read-modify-write stays on chip,
//    and the range's top-k is selected straight out of shared memory.  A small merge kernel
//    combines the per-range candidates.  Traffic per query = its postings, 12 B each (16 B with the
//    doc-norm gather when no factor table is given) - the algorithmic minimum; at C2 the postings
//    live in L2 and the kernel is issue/latency-bound, not bandwidth-bound (profiles/).
//  * bm25_score_kernel (ragarc_bm25_scores and the fallback for huge k): one CTA per query with a
//    dense fp64 accumulator row in global memory (L2), followed by bm25_topk_kernel.
//
// Selection (block_topk_f64): linear 256-bin histogram over [min,max] -> the bin holding the k-th
// largest and everything above survive (3 passes); if too many survive (massive ties, e.g. the
// zero-score tail) an MSB-first radix select over the 96-bit (score, ~doc id) key decides.  Either
// way ties resolve to the lowest doc id and the winners are bitonic-sorted.
#include <cstdlib>
#include "common.cuh"

namespace ragarc {

namespace bm25 {
constexpr int THREADS = 1024;
constexpr int KMAX = 1024;
constexpr int RANGE = 24576;                      // docs per CTA in the shared-memory schedule
constexpr int RANGE_SMEM = RANGE * 8;             // 192 KB
}

__device__ __forceinline__ uint64_t f64_to_ord(double v) {
  v = __dadd_rn(v, 0.0);   // canonicalise -0.0
  uint64_t u = (uint64_t)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double ord_to_f64(uint64_t o) {
  uint64_t u = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
  return __longlong_as_double((long long)u);
}
__device__ __forceinline__ double bm25_term(double w, double tf, double k1p1, double norm) {
  return __dmul_rn(w, __ddiv_rn(__dmul_rn(tf, k1p1), __dadd_rn(tf, norm)));
}

// finds the bucket that contains the rem-th largest element of a 256-bin histogram (warp 0 only;
// lane owns bins [8*lane, 8*lane+8)): out3 = (digit, count above, count in digit)
__device__ __forceinline__ void find_bucket(const uint32_t* hist, uint32_t rem, uint32_t* out3) {
  const int lane = threadIdx.x & 31;
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
      if (a + h[j] >= rem) { out3[0] = 8 * lane + j; out3[1] = a; out3[2] = h[j]; break; }
      a += h[j];
    }
  }
}

struct TopkScratch {          // shared-memory scratch of block_topk_f64
  uint32_t hist[256];
  uint32_t sel[3];
  uint32_t nwin;
  unsigned long long omin, omax;
  uint64_t win_ord[bm25::KMAX];
  uint32_t win_id[bm25::KMAX];
};

// Top-kk of acc[0..n) (global or shared memory), ids reported as id_base + index.  On return
// win_ord/win_id[0..kk) hold the winners sorted by (score desc, id asc).  All threads must call.
__device__ __forceinline__ void block_topk_f64(const double* acc, int n, int kk, uint32_t id_base, TopkScratch& S) {
  bool selected = false;
  if (threadIdx.x == 0) { S.omin = ~0ull; S.omax = 0ull; S.nwin = 0; }
  __syncthreads();
  if (kk > 0) {
    uint64_t lmin = ~0ull, lmax = 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t o = f64_to_ord(acc[i]);
      lmin = o < lmin ? o : lmin; lmax = o > lmax ? o : lmax;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const uint64_t a = __shfl_xor_sync(FULL, lmin, off), b = __shfl_xor_sync(FULL, lmax, off);
      lmin = a < lmin ? a : lmin; lmax = b > lmax ? b : lmax;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&S.omin, (unsigned long long)lmin); atomicMax(&S.omax, (unsigned long long)lmax); }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) S.hist[i] = 0;
    __syncthreads();
    const double mn = ord_to_f64(S.omin), mx = ord_to_f64(S.omax);
    if (mx > mn && isfinite(mx) && isfinite(mn)) {
      const double scale = 256.0 / (mx - mn);
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int b = (int)((acc[i] - mn) * scale);
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        atomicAdd(&S.hist[b], 1u);
      }
      __syncthreads();
      if (threadIdx.x < 32) find_bucket(S.hist, (uint32_t)kk, S.sel);
      __syncthreads();
      const int bstar = (int)S.sel[0];
      if (S.sel[1] + S.sel[2] <= (uint32_t)bm25::KMAX) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
          const double v = acc[i];
          int b = (int)((v - mn) * scale);
          b = b < 0 ? 0 : (b > 255 ? 255 : b);
          if (b >= bstar) {
            const uint32_t slot = atomicAdd(&S.nwin, 1u);
            S.win_ord[slot] = f64_to_ord(v); S.win_id[slot] = id_base + (uint32_t)i;
          }
        }
        selected = true;
      }
      __syncthreads();
    }
  }
  int have = kk;
  if (selected) {
    have = (int)S.nwin;                // >= kk survivors; the sort below puts the kk best first
  } else {
    // composite key = (ord64(score), ~index): 12 digits of 8 bits, MSB first
    uint64_t pre_hi = 0, mask_hi = 0;
    uint32_t pre_lo = 0, mask_lo = 0;
    uint32_t rem = (uint32_t)kk;
    for (int pass = 0; pass < 12 && kk > 0; ++pass) {
      for (int i = threadIdx.x; i < 256; i += blockDim.x) S.hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t o = f64_to_ord(acc[i]);
        const uint32_t nid = ~(uint32_t)i;
        if ((o & mask_hi) == pre_hi && (nid & mask_lo) == pre_lo) {
          const uint32_t dig = pass < 8 ? (uint32_t)(o >> (56 - 8 * pass)) & 0xFFu
                                        : (nid >> (24 - 8 * (pass - 8))) & 0xFFu;
          atomicAdd(&S.hist[dig], 1u);
        }
      }
      __syncthreads();
      if (threadIdx.x < 32) find_bucket(S.hist, rem, S.sel);
      __syncthreads();
      const uint32_t D = S.sel[0], above = S.sel[1], hD = S.sel[2];
      rem -= above;
      if (pass < 8) { pre_hi |= (uint64_t)D << (56 - 8 * pass); mask_hi |= (uint64_t)0xFF << (56 - 8 * pass); }
      else { pre_lo |= D << (24 - 8 * (pass - 8)); mask_lo |= 0xFFu << (24 - 8 * (pass - 8)); }
      __syncthreads();
      if (hD == rem) break;
    }
    if (threadIdx.x == 0) S.nwin = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t o = f64_to_ord(acc[i]);
      const uint32_t nid = ~(uint32_t)i;
      const uint64_t oh = o & mask_hi;
      const bool keep = kk > 0 && (oh > pre_hi || (oh == pre_hi && (nid & mask_lo) >= pre_lo));
      if (keep) {
        const uint32_t slot = atomicAdd(&S.nwin, 1u);
        if (slot < bm25::KMAX) { S.win_ord[slot] = o; S.win_id[slot] = id_base + (uint32_t)i; }
      }
    }
    __syncthreads();
  }
  // sort winners: descending score, ascending id.  P = pow2 >= have, pad with (0, max id)
  int P = 1; while (P < have) P <<= 1;
  for (int i = have + threadIdx.x; i < P; i += blockDim.x) { S.win_ord[i] = 0; S.win_id[i] = 0xFFFFFFFFu; }
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t ao = S.win_ord[lo], bo = S.win_ord[hi];
        const uint32_t ai = S.win_id[lo], bi = S.win_id[hi];
        const bool a_less = (ao < bo) || (ao == bo && ai > bi);   // "a ranks below b"
        if (a_less == desc) { S.win_ord[lo] = bo; S.win_ord[hi] = ao; S.win_id[lo] = bi; S.win_id[hi] = ai; }
      }
    }
  }
  __syncthreads();
}

// ---- dense-accumulator schedule ------------------------------------------------------------------
__global__ void __launch_bounds__(bm25::THREADS)
bm25_score_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ post_doc,
                  const int32_t* __restrict__ post_tf, const double* __restrict__ post_val,
                  const double* __restrict__ idf,
                  const double* __restrict__ doc_norm, double k1p1, const int32_t* __restrict__ q_terms,
                  const int32_t* __restrict__ q_len, int tmax, int64_t n_docs, int q_base,
                  double* __restrict__ acc_all) {
  const int q = q_base + blockIdx.x;
  double* acc = acc_all + (size_t)blockIdx.x * n_docs;
  for (int64_t i = threadIdx.x; i < n_docs; i += blockDim.x) acc[i] = 0.0;
  __syncthreads();
  const int len = q_len[q];
  for (int t = 0; t < len && t < tmax; ++t) {
    const int term = q_terms[(size_t)q * tmax + t];
    if (term >= 0) {
      const int64_t lo = indptr[term], hi = indptr[term + 1];
      const double w = idf[term];
      for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const int d = post_doc[i];
        const double c = post_val ? __dmul_rn(w, post_val[i]) : bm25_term(w, (double)post_tf[i], k1p1, doc_norm[d]);
        acc[d] = __dadd_rn(acc[d], c);
      }
    }
    __syncthreads();   // orders this term's accumulator updates before the next term's
  }
}

__global__ void __launch_bounds__(bm25::THREADS)
bm25_topk_kernel(const double* __restrict__ acc_all, int64_t n_docs, int k, int q_base,
                 double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ TopkScratch S;
  const double* acc = acc_all + (size_t)blockIdx.x * n_docs;
  const int q = q_base + blockIdx.x;
  const int kk = (int64_t)k < n_docs ? k : (int)n_docs;
  block_topk_f64(acc, (int)n_docs, kk, 0u, S);
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool ok = j < kk;
    out_scores[(size_t)q * k + j] = ok ? ord_to_f64(S.win_ord[j]) : -INFINITY;
    out_ids[(size_t)q * k + j] = ok ? (int64_t)S.win_id[j] : -1;
  }
}

// Warp-cooperative lower bound: first index in [lo,hi) whose value is >= key (hi if none).  Every
// round probes 32 evenly spaced entries at once, so the dependent-load chain is log33(n) long.
__device__ __forceinline__ int64_t warp_lower_bound(const int32_t* __restrict__ p, int64_t lo, int64_t hi, int64_t key) {
  const int lane = threadIdx.x & 31;
  while (hi - lo > 32) {
    const int64_t step = (hi - lo + 32) / 33;
    const int64_t idx = lo + (int64_t)(lane + 1) * step - 1;
    const bool lt = idx < hi && (int64_t)p[idx] < key;
    const int c = __popc(__ballot_sync(FULL, lt));            // probes are monotone: the first c are < key
    const int64_t nhi = lo + (int64_t)(c + 1) * step - 1;
    if (c < 32 && nhi < hi) hi = nhi;                         // probe c is >= key: the answer is at or before it
    lo += (int64_t)c * step;                                  // probe c-1 is < key: the answer is after it
  }
  const int64_t idx = lo + lane;
  const bool lt = idx < hi && (int64_t)p[idx] < key;
  return lo + __popc(__ballot_sync(FULL, lt));
}

// ---- shared-memory range schedule ----------------------------------------------------------------
// grid = (n_ranges, nq).  cand_ord/cand_id: [nq, n_ranges, k] per-range winners (0 / 0xFFFFFFFF pad).
__global__ void __launch_bounds__(bm25::THREADS, 1)
bm25_range_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ post_doc,
                  const int32_t* __restrict__ post_tf, const double* __restrict__ post_val,
                  const double* __restrict__ idf,
                  const double* __restrict__ doc_norm, double k1p1, const int32_t* __restrict__ q_terms,
                  const int32_t* __restrict__ q_len, int tmax, int64_t n_docs, int k, int range_docs,
                  uint64_t* __restrict__ cand_ord, uint32_t* __restrict__ cand_id) {
  extern __shared__ double racc[];                 // [range_docs]
  __shared__ TopkScratch S;
  __shared__ int64_t span_lo[32], span_hi[32];
  __shared__ double span_w[32];
  const int range = blockIdx.x, n_ranges = gridDim.x, q = blockIdx.y;
  const int64_t d0 = (int64_t)range * range_docs;
  const int nd = (int)((n_docs - d0) < range_docs ? (n_docs - d0) : range_docs);
  for (int i = threadIdx.x; i < nd; i += blockDim.x) racc[i] = 0.0;
  const int len = q_len[q] < tmax ? q_len[q] : tmax;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int tb = 0; tb < len; tb += 32) {
    // the posting span of each of (up to) 32 terms that falls into this doc range: one warp per
    // (term, bound), 32-ary search (4 dependent loads for a 1M-entry list instead of 20)
    const int nt = len - tb < 32 ? len - tb : 32;
    for (int p = warp; p < 2 * nt; p += nwarps) {
      const int t = p >> 1, hi_bound = p & 1;
      const int term = q_terms[(size_t)q * tmax + tb + t];
      int64_t r = 0;
      if (term >= 0) r = warp_lower_bound(post_doc, indptr[term], indptr[term + 1], hi_bound ? d0 + nd : d0);
      if ((threadIdx.x & 31) == 0) {
        if (hi_bound) span_hi[t] = r;
        else { span_lo[t] = r; span_w[t] = term >= 0 ? idf[term] : 0.0; }
      }
    }
    __syncthreads();
    for (int t = 0; t < nt; ++t) {
      const int64_t a = span_lo[t], b = span_hi[t];
      const double w = span_w[t];
      // four postings per thread in flight: all loads first, then the shared-memory updates
      // (32-bit offsets inside the span; prefetching the next round across the term barrier was
      // tried and measured slower: 0.33 vs 0.29 ms at C2)
      const int n_span = (int)(b - a);
      const int32_t* pd = post_doc + a;
      for (int i0 = threadIdx.x; i0 < n_span; i0 += 4 * (int)blockDim.x) {
        int l[4]; double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * (int)blockDim.x;
          l[u] = -1; v[u] = 0.0;
          if (i < n_span) {
            const int d = pd[i];
            l[u] = (int)(d - d0);
            v[u] = post_val ? post_val[a + i]
                            : __ddiv_rn(__dmul_rn((double)post_tf[a + i], k1p1), __dadd_rn((double)post_tf[a + i], doc_norm[d]));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (l[u] >= 0) racc[l[u]] = __dadd_rn(racc[l[u]], __dmul_rn(w, v[u]));
      }
      __syncthreads();   // term order
    }
  }
  __syncthreads();
  const int kk = k < nd ? k : nd;
  block_topk_f64(racc, nd, kk, (uint32_t)d0, S);
  const size_t base = ((size_t)q * n_ranges + range) * k;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    cand_ord[base + j] = j < kk ? S.win_ord[j] : 0ull;
    cand_id[base + j] = j < kk ? S.win_id[j] : 0xFFFFFFFFu;
  }
}

// bitonic sort of P (power of two) candidates in shared memory by (score desc, id asc); all threads
__device__ __forceinline__ void sort_candidates(uint64_t* mo, uint32_t* mi, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t ao = mo[lo], bo = mo[hi];
        const uint32_t ai = mi[lo], bi = mi[hi];
        const bool a_less = (ao < bo) || (ao == bo && ai > bi);
        if (a_less == desc) { mo[lo] = bo; mo[hi] = ao; mi[lo] = bi; mi[hi] = ai; }
      }
    }
  }
  __syncthreads();
}

// one CTA per query: sort the n_ranges*k candidates by (score desc, id asc), emit the first k
__global__ void bm25_merge_kernel(const uint64_t* __restrict__ cand_ord, const uint32_t* __restrict__ cand_id,
                                  int n_ranges, int k, int64_t n_docs, int P, double* __restrict__ out_scores,
                                  int64_t* __restrict__ out_ids) {
  extern __shared__ uint64_t mo[];                  // [P] ords then [P] ids (uint32)
  uint32_t* mi = (uint32_t*)(mo + P);
  const int q = blockIdx.x, total = n_ranges * k;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    mo[i] = i < total ? cand_ord[(size_t)q * total + i] : 0ull;
    mi[i] = i < total ? cand_id[(size_t)q * total + i] : 0xFFFFFFFFu;
  }
  sort_candidates(mo, mi, P);
  const int kk = (int64_t)k < n_docs ? k : (int)n_docs;
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool ok = j < kk && mi[j] != 0xFFFFFFFFu;
    out_scores[(size_t)q * k + j] = ok ? ord_to_f64(mo[j]) : -INFINITY;
    out_ids[(size_t)q * k + j] = ok ? (int64_t)mi[j] : -1;
  }
}

// Doc-range sharded BM25: merge the per-shard results [n_lists, nq, k_in] (fp64 scores, global doc
// ids, -1 = padding) of each query into its global top-k_out, same order rule.
__global__ void bm25_merge_shards_kernel(const double* __restrict__ scores, const int64_t* __restrict__ ids,
                                         int n_lists, int nq, int k_in, int k_out, int P,
                                         double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  extern __shared__ uint64_t mo[];
  uint32_t* mi = (uint32_t*)(mo + P);
  const int q = blockIdx.x, total = n_lists * k_in;
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    uint64_t o = 0ull; uint32_t id = 0xFFFFFFFFu;
    if (i < total) {
      const int g = i / k_in, j = i - g * k_in;
      const size_t src = ((size_t)g * nq + q) * k_in + j;
      const int64_t d = ids[src];
      if (d >= 0) { o = f64_to_ord(scores[src]); id = (uint32_t)d; }
    }
    mo[i] = o; mi[i] = id;
  }
  sort_candidates(mo, mi, P);
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    const bool ok = j < P && mi[j] != 0xFFFFFFFFu;
    out_scores[(size_t)q * k_out + j] = ok ? ord_to_f64(mo[j]) : -INFINITY;
    out_ids[(size_t)q * k_out + j] = ok ? (int64_t)mi[j] : -1;
  }
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

size_t ragarc_bm25_workspace_bytes(int64_t n_docs, int nq) {
  // dense schedule: one fp64 accumulator row per concurrently scored query, capped near 512 MB;
  // range schedule: 12 bytes per (query, range, k<=1024) candidate - covered by the same bound
  if (n_docs <= 0 || nq <= 0) return 256;
  size_t row = (size_t)n_docs * 8;
  size_t rows = (size_t)nq;
  size_t cap_rows = (size_t)(512ull << 20) / row;
  if (cap_rows < 1) cap_rows = 1;
  if (rows > cap_rows) rows = cap_rows;
  size_t dense = rows * row;
  size_t n_ranges = (size_t)ceil_div(n_docs, bm25::RANGE / 2);       // the finer of the two geometries
  size_t cand = (size_t)nq * n_ranges * bm25::KMAX * 12 + 512;
  if (cand > (size_t)(1ull << 30)) cand = 0;       // range schedule not used then
  return align_up(dense > cand ? dense : cand, 256);
}

// Geometry of the range schedule: (docs per CTA, threads per CTA) -> number of ranges.
//  * half ranges (<= 12 288 docs = 96 KB of accumulators, 512 threads): TWO CTAs per SM, so one CTA's per-term
//    barriers and its selection overlap the other's posting loop - 0.259 ms against 0.289 ms at C2;
//  * full ranges (<= 24 576 docs, 1024 threads, one CTA per SM) when the half ranges would make more than
//    8192 candidates per query for the merge sort.
// Ranges are balanced (ceil(n / n_ranges) rounded up to 256 docs) so that no CTA is left with a sliver.
// RAGARC_BM25_RANGE / RAGARC_BM25_THREADS override it for experiments.
static int64_t bm25_range_geometry(int64_t n_docs, int k, int* range_docs, int* range_threads) {
  static const int env_range = getenv("RAGARC_BM25_RANGE") ? atoi(getenv("RAGARC_BM25_RANGE")) : 0;
  static const int env_threads = getenv("RAGARC_BM25_THREADS") ? atoi(getenv("RAGARC_BM25_THREADS")) : 0;
  int rmax = bm25::RANGE / 2, threads = bm25::THREADS / 2;
  if (ceil_div(n_docs, rmax) * (int64_t)k > 8192) { rmax = bm25::RANGE; threads = bm25::THREADS; }
  if (env_range >= 1024 && env_range <= bm25::RANGE) rmax = env_range / 256 * 256;
  if (env_threads >= 128 && env_threads <= bm25::THREADS) threads = env_threads / 32 * 32;
  const int64_t nr = ceil_div(n_docs > 0 ? n_docs : 1, rmax);
  int64_t r = ceil_div(ceil_div(n_docs > 0 ? n_docs : 1, nr), 256) * 256;
  if (r > rmax) r = rmax;
  *range_docs = (int)r; *range_threads = threads;
  return ceil_div(n_docs > 0 ? n_docs : 1, r);
}

static int bm25_check(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                      const double* idf, const double* doc_norm, const int32_t* q_terms,
                      const int32_t* q_len, int nq, int tmax, int64_t n_docs) {
  RA_REQUIRE(indptr && post_doc && post_tf && idf && doc_norm, "bm25: null index pointer");
  RA_REQUIRE(nq >= 0 && tmax >= 0 && n_docs > 0, "bm25: bad shape nq=%d tmax=%d n_docs=%lld", nq, tmax,
             (long long)n_docs);
  RA_REQUIRE(nq == 0 || (q_terms && q_len), "bm25: null query pointer");
  RA_REQUIRE(n_docs < 0x7FFFFFFFll, "bm25: n_docs must fit int32");
  return RAGARC_OK;
}

int ragarc_bm25_scores(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                       const double* post_val, const double* idf, const double* doc_norm, double k1_plus_1,
                       const int32_t* q_terms, const int32_t* q_len, int nq, int tmax, int64_t n_docs,
                       double* out_scores, void* stream) {
  int rc = bm25_check(indptr, post_doc, post_tf, idf, doc_norm, q_terms, q_len, nq, tmax, n_docs);
  if (rc) return rc;
  RA_REQUIRE(out_scores || nq == 0, "bm25_scores: null output");
  if (nq == 0) return RAGARC_OK;
  bm25_score_kernel<<<nq, bm25::THREADS, 0, (cudaStream_t)stream>>>(
      indptr, post_doc, post_tf, post_val, idf, doc_norm, k1_plus_1, q_terms, q_len, tmax, n_docs, 0, out_scores);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_bm25_topk(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                     const double* post_val, const double* idf, const double* doc_norm, double k1_plus_1, const int32_t* q_terms,
                     const int32_t* q_len, int nq, int tmax, int64_t n_docs, int k, double* out_scores,
                     int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = bm25_check(indptr, post_doc, post_tf, idf, doc_norm, q_terms, q_len, nq, tmax, n_docs);
  if (rc) return rc;
  RA_REQUIRE(k > 0 && k <= bm25::KMAX, "bm25_topk: k=%d must be in [1,%d]", k, bm25::KMAX);
  RA_REQUIRE(out_scores && out_ids, "bm25_topk: null outputs");
  if (nq == 0) return RAGARC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // shared-memory range schedule whenever its candidate buffer fits and the merge is one sort
  int range_docs, range_threads;
  const int64_t n_ranges = bm25_range_geometry(n_docs, k, &range_docs, &range_threads);
  int P = 32; while (P < n_ranges * k) P <<= 1;
  const size_t ord_bytes = align_up((size_t)nq * n_ranges * k * 8, 256);
  const size_t cand_bytes = ord_bytes + (size_t)nq * n_ranges * k * 4;
  static const bool force_dense = getenv("RAGARC_BM25_DENSE") != nullptr;
  if (!force_dense && P <= 8192 && n_ranges <= 65535 && nq <= 65535 && workspace && workspace_bytes >= cand_bytes) {
    uint64_t* cand_ord = (uint64_t*)workspace;
    uint32_t* cand_id = (uint32_t*)((char*)workspace + ord_bytes);
    RA_CUDA(cudaFuncSetAttribute(bm25_range_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bm25::RANGE_SMEM));
    bm25_range_kernel<<<dim3((unsigned)n_ranges, (unsigned)nq), range_threads, (size_t)range_docs * 8, st>>>(
        indptr, post_doc, post_tf, post_val, idf, doc_norm, k1_plus_1, q_terms, q_len, tmax, n_docs, k, range_docs,
        cand_ord, cand_id);
    RA_LAUNCH_CHECK();
    const size_t msm = (size_t)P * 12;
    if (msm > 48 * 1024)
      RA_CUDA(cudaFuncSetAttribute(bm25_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
    int mthreads = P / 2 < 1024 ? P / 2 : 1024;
    if (mthreads < 32) mthreads = 32;
    bm25_merge_kernel<<<nq, mthreads, msm, st>>>(cand_ord, cand_id, (int)n_ranges, k, n_docs, P, out_scores, out_ids);
    RA_LAUNCH_CHECK();
    return RAGARC_OK;
  }
  const size_t row = (size_t)n_docs * 8;
  RA_REQUIRE(workspace && workspace_bytes >= row, "bm25_topk: workspace %zu < one accumulator row %zu",
             workspace_bytes, row);
  int chunk = (int)(workspace_bytes / row < (size_t)nq ? workspace_bytes / row : (size_t)nq);
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = nq - q0 < chunk ? nq - q0 : chunk;
    bm25_score_kernel<<<c, bm25::THREADS, 0, st>>>(indptr, post_doc, post_tf, post_val, idf, doc_norm, k1_plus_1,
                                                   q_terms, q_len, tmax, n_docs, q0, (double*)workspace);
    RA_LAUNCH_CHECK();
    bm25_topk_kernel<<<c, bm25::THREADS, 0, st>>>((const double*)workspace, n_docs, k, q0, out_scores, out_ids);
    RA_LAUNCH_CHECK();
  }
  return RAGARC_OK;
}

int ragarc_bm25_merge_topk(const double* scores, const int64_t* ids, int n_lists, int nq, int k_in, int k_out,
                           double* out_scores, int64_t* out_ids, void* stream) {
  RA_REQUIRE(n_lists >= 1 && nq >= 0 && k_in >= 1 && k_out >= 1, "bm25_merge_topk: bad shape");
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(scores && ids && out_scores && out_ids, "bm25_merge_topk: null pointer");
  const int64_t total = (int64_t)n_lists * k_in;
  RA_REQUIRE(total <= 8192, "bm25_merge_topk: n_lists*k_in=%lld > 8192", (long long)total);
  int P = 32;
  while (P < total) P <<= 1;
  const size_t msm = (size_t)P * 12;
  if (msm > 48 * 1024)
    RA_CUDA(cudaFuncSetAttribute(bm25_merge_shards_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
  int threads = P / 2 < 1024 ? P / 2 : 1024;
  if (threads < 32) threads = 32;
  bm25_merge_shards_kernel<<<nq, threads, msm, (cudaStream_t)stream>>>(scores, ids, n_lists, nq, k_in, k_out, P,
                                                                      out_scores, out_ids);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

}  // extern "C"
