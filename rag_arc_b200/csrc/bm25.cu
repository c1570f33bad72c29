// BM25 (Okapi) scoring over CSR postings + exact top-k, fp64, bit-identical to the numpy
// expression the reference evaluates through rank_bm25 (core/retrieval/bm25.py:306):
//
//   score[d] += idf[t] * ( tf * (k1 + 1) / ( tf + k1 * (1 - b + b * dl[d] / avgdl) ) )
//
// One CTA owns one query's dense fp64 accumulator (n_docs doubles in the workspace, L2-resident
// at 100k docs).  Query terms are processed SEQUENTIALLY in query order (duplicates repeat),
// the postings of one term in parallel: a doc id occurs at most once per posting list, so no
// two threads touch the same accumulator inside a term and there are no atomics - that, plus
// the explicit round-to-nearest intrinsics (no FMA contraction), is what makes the fp64 sums
// reproduce numpy's bit for bit.  HBM/L2-bound gather-scatter: loads of (doc,tf) are coalesced
// and streamed, the accumulator traffic is random 8-byte read-modify-write.
//
// Selection: MSB-first radix select over the 96-bit composite key (orderable fp64 score, ~doc id)
// so that ties resolve to the lowest doc id, then a bitonic sort of the k winners.
#include "common.cuh"

namespace ragarc {

namespace bm25 {
constexpr int THREADS = 1024;
constexpr int KMAX = 1024;
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

__global__ void __launch_bounds__(bm25::THREADS)
bm25_score_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ post_doc,
                  const int32_t* __restrict__ post_tf, const double* __restrict__ idf,
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
        const double tf = (double)post_tf[i];
        const double c = __dmul_rn(w, __ddiv_rn(__dmul_rn(tf, k1p1), __dadd_rn(tf, doc_norm[d])));
        acc[d] = __dadd_rn(acc[d], c);
      }
    }
    __syncthreads();   // orders this term's accumulator updates before the next term's
  }
}

// block-wide exclusive scan helper over 256 bins held in shared memory: finds the bucket that
// contains the rem-th largest element.  Returns (digit, count above, count in digit).
__device__ __forceinline__ void find_bucket(const uint32_t* hist, uint32_t rem, uint32_t* out3) {
  // executed by warp 0 only; lane owns bins [8*lane, 8*lane+8)
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

__global__ void __launch_bounds__(bm25::THREADS)
bm25_topk_kernel(const double* __restrict__ acc_all, int64_t n_docs, int k, int q_base,
                 double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel[3];
  __shared__ uint64_t win_ord[bm25::KMAX];
  __shared__ uint32_t win_id[bm25::KMAX];
  __shared__ uint32_t nwin;
  const double* acc = acc_all + (size_t)blockIdx.x * n_docs;
  const int q = q_base + blockIdx.x;
  const int kk = (int64_t)k < n_docs ? k : (int)n_docs;

  __shared__ unsigned long long s_omin, s_omax;
  bool selected = false;
  if (threadIdx.x == 0) { s_omin = ~0ull; s_omax = 0ull; nwin = 0; }
  __syncthreads();
  if (kk > 0) {
    // Fast path (3 passes over the accumulator instead of up to 13): linear 256-bin histogram of the
    // scores over [min, max]; the bin holding the k-th largest score and everything above it
    // survive, provided that is at most KMAX documents.  Binning is monotone in the score, so all
    // ties of a survivor survive with it and the final (score, id) sort is exact.
    uint64_t lmin = ~0ull, lmax = 0ull;
    for (int64_t i = threadIdx.x; i < n_docs; i += blockDim.x) {
      const uint64_t o = f64_to_ord(acc[i]);
      lmin = o < lmin ? o : lmin; lmax = o > lmax ? o : lmax;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const uint64_t a = __shfl_xor_sync(FULL, lmin, off), b = __shfl_xor_sync(FULL, lmax, off);
      lmin = a < lmin ? a : lmin; lmax = b > lmax ? b : lmax;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&s_omin, (unsigned long long)lmin); atomicMax(&s_omax, (unsigned long long)lmax); }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const double mn = ord_to_f64(s_omin), mx = ord_to_f64(s_omax);
    if (mx > mn && isfinite(mx) && isfinite(mn)) {
      const double scale = 256.0 / (mx - mn);
      for (int64_t i = threadIdx.x; i < n_docs; i += blockDim.x) {
        int b = (int)((acc[i] - mn) * scale);
        b = b < 0 ? 0 : (b > 255 ? 255 : b);
        atomicAdd(&hist[b], 1u);
      }
      __syncthreads();
      if (threadIdx.x < 32) find_bucket(hist, (uint32_t)kk, sel);
      __syncthreads();
      const int bstar = (int)sel[0];
      if (sel[1] + sel[2] <= (uint32_t)bm25::KMAX) {
        for (int64_t i = threadIdx.x; i < n_docs; i += blockDim.x) {
          const double v = acc[i];
          int b = (int)((v - mn) * scale);
          b = b < 0 ? 0 : (b > 255 ? 255 : b);
          if (b >= bstar) {
            const uint32_t slot = atomicAdd(&nwin, 1u);
            win_ord[slot] = f64_to_ord(v); win_id[slot] = (uint32_t)i;
          }
        }
        selected = true;
      }
      __syncthreads();
    }
  }
  int have = kk;
  if (selected) {
    have = (int)nwin;                  // >= kk survivors; the sort below puts the kk best first
  } else {
  // composite key = (ord64(score), ~id): 12 digits of 8 bits, MSB first
  uint64_t pre_hi = 0, mask_hi = 0;
  uint32_t pre_lo = 0, mask_lo = 0;
  uint32_t rem = (uint32_t)kk;
  for (int pass = 0; pass < 12 && kk > 0; ++pass) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int64_t b = 0; b < n_docs; b += blockDim.x) {
      const int64_t i = b + threadIdx.x;
      const uint64_t o = i < n_docs ? f64_to_ord(acc[i]) : 0ull;
      const uint32_t nid = ~(uint32_t)i;
      const uint32_t dig = pass < 8 ? (uint32_t)(o >> (56 - 8 * pass)) & 0xFFu
                                    : (nid >> (24 - 8 * (pass - 8))) & 0xFFu;
      hist_add_agg(hist, dig, i < n_docs && (o & mask_hi) == pre_hi && (nid & mask_lo) == pre_lo);
    }
    __syncthreads();
    if (threadIdx.x < 32) find_bucket(hist, rem, sel);
    __syncthreads();
    const uint32_t D = sel[0], above = sel[1], hD = sel[2];
    rem -= above;
    if (pass < 8) { pre_hi |= (uint64_t)D << (56 - 8 * pass); mask_hi |= (uint64_t)0xFF << (56 - 8 * pass); }
    else { pre_lo |= D << (24 - 8 * (pass - 8)); mask_lo |= 0xFFu << (24 - 8 * (pass - 8)); }
    __syncthreads();
    if (hD == rem) break;
  }
  // collect: key >= prefix under the masks  -> exactly kk winners
  if (threadIdx.x == 0) nwin = 0;
  __syncthreads();
  for (int64_t i = threadIdx.x; i < n_docs; i += blockDim.x) {
    const uint64_t o = f64_to_ord(acc[i]);
    const uint32_t nid = ~(uint32_t)i;
    const uint64_t oh = o & mask_hi;
    const bool keep = kk > 0 && (oh > pre_hi || (oh == pre_hi && (nid & mask_lo) >= pre_lo));
    if (keep) {
      const uint32_t slot = atomicAdd(&nwin, 1u);
      if (slot < bm25::KMAX) { win_ord[slot] = o; win_id[slot] = (uint32_t)i; }
    }
  }
  __syncthreads();
  }
  // sort winners: descending score, ascending id.  P = pow2 >= kk, pad with (0, max id)
  int P = 1; while (P < have) P <<= 1;
  for (int i = have + threadIdx.x; i < P; i += blockDim.x) { win_ord[i] = 0; win_id[i] = 0xFFFFFFFFu; }
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t ao = win_ord[lo], bo = win_ord[hi];
        const uint32_t ai = win_id[lo], bi = win_id[hi];
        const bool a_less = (ao < bo) || (ao == bo && ai > bi);   // "a ranks below b"
        if (a_less == desc) { win_ord[lo] = bo; win_ord[hi] = ao; win_id[lo] = bi; win_id[hi] = ai; }
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const bool ok = j < kk;
    out_scores[(size_t)q * k + j] = ok ? ord_to_f64(win_ord[j]) : -INFINITY;
    out_ids[(size_t)q * k + j] = ok ? (int64_t)win_id[j] : -1;
  }
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

size_t ragarc_bm25_workspace_bytes(int64_t n_docs, int nq) {
  // one fp64 accumulator row per concurrently scored query; at least one row, at most nq,
  // capped so that the default workspace stays around 512 MB
  if (n_docs <= 0 || nq <= 0) return 256;
  size_t row = (size_t)n_docs * 8;
  size_t rows = (size_t)nq;
  size_t cap_rows = (size_t)(512ull << 20) / row;
  if (cap_rows < 1) cap_rows = 1;
  if (rows > cap_rows) rows = cap_rows;
  return align_up(rows * row, 256);
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
                       const double* idf, const double* doc_norm, double k1_plus_1,
                       const int32_t* q_terms, const int32_t* q_len, int nq, int tmax, int64_t n_docs,
                       double* out_scores, void* stream) {
  int rc = bm25_check(indptr, post_doc, post_tf, idf, doc_norm, q_terms, q_len, nq, tmax, n_docs);
  if (rc) return rc;
  RA_REQUIRE(out_scores || nq == 0, "bm25_scores: null output");
  if (nq == 0) return RAGARC_OK;
  bm25_score_kernel<<<nq, bm25::THREADS, 0, (cudaStream_t)stream>>>(
      indptr, post_doc, post_tf, idf, doc_norm, k1_plus_1, q_terms, q_len, tmax, n_docs, 0, out_scores);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_bm25_topk(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                     const double* idf, const double* doc_norm, double k1_plus_1, const int32_t* q_terms,
                     const int32_t* q_len, int nq, int tmax, int64_t n_docs, int k, double* out_scores,
                     int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = bm25_check(indptr, post_doc, post_tf, idf, doc_norm, q_terms, q_len, nq, tmax, n_docs);
  if (rc) return rc;
  RA_REQUIRE(k > 0 && k <= bm25::KMAX, "bm25_topk: k=%d must be in [1,%d]", k, bm25::KMAX);
  RA_REQUIRE(out_scores && out_ids, "bm25_topk: null outputs");
  if (nq == 0) return RAGARC_OK;
  const size_t row = (size_t)n_docs * 8;
  RA_REQUIRE(workspace && workspace_bytes >= row, "bm25_topk: workspace %zu < one accumulator row %zu",
             workspace_bytes, row);
  int chunk = (int)(workspace_bytes / row < (size_t)nq ? workspace_bytes / row : (size_t)nq);
  cudaStream_t st = (cudaStream_t)stream;
  for (int q0 = 0; q0 < nq; q0 += chunk) {
    const int c = nq - q0 < chunk ? nq - q0 : chunk;
    bm25_score_kernel<<<c, bm25::THREADS, 0, st>>>(indptr, post_doc, post_tf, idf, doc_norm, k1_plus_1,
                                                   q_terms, q_len, tmax, n_docs, q0, (double*)workspace);
    RA_LAUNCH_CHECK();
    bm25_topk_kernel<<<c, bm25::THREADS, 0, st>>>((const double*)workspace, n_docs, k, q0, out_scores, out_ids);
    RA_LAUNCH_CHECK();
  }
  return RAGARC_OK;
}

}  // extern "C"
