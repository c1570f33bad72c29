// Small kernels either side of the search: L2-normalise + cast (index add / query staging),
// pool + normalise of encoder outputs, reciprocal-rank fusion, greedy MMR selection.
#include "common.cuh"

namespace ragarc {

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

template <typename T> __device__ __forceinline__ float load_f32(const T* p);
template <> __device__ __forceinline__ float load_f32<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float load_f32<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float load_f32<__half>(const __half* p) { return __half2float(*p); }

template <typename T> __device__ __forceinline__ float load_f32_round(float v) { T r = from_f32<T>(v); return load_f32(&r); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// ---- normalize + cast: one warp per row (faiss.normalize_L2 semantics) ------------------------
template <typename T>
__global__ void normalize_cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n, int d,
                                      int normalize) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* s = src + (size_t)row * d;
  T* o = dst + (size_t)row * d;
  float scale = 1.0f;
  if (normalize) {
    float nr = 0.f;
    for (int i = lane; i < d; i += 32) { float v = s[i]; nr = fmaf(v, v, nr); }
    nr = warp_sum(nr);
    if (nr > 0.f) scale = __fdiv_rn(1.0f, __fsqrt_rn(nr));
  }
  for (int i = lane; i < d; i += 32) o[i] = from_f32<T>(__fmul_rn(s[i], scale));
}

// ---- squared-L2 search through the inner-product kernels -----------------------------------------
// ||q - x||^2 = ||q||^2 - 2 (q.x - ||x||^2 / 2): a row is stored as [x | h] with h = -||x||^2/2 and a
// query as [q | 1], so that the UNCHANGED scoring + selection kernels rank by q.x + h (largest first
// = nearest first) and a tiny epilogue turns the k kept values into distances.  For half-precision
// storage h is carried as three pieces h1 + h2 + h3 (each the rounding of what is left, as in
// normalize_split3) against three ones in the query, so it reaches the fp32 accumulator with ~2^-24
// relative error instead of the storage type's 2^-8 / 2^-11; ||x||^2 is taken over the STORED
// (rounded) values, i.e. distances are exact for what the index holds.  One warp per row.
template <typename T>
__global__ void l2_augment_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n, int d, int d_aug,
                                  int is_query, int normalize, float* __restrict__ sqnorm) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* s = src + (size_t)row * d;
  T* o = dst + (size_t)row * d_aug;
  float scale = 1.0f;
  if (normalize) {
    float nr = 0.f;
    for (int i = lane; i < d; i += 32) { float v = s[i]; nr = fmaf(v, v, nr); }
    nr = warp_sum(nr);
    if (nr > 0.f) scale = __fdiv_rn(1.0f, __fsqrt_rn(nr));
  }
  float xn = 0.f;
  for (int i = lane; i < d; i += 32) {
    const T t = from_f32<T>(__fmul_rn(s[i], scale));
    o[i] = t;
    const float v = load_f32(&t);
    xn = fmaf(v, v, xn);
  }
  xn = warp_sum(xn);
  if (lane == 0) {
    if (sqnorm) sqnorm[row] = xn;
    const float h = -0.5f * xn;
    if (sizeof(T) == 4) {
      o[d] = from_f32<T>(is_query ? 1.0f : h);
    } else {
      const T h1 = from_f32<T>(h);
      const float r1 = h - load_f32(&h1);
      const T h2 = from_f32<T>(r1);
      const float r2 = r1 - load_f32(&h2);
      const T one = from_f32<T>(1.0f), zero = from_f32<T>(0.0f);
      o[d] = is_query ? one : h1; o[d + 1] = is_query ? one : h2; o[d + 2] = is_query ? one : from_f32<T>(r2);
      for (int i = d + 3; i < d_aug; ++i) o[i] = zero;
    }
  }
}

// scores[q][j] (= q.x - ||x||^2/2 of the j-th best row, descending) -> squared distances (ascending);
// ||q||^2 is recomputed from the stored query values.  FAISS clamps its BLAS path at 0 as well.
template <typename T>
__global__ void l2_finish_kernel(float* __restrict__ scores, const T* __restrict__ q_aug, int nq, int k, int d, int d_aug) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  const int lane = threadIdx.x & 31;
  const T* s = q_aug + (size_t)q * d_aug;
  float qn = 0.f;
  for (int i = lane; i < d; i += 32) { const float v = load_f32(s + i); qn = fmaf(v, v, qn); }
  qn = warp_sum(qn);
  for (int j = lane; j < k; j += 32) {
    const float sc = scores[(size_t)q * k + j];
    scores[(size_t)q * k + j] = sc == -INFINITY ? INFINITY : fmaxf(0.0f, fmaf(-2.0f, sc, qn));
  }
}

// ---- normalize + split into three bf16 planes: v = b1 + b2 + b3 up to 2^-24 relative ------------
__global__ void normalize_split3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n,
                                        int d, int normalize) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  const float* s = src + (size_t)row * d;
  __nv_bfloat16* o = dst + (size_t)row * 3 * d;
  float scale = 1.0f;
  if (normalize) {
    float nr = 0.f;
    for (int i = lane; i < d; i += 32) { float v = s[i]; nr = fmaf(v, v, nr); }
    nr = warp_sum(nr);
    if (nr > 0.f) scale = __fdiv_rn(1.0f, __fsqrt_rn(nr));
  }
  for (int i = lane; i < d; i += 32) {
    const float v = __fmul_rn(s[i], scale);
    const __nv_bfloat16 b1 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(b1);          // exact
    const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(b2);         // exact
    o[i] = b1; o[d + i] = b2; o[2 * d + i] = __float2bfloat16_rn(r2);
  }
}

// ---- pool + normalize: one CTA per sequence ---------------------------------------------------
// HBM-bound: every kept token row is read once.  The CTA is a (TS token-slices) x (column groups)
// grid of threads; a thread owns VEC consecutive columns (one 16-byte load per token) and sums the
// tokens t = slice, slice+TS, ... of its slice with several loads in flight; masked tokens are
// never read.  Slices are combined through shared memory in a fixed order (deterministic).
template <typename T> struct PoolVec;
template <> struct PoolVec<float> { static constexpr int VEC = 4; };
template <> struct PoolVec<__nv_bfloat16> { static constexpr int VEC = 8; };
template <> struct PoolVec<__half> { static constexpr int VEC = 8; };

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&o)[VEC]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < VEC; ++i) o[i] = load_f32<T>(e + i);
}

template <typename T>
__global__ void pool_normalize_kernel(const T* __restrict__ x, const int32_t* __restrict__ mask, int T_len,
                                      int H, int mode, int normalize, int TS, float* __restrict__ out) {
  constexpr int VEC = PoolVec<T>::VEC;
  extern __shared__ float psm[];          // [TS][Hpad] partial sums, then [nwarp] norm partials
  __shared__ float s_cnt, s_norm;
  __shared__ int s_last;
  const int b = blockIdx.x;
  const T* xb = x + (size_t)b * T_len * H;
  const int32_t* mb = mask + (size_t)b * T_len;
  const int groups = (H + VEC - 1) / VEC;
  const int Hpad = groups * VEC;
  if (threadIdx.x < 32) {
    float c = 0.f; int last = -1;
    for (int t = threadIdx.x; t < T_len; t += 32) if (mb[t] != 0) { c += 1.f; last = t; }
    c = warp_sum(c);
#pragma unroll
    for (int o = 16; o; o >>= 1) last = max(last, __shfl_xor_sync(FULL, last, o));
    if (threadIdx.x == 0) { s_cnt = c; s_last = last; }
  }
  __syncthreads();
  const int g = threadIdx.x % groups, slice = threadIdx.x / groups;
  const bool vec_ok = (H % VEC == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (slice < TS) {
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    const int h0 = g * VEC;
    if (mode == RAGARC_POOL_MEAN) {
      for (int t = slice; t < T_len; t += TS) {
        if (mb[t] == 0) continue;
        const T* row = xb + (size_t)t * H + h0;
        if (vec_ok) {
          float v[VEC];
          load_vec<T, VEC>(row, v);
#pragma unroll
          for (int i = 0; i < VEC; ++i) acc[i] += v[i];
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) if (h0 + i < H) acc[i] += load_f32<T>(row + i);
        }
      }
    } else if (slice == 0) {
      const int t = mode == RAGARC_POOL_CLS ? 0 : (s_last < 0 ? 0 : s_last);
      const T* row = xb + (size_t)t * H + h0;
#pragma unroll
      for (int i = 0; i < VEC; ++i) if (h0 + i < H) acc[i] = load_f32<T>(row + i);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) psm[slice * Hpad + h0 + i] = acc[i];
  }
  __syncthreads();
  // combine slices in order, divide, square-sum
  const float cnt = s_cnt;
  float sq = 0.f;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    float v = 0.f;
    if (mode == RAGARC_POOL_MEAN) {
      for (int sl = 0; sl < TS; ++sl) v += psm[sl * Hpad + h];
      v = __fdiv_rn(v, fmaxf(cnt, 1e-9f));
    } else {
      v = psm[h];
    }
    out[(size_t)b * H + h] = v;
    sq = fmaf(v, v, sq);
  }
  if (!normalize) return;
  __syncthreads();
  sq = warp_sum(sq);
  float* red = psm;                       // reuse
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    s_norm = fmaxf(__fsqrt_rn(tot), 1e-12f);
  }
  __syncthreads();
  const float den = s_norm;
  for (int h = threadIdx.x; h < H; h += blockDim.x)
    out[(size_t)b * H + h] = __fdiv_rn(out[(size_t)b * H + h], den);
}

// ---- reciprocal-rank fusion: one CTA per query ------------------------------------------------
// Positions p = l*kl + i walk the lists in the reference's order (Fusion.py:55-61).  The first
// position holding a key owns its score; contributions are added in position order (fp64, each
// 1.0/(k+rank) correctly rounded) so the sums equal Python's.  Output order = Python's stable
// sorted(reverse=True): descending score, ties by first appearance.
//
// rows != nullptr (ragarc_rrf_fuse_rows): the lists arrive as corpus ROWS of each retriever; the content key
// of a row comes from that retriever's row -> key table on the way in, and every fused key leaves with the
// (list, row) of the Document the reference would hand back for it: the one at the LAST position holding
// the key (document_map[content] is overwritten while walking the lists in order, Fusion.py:61).
constexpr int RRF_MAX_LISTS = 8;
struct RrfRows {
  const int64_t* rows[RRF_MAX_LISTS];       // [nq, kl_each[l]] corpus rows of list l, -1 = padding
  const int32_t* row_to_key[RRF_MAX_LISTS]; // [corpus rows of list l] content key of a row
  int kl_each[RRF_MAX_LISTS];
  int32_t* out_list;                        // [nq, top_k]
  int64_t* out_row;                         // [nq, top_k]
};

__global__ void rrf_fuse_kernel(const int32_t* __restrict__ ids, int L, int nq, int kl, double rrf_k,
                                int top_k, int32_t* __restrict__ out_ids, double* __restrict__ out_scores,
                                int32_t* __restrict__ out_count, const RrfRows rr, const bool by_rows) {
  extern __shared__ unsigned char rrf_smem[];
  const int n = L * kl;
  double* score = (double*)rrf_smem;                    // [n], valid for owners (< 0: not an owner)
  double* recip = score + n;                            // [kl]: 1/(k + rank), each correctly rounded once
  int32_t* key = (int32_t*)(recip + kl);                // [n]
  __shared__ int n_owner;
  const int q = blockIdx.x;
  if (threadIdx.x == 0) n_owner = 0;
  for (int i = threadIdx.x; i < kl; i += blockDim.x) recip[i] = __ddiv_rn(1.0, __dadd_rn(rrf_k, (double)(i + 1)));
  if (by_rows) {
    for (int l = 0; l < L; ++l) {
      const int kle = rr.kl_each[l];
      for (int i = threadIdx.x; i < kl; i += blockDim.x) {
        int64_t row = -1;
        if (i < kle && rr.rows[l]) row = rr.rows[l][(size_t)q * kle + i];
        key[l * kl + i] = row >= 0 ? rr.row_to_key[l][row] : -1;
      }
    }
  } else {
    for (int l = 0; l < L; ++l)
      for (int i = threadIdx.x; i < kl; i += blockDim.x) key[l * kl + i] = ids[((size_t)l * nq + q) * kl + i];
  }
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const int32_t me = key[p];
    double s = -1.0;
    if (me >= 0) {
      bool first = true;
      for (int j = 0; j < p; ++j) first = first && (key[j] != me);
      if (first) {
        s = 0.0;
        int i = p % kl;                                 // rank-1 of position j inside its list
        for (int j = p; j < n; ++j) {
          if (key[j] == me) s = __dadd_rn(s, recip[i]);
          if (++i == kl) i = 0;
        }
        atomicAdd(&n_owner, 1);
      }
    }
    score[p] = s;
  }
  __syncthreads();
  for (int p = threadIdx.x; p < n; p += blockDim.x) {
    const double s = score[p];
    if (s < 0.0) continue;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double t = score[j];                        // non-owners hold -1 and never outrank s >= 0
      rank += (t > s || (t == s && j < p)) ? 1 : 0;
    }
    if (rank < top_k) {
      out_ids[(size_t)q * top_k + rank] = key[p];
      out_scores[(size_t)q * top_k + rank] = s;
      if (by_rows) {
        int last = p;
        for (int j = n - 1; j > p; --j)
          if (key[j] == key[p]) { last = j; break; }
        const int l = last / kl;
        rr.out_list[(size_t)q * top_k + rank] = l;
        rr.out_row[(size_t)q * top_k + rank] = rr.rows[l][(size_t)q * rr.kl_each[l] + (last - l * kl)];
      }
    }
  }
  const int cnt = n_owner < top_k ? n_owner : top_k;
  for (int j = cnt + threadIdx.x; j < top_k; j += blockDim.x) {
    out_ids[(size_t)q * top_k + j] = -1;
    out_scores[(size_t)q * top_k + j] = 0.0;
    if (by_rows) { rr.out_list[(size_t)q * top_k + j] = -1; rr.out_row[(size_t)q * top_k + j] = -1; }
  }
  if (threadIdx.x == 0) out_count[q] = cnt;
}

// ---- greedy MMR: one CTA per query ------------------------------------------------------------
// Restates _mmr_select (VectorStore_Faiss.py:16-62): first pick = candidate 0; then k-1 times pick
// argmax over the remaining of  lambda*<q,c> - (1-lambda)*max(0, max_sel <s,c>)  (first maximum wins,
// python max()).  Dots are fp64 over the stored rows.
template <typename T>
__global__ void mmr_select_kernel(const T* __restrict__ X, int d, const T* __restrict__ Q,
                                  const int64_t* __restrict__ cand, int fetch_k, int k, double lambda,
                                  int32_t* __restrict__ out_sel) {
  extern __shared__ double mmr_smem[];
  double* qsim = mmr_smem;                 // [fetch_k]
  double* maxsim = qsim + fetch_k;         // [fetch_k]
  double* red = maxsim + fetch_k;          // [fetch_k] scratch scores
  __shared__ int s_pick, s_nvalid;
  int* taken = (int*)(red + fetch_k);      // [fetch_k]
  const int q = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int64_t* c = cand + (size_t)q * fetch_k;
  if (threadIdx.x == 0) {
    int nv = 0;
    while (nv < fetch_k && c[nv] >= 0) ++nv;
    s_nvalid = nv;
  }
  __syncthreads();
  const int nv = s_nvalid;
  for (int i = warp; i < nv; i += nwarp) {
    const T* row = X + (size_t)c[i] * d;
    const T* qq = Q + (size_t)q * d;
    double acc = 0.0;
    for (int e = lane; e < d; e += 32) acc += (double)load_f32<T>(row + e) * (double)load_f32<T>(qq + e);
    acc = warp_sum_f64(acc);
    if (lane == 0) { qsim[i] = acc; maxsim[i] = 0.0; taken[i] = 0; }
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) out_sel[(size_t)q * k + j] = -1;
  __syncthreads();
  if (nv == 0) return;
  if (k >= nv) {   // reference returns all candidates in order
    for (int j = threadIdx.x; j < nv && j < k; j += blockDim.x) out_sel[(size_t)q * k + j] = j;
    return;
  }
  int pick = 0;
  for (int step = 0; step < k; ++step) {
    if (threadIdx.x == 0) { out_sel[(size_t)q * k + step] = pick; taken[pick] = 1; }
    __syncthreads();
    if (step == k - 1) break;
    const T* prow = X + (size_t)c[pick] * d;
    for (int i = warp; i < nv; i += nwarp) {
      if (taken[i]) continue;
      const T* row = X + (size_t)c[i] * d;
      double acc = 0.0;
      for (int e = lane; e < d; e += 32) acc += (double)load_f32<T>(row + e) * (double)load_f32<T>(prow + e);
      acc = warp_sum_f64(acc);
      if (lane == 0) {
        if (acc > maxsim[i]) maxsim[i] = acc;
        red[i] = lambda * qsim[i] - (1.0 - lambda) * maxsim[i];
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int best = -1; double bs = 0.0;
      for (int i = 0; i < nv; ++i) {
        if (taken[i]) continue;
        if (best < 0 || red[i] > bs) { best = i; bs = red[i]; }
      }
      s_pick = best;
    }
    __syncthreads();
    pick = s_pick;
  }
}

// ---- adjacent-row cosine distance (SemanticChunker): one warp per pair (i, i+1), fp64 ------------
template <typename T>
__global__ void adjacent_cosine_kernel(const T* __restrict__ x, int64_t n, int d, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= n - 1) return;
  const int lane = threadIdx.x & 31;
  const T* a = x + (size_t)i * d;
  const T* b = a + d;
  double dot = 0.0, na = 0.0, nb = 0.0;
  for (int j = lane; j < d; j += 32) {
    const double u = (double)a[j], v = (double)b[j];
    dot = __dadd_rn(dot, __dmul_rn(u, v));
    na = __dadd_rn(na, __dmul_rn(u, u));
    nb = __dadd_rn(nb, __dmul_rn(v, v));
  }
  dot = warp_sum_f64(dot); na = warp_sum_f64(na); nb = warp_sum_f64(nb);
  if (lane == 0) {
    double sim = __ddiv_rn(dot, __dmul_rn(__dsqrt_rn(na), __dsqrt_rn(nb)));
    if (!isfinite(sim)) sim = 0.0;                     // nan / inf -> 0, as the reference patches them
    out[i] = __dsub_rn(1.0, sim);
  }
}

// ---- reranker tail: P(yes) from the two answer-token logits of the last position ----------------
// torch: log_softmax over [false, true] in the logits' dtype (fp32 math, result rounded to the
// dtype), then exp (again rounded to the dtype).
template <typename T>
__global__ void yes_no_score_kernel(const T* __restrict__ logits, int B, int64_t row_stride, int true_id,
                                    int false_id, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float f = load_f32(logits + (size_t)b * row_stride + false_id);
  const float t = load_f32(logits + (size_t)b * row_stride + true_id);
  const float m = fmaxf(f, t);
  const float ls = load_f32_round<T>((t - m) - logf(expf(f - m) + expf(t - m)));   // torch's operation order
  out[b] = load_f32_round<T>(expf(ls));
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

int ragarc_normalize_cast(const float* src, void* dst, int64_t n, int d, int dst_dtype, int normalize,
                          void* stream) {
  RA_REQUIRE(n >= 0 && d > 0, "normalize_cast: bad shape n=%lld d=%d", (long long)n, d);
  if (n == 0) return RAGARC_OK;
  RA_REQUIRE(src && dst, "normalize_cast: null pointer");
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(n, warps);
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == RAGARC_F32) normalize_cast_kernel<float><<<grid, warps * 32, 0, st>>>(src, (float*)dst, n, d, normalize);
  else if (dst_dtype == RAGARC_BF16) normalize_cast_kernel<__nv_bfloat16><<<grid, warps * 32, 0, st>>>(src, (__nv_bfloat16*)dst, n, d, normalize);
  else if (dst_dtype == RAGARC_F16) normalize_cast_kernel<__half><<<grid, warps * 32, 0, st>>>(src, (__half*)dst, n, d, normalize);
  else { set_error("normalize_cast: bad dtype %d", dst_dtype); return RAGARC_ERR_INVALID; }
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_normalize_split3(const float* src, void* dst, int64_t n, int d, int normalize, void* stream) {
  RA_REQUIRE(n >= 0 && d > 0, "normalize_split3: bad shape n=%lld d=%d", (long long)n, d);
  if (n == 0) return RAGARC_OK;
  RA_REQUIRE(src && dst, "normalize_split3: null pointer");
  const int warps = 8;
  normalize_split3_kernel<<<(unsigned)ceil_div(n, warps), warps * 32, 0, (cudaStream_t)stream>>>(
      src, (__nv_bfloat16*)dst, n, d, normalize);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_l2_aug_dim(int d, int dtype) {
  if (d <= 0) return 0;
  return dtype == RAGARC_F32 ? d + 1 : (d + 3 + 7) / 8 * 8;
}

int ragarc_l2_augment(const float* src, void* dst, int64_t n, int d, int dst_dtype, int is_query, int normalize,
                      float* sqnorm_out, void* stream) {
  RA_REQUIRE(n >= 0 && d > 0, "l2_augment: bad shape n=%lld d=%d", (long long)n, d);
  RA_REQUIRE(dst_dtype == RAGARC_F32 || dst_dtype == RAGARC_BF16 || dst_dtype == RAGARC_F16, "l2_augment: bad dtype %d", dst_dtype);
  if (n == 0) return RAGARC_OK;
  RA_REQUIRE(src && dst, "l2_augment: null pointer");
  const int d_aug = ragarc_l2_aug_dim(d, dst_dtype);
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(n, warps);
  cudaStream_t st = (cudaStream_t)stream;
  if (dst_dtype == RAGARC_F32) l2_augment_kernel<float><<<grid, warps * 32, 0, st>>>(src, (float*)dst, n, d, d_aug, is_query, normalize, sqnorm_out);
  else if (dst_dtype == RAGARC_BF16) l2_augment_kernel<__nv_bfloat16><<<grid, warps * 32, 0, st>>>(src, (__nv_bfloat16*)dst, n, d, d_aug, is_query, normalize, sqnorm_out);
  else l2_augment_kernel<__half><<<grid, warps * 32, 0, st>>>(src, (__half*)dst, n, d, d_aug, is_query, normalize, sqnorm_out);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_l2_distances(float* scores, const void* queries_aug, int dtype, int nq, int k, int d, void* stream) {
  RA_REQUIRE(nq >= 0 && k > 0 && d > 0, "l2_distances: bad shape nq=%d k=%d d=%d", nq, k, d);
  RA_REQUIRE(dtype == RAGARC_F32 || dtype == RAGARC_BF16 || dtype == RAGARC_F16, "l2_distances: bad dtype %d", dtype);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(scores && queries_aug, "l2_distances: null pointer");
  const int d_aug = ragarc_l2_aug_dim(d, dtype);
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(nq, warps);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RAGARC_F32) l2_finish_kernel<float><<<grid, warps * 32, 0, st>>>(scores, (const float*)queries_aug, nq, k, d, d_aug);
  else if (dtype == RAGARC_BF16) l2_finish_kernel<__nv_bfloat16><<<grid, warps * 32, 0, st>>>(scores, (const __nv_bfloat16*)queries_aug, nq, k, d, d_aug);
  else l2_finish_kernel<__half><<<grid, warps * 32, 0, st>>>(scores, (const __half*)queries_aug, nq, k, d, d_aug);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_pool_normalize(const void* x, int dtype, const int32_t* mask, int B, int T, int H, int mode,
                          int normalize, float* out, void* stream) {
  RA_REQUIRE(B >= 0 && T > 0 && H > 0, "pool_normalize: bad shape B=%d T=%d H=%d", B, T, H);
  RA_REQUIRE(mode >= RAGARC_POOL_MEAN && mode <= RAGARC_POOL_LAST, "pool_normalize: bad mode %d", mode);
  if (B == 0) return RAGARC_OK;
  RA_REQUIRE(x && mask && out, "pool_normalize: null pointer");
  const int vec = dtype == RAGARC_F32 ? 4 : 8;
  const int groups = (H + vec - 1) / vec;
  RA_REQUIRE(groups <= 1024, "pool_normalize: H=%d too large", H);
  int TS = 1024 / groups;                 // token slices that fit a 1024-thread CTA
  if (TS > 16) TS = 16;
  if (TS > T) TS = T;
  if (TS < 1) TS = 1;
  int threads = ((groups * TS + 31) / 32) * 32;
  size_t smem = (size_t)TS * groups * vec * sizeof(float);
  if (smem < 32 * sizeof(float)) smem = 32 * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define RA_POOL(TT)                                                                                   \
  do {                                                                                                \
    if (smem > 48 * 1024)                                                                             \
      RA_CUDA(cudaFuncSetAttribute(pool_normalize_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    pool_normalize_kernel<TT><<<B, threads, smem, st>>>((const TT*)x, mask, T, H, mode, normalize, TS, out); \
  } while (0)
  if (dtype == RAGARC_F32) RA_POOL(float);
  else if (dtype == RAGARC_BF16) RA_POOL(__nv_bfloat16);
  else if (dtype == RAGARC_F16) RA_POOL(__half);
  else { set_error("pool_normalize: bad dtype %d", dtype); return RAGARC_ERR_INVALID; }
#undef RA_POOL
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_rrf_fuse(const int32_t* ids, int n_lists, int nq, int kl, double rrf_k, int top_k,
                    int32_t* out_ids, double* out_scores, int32_t* out_count, void* stream) {
  RA_REQUIRE(n_lists > 0 && nq >= 0 && kl > 0 && top_k > 0, "rrf_fuse: bad shape L=%d nq=%d kl=%d top_k=%d",
             n_lists, nq, kl, top_k);
  RA_REQUIRE((int64_t)n_lists * kl <= 4096, "rrf_fuse: L*kl=%d exceeds 4096", n_lists * kl);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(ids && out_ids && out_scores && out_count, "rrf_fuse: null pointer");
  const int n = n_lists * kl;
  size_t smem = (size_t)n * 8 + (size_t)kl * 8 + (size_t)n * 4;
  int threads = n <= 128 ? 128 : 256;
  if (smem > 48 * 1024) RA_CUDA(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rrf_fuse_kernel<<<nq, threads, smem, (cudaStream_t)stream>>>(ids, n_lists, nq, kl, rrf_k, top_k, out_ids,
                                                              out_scores, out_count, RrfRows{}, false);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_rrf_fuse_rows(const int64_t* const* rows, const int* kl_each, const int32_t* const* row_to_key,
                         int n_lists, int nq, int kl, double rrf_k, int top_k, int32_t* out_ids, double* out_scores,
                         int32_t* out_count, int32_t* out_list, int64_t* out_row, void* stream) {
  RA_REQUIRE(n_lists > 0 && n_lists <= RRF_MAX_LISTS && nq >= 0 && kl > 0 && top_k > 0,
             "rrf_fuse_rows: bad shape L=%d (max %d) nq=%d kl=%d top_k=%d", n_lists, RRF_MAX_LISTS, nq, kl, top_k);
  RA_REQUIRE((int64_t)n_lists * kl <= 4096, "rrf_fuse_rows: L*kl=%d exceeds 4096", n_lists * kl);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(rows && kl_each && row_to_key && out_ids && out_scores && out_count && out_list && out_row,
             "rrf_fuse_rows: null pointer");
  RrfRows rr{};
  for (int l = 0; l < n_lists; ++l) {
    RA_REQUIRE(kl_each[l] >= 0 && kl_each[l] <= kl, "rrf_fuse_rows: list %d has %d columns, kl=%d", l, kl_each[l], kl);
    RA_REQUIRE(kl_each[l] == 0 || rows[l] == nullptr || row_to_key[l] != nullptr, "rrf_fuse_rows: list %d has rows but no key table", l);
    rr.rows[l] = rows[l]; rr.row_to_key[l] = row_to_key[l]; rr.kl_each[l] = rows[l] ? kl_each[l] : 0;
  }
  rr.out_list = out_list; rr.out_row = out_row;
  const int n = n_lists * kl;
  size_t smem = (size_t)n * 8 + (size_t)kl * 8 + (size_t)n * 4;
  int threads = n <= 128 ? 128 : 256;
  if (smem > 48 * 1024) RA_CUDA(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rrf_fuse_kernel<<<nq, threads, smem, (cudaStream_t)stream>>>(nullptr, n_lists, nq, kl, rrf_k, top_k, out_ids,
                                                              out_scores, out_count, rr, true);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_mmr_select(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                      const int64_t* cand_rows, int fetch_k, int k, double lambda_mult, int32_t* out_sel,
                      void* stream) {
  RA_REQUIRE(n > 0 && d > 0 && nq >= 0 && fetch_k > 0 && k > 0, "mmr_select: bad shape");
  RA_REQUIRE(fetch_k <= 1024, "mmr_select: fetch_k=%d exceeds 1024", fetch_k);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(corpus && queries && cand_rows && out_sel, "mmr_select: null pointer");
  size_t smem = (size_t)fetch_k * (3 * 8 + 4);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RAGARC_F32) mmr_select_kernel<float><<<nq, 256, smem, st>>>((const float*)corpus, d, (const float*)queries, cand_rows, fetch_k, k, lambda_mult, out_sel);
  else if (dtype == RAGARC_BF16) mmr_select_kernel<__nv_bfloat16><<<nq, 256, smem, st>>>((const __nv_bfloat16*)corpus, d, (const __nv_bfloat16*)queries, cand_rows, fetch_k, k, lambda_mult, out_sel);
  else if (dtype == RAGARC_F16) mmr_select_kernel<__half><<<nq, 256, smem, st>>>((const __half*)corpus, d, (const __half*)queries, cand_rows, fetch_k, k, lambda_mult, out_sel);
  else { set_error("mmr_select: bad dtype %d", dtype); return RAGARC_ERR_INVALID; }
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_adjacent_cosine_distance(const void* x, int dtype, int64_t n, int d, double* out, void* stream) {
  RA_REQUIRE(n >= 0 && d > 0, "adjacent_cosine_distance: bad shape n=%lld d=%d", (long long)n, d);
  if (n < 2) return RAGARC_OK;
  RA_REQUIRE(x && out, "adjacent_cosine_distance: null pointer");
  const int warps = 8;
  const unsigned grid = (unsigned)ceil_div(n - 1, warps);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RAGARC_F32) adjacent_cosine_kernel<float><<<grid, warps * 32, 0, st>>>((const float*)x, n, d, out);
  else if (dtype == RAGARC_F64) adjacent_cosine_kernel<double><<<grid, warps * 32, 0, st>>>((const double*)x, n, d, out);
  else { set_error("adjacent_cosine_distance: dtype must be RAGARC_F32 or RAGARC_F64"); return RAGARC_ERR_INVALID; }
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

int ragarc_yes_no_score(const void* logits, int dtype, int B, int64_t row_stride, int vocab, int true_id, int false_id,
                        float* out, void* stream) {
  RA_REQUIRE(B >= 0 && vocab > 0 && row_stride >= vocab, "yes_no_score: bad shape B=%d vocab=%d stride=%lld", B, vocab,
             (long long)row_stride);
  RA_REQUIRE(true_id >= 0 && true_id < vocab && false_id >= 0 && false_id < vocab, "yes_no_score: token id out of range");
  if (B == 0) return RAGARC_OK;
  RA_REQUIRE(logits && out, "yes_no_score: null pointer");
  const unsigned grid = (unsigned)ceil_div(B, 128);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == RAGARC_F32) yes_no_score_kernel<float><<<grid, 128, 0, st>>>((const float*)logits, B, row_stride, true_id, false_id, out);
  else if (dtype == RAGARC_BF16) yes_no_score_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>((const __nv_bfloat16*)logits, B, row_stride, true_id, false_id, out);
  else if (dtype == RAGARC_F16) yes_no_score_kernel<__half><<<grid, 128, 0, st>>>((const __half*)logits, B, row_stride, true_id, false_id, out);
  else { set_error("yes_no_score: bad dtype %d", dtype); return RAGARC_ERR_INVALID; }
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

}  // extern "C"
