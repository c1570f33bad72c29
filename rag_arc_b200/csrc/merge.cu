// Final selection: merge the per-slice candidate lists of one query (or the all-gathered
// per-GPU top-k lists) into one sorted top-k.  One CTA per query; the keys (64-bit, larger =
// better, unique) are bitonic-sorted in shared memory, so the result is deterministic: score
// descending, ties by ascending row id, independent of how the corpus was tiled or sharded.
#include "common.cuh"

namespace ragarc {

__device__ __forceinline__ void block_bitonic_desc(uint64_t* s, int P) {
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool desc = (lo & size) == 0;
        uint64_t a = s[lo], b = s[hi];
        if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void emit_topk(const uint64_t* s, int q, int k, uint64_t id_base,
                                          uint64_t* out_keys, float* out_scores, int64_t* out_ids) {
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    uint64_t key = s[j];
    if (out_keys) {
      uint64_t gk = 0;
      if (key) gk = (key & 0xFFFFFFFF00000000ull) | uint64_t(0xFFFFFFFFu - (uint32_t)(key_row(key) + id_base));
      out_keys[(size_t)q * k + j] = gk;
    }
    if (out_scores) out_scores[(size_t)q * k + j] = key ? key_score(key) : -INFINITY;
    if (out_ids) out_ids[(size_t)q * k + j] = key ? (int64_t)(key_row(key) + id_base) : -1;
  }
}

// lists[(slice*MB + qb) * rows + r][cap], counts[(slice*MB + qb) * rows + r]
__global__ void merge_lists_kernel(const uint64_t* __restrict__ lists, const int* __restrict__ counts,
                                   int MB, int S, int rows, int cap, int k, int P, uint64_t id_base,
                                   uint64_t* out_keys, float* out_scores, int64_t* out_ids) {
  extern __shared__ uint64_t skeys[];
  __shared__ int offs[1025];
  const int q = blockIdx.x;
  const int qb = q / rows, r = q % rows;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < S; ++s) {
      offs[s] = acc;
      int c = counts[((size_t)s * MB + qb) * rows + r];
      acc += c < k ? c : k;       // lists are pruned to <= k at item end
    }
    offs[S] = acc;
  }
  __syncthreads();
  const int total = offs[S];
  for (int s = 0; s < S; ++s) {
    const size_t li = ((size_t)s * MB + qb) * rows + r;
    const int c = offs[s + 1] - offs[s];
    const uint64_t* src = lists + li * (size_t)cap;
    for (int j = threadIdx.x; j < c; j += blockDim.x) skeys[offs[s] + j] = src[j];
  }
  for (int j = total + threadIdx.x; j < P; j += blockDim.x) skeys[j] = 0;
  block_bitonic_desc(skeys, P);
  emit_topk(skeys, q, k, id_base, out_keys, out_scores, out_ids);
}

// keys[g][q][k_in] -> top k_out
__global__ void merge_keys_kernel(const uint64_t* __restrict__ keys, int G, int nq, int k_in, int k_out,
                                  int P, float* out_scores, int64_t* out_ids) {
  extern __shared__ uint64_t skeys[];
  const int q = blockIdx.x;
  const int total = G * k_in;
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
    uint64_t v = 0;
    if (j < total) {
      int g = j / k_in, i = j % k_in;
      v = keys[((size_t)g * nq + q) * k_in + i];
    }
    skeys[j] = v;
  }
  block_bitonic_desc(skeys, P);
  emit_topk(skeys, q, k_out, 0, nullptr, out_scores, out_ids);
}

static int next_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

int launch_merge_lists(const uint64_t* lists, const int* counts, const DensePlan& pl, int nq, int k,
                       uint64_t id_base, uint64_t* out_keys, float* out_scores, int64_t* out_ids,
                       cudaStream_t stream) {
  int P = next_pow2(pl.S * k > k ? pl.S * k : k);
  RA_REQUIRE(P <= 8192 && pl.S <= 1024, "merge: S*k=%d too large", pl.S * k);
  int threads = P / 2 < 1024 ? P / 2 : 1024;
  if (threads < 32) threads = 32;
  size_t smem = (size_t)P * 8;
  if (smem > 48 * 1024)
    RA_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  merge_lists_kernel<<<nq, threads, smem, stream>>>(lists, counts, pl.MB, pl.S, pl.rows_per_item, pl.cap,
                                                   k, P, id_base, out_keys, out_scores, out_ids);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

}  // namespace ragarc

using namespace ragarc;

extern "C" int ragarc_merge_topk_keys(const uint64_t* keys, int nlists, int nq, int k_in, int k_out,
                                      float* out_scores, int64_t* out_ids, void* stream) {
  RA_REQUIRE(keys && out_scores && out_ids, "merge_topk_keys: null pointer");
  RA_REQUIRE(nlists > 0 && nq >= 0 && k_in > 0 && k_out > 0 && k_out <= nlists * k_in,
             "merge_topk_keys: bad shape G=%d nq=%d k_in=%d k_out=%d", nlists, nq, k_in, k_out);
  if (nq == 0) return RAGARC_OK;
  int P = next_pow2(nlists * k_in);
  RA_REQUIRE(P <= 8192, "merge_topk_keys: nlists*k_in=%d exceeds 8192", nlists * k_in);
  int threads = P / 2 < 1024 ? P / 2 : 1024;
  if (threads < 32) threads = 32;
  size_t smem = (size_t)P * 8;
  if (smem > 48 * 1024)
    RA_CUDA(cudaFuncSetAttribute(merge_keys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  merge_keys_kernel<<<nq, threads, smem, (cudaStream_t)stream>>>(keys, nlists, nq, k_in, k_out, P,
                                                                 out_scores, out_ids);
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}
