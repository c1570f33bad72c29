// Dense scoring + fused per-query selection on the CUDA cores (fp32 FMA accumulate).
// This is the path for fp32 corpora (the reference's own storage type: FAISS IndexFlatIP holds
// fp32, VectorStore_Faiss.py:170) and for shapes the tensor-core path cannot take (d % 8 != 0).
// Same work decomposition and candidate-list epilogue as the tcgen05 kernel: a work item is
// (query block of 64, corpus slice); scores of a 64x64 tile live only in shared memory.
#include "common.cuh"

namespace ragarc {

namespace simt {
constexpr int BQ = 64, BN = 64, BK = 32, THREADS = 256, LD = 68, LDS_S = 65;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(simt::THREADS)
dense_simt_kernel(const T* __restrict__ X, int64_t n, int d, const T* __restrict__ Q, int nq, int k,
                  int MB, int S, int64_t tiles, int cap, int keep, uint64_t* __restrict__ lists,
                  int* __restrict__ counts, uint32_t* __restrict__ gthr) {
  using namespace simt;
  __shared__ __align__(16) float Qs[BK][LD];
  __shared__ __align__(16) float Xs[BK][LD];
  __shared__ float Ss[BQ][LDS_S];
  __shared__ uint32_t hist[2][256];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ty = tid >> 4, tx = tid & 15;
  const int lrow = tid >> 2, lk = (tid & 3) * 8;   // tile loader mapping: row, first k
  const int64_t items = (int64_t)MB * S;

  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int qb = (int)(item % MB);
    const int64_t s = item / MB;
    const int64_t t0 = s * tiles / S, t1 = (s + 1) * tiles / S;
    const int q0 = qb * BQ;

    RowState st;
    uint32_t* grow = nullptr;
    if (warp < 2) {
      const int qrow = q0 + tid;
      st.list = lists + ((size_t)item * BQ + tid) * (size_t)cap;
      st.cnt = 0;
      st.ord_local = 0;
      st.ord_global = 0;
      st.thr = (qrow < nq) ? -INFINITY : INFINITY;
      if (qrow < nq) grow = gthr + qrow;
    }

    for (int64_t t = t0; t < t1; ++t) {
      const int64_t r0 = t * BN;
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

      for (int k0 = 0; k0 < d; k0 += BK) {
        {
          const int qr = q0 + lrow;
          const int64_t xr = r0 + lrow;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int kk = k0 + lk + e;
            float qv = 0.f, xv = 0.f;
            if (kk < d) {
              if (qr < nq) qv = to_f32<T>(Q[(size_t)qr * d + kk]);
              if (xr < n) xv = to_f32<T>(X[(size_t)xr * d + kk]);
            }
            Qs[lk + e][lrow] = qv;
            Xs[lk + e][lrow] = xv;
          }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(&Qs[kk][ty * 4]);
          const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Ss[ty * 4 + i][tx * 4 + j] = acc[i][j];
      __syncthreads();

      if (warp < 2) {
        if (grow) {
          uint32_t g = *(volatile uint32_t*)grow;
          if (g > st.ord_global) { st.ord_global = g; st.thr = combine_thr(st.ord_local, g); }
        }
        const int64_t nvalid = n - r0;   // columns >= nvalid are padding
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
#pragma unroll 4
          for (int c = c0; c < c0 + 32; ++c) {
            const float v = Ss[tid][c];
            if (v > st.thr && c < nvalid) {
              st.list[st.cnt++] = make_key(v, (uint32_t)(r0 + c));
            }
          }
          prune_if_needed(st, k, cap, cap - 32, grow, hist[warp]);
        }
      }
      // Ss is rewritten only after the next tile's k-loop barriers, which warps 0/1 also reach.
    }
    if (warp < 2) {
      prune_if_needed(st, k, cap, keep, grow, hist[warp]);
      counts[(size_t)item * BQ + tid] = st.cnt;
    }
    __syncthreads();
  }
}

int launch_dense_simt(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                      int k, const DensePlan& pl, uint64_t* lists, int* counts, uint32_t* gthr,
                      cudaStream_t stream) {
  int64_t items = (int64_t)pl.MB * pl.S;
  int grid = (int)(items < (int64_t)sm_count() * 8 ? items : (int64_t)sm_count() * 8);
  if (grid < 1) grid = 1;
#define RA_SIMT_LAUNCH(T)                                                                        \
  dense_simt_kernel<T><<<grid, simt::THREADS, 0, stream>>>((const T*)corpus, n, d, (const T*)queries, \
      nq, k, pl.MB, pl.S, pl.tiles, pl.cap, pl.keep, lists, counts, gthr)
  if (dtype == RAGARC_F32) RA_SIMT_LAUNCH(float);
  else if (dtype == RAGARC_BF16) RA_SIMT_LAUNCH(__nv_bfloat16);
  else RA_SIMT_LAUNCH(__half);
#undef RA_SIMT_LAUNCH
  RA_LAUNCH_CHECK();
  return RAGARC_OK;
}

}  // namespace ragarc
