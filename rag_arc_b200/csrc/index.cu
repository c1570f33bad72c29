// ragarc_index_*: a flat exact index object behind the C ABI - the library-owned counterpart of
// faiss.IndexFlatIP as the reference drives it (VectorStore_Faiss.py:110-148 create, :169-178,202
// add, :258-263 search, :385-419 delete -> remove_ids).  It owns the row matrix in HBM (storage
// dtype, rows normalised at add time when the metric is cosine), grows it geometrically, keeps its
// own workspace and staging buffers, and accepts fp32 inputs and outputs either as device pointers
// or as plain HOST buffers - a host that is not PyTorch (or not Python) needs nothing else.
// All arithmetic is the stateless entry points' (ragarc_normalize_cast, ragarc_dense_topk).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>
#include "common.cuh"

struct ragarc_index {
  int d = 0, dtype = RAGARC_F32, metric = RAGARC_METRIC_IP, device = 0;
  int ds = 0;                     // stored row width: d, or ragarc_l2_aug_dim(d, dtype) for RAGARC_METRIC_L2
  void* rows = nullptr;           // [cap, ds] storage dtype
  int64_t n = 0, cap = 0;
  void* ws = nullptr;             // dense_topk workspace
  size_t ws_bytes = 0;
  void* stage = nullptr;          // device staging: fp32 inputs, prepared queries, results
  size_t stage_bytes = 0;
  cudaEvent_t last = nullptr;     // end of the previous call's GPU work (calls may come on different streams)
  std::mutex mu;
};

// Row-sharded flat index driven by ONE host process: shard g is a ragarc_index on devices[g] holding
// a contiguous range of global rows; a search fans the queries out to every shard on its own
// stream, gathers the per-shard packed keys onto shard 0's device and merges them there.
struct ragarc_sharded_index {
  int d = 0, dtype = RAGARC_F32, metric = RAGARC_METRIC_IP;
  std::vector<ragarc_index*> shard;
  std::vector<int64_t> base;             // global id of each shard's first row
  std::vector<cudaStream_t> stream;      // one per shard, on the shard's device
  std::vector<cudaEvent_t> done;
  std::vector<void*> keys;               // per shard: [nq,k] packed keys on the shard's device
  std::vector<size_t> keys_bytes;
  void* gather = nullptr;                // shard 0's device: [G,nq,k] keys | scores | ids
  size_t gather_bytes = 0;
  std::mutex mu;
};

namespace ragarc {

static size_t esize(int dtype) { return dtype == RAGARC_F32 ? 4 : 2; }

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    ok = cudaGetDevice(&prev) == cudaSuccess && (prev == dev || cudaSetDevice(dev) == cudaSuccess);
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int grow_buffer(void** buf, size_t* have, size_t need, cudaStream_t st) {
  if (*have >= need) return RAGARC_OK;
  (void)st;
  RA_CUDA(cudaDeviceSynchronize());            // nothing in flight may still use the old buffer
  if (*buf) RA_CUDA(cudaFree(*buf));
  *buf = nullptr; *have = 0;
  size_t want = need + need / 4;
  RA_CUDA(cudaMalloc(buf, want));
  *have = want;
  return RAGARC_OK;
}

static int reserve_rows(ragarc_index* ix, int64_t capacity, cudaStream_t st) {
  if (capacity <= ix->cap) return RAGARC_OK;
  int64_t cap = ix->cap > 0 ? ix->cap : 1024;
  while (cap < capacity) cap += cap / 2 + 1024;          // geometric growth: O(1) amortised copies
  if (capacity > ix->cap * 4) cap = capacity;            // bulk load: exact size, no slack
  void* fresh = nullptr;
  const size_t row_bytes = (size_t)ix->ds * esize(ix->dtype);
  RA_CUDA(cudaMalloc(&fresh, (size_t)cap * row_bytes));
  if (ix->n > 0) RA_CUDA(cudaMemcpyAsync(fresh, ix->rows, (size_t)ix->n * row_bytes, cudaMemcpyDeviceToDevice, st));
  RA_CUDA(cudaDeviceSynchronize());
  if (ix->rows) RA_CUDA(cudaFree(ix->rows));
  ix->rows = fresh;
  ix->cap = cap;
  return RAGARC_OK;
}

// calls on one index are ordered on the GPU even when they arrive on different streams
static int order_begin(ragarc_index* ix, cudaStream_t st) {
  RA_CUDA(cudaStreamWaitEvent(st, ix->last, 0));
  return RAGARC_OK;
}
static int order_end(ragarc_index* ix, cudaStream_t st) {
  RA_CUDA(cudaEventRecord(ix->last, st));
  return RAGARC_OK;
}

// rows of a new matrix gathered from an old one: dst[i] = src[map[i]]   (16-byte words when possible)
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                   const int64_t* __restrict__ map, int64_t n_out, int row_bytes) {
  const int64_t row = blockIdx.x;
  if (row >= n_out) return;
  const uint8_t* s = src + (size_t)map[row] * row_bytes;
  uint8_t* o = dst + (size_t)row * row_bytes;
  if ((row_bytes & 15) == 0) {
    for (int i = threadIdx.x; i < (row_bytes >> 4); i += blockDim.x)
      reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) o[i] = s[i];
  }
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

int ragarc_host_alloc(size_t bytes, void** out_host) {
  RA_REQUIRE(out_host != nullptr, "host_alloc: out is NULL");
  *out_host = nullptr;
  RA_CUDA(cudaHostAlloc(out_host, bytes > 0 ? bytes : 1, cudaHostAllocDefault));
  return RAGARC_OK;
}

int ragarc_host_free(void* host) {
  if (host) RA_CUDA(cudaFreeHost(host));
  return RAGARC_OK;
}

int ragarc_index_create(int d, int dtype, int metric, ragarc_index_t** out) {
  RA_REQUIRE(out != nullptr, "index_create: out is NULL");
  *out = nullptr;
  RA_REQUIRE(d > 0, "index_create: d=%d", d);
  RA_REQUIRE(dtype == RAGARC_F32 || dtype == RAGARC_BF16 || dtype == RAGARC_F16, "index_create: bad dtype %d", dtype);
  RA_REQUIRE(metric == RAGARC_METRIC_IP || metric == RAGARC_METRIC_COSINE || metric == RAGARC_METRIC_L2,
             "index_create: bad metric %d", metric);
  int dev = 0;
  RA_CUDA(cudaGetDevice(&dev));
  ragarc_index* ix = new (std::nothrow) ragarc_index();
  RA_REQUIRE(ix != nullptr, "index_create: out of host memory");
  ix->d = d; ix->dtype = dtype; ix->metric = metric; ix->device = dev;
  ix->ds = metric == RAGARC_METRIC_L2 ? ragarc_l2_aug_dim(d, dtype) : d;
  if (cudaEventCreateWithFlags(&ix->last, cudaEventDisableTiming) != cudaSuccess) {
    delete ix;
    set_error("index_create: cannot create a CUDA event: %s", cudaGetErrorString(cudaGetLastError()));
    return RAGARC_ERR_CUDA;
  }
  *out = ix;
  return RAGARC_OK;
}

int ragarc_index_free(ragarc_index_t* ix) {
  if (!ix) return RAGARC_OK;
  {
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    if (ix->rows) cudaFree(ix->rows);
    if (ix->ws) cudaFree(ix->ws);
    if (ix->stage) cudaFree(ix->stage);
    if (ix->last) cudaEventDestroy(ix->last);
  }
  delete ix;
  return RAGARC_OK;
}

int64_t ragarc_index_ntotal(const ragarc_index_t* ix) { return ix ? ix->n : -1; }
int ragarc_index_dim(const ragarc_index_t* ix) { return ix ? ix->d : -1; }
const void* ragarc_index_rows(const ragarc_index_t* ix) { return ix ? ix->rows : nullptr; }

int ragarc_index_reserve(ragarc_index_t* ix, int64_t capacity, void* stream) {
  RA_REQUIRE(ix != nullptr, "index_reserve: null index");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  RA_REQUIRE(g.ok, "index_reserve: cannot select device %d", ix->device);
  return reserve_rows(ix, capacity, (cudaStream_t)stream);
}

int ragarc_index_add(ragarc_index_t* ix, const float* rows, int64_t n, int rows_on_host, void* stream) {
  RA_REQUIRE(ix != nullptr, "index_add: null index");
  RA_REQUIRE(n >= 0, "index_add: n=%lld", (long long)n);
  if (n == 0) return RAGARC_OK;
  RA_REQUIRE(rows != nullptr, "index_add: null rows");
  RA_REQUIRE(ix->n + n < (int64_t)0xFFFFFFF0ll, "index_add: at most 2^32-16 rows per index");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  RA_REQUIRE(g.ok, "index_add: cannot select device %d", ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = order_begin(ix, st);
  if (rc) return rc;
  rc = reserve_rows(ix, ix->n + n, st);
  if (rc) return rc;
  const size_t row_bytes = (size_t)ix->ds * esize(ix->dtype);
  const int normalize = ix->metric == RAGARC_METRIC_COSINE;
  const bool l2 = ix->metric == RAGARC_METRIC_L2;
  if (!rows_on_host) {
    void* dst = (char*)ix->rows + (size_t)ix->n * row_bytes;
    rc = l2 ? ragarc_l2_augment(rows, dst, n, ix->d, ix->dtype, 0, 0, nullptr, stream)
            : ragarc_normalize_cast(rows, dst, n, ix->d, ix->dtype, normalize, stream);
    if (rc) return rc;
  } else {
    // host rows travel through a bounded device staging buffer (<= 256 MB of fp32 at a time)
    int64_t chunk = (int64_t)((256ull << 20) / ((size_t)ix->d * 4));
    if (chunk < 1) chunk = 1;
    if (chunk > n) chunk = n;
    rc = grow_buffer(&ix->stage, &ix->stage_bytes, (size_t)chunk * ix->d * 4, st);
    if (rc) return rc;
    for (int64_t r0 = 0; r0 < n; r0 += chunk) {
      const int64_t c = n - r0 < chunk ? n - r0 : chunk;
      RA_CUDA(cudaMemcpyAsync(ix->stage, rows + (size_t)r0 * ix->d, (size_t)c * ix->d * 4, cudaMemcpyHostToDevice, st));
      void* dst = (char*)ix->rows + (size_t)(ix->n + r0) * row_bytes;
      rc = l2 ? ragarc_l2_augment((const float*)ix->stage, dst, c, ix->d, ix->dtype, 0, 0, nullptr, stream)
              : ragarc_normalize_cast((const float*)ix->stage, dst, c, ix->d, ix->dtype, normalize, stream);
      if (rc) return rc;
    }
    RA_CUDA(cudaStreamSynchronize(st));        // the caller may reuse its host buffer on return
  }
  ix->n += n;
  return order_end(ix, st);
}

int ragarc_index_search(ragarc_index_t* ix, const float* queries, int nq, int k, float* out_scores,
                        int64_t* out_ids, int buffers_on_host, void* stream) {
  RA_REQUIRE(ix != nullptr, "index_search: null index");
  RA_REQUIRE(nq >= 0 && k > 0, "index_search: nq=%d k=%d", nq, k);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(queries && out_scores && out_ids, "index_search: null pointer");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  RA_REQUIRE(g.ok, "index_search: cannot select device %d", ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t q32_bytes = align_up((size_t)nq * ix->d * 4, 256);
  const size_t qs_bytes = align_up((size_t)nq * ix->ds * esize(ix->dtype), 256);
  const size_t os_bytes = align_up((size_t)nq * k * 4, 256);
  const size_t oi_bytes = align_up((size_t)nq * k * 8, 256);
  int rc = order_begin(ix, st);
  if (rc) return rc;
  rc = grow_buffer(&ix->stage, &ix->stage_bytes, q32_bytes + qs_bytes + os_bytes + oi_bytes, st);
  if (rc) return rc;
  const size_t need_ws = ragarc_dense_topk_workspace_bytes(ix->n, ix->ds, ix->dtype, nq, k);
  RA_REQUIRE(need_ws > 0, "index_search: unsupported shape (k=%d)", k);
  rc = grow_buffer(&ix->ws, &ix->ws_bytes, need_ws, st);
  if (rc) return rc;
  char* sg = (char*)ix->stage;
  float* q32 = (float*)sg;
  void* qs = sg + q32_bytes;
  float* d_scores = buffers_on_host ? (float*)(sg + q32_bytes + qs_bytes) : out_scores;
  int64_t* d_ids = buffers_on_host ? (int64_t*)(sg + q32_bytes + qs_bytes + os_bytes) : out_ids;
  const float* q_src = queries;
  static const bool trace = getenv("RAGARC_INDEX_TRACE") != nullptr;      // experiments: per-phase wall clock on stderr
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_a = 0, t_b = 0, t_c = 0, t_d = 0;
  if (trace) { cudaStreamSynchronize(st); t_a = now(); }
  if (buffers_on_host) {
    RA_CUDA(cudaMemcpyAsync(q32, queries, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, st));
    q_src = q32;
  }
  if (trace) { cudaStreamSynchronize(st); t_b = now(); }
  // faiss.normalize_L2 on the query when the metric is cosine (VectorStore_Faiss.py:259), then the
  // cast to the storage dtype; for fp32 storage without normalisation the queries are used in place
  const void* q_use = q_src;
  const bool l2 = ix->metric == RAGARC_METRIC_L2;
  if (l2) {
    rc = ragarc_l2_augment(q_src, qs, nq, ix->d, ix->dtype, 1, 0, nullptr, stream);     // [q | 1]
    if (rc) return rc;
    q_use = qs;
  } else if (ix->dtype != RAGARC_F32 || ix->metric == RAGARC_METRIC_COSINE) {
    rc = ragarc_normalize_cast(q_src, qs, nq, ix->d, ix->dtype, ix->metric == RAGARC_METRIC_COSINE, stream);
    if (rc) return rc;
    q_use = qs;
  }
  rc = ragarc_dense_topk(ix->rows, ix->n, ix->ds, ix->dtype, q_use, nq, k, d_scores, d_ids, ix->ws, ix->ws_bytes,
                         RAGARC_DENSE_AUTO, nullptr, stream);
  if (rc) return rc;
  if (l2) {                                                                              // kept values -> squared distances
    rc = ragarc_l2_distances(d_scores, qs, ix->dtype, nq, k, ix->d, stream);
    if (rc) return rc;
  }
  if (trace) { t_c = now(); cudaStreamSynchronize(st); t_d = now(); }
  if (buffers_on_host) {
    RA_CUDA(cudaMemcpyAsync(out_scores, d_scores, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    RA_CUDA(cudaMemcpyAsync(out_ids, d_ids, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    RA_CUDA(cudaStreamSynchronize(st));
  }
  if (trace)
    fprintf(stderr, "[index_search] h2d %.3f ms, enqueue kernels %.3f ms, kernels done +%.3f ms, d2h %.3f ms\n",
            t_b - t_a, t_c - t_b, t_d - t_c, now() - t_d);
  return order_end(ix, st);
}

int ragarc_index_remove(ragarc_index_t* ix, const int64_t* rows_host, int64_t n_remove, void* stream) {
  RA_REQUIRE(ix != nullptr, "index_remove: null index");
  RA_REQUIRE(n_remove >= 0, "index_remove: n_remove=%lld", (long long)n_remove);
  if (n_remove == 0) return RAGARC_OK;
  RA_REQUIRE(rows_host != nullptr, "index_remove: null rows");
  std::lock_guard<std::mutex> lk(ix->mu);
  DeviceGuard g(ix->device);
  RA_REQUIRE(g.ok, "index_remove: cannot select device %d", ix->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = order_begin(ix, st);
  if (rc) return rc;
  std::vector<uint8_t> drop((size_t)ix->n, 0);
  for (int64_t i = 0; i < n_remove; ++i) {
    RA_REQUIRE(rows_host[i] >= 0 && rows_host[i] < ix->n, "index_remove: row %lld out of range [0,%lld)",
               (long long)rows_host[i], (long long)ix->n);
    drop[(size_t)rows_host[i]] = 1;
  }
  // survivors keep their relative order and are renumbered densely (faiss remove_ids semantics,
  // which VectorStore_Faiss.py:403-412 mirrors in index_to_docstore_id)
  std::vector<int64_t> map;
  map.reserve((size_t)ix->n);
  for (int64_t r = 0; r < ix->n; ++r) if (!drop[(size_t)r]) map.push_back(r);
  const int64_t n_out = (int64_t)map.size();
  const size_t row_bytes = (size_t)ix->ds * esize(ix->dtype);
  void* fresh = nullptr;
  int64_t* d_map = nullptr;
  const int64_t cap = n_out > 0 ? n_out : 1;
  RA_CUDA(cudaMalloc(&fresh, (size_t)cap * row_bytes));
  if (n_out > 0) {
    RA_CUDA(cudaMalloc((void**)&d_map, (size_t)n_out * 8));
    RA_CUDA(cudaMemcpyAsync(d_map, map.data(), (size_t)n_out * 8, cudaMemcpyHostToDevice, st));
    gather_rows_kernel<<<(unsigned)n_out, 128, 0, st>>>((const uint8_t*)ix->rows, (uint8_t*)fresh, d_map, n_out, (int)row_bytes);
    RA_LAUNCH_CHECK();
  }
  RA_CUDA(cudaDeviceSynchronize());
  if (d_map) RA_CUDA(cudaFree(d_map));
  if (ix->rows) RA_CUDA(cudaFree(ix->rows));
  ix->rows = fresh;
  ix->cap = cap;
  ix->n = n_out;
  return order_end(ix, st);
}

}  // extern "C"

namespace ragarc {

// per-shard leg of a sharded search: fp32 HOST queries -> this shard's top-k as packed keys carrying
// global row ids (id_base + local row), written to `keys` (device, [nq,k]); asynchronous on `st`
static int index_search_keys(ragarc_index* ix, const float* queries_host, int nq, int k, uint64_t id_base,
                             uint64_t* keys, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(ix->mu);
  int rc = order_begin(ix, st);
  if (rc) return rc;
  if (ix->n == 0) {
    RA_CUDA(cudaMemsetAsync(keys, 0, (size_t)nq * k * 8, st));     // key 0 = padding
    return order_end(ix, st);
  }
  const size_t q32_bytes = align_up((size_t)nq * ix->d * 4, 256);
  const size_t qs_bytes = align_up((size_t)nq * ix->ds * esize(ix->dtype), 256);
  rc = grow_buffer(&ix->stage, &ix->stage_bytes, q32_bytes + qs_bytes, st);
  if (rc) return rc;
  const size_t need_ws = ragarc_dense_topk_workspace_bytes(ix->n, ix->ds, ix->dtype, nq, k);
  RA_REQUIRE(need_ws > 0, "sharded_search: unsupported shape (k=%d)", k);
  rc = grow_buffer(&ix->ws, &ix->ws_bytes, need_ws, st);
  if (rc) return rc;
  float* q32 = (float*)ix->stage;
  void* qs = (char*)ix->stage + q32_bytes;
  RA_CUDA(cudaMemcpyAsync(q32, queries_host, (size_t)nq * ix->d * 4, cudaMemcpyHostToDevice, st));
  const void* q_use = q32;
  if (ix->metric == RAGARC_METRIC_L2) {
    // keys carry q.x - |x|^2/2 (descending = ascending distance); the caller turns the merged values into distances
    rc = ragarc_l2_augment(q32, qs, nq, ix->d, ix->dtype, 1, 0, nullptr, st);
    if (rc) return rc;
    q_use = qs;
  } else if (ix->dtype != RAGARC_F32 || ix->metric == RAGARC_METRIC_COSINE) {
    rc = ragarc_normalize_cast(q32, qs, nq, ix->d, ix->dtype, ix->metric == RAGARC_METRIC_COSINE, st);
    if (rc) return rc;
    q_use = qs;
  }
  rc = ragarc_dense_topk_keys(ix->rows, ix->n, ix->ds, ix->dtype, q_use, nq, k, id_base, keys, ix->ws, ix->ws_bytes,
                              RAGARC_DENSE_AUTO, nullptr, st);
  if (rc) return rc;
  return order_end(ix, st);
}

}  // namespace ragarc

extern "C" {

int ragarc_sharded_create(int d, int dtype, int metric, int n_shards, const int* devices,
                          ragarc_sharded_index_t** out) {
  RA_REQUIRE(out != nullptr, "sharded_create: out is NULL");
  *out = nullptr;
  RA_REQUIRE(n_shards >= 1 && n_shards <= 64, "sharded_create: n_shards=%d", n_shards);
  int ndev = 0;
  RA_CUDA(cudaGetDeviceCount(&ndev));
  ragarc_sharded_index* sh = new (std::nothrow) ragarc_sharded_index();
  RA_REQUIRE(sh != nullptr, "sharded_create: out of host memory");
  sh->d = d; sh->dtype = dtype; sh->metric = metric;
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = RAGARC_OK;
  for (int g = 0; g < n_shards && rc == RAGARC_OK; ++g) {
    const int dev = devices ? devices[g] : g;
    if (dev < 0 || dev >= ndev) { set_error("sharded_create: device %d of shard %d does not exist (%d devices)", dev, g, ndev); rc = RAGARC_ERR_INVALID; break; }
    if (cudaSetDevice(dev) != cudaSuccess) { set_error("sharded_create: cannot select device %d", dev); rc = RAGARC_ERR_CUDA; break; }
    ragarc_index* ix = nullptr;
    rc = ragarc_index_create(d, dtype, metric, &ix);
    if (rc) break;
    sh->shard.push_back(ix);
    sh->base.push_back(0);
    sh->keys.push_back(nullptr);
    sh->keys_bytes.push_back(0);
    cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
      set_error("sharded_create: cannot create stream/event on device %d", dev); rc = RAGARC_ERR_CUDA;
    }
    sh->stream.push_back(st);
    sh->done.push_back(ev);
  }
  cudaSetDevice(prev);
  if (rc) { ragarc_sharded_free(sh); return rc; }
  *out = sh;
  return RAGARC_OK;
}

int ragarc_sharded_free(ragarc_sharded_index_t* sh) {
  if (!sh) return RAGARC_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  for (size_t g = 0; g < sh->shard.size(); ++g) {
    cudaSetDevice(sh->shard[g]->device);
    cudaDeviceSynchronize();
    if (g < sh->keys.size() && sh->keys[g]) cudaFree(sh->keys[g]);
    if (g == 0 && sh->gather) cudaFree(sh->gather);
    if (g < sh->stream.size() && sh->stream[g]) cudaStreamDestroy(sh->stream[g]);
    if (g < sh->done.size() && sh->done[g]) cudaEventDestroy(sh->done[g]);
    ragarc_index_free(sh->shard[g]);
  }
  cudaSetDevice(prev);
  delete sh;
  return RAGARC_OK;
}

int64_t ragarc_sharded_ntotal(const ragarc_sharded_index_t* sh) {
  if (!sh) return -1;
  int64_t n = 0;
  for (auto* ix : sh->shard) n += ix->n;
  return n;
}

int ragarc_sharded_add(ragarc_sharded_index_t* sh, const float* rows_host, int64_t n) {
  RA_REQUIRE(sh != nullptr, "sharded_add: null index");
  RA_REQUIRE(n >= 0, "sharded_add: n=%lld", (long long)n);
  if (n == 0) return RAGARC_OK;
  RA_REQUIRE(rows_host != nullptr, "sharded_add: null rows");
  std::lock_guard<std::mutex> lk(sh->mu);
  const int G = (int)sh->shard.size();
  const int64_t total = ragarc_sharded_ntotal(sh);
  RA_REQUIRE(total + n < (int64_t)0xFFFFFFF0ll, "sharded_add: global row ids must fit 32 bits");
  int prev = 0;
  cudaGetDevice(&prev);
  int rc = RAGARC_OK;
  if (total == 0) {
    // first load: contiguous ranges of ceil(n/G) rows, shard g = rows [g*per, (g+1)*per)
    const int64_t per = (n + G - 1) / G;
    for (int g = 0; g < G && rc == RAGARC_OK; ++g) {
      const int64_t lo = g * per < n ? g * per : n, hi = (g + 1) * per < n ? (g + 1) * per : n;
      sh->base[g] = lo;
      if (hi > lo) {
        cudaSetDevice(sh->shard[g]->device);
        rc = ragarc_index_add(sh->shard[g], rows_host + (size_t)lo * sh->d, hi - lo, 1, sh->stream[g]);
      }
    }
  } else {
    // later rows extend the last shard, whose range stays contiguous (global id = base + local row)
    ragarc_index* last = sh->shard[G - 1];
    if (last->n == 0) sh->base[G - 1] = total;
    cudaSetDevice(last->device);
    rc = ragarc_index_add(last, rows_host, n, 1, sh->stream[G - 1]);
  }
  cudaSetDevice(prev);
  return rc;
}

int ragarc_sharded_search(ragarc_sharded_index_t* sh, const float* queries_host, int nq, int k,
                          float* out_scores_host, int64_t* out_ids_host) {
  RA_REQUIRE(sh != nullptr, "sharded_search: null index");
  RA_REQUIRE(nq >= 0 && k > 0, "sharded_search: nq=%d k=%d", nq, k);
  if (nq == 0) return RAGARC_OK;
  RA_REQUIRE(queries_host && out_scores_host && out_ids_host, "sharded_search: null pointer");
  std::lock_guard<std::mutex> lk(sh->mu);
  const int G = (int)sh->shard.size();
  const size_t kb = (size_t)nq * k * 8;
  int prev = 0;
  cudaGetDevice(&prev);
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
  // fan out: every shard scores the batch on its own device and stream
  for (int g = 0; g < G; ++g) {
    RA_CUDA(cudaSetDevice(sh->shard[g]->device));
    int rc = grow_buffer(&sh->keys[g], &sh->keys_bytes[g], kb, sh->stream[g]);
    if (rc) return rc;
    rc = index_search_keys(sh->shard[g], queries_host, nq, k, (uint64_t)sh->base[g], (uint64_t*)sh->keys[g], sh->stream[g]);
    if (rc) return rc;
    RA_CUDA(cudaEventRecord(sh->done[g], sh->stream[g]));
  }
  // gather the G key blocks on shard 0's device and merge there (keys order by score, then lowest
  // global row: the result does not depend on G)
  const int dev0 = sh->shard[0]->device;
  cudaStream_t s0 = sh->stream[0];
  RA_CUDA(cudaSetDevice(dev0));
  const size_t keys_all = align_up((size_t)G * kb, 256), sc_bytes = align_up((size_t)nq * k * 4, 256);
  const size_t id_bytes = align_up((size_t)nq * k * 8, 256);
  const bool l2 = sh->metric == RAGARC_METRIC_L2;
  const int d_aug = l2 ? ragarc_l2_aug_dim(sh->d, sh->dtype) : 0;
  const size_t q32_bytes = l2 ? align_up((size_t)nq * sh->d * 4, 256) : 0;
  const size_t qa_bytes = l2 ? align_up((size_t)nq * d_aug * esize(sh->dtype), 256) : 0;
  int rc = grow_buffer(&sh->gather, &sh->gather_bytes, keys_all + sc_bytes + id_bytes + q32_bytes + qa_bytes, s0);
  if (rc) return rc;
  char* gb = (char*)sh->gather;
  for (int g = 0; g < G; ++g) {
    RA_CUDA(cudaStreamWaitEvent(s0, sh->done[g], 0));
    RA_CUDA(cudaMemcpyPeerAsync(gb + (size_t)g * kb, dev0, sh->keys[g], sh->shard[g]->device, kb, s0));
  }
  float* d_scores = (float*)(gb + keys_all);
  int64_t* d_ids = (int64_t*)(gb + keys_all + sc_bytes);
  rc = ragarc_merge_topk_keys((const uint64_t*)gb, G, nq, k, k, d_scores, d_ids, s0);
  if (rc) return rc;
  if (l2) {
    // merged values q.x - |x|^2/2 -> squared distances |q|^2 - 2(...), with |q|^2 from the queries as every
    // shard rounded them (the augmented batch is rebuilt here: shard 0 may hold no rows and skip its leg)
    float* q32 = (float*)(gb + keys_all + sc_bytes + id_bytes);
    void* qa = gb + keys_all + sc_bytes + id_bytes + q32_bytes;
    RA_CUDA(cudaMemcpyAsync(q32, queries_host, (size_t)nq * sh->d * 4, cudaMemcpyHostToDevice, s0));
    rc = ragarc_l2_augment(q32, qa, nq, sh->d, sh->dtype, 1, 0, nullptr, s0);
    if (rc) return rc;
    rc = ragarc_l2_distances(d_scores, qa, sh->dtype, nq, k, sh->d, s0);
    if (rc) return rc;
  }
  RA_CUDA(cudaMemcpyAsync(out_scores_host, d_scores, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, s0));
  RA_CUDA(cudaMemcpyAsync(out_ids_host, d_ids, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, s0));
  RA_CUDA(cudaStreamSynchronize(s0));
  return RAGARC_OK;
}

}  // extern "C"
