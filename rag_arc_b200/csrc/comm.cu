// NCCL-backed multi-GPU search inside the C ABI: for hosts that are not
// PyTorch.  ragarc_sharded_topk = the fused local scoring + selection on this rank's row shard
// (packed keys carrying global row ids), ONE ncclAllGather of the [nq,k] key blocks over NVLink, and
// the G-way merge on every rank - bit-identical to the single-GPU result for any G.  (The Python
// package's default exchange, rag_arc_b200/sharded.py, goes through peer memory instead; this is the
// collective form of the same step.)
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy already loaded into the process, e.g.
// PyTorch's, is reused), so the library itself carries no link-time dependency on it.
#include <dlfcn.h>
#include <mutex>
#include <new>
#include <vector>
#include "common.cuh"

namespace ragarc {

// the slice of nccl.h this file needs (layout-compatible declarations: NCCL 2.x ABI)
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { NCCL_SUCCESS = 0, NCCL_UINT64 = 5 };

struct NcclApi {
  int (*GetUniqueId)(NcclUniqueId*);
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int);
  int (*CommInitAll)(NcclComm*, int, const int*);
  int (*CommDestroy)(NcclComm);
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
  const char* (*GetErrorString)(int);
  int (*GetVersion)(int*);
  void* handle = nullptr;
  bool ok = false;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (PyTorch's)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.handle = h;
#define RA_SYM(field, name) *(void**)(&api.field) = dlsym(h, name)
    RA_SYM(GetUniqueId, "ncclGetUniqueId"); RA_SYM(CommInitRank, "ncclCommInitRank"); RA_SYM(CommInitAll, "ncclCommInitAll");
    RA_SYM(CommDestroy, "ncclCommDestroy"); RA_SYM(AllGather, "ncclAllGather"); RA_SYM(GetErrorString, "ncclGetErrorString");
    RA_SYM(GetVersion, "ncclGetVersion");
#undef RA_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.GetErrorString;
  });
  return api.ok ? &api : nullptr;
}

#define RA_NCCL(api, expr)                                                                 \
  do {                                                                                      \
    int _r = (expr);                                                                        \
    if (_r != NCCL_SUCCESS) {                                                               \
      set_error("%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
      return RAGARC_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)

}  // namespace ragarc

struct ragarc_comm {
  ragarc::NcclComm comm = nullptr;
  int nranks = 1, rank = 0, device = 0;
};

using namespace ragarc;

extern "C" {

int ragarc_comm_nccl_version(void) {
  NcclApi* api = nccl_api();
  int v = 0;
  if (!api || !api->GetVersion || api->GetVersion(&v) != NCCL_SUCCESS) return 0;
  return v;
}

int ragarc_comm_unique_id(char* out128_host) {
  RA_REQUIRE(out128_host != nullptr, "comm_unique_id: out is NULL");
  NcclApi* api = nccl_api();
  RA_REQUIRE(api != nullptr, "comm: libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  RA_NCCL(api, api->GetUniqueId(&id));
  memcpy(out128_host, id.internal, 128);
  return RAGARC_OK;
}

int ragarc_comm_init_rank(const char* id128_host, int nranks, int rank, ragarc_comm_t** out) {
  RA_REQUIRE(out != nullptr && id128_host != nullptr, "comm_init_rank: null pointer");
  *out = nullptr;
  RA_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_init_rank: nranks=%d rank=%d", nranks, rank);
  NcclApi* api = nccl_api();
  RA_REQUIRE(api != nullptr, "comm: libnccl.so.2 could not be loaded");
  ragarc_comm* c = new (std::nothrow) ragarc_comm();
  RA_REQUIRE(c != nullptr, "comm_init_rank: out of host memory");
  c->nranks = nranks; c->rank = rank;
  if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; set_error("comm_init_rank: no CUDA device"); return RAGARC_ERR_CUDA; }
  NcclUniqueId id;
  memcpy(id.internal, id128_host, 128);
  const int r = api->CommInitRank(&c->comm, nranks, id, rank);
  if (r != NCCL_SUCCESS) { set_error("comm_init_rank: NCCL: %s", api->GetErrorString(r)); delete c; return RAGARC_ERR_CUDA; }
  *out = c;
  return RAGARC_OK;
}

int ragarc_comm_init_all(int ndev, const int* devices, ragarc_comm_t** out_comms) {
  RA_REQUIRE(out_comms != nullptr && ndev >= 1 && ndev <= 64, "comm_init_all: ndev=%d", ndev);
  NcclApi* api = nccl_api();
  RA_REQUIRE(api != nullptr, "comm: libnccl.so.2 could not be loaded");
  std::vector<int> devs(ndev);
  for (int i = 0; i < ndev; ++i) devs[i] = devices ? devices[i] : i;
  std::vector<NcclComm> comms(ndev, nullptr);
  RA_NCCL(api, api->CommInitAll(comms.data(), ndev, devs.data()));
  for (int i = 0; i < ndev; ++i) {
    ragarc_comm* c = new (std::nothrow) ragarc_comm();
    RA_REQUIRE(c != nullptr, "comm_init_all: out of host memory");
    c->comm = comms[i]; c->nranks = ndev; c->rank = i; c->device = devs[i];
    out_comms[i] = c;
  }
  return RAGARC_OK;
}

int ragarc_comm_free(ragarc_comm_t* c) {
  if (!c) return RAGARC_OK;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  delete c;
  return RAGARC_OK;
}

int ragarc_comm_rank(const ragarc_comm_t* c) { return c ? c->rank : -1; }
int ragarc_comm_nranks(const ragarc_comm_t* c) { return c ? c->nranks : -1; }

size_t ragarc_sharded_topk_workspace_bytes(int64_t n_local, int d, int dtype, int nq, int k, int nranks) {
  const size_t base = ragarc_dense_topk_workspace_bytes(n_local, d, dtype, nq, k);
  if (base == 0 || nranks < 1) return 0;
  // + this rank's key block and the gathered [nranks, nq, k] block
  return align_up(base, 256) + align_up((size_t)nq * k * 8, 256) + align_up((size_t)nranks * nq * k * 8, 256);
}

int ragarc_sharded_topk(ragarc_comm_t* c, const void* corpus_shard, int64_t n_local, int d, int dtype,
                        const void* queries, int nq, int k, uint64_t id_base, float* out_scores, int64_t* out_ids,
                        void* workspace, size_t workspace_bytes, void* stream) {
  RA_REQUIRE(c != nullptr, "sharded_topk: null communicator");
  RA_REQUIRE(out_scores && out_ids, "sharded_topk: null outputs");
  NcclApi* api = nccl_api();
  RA_REQUIRE(api != nullptr, "comm: libnccl.so.2 could not be loaded");
  const size_t base = ragarc_dense_topk_workspace_bytes(n_local, d, dtype, nq, k);
  RA_REQUIRE(base > 0, "sharded_topk: unsupported shape (k=%d)", k);
  const size_t need = ragarc_sharded_topk_workspace_bytes(n_local, d, dtype, nq, k, c->nranks);
  RA_REQUIRE(workspace && workspace_bytes >= need, "sharded_topk: workspace %zu < required %zu", workspace_bytes, need);
  if (nq == 0) return RAGARC_OK;
  char* ws = (char*)workspace;
  uint64_t* mine = (uint64_t*)(ws + align_up(base, 256));
  uint64_t* all = (uint64_t*)(ws + align_up(base, 256) + align_up((size_t)nq * k * 8, 256));
  int rc = ragarc_dense_topk_keys(corpus_shard, n_local, d, dtype, queries, nq, k, id_base, mine, workspace, base,
                                  RAGARC_DENSE_AUTO, nullptr, stream);
  if (rc) return rc;
  RA_NCCL(api, api->AllGather(mine, all, (size_t)nq * k, NCCL_UINT64, c->comm, (cudaStream_t)stream));
  return ragarc_merge_topk_keys(all, c->nranks, nq, k, k, out_scores, out_ids, stream);
}

}  // extern "C"
