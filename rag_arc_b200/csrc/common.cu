// Error plumbing, launch accounting, the dense-search planner and the dense C-ABI entry points.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace ragarc {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}

struct ProfRec { cudaEvent_t e0, es, e1, e2; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

static int cap_for_k(int k) {
  if (k <= 224) return 512;
  if (k <= 480) return 1024;
  if (k <= 992) return 2048;
  if (k <= 2016) return 4096;
  return 0;
}

#ifndef RAGARC_DEFAULT_TC_CL
#define RAGARC_DEFAULT_TC_CL 2
#endif

int plan_dense(int64_t n, int d, int dtype, int nq, int k, int path, DensePlan* pl, int* path_out) {
  RA_REQUIRE(n >= 0 && d > 0 && nq >= 0 && k > 0, "dense: bad shape n=%lld d=%d nq=%d k=%d",
             (long long)n, d, nq, k);
  RA_REQUIRE(dtype == RAGARC_F32 || dtype == RAGARC_BF16 || dtype == RAGARC_F16, "dense: bad dtype %d", dtype);
  RA_REQUIRE(n < (int64_t)0xFFFFFFF0ll, "dense: at most 2^32-16 rows per shard");
  int cap = cap_for_k(k);
  if (!cap) { set_error("dense: k=%d > 2016 unsupported", k); return RAGARC_ERR_UNSUPPORTED; }
  int use = path;
  if (use == RAGARC_DENSE_AUTO)
    use = (dtype != RAGARC_F32 && d % 8 == 0) ? RAGARC_DENSE_TCGEN05 : RAGARC_DENSE_SIMT;
  if (use == RAGARC_DENSE_TCGEN05 && (dtype == RAGARC_F32 || d % 8 != 0)) {
    set_error("dense: tcgen05 path needs bf16/fp16 and d %% 8 == 0");
    return RAGARC_ERR_UNSUPPORTED;
  }
  *path_out = use;
  const bool tc = use == RAGARC_DENSE_TCGEN05;
  // tcgen05 path: CTA pairs (cta_group::2, 256 query rows per work item) once there are enough
  // queries to fill both halves; RAGARC_TC_CG=1|2 forces a variant (A/B measurements)
  int cg = nq > 128 ? 2 : 1;
  {
    static const char* env = getenv("RAGARC_TC_CG");
    if (env && (env[0] == '1' || env[0] == '2')) cg = env[0] - '0';
  }
  pl->rows_per_item = tc ? 128 * cg : 64;
  pl->tile_n = tc ? 256 : 64;
  pl->MB = (int)ceil_div(nq > 0 ? nq : 1, pl->rows_per_item);
  // pairs per cluster that share each corpus tile through TMA multicast (needs that many query
  // blocks); RAGARC_TC_CL=1|2|4 forces a variant
  int cl = 1;
  if (tc && cg == 2) {
    static const char* envc = getenv("RAGARC_TC_CL");
    const int want = envc ? atoi(envc) : RAGARC_DEFAULT_TC_CL;
    if (want >= 4 && pl->MB % 4 == 0) cl = 4;
    else if (want >= 2 && pl->MB % 2 == 0) cl = 2;
  }
  pl->cl = cl;
  pl->tiles = ceil_div(n > 0 ? n : 1, pl->tile_n);
  pl->cap = cap;
  // Slice count: work items (MB x S) should fill the persistent workers (SMs, or SM pairs for
  // cta_group::2) in whole waves - two waves when the merge capacity allows it (the second item of
  // a worker starts with the thresholds the first wave published), else one.  A ragged last wave
  // costs a full item time (measured: 163 items on 148 SMs ran 1.3x slower than 148).
  int64_t merge_cap = 8192;
  int64_t S = 0;
  pl->S_tail = 0;
  pl->tiles_main = pl->tiles;
  if (tc && cl > 1) {
    // Multicast clusters (cl pairs) only fit where a GPC has 2*cl free SMs; the SMs left over host
    // plain pairs.  Two concurrent launches share the corpus: the clusters take the first
    // tiles_main tiles in S - S_tail slices, the left-over pairs the rest in S_tail slices, sized by
    // the clusters' per-pair speed advantage so that both launches end together.
    const int64_t clusters = dense_tc_units(cg, cl) / cl;
    const int64_t spare = (sm_count() - clusters * cg * cl) / cg;
    const int64_t units_all = clusters * cl + spare;
    if ((units_all / pl->MB) * (int64_t)k > merge_cap) merge_cap = 16384;
    int64_t smax = merge_cap / k;
    if (smax > pl->tiles / 2) smax = pl->tiles / 2;
    static const char* envr = getenv("RAGARC_TC_RHO");      // per-pair speed of a cluster vs a plain pair
    // measured with the rung thresholds in place: clusters of two pairs are ~3 % faster per pair than plain
    // pairs (1.09 with the seeding pass; profiles/r02_rungs_timeline.txt, A/B in profiles/r02_split_ab.txt)
    const double rho = envr ? atof(envr) : (cl == 2 ? 1.03 : 1.15);
    for (int waves = 2; waves >= 1 && S == 0; --waves) {
      const int64_t sm = waves * clusters * cl / pl->MB, st = waves * spare / pl->MB;
      if (sm < 1 || sm + st > smax) continue;
      S = sm + st;
      pl->S_tail = (int)st;
      int64_t tm = (int64_t)((double)pl->tiles * (sm * rho) / (sm * rho + st) + 0.5);
      if (tm > pl->tiles - st) tm = pl->tiles - st;
      if (tm < sm) tm = sm;
      pl->tiles_main = st > 0 ? tm : pl->tiles;
    }
    if (S == 0) { cl = 1; pl->cl = 1; merge_cap = 8192; }   // too few tiles / too many slices: plain pairs
  }
  if (S == 0) {
    const int64_t units = tc ? dense_tc_units(cg, 1) : (int64_t)sm_count() * 2;
    // the merge kernel holds S*keep candidate keys in shared memory: 8192 normally, 16384 when that is
    // what it takes to give every worker a slice (small batches: the merge grid is tiny then anyway)
    if ((units / pl->MB) * (int64_t)k > merge_cap) merge_cap = 16384;
    int64_t smax = merge_cap / k;
    if (smax > 1024) smax = 1024;
    if (smax > pl->tiles) smax = pl->tiles;
    if (smax < 1) smax = 1;
    double best_eff = -1.0;
    for (int waves = 2; waves >= 1; --waves) {
      int64_t cand = waves * units / pl->MB;
      if (cand < 1) cand = 1;
      if (cand > smax) continue;
      const int64_t items_c = cand * pl->MB;
      const double eff = (double)items_c / (double)(ceil_div(items_c, units) * units);
      if (eff > best_eff + 1e-9) { best_eff = eff; S = cand; }
    }
    if (S == 0) S = smax;                 // cannot fill a wave within the merge capacity
    {
      static const char* envs = getenv("RAGARC_DENSE_S");     // experiments: force the slice count
      if (envs && atoi(envs) > 0) S = atoi(envs);
    }
    if (S < 1) S = 1;
    if (S > pl->tiles) S = pl->tiles;
  }
  pl->S = (int)S;
  // Two epilogue warp sets per CTA (tensor-core path), each draining every other tile into its own
  // candidate list: twice the lists per query, so the merge capacity must hold them, and every slice
  // needs a few tiles for both sets to have work.  Measured (profiles/r02_epilogue_sets_ab.txt): no
  // faster than one set - the epilogue is not the pace setter - so it is opt-in: RAGARC_TC_SETS=2.
  pl->sets = 1;
  if (tc) {
    static const char* envs2 = getenv("RAGARC_TC_SETS");
    const bool want2 = envs2 && envs2[0] == '2' && dense_tc_max_sets() >= 2;
    const int64_t s_main = S - pl->S_tail;
    if (want2 && S * 2 * (int64_t)k <= merge_cap && pl->tiles_main >= 4 * s_main &&
        (pl->S_tail == 0 || pl->tiles - pl->tiles_main >= 4 * (int64_t)pl->S_tail))
      pl->sets = 2;
  }
  pl->keep = (int)(merge_cap / (S * pl->sets));
  if (pl->keep < k) pl->keep = k;
  if (pl->keep > cap - 32) pl->keep = cap - 32;
  // Published order statistics (tensor-core path, see common.cuh): the lists of the first pub_n
  // slices - all of them running in the first wave of the main launch - publish their pub_m-th best
  // score after every tile.  Needs pub_n * pub_m >= k with pub_m <= 8 and at least two tiles per
  // slice (so that the first tile of every publishing list is a full one).  Replaces the seeding
  // pass below where it applies; RAGARC_TC_PUB=0 switches it off (A/B measurements).
  pl->pub_n = 0;
  pl->pub_m = 0;
  if (tc) {
    static const char* envp = getenv("RAGARC_TC_PUB");
    const bool want = !(envp && envp[0] == '0');
    const int64_t mbq = pl->MB / cl, s_main = pl->S - pl->S_tail;
    const int64_t workers = cl > 1 ? dense_tc_units(cg, cl) / cl : dense_tc_units(cg, 1);
    const int64_t items_main = mbq * s_main;
    const int64_t units = items_main < workers ? items_main : workers;
    int64_t pn = units / mbq;                      // slices fully covered by the first wave
    if (pn > s_main) pn = s_main;
    if (pn > PUB_LD / pl->sets) pn = PUB_LD / pl->sets;
    if (want && pn >= 1 && pl->tiles_main >= 2 * pl->sets * s_main) {
      const int64_t m = ceil_div(k, pn * pl->sets);          // one publishing list per (slice, epilogue set)
      if (m <= PUB_MAX_M) { pl->pub_n = (int)(pn * pl->sets); pl->pub_m = (int)m; }
    }
  }
  // threshold seeding (tensor-core path): the first seed_rows rows are scored, the maximum of every
  // 16-row group is kept, and the k-th largest group maximum - a score that at least k distinct
  // rows reach - becomes the initial shared threshold of each query.  Needs >= 4k groups for a
  // tight bound and is only worth it when the seed rows are a small fraction of the corpus.
  pl->seed_rows = 0;
  pl->seed_S = 0;
  if (tc && k <= 256 && pl->pub_n == 0) {
    int64_t sr = n / 64;
    if (sr < 64 * (int64_t)k) sr = 64 * (int64_t)k;
    if (sr > 16384) sr = 16384;
    {
      static const char* envr = getenv("RAGARC_SEED_ROWS");    // experiments: force the seed sample size
      if (envr && atoi(envr) > 0 && atoi(envr) <= 262144) sr = atoi(envr);
    }
    sr = (sr + pl->tile_n - 1) / pl->tile_n * pl->tile_n;
    if (sr / 16 >= 4 * (int64_t)k && n >= 8 * sr) {
      pl->seed_rows = (int)sr;
      int64_t st = sr / pl->tile_n;
      int64_t ss = dense_tc_units(cg, 1) / pl->MB;
      pl->seed_S = (int)(ss < 1 ? 1 : (ss > st ? st : ss));
    }
  }
  size_t items = (size_t)pl->MB * pl->S;
  size_t off = 0;
  pl->off_lists = off;  off = align_up(off + items * pl->rows_per_item * pl->sets * (size_t)cap * 8, 256);
  pl->off_counts = off; off = align_up(off + items * pl->rows_per_item * pl->sets * 4, 256);
  pl->off_gthr = off;   off = align_up(off + (size_t)(nq > 0 ? nq : 1) * 4, 256);
  pl->off_pub = off;    off = align_up(off + (tc ? (size_t)(nq > 0 ? nq : 1) * PUB_LD * 4 : 0), 256);
  pl->off_keys = off;   off = align_up(off + (size_t)(nq > 0 ? nq : 1) * k * 8, 256);
  pl->off_qpad = off;   off = align_up(off + (tc ? (size_t)pl->MB * pl->rows_per_item * d * 2 : 0), 256);
  pl->off_seed = off;   off = align_up(off + (size_t)(nq > 0 ? nq : 1) * (pl->seed_rows / 16) * 4, 256);
  // merge kernel: global gather rows for the (rare) queries whose raw candidates exceed its shared-memory array
  pl->off_mscratch = off; off = align_up(off + (size_t)(nq > 0 ? nq : 1) * pl->S * pl->sets * pl->keep * 8, 256);
  pl->total = off;
  return RAGARC_OK;
}

static int dense_common(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                        int k, uint64_t id_base, uint64_t* out_keys, float* out_scores,
                        int64_t* out_ids, void* workspace, size_t workspace_bytes, int path,
                        int* path_used_host, cudaStream_t stream, int x3_d = 0,
                        const MergePush* push = nullptr, int phase = RAGARC_PHASE_BOTH, bool ws_clean = false) {
  DensePlan pl;
  pl.x3_d = x3_d;
  int use = 0;
  if (path == RAGARC_DENSE_AUTO && !dense_tc_supported(corpus, n, d, dtype, queries)) path = RAGARC_DENSE_SIMT;
  int rc = plan_dense(n, d, dtype, nq, k, path, &pl, &use);
  if (rc) return rc;
  if (path_used_host) *path_used_host = use;
  if (nq == 0) return RAGARC_OK;
  const bool do_score = phase != RAGARC_PHASE_SELECT, do_select = phase != RAGARC_PHASE_SCORE;
  RA_REQUIRE(!do_score || corpus || n == 0, "dense: null corpus");
  RA_REQUIRE(!do_score || queries, "dense: null queries");
  RA_REQUIRE(workspace && workspace_bytes >= pl.total, "dense: workspace %zu < required %zu",
             workspace_bytes, pl.total);
  RA_REQUIRE(((uintptr_t)workspace & 255) == 0, "dense: workspace must be 256-byte aligned");
  char* ws = (char*)workspace;
  uint64_t* lists = (uint64_t*)(ws + pl.off_lists);
  int* counts = (int*)(ws + pl.off_counts);
  uint32_t* gthr = (uint32_t*)(ws + pl.off_gthr);
  // list lengths are written for every (item,row) by the scoring kernels, so only the shared
  // thresholds need a reset - and not even those when the seed pass overwrites all of them
  uint32_t* pub = (uint32_t*)(ws + pl.off_pub);
  if (!do_score) {
    // selection only: the candidate lists of an earlier RAGARC_PHASE_SCORE call with the same shape are
    // in the workspace (the caller orders the two calls; they may be on different streams).  The
    // thresholds and published rungs are reset BEHIND the merge, so that the next SCORE on this
    // workspace can start with its scoring kernel instead of a memset (workspace_clean).
    rc = launch_merge_lists(lists, counts, pl, nq, k, id_base, gthr, (uint64_t*)(ws + pl.off_mscratch), out_keys,
                            out_scores, out_ids, push, stream);
    if (rc) return rc;
    RA_CUDA(cudaMemsetAsync(gthr, 0, pl.off_keys - pl.off_gthr, stream));
    return RAGARC_OK;
  }
  if (ws_clean && phase == RAGARC_PHASE_SCORE && !(use == RAGARC_DENSE_TCGEN05 && pl.seed_rows > 0)) {
    // the caller vouches that a SELECT (which resets them) ran on this workspace since the last SCORE
  } else if (use == RAGARC_DENSE_TCGEN05 && pl.pub_n > 0)
    RA_CUDA(cudaMemsetAsync(gthr, 0, pl.off_keys - pl.off_gthr, stream));   // thresholds + published rungs
  else if (!(use == RAGARC_DENSE_TCGEN05 && pl.seed_rows > 0 && n > 0))
    RA_CUDA(cudaMemsetAsync(gthr, 0, (size_t)nq * 4, stream));
  if (n == 0) RA_CUDA(cudaMemsetAsync(counts, 0, pl.off_gthr - pl.off_counts, stream));
  ProfRec pr{};
  const bool prof = g_prof_on.load() != 0 && do_select;
  if (prof) {
    RA_CUDA(cudaEventCreate(&pr.e0)); RA_CUDA(cudaEventCreate(&pr.e1)); RA_CUDA(cudaEventCreate(&pr.e2));
    RA_CUDA(cudaEventCreate(&pr.es));
    RA_CUDA(cudaEventRecord(pr.e0, stream));
  }
  if (n > 0) {
    if (use == RAGARC_DENSE_TCGEN05)
      rc = launch_dense_tc(corpus, n, d, dtype, queries, nq, k, pl, lists, counts, gthr, pub,
                           (float*)(ws + pl.off_seed), ws + pl.off_qpad, prof ? pr.es : nullptr, stream);
    else {
      if (prof) RA_CUDA(cudaEventRecord(pr.es, stream));
      rc = launch_dense_simt(corpus, n, d, dtype, queries, nq, k, pl, lists, counts, gthr, stream);
    }
    if (rc) return rc;
  } else if (prof) {
    RA_CUDA(cudaEventRecord(pr.es, stream));
  }
  if (prof) RA_CUDA(cudaEventRecord(pr.e1, stream));
  if (!do_select) return RAGARC_OK;
  rc = launch_merge_lists(lists, counts, pl, nq, k, id_base, gthr, (uint64_t*)(ws + pl.off_mscratch), out_keys,
                          out_scores, out_ids, push, stream);
  if (rc) return rc;
  if (prof) {
    RA_CUDA(cudaEventRecord(pr.e2, stream));
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(pr);
  }
  return RAGARC_OK;
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

int ragarc_abi_version(void) { return RAGARC_ABI_VERSION; }
const char* ragarc_last_error(void) { return g_err; }
uint64_t ragarc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int ragarc_profile_enable(int on) { g_prof_on.store(on ? 1 : 0); return RAGARC_OK; }

int ragarc_profile_read(double* seed_ms_sum_host, double* score_ms_sum_host, double* merge_ms_sum_host,
                        int* n_host) {
  std::vector<ProfRec> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    recs.swap(g_prof);
  }
  double a = 0, b = 0, c = 0;
  for (auto& r : recs) {
    RA_CUDA(cudaEventSynchronize(r.e2));
    float ms = 0, m0 = 0, m1 = 0;
    RA_CUDA(cudaEventElapsedTime(&ms, r.e0, r.es));
    RA_CUDA(cudaEventElapsedTime(&m0, r.es, r.e1));
    RA_CUDA(cudaEventElapsedTime(&m1, r.e1, r.e2));
    c += ms; a += m0; b += m1;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.es); cudaEventDestroy(r.e1); cudaEventDestroy(r.e2);
  }
  if (seed_ms_sum_host) *seed_ms_sum_host = c;
  if (score_ms_sum_host) *score_ms_sum_host = a;
  if (merge_ms_sum_host) *merge_ms_sum_host = b;
  if (n_host) *n_host = (int)recs.size();
  return RAGARC_OK;
}

size_t ragarc_dense_topk_x3_workspace_bytes(int64_t n, int d, int nq, int k) {
  DensePlan pl;
  int use;
  if (d % 64 != 0) return 0;
  if (plan_dense(n, 3 * d, RAGARC_BF16, nq, k, RAGARC_DENSE_TCGEN05, &pl, &use) != RAGARC_OK) return 0;
  return pl.total;
}

int ragarc_dense_topk_x3(const void* corpus_planes, int64_t n, int d, const void* query_planes, int nq, int k,
                         float* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                         void* stream) {
  RA_REQUIRE(out_scores && out_ids, "dense_topk_x3: null outputs");
  RA_REQUIRE(d > 0 && d % 64 == 0, "dense_topk_x3: d=%d must be a multiple of 64", d);
  return dense_common(corpus_planes, n, 3 * d, RAGARC_BF16, query_planes, nq, k, 0, nullptr, out_scores, out_ids,
                      workspace, workspace_bytes, RAGARC_DENSE_TCGEN05, nullptr, (cudaStream_t)stream, d);
}

size_t ragarc_dense_topk_workspace_bytes(int64_t n, int d, int dtype, int nq, int k) {
  // worst case over the two paths so that the caller can size once
  size_t best = 0;
  for (int path : {RAGARC_DENSE_SIMT, RAGARC_DENSE_TCGEN05}) {
    DensePlan pl;
    int use;
    if (path == RAGARC_DENSE_TCGEN05 && (dtype == RAGARC_F32 || d % 8 != 0)) continue;
    if (plan_dense(n, d, dtype, nq, k, path, &pl, &use) == RAGARC_OK && pl.total > best) best = pl.total;
  }
  return best;
}

int ragarc_dense_topk_plan(int64_t n, int d, int dtype, int nq, int k, int path, int* out16) {
  int* out8 = out16;
  RA_REQUIRE(out8 != nullptr, "dense_topk_plan: out is NULL");
  DensePlan pl;
  int use = 0;
  int rc = plan_dense(n, d, dtype, nq, k, path, &pl, &use);
  if (rc) return rc;
  const bool tc = use == RAGARC_DENSE_TCGEN05;
  out8[0] = use; out8[1] = pl.rows_per_item; out8[2] = pl.cl; out8[3] = pl.MB; out8[4] = pl.S;
  out8[5] = tc ? dense_tc_units(pl.rows_per_item / 128, pl.cl) : sm_count() * 2;
  out8[6] = pl.seed_rows; out8[7] = pl.keep; out8[8] = pl.S_tail; out8[9] = (int)pl.tiles_main;
  out8[10] = pl.pub_n; out8[11] = pl.pub_m; out8[12] = pl.sets; out8[13] = out8[14] = out8[15] = 0;
  return RAGARC_OK;
}

int ragarc_dense_topk(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq,
                      int k, float* out_scores, int64_t* out_ids, void* workspace,
                      size_t workspace_bytes, int path, int* path_used_host, void* stream) {
  RA_REQUIRE(out_scores && out_ids, "dense_topk: null outputs");
  return dense_common(corpus, n, d, dtype, queries, nq, k, 0, nullptr, out_scores, out_ids, workspace,
                      workspace_bytes, path, path_used_host, (cudaStream_t)stream);
}

int ragarc_dense_topk_keys(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                           int nq, int k, uint64_t id_base, uint64_t* out_keys, void* workspace,
                           size_t workspace_bytes, int path, int* path_used_host, void* stream) {
  RA_REQUIRE(out_keys, "dense_topk_keys: null output");
  RA_REQUIRE(id_base + (uint64_t)n < 0xFFFFFFF0ull, "dense_topk_keys: global ids must fit 32 bits");
  return dense_common(corpus, n, d, dtype, queries, nq, k, id_base, out_keys, nullptr, nullptr,
                      workspace, workspace_bytes, path, path_used_host, (cudaStream_t)stream);
}

int ragarc_dense_topk_ex(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq, int k,
                         const ragarc_dense_opts_t* opts, void* workspace, size_t workspace_bytes, int path,
                         int* path_used_host, void* stream) {
  RA_REQUIRE(opts != nullptr, "dense_topk_ex: null options");
  RA_REQUIRE(opts->phase == RAGARC_PHASE_BOTH || opts->phase == RAGARC_PHASE_SCORE || opts->phase == RAGARC_PHASE_SELECT,
             "dense_topk_ex: bad phase %d", opts->phase);
  const bool wants_out = opts->phase != RAGARC_PHASE_SCORE;
  MergePush mp{nullptr, 0, 1, 0};
  const MergePush* push = nullptr;
  if (wants_out && opts->inboxes) {
    RA_REQUIRE(opts->n_ranks > 0 && opts->rank >= 0 && opts->rank < opts->n_ranks && opts->nq_per_rank > 0 &&
               (int64_t)opts->nq_per_rank * opts->n_ranks >= nq, "dense_topk_ex: bad query-owner partition");
    mp = MergePush{opts->inboxes, opts->rank, opts->nq_per_rank, opts->signal ? opts->n_ranks : 0};
    push = &mp;
  } else if (wants_out) {
    RA_REQUIRE(opts->out_keys || (opts->out_scores && opts->out_ids), "dense_topk_ex: no output given");
  }
  RA_REQUIRE(opts->id_base + (uint64_t)n < 0xFFFFFFF0ull, "dense_topk_ex: global ids must fit 32 bits");
  return dense_common(corpus, n, d, dtype, queries, nq, k, opts->id_base, push ? nullptr : opts->out_keys,
                      push ? nullptr : opts->out_scores, push ? nullptr : opts->out_ids, workspace, workspace_bytes,
                      path, path_used_host, (cudaStream_t)stream, 0, push, opts->phase, opts->workspace_clean != 0);
}

int ragarc_dense_topk_keys_push(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                                int nq, int k, uint64_t id_base, uint64_t* const* inboxes, int n_ranks,
                                int rank, int nq_per_rank, int signal, void* workspace, size_t workspace_bytes,
                                int path, int* path_used_host, void* stream) {
  RA_REQUIRE(inboxes, "dense_topk_keys_push: null inbox table");
  RA_REQUIRE(n_ranks > 0 && rank >= 0 && rank < n_ranks && nq_per_rank > 0 &&
             (int64_t)nq_per_rank * n_ranks >= nq,
             "dense_topk_keys_push: bad partition n_ranks=%d rank=%d nq_per_rank=%d nq=%d", n_ranks, rank,
             nq_per_rank, nq);
  RA_REQUIRE(id_base + (uint64_t)n < 0xFFFFFFF0ull, "dense_topk_keys_push: global ids must fit 32 bits");
  MergePush mp{inboxes, rank, nq_per_rank, signal ? n_ranks : 0};
  return dense_common(corpus, n, d, dtype, queries, nq, k, id_base, nullptr, nullptr, nullptr, workspace,
                      workspace_bytes, path, path_used_host, (cudaStream_t)stream, 0, &mp);
}

}  // extern "C"
