/*
 * ragarc_b200.h - C ABI of libragarc_b200.so: the B200 (sm_100a) retrieval hot path of RAG-ARC.
 *
 * Every entry point replaces a piece of arithmetic that the reference obtains from a Python
 * third-party package on the CPU; the citation after each prototype names the reference call
 * site (relative to /root/reference) a maintainer would re-point at this library (see
 * INTEGRATION.md for the ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host (or a flag says so); the
 *     caller owns all buffers and the compute entry points keep no state between calls (no
 *     globals except the thread-local error string).  The one object the library owns is the
 *     optional ragarc_index_t (ragarc_index_*), a flat index that holds its row matrix,
 *     workspace and staging buffers itself for hosts that do not want to manage device memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are
 *     asynchronous on that stream, re-entrant, and may be made from any host thread
 *     (core/retrieval/base.py:82-96 runs retrievers from a thread pool);
 *   - return value 0 = ok; anything else is a RAGARC_ERR_* code and ragarc_last_error() holds a
 *     message for the calling thread.  Nothing throws across the ABI;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RAGARC_ERR_CUDA.
 */
#ifndef RAGARC_B200_H
#define RAGARC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RAGARC_ABI_VERSION 1

enum {
  RAGARC_OK = 0,
  RAGARC_ERR_INVALID = 1,   /* bad argument */
  RAGARC_ERR_CUDA = 2,      /* CUDA runtime / driver error */
  RAGARC_ERR_WORKSPACE = 3, /* workspace too small */
  RAGARC_ERR_UNSUPPORTED = 4
};

/* storage dtype of corpus / query / encoder matrices */
enum { RAGARC_F32 = 0, RAGARC_BF16 = 1, RAGARC_F16 = 2, RAGARC_F64 = 3 /* adjacent_cosine_distance only */ };

/* pooling modes (sentence-transformers Pooling module) */
enum { RAGARC_POOL_MEAN = 0, RAGARC_POOL_CLS = 1, RAGARC_POOL_LAST = 2 };

/* similarity of a ragarc_index_t: raw inner product, cosine (rows and queries L2-normalised), or
 * squared L2 distance (results ascending, like faiss.IndexFlatL2) */
enum { RAGARC_METRIC_IP = 0, RAGARC_METRIC_COSINE = 1, RAGARC_METRIC_L2 = 2 };

/* which dense scoring kernel ran / should run */
enum { RAGARC_DENSE_AUTO = 0, RAGARC_DENSE_SIMT = 1, RAGARC_DENSE_TCGEN05 = 2 };

int ragarc_abi_version(void);
const char* ragarc_last_error(void);

/* Number of GPU kernels this library has launched from the calling process so far (all
 * threads).  bench.py reports the delta over its timed region as "gpu_launches". */
uint64_t ragarc_launch_count(void);

/* Measurement hooks (bench.py): when enabled, every dense search records CUDA events on the
 * caller's stream around its three phases: threshold seeding (tcgen05 path on large corpora
 * only, else ~0), the main scoring + selection kernel, and the merge kernel.
 * ragarc_profile_read synchronises on the recorded events, returns the summed device times in
 * milliseconds and the number of searches recorded since the last read, and clears the record. */
int ragarc_profile_enable(int on);
int ragarc_profile_read(double* seed_ms_sum_host, double* score_ms_sum_host,
                        double* merge_ms_sum_host, int* n_host);

/* ---------------------------------------------------------------------------------------------
 * L2 normalise + cast.   Replaces faiss.normalize_L2 at
 *   encapsulation/database/vector_db/VectorStore_Faiss.py:150-154 (called :178 on add, :259 on
 *   query) and the .astype(np.float32) staging at :170,:258.
 * src: fp32 [n,d] row-major.  dst: [n,d] of dst_dtype (may alias src when dst_dtype==F32).
 * normalize!=0: per row nr=sum(x^2) in fp32; if nr>0 row *= 1/sqrt(nr) (zero rows untouched).
 */
int ragarc_normalize_cast(const float* src, void* dst, int64_t n, int d, int dst_dtype,
                          int normalize, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Exact dense top-k.   Replaces faiss.IndexFlatIP.search at VectorStore_Faiss.py:263
 *   (D float32[nq,k] descending inner product, I int64[nq,k], -1 padding when k > n).
 * corpus: [n,d] row-major, queries: [nq,d] row-major, both of `dtype`; fp32 accumulation.
 * Equal scores are ordered by ascending row id (FAISS leaves tie order unspecified).
 * The nq x n score matrix is never written to memory: scoring and per-query selection are
 * fused (tcgen05 tensor-core tiles for bf16/fp16 with d % 8 == 0, fp32 SIMT tiles otherwise).
 * `path` = RAGARC_DENSE_AUTO lets the library choose; *path_used (host, may be NULL) reports it.
 */
size_t ragarc_dense_topk_workspace_bytes(int64_t n, int d, int dtype, int nq, int k);
/* Introspection: the schedule the library would use for this shape (for logs and benchmarks).
 * out[0..16) = path, query rows per work item, CTA pairs per multicast cluster, query blocks,
 * corpus slices, resident work-item slots, seed rows, candidates kept per list, slices given to the
 * concurrent plain-pair launch, corpus tiles given to the multicast-cluster launch, number of
 * candidate lists per query that publish an order statistic as a running threshold (0 = threshold
 * seeding pass instead), which order statistic (m-th best) they publish, and the number of epilogue
 * warp sets per CTA (= candidate lists per work item and query); the remaining entries are 0. */
int ragarc_dense_topk_plan(int64_t n, int d, int dtype, int nq, int k, int path, int* out16);
int ragarc_dense_topk(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                      int nq, int k, float* out_scores, int64_t* out_ids, void* workspace,
                      size_t workspace_bytes, int path, int* path_used_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Flat exact index object.   Replaces faiss.IndexFlatIP as FaissVectorStore drives it:
 *   faiss.IndexFlatIP(d)            VectorStore_Faiss.py:114-115,125-126,135-136  -> ragarc_index_create
 *   normalize_L2 + index.add        :169-178,202                                   -> ragarc_index_add
 *   normalize_L2 + index.search     :258-263                                       -> ragarc_index_search
 *   index.ntotal                    :201,262                                       -> ragarc_index_ntotal
 *   index.remove_ids + renumbering  :385-419                                       -> ragarc_index_remove
 * The index lives on the CUDA device that is current at creation.  Rows are given as fp32 (what the
 * reference hands to FAISS), L2-normalised on add when metric == RAGARC_METRIC_COSINE (zero rows
 * untouched) and stored as `dtype`; queries are fp32 and get the same treatment.  `*_on_host` != 0
 * means the fp32 / result pointers are plain host memory: the call stages them through the device
 * itself and returns when the host buffers are reusable / filled; otherwise they are device
 * pointers and the call is asynchronous on `stream`.  Calls on one index are serialised (host
 * mutex + GPU ordering across streams); different indexes are independent.
 * Search results follow ragarc_dense_topk (descending score, ties by ascending row, -1 padding).
 */
/* Page-locked host memory for the host-buffer forms below (cudaHostAlloc / cudaFreeHost): with
 * pinned query and result buffers the copies of ragarc_index_search are direct DMA transfers
 * (3 MB of fp32 queries in ~60 us over PCIe 5) instead of being staged through the driver's own
 * bounce buffers, which costs several hundred microseconds per call for pageable memory. */
int ragarc_host_alloc(size_t bytes, void** out_host);
int ragarc_host_free(void* host);

typedef struct ragarc_index ragarc_index_t;
int ragarc_index_create(int d, int dtype, int metric, ragarc_index_t** out);
int ragarc_index_free(ragarc_index_t* index);
int ragarc_index_reserve(ragarc_index_t* index, int64_t capacity, void* stream);
int ragarc_index_add(ragarc_index_t* index, const float* rows, int64_t n, int rows_on_host, void* stream);
int ragarc_index_search(ragarc_index_t* index, const float* queries, int nq, int k, float* out_scores,
                        int64_t* out_ids, int buffers_on_host, void* stream);
/* drops the given rows (host array, any order, duplicates allowed); survivors keep their order and
 * are renumbered densely, exactly as faiss remove_ids does */
int ragarc_index_remove(ragarc_index_t* index, const int64_t* rows_host, int64_t n_remove, void* stream);
int64_t ragarc_index_ntotal(const ragarc_index_t* index);
int ragarc_index_dim(const ragarc_index_t* index);
/* device pointer to the [ntotal, d] row matrix in the storage dtype (valid until the next add /
 * reserve / remove), e.g. as the candidate source of ragarc_mmr_select */
const void* ragarc_index_rows(const ragarc_index_t* index);

/* ---------------------------------------------------------------------------------------------
 * Row-sharded flat index driven by one host process (the multi-GPU form of the object above; the
 * reference has no distributed code).  Shard g is a flat index on CUDA device devices[g] (NULL =
 * devices 0..n_shards-1; a device may appear more than once) holding a contiguous range of global
 * rows: the first add splits its rows into ceil(n/G)-row ranges, later adds extend the last shard.
 * A search copies the fp32 host queries to every shard, runs the fused scoring + selection there
 * concurrently (one stream per shard), gathers the per-shard packed keys (nq*k*8 bytes each) on
 * shard 0's device and merges them with the same key order as a single index - the result is
 * identical for any number of shards.  All buffers are HOST memory; calls are synchronous.
 * All three metrics are offered; with RAGARC_METRIC_L2 the shards exchange the kept values
 * q.x - ||x||^2/2 (same order as ascending distance) and the merged values become squared distances.
 * (One process per GPU with a collective in between - rag_arc_b200/sharded.py - is the other form:
 * ragarc_dense_topk_keys per rank, then ragarc_merge_topk_keys[_p2p].)
 */
typedef struct ragarc_sharded_index ragarc_sharded_index_t;
int ragarc_sharded_create(int d, int dtype, int metric, int n_shards, const int* devices,
                          ragarc_sharded_index_t** out);
int ragarc_sharded_free(ragarc_sharded_index_t* index);
int ragarc_sharded_add(ragarc_sharded_index_t* index, const float* rows_host, int64_t n);
int ragarc_sharded_search(ragarc_sharded_index_t* index, const float* queries_host, int nq, int k,
                          float* out_scores_host, int64_t* out_ids_host);
int64_t ragarc_sharded_ntotal(const ragarc_sharded_index_t* index);

/* ---------------------------------------------------------------------------------------------
 * Exact squared-L2 search.   Replaces faiss.IndexFlatL2 (VectorStore_Faiss.py:125-126 metric "l2";
 *   search at :263 returns squared distances ascending) on the SAME scoring + selection kernels:
 *   ||q - x||^2 = ||q||^2 - 2 (q.x - ||x||^2/2).  ragarc_l2_augment stores a row as [x | -||x||^2/2]
 *   (half precision: the extra term as three pieces h1+h2+h3, padded with zeros to a multiple of 8
 *   columns) and a query as [q | 1] ([q | 1 1 1 0..]); ragarc_dense_topk over the augmented matrices
 *   (width ragarc_l2_aug_dim(d, dtype)) ranks nearest first; ragarc_l2_distances turns the k kept
 *   values into distances in place (clamped at 0, +inf for -1 padding).  ||x||^2 is taken over the
 *   stored (rounded) values.  sqnorm_out (may be NULL): fp32 [n] squared norms.
 *   normalize != 0 applies faiss.normalize_L2 first (the reference's normalize_L2 flag, :150-154).
 */
int ragarc_l2_aug_dim(int d, int dtype);
int ragarc_l2_augment(const float* src, void* dst, int64_t n, int d, int dst_dtype, int is_query,
                      int normalize, float* sqnorm_out, void* stream);
int ragarc_l2_distances(float* scores, const void* queries_aug, int dtype, int nq, int k, int d,
                        void* stream);

/* ---------------------------------------------------------------------------------------------
 * One process per GPU, NCCL between them (the reference has no distributed code).
 * A communicator is created either from a unique id the host distributes to its ranks
 * (ragarc_comm_unique_id on one rank, ragarc_comm_init_rank on every rank with its CUDA device
 * current) or, for one process driving several GPUs, by ragarc_comm_init_all.  NCCL is bound at
 * run time (libnccl.so.2, the copy already loaded into the process is reused); without it the calls
 * fail with RAGARC_ERR_CUDA.  ragarc_comm_nccl_version() returns NCCL's version code (0 = not loadable).
 * ragarc_sharded_topk: every rank holds a contiguous row shard (id_base = global id of its first
 * row) and the same queries; local fused scoring + selection -> ncclAllGather of the [nq,k] packed
 * keys -> merge; every rank gets the global result, identical to the single-GPU one for any number
 * of ranks.  Collective: all ranks must call it with the same nq and k, on their own streams. */
typedef struct ragarc_comm ragarc_comm_t;
int ragarc_comm_nccl_version(void);
int ragarc_comm_unique_id(char* out128_host);
int ragarc_comm_init_rank(const char* id128_host, int nranks, int rank, ragarc_comm_t** out);
int ragarc_comm_init_all(int ndev, const int* devices, ragarc_comm_t** out_comms);
int ragarc_comm_free(ragarc_comm_t* comm);
int ragarc_comm_rank(const ragarc_comm_t* comm);
int ragarc_comm_nranks(const ragarc_comm_t* comm);
size_t ragarc_sharded_topk_workspace_bytes(int64_t n_local, int d, int dtype, int nq, int k, int nranks);
int ragarc_sharded_topk(ragarc_comm_t* comm, const void* corpus_shard, int64_t n_local, int d, int dtype,
                        const void* queries, int nq, int k, uint64_t id_base, float* out_scores,
                        int64_t* out_ids, void* workspace, size_t workspace_bytes, void* stream);

/* fp32-accurate search on the tensor cores ("bf16x3").  An fp32 vector v is stored as three bf16
 * planes v1+v2+v3 (v1 = bf16(v), v2 = bf16(v-v1), v3 = bf16(v-v1-v2); exact to 2^-24 relative), a
 * row being [v1 | v2 | v3] (3*d bf16).  The inner product is accumulated in fp32 over the six
 * largest plane-pair products - six passes of the bf16 kernel - and matches the fp32 inner product
 * the reference computes (IndexFlatIP holds fp32, VectorStore_Faiss.py:170) to ~1e-6, at 1/6 of the
 * bf16 rate instead of CUDA-core rate.  d must be a multiple of 64.
 * ragarc_normalize_split3: fp32 [n,d] -> (optionally L2-normalised, faiss semantics) planes [n,3d]. */
int ragarc_normalize_split3(const float* src, void* dst_planes, int64_t n, int d, int normalize,
                            void* stream);
size_t ragarc_dense_topk_x3_workspace_bytes(int64_t n, int d, int nq, int k);
int ragarc_dense_topk_x3(const void* corpus_planes, int64_t n, int d, const void* query_planes,
                         int nq, int k, float* out_scores, int64_t* out_ids, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Same search, but returns packed sortable keys for the multi-GPU merge:
 *   key = (orderable_fp32(score) << 32) | (0xFFFFFFFF - (id_base + row)),  0 = empty slot,
 * sorted descending per query.  One shard per GPU; id_base = first global row of the shard. */
int ragarc_dense_topk_keys(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                           int nq, int k, uint64_t id_base, uint64_t* out_keys, void* workspace,
                           size_t workspace_bytes, int path, int* path_used_host, void* stream);

/* The general form: the two phases of a search as separate calls, so that a host can overlap the
 * SELECTION of batch i (merge of the per-slice candidate lists: small, latency-bound, fits beside the
 * persistent scoring CTAs on every SM) with the SCORING of batch i+1 on another stream.
 *   phase RAGARC_PHASE_SCORE : scoring + fused per-slice selection into the workspace, no output
 *   phase RAGARC_PHASE_SELECT: merge of what a SCORE call with the same (n, d, dtype, nq, k) left in the
 *                              same workspace -> outputs; corpus / queries are not read (may be NULL)
 *   phase RAGARC_PHASE_BOTH  : the two back to back (what ragarc_dense_topk does)
 * Outputs (SELECT / BOTH): out_scores + out_ids, and/or out_keys (packed keys carrying id_base + row),
 * or - inboxes != NULL - the query-owner push of ragarc_dense_topk_keys_push.  The caller orders a
 * SELECT after its SCORE and keeps the workspace untouched in between (use one workspace per batch in
 * flight). */
enum { RAGARC_PHASE_BOTH = 0, RAGARC_PHASE_SCORE = 1, RAGARC_PHASE_SELECT = 2 };
typedef struct ragarc_dense_opts {
  int phase;
  uint64_t id_base;
  uint64_t* out_keys;             /* [nq,k] or NULL */
  float* out_scores;              /* [nq,k] or NULL */
  int64_t* out_ids;               /* [nq,k] or NULL */
  uint64_t* const* inboxes;       /* DEVICE array of n_ranks inbox pointers, or NULL */
  int n_ranks, rank, nq_per_rank, signal;
  int workspace_clean;            /* SCORE only: a SELECT ran on this workspace since its last SCORE (SELECT resets
                                     the thresholds behind its merge), so the scoring kernel needs no memset in front */
} ragarc_dense_opts_t;
int ragarc_dense_topk_ex(const void* corpus, int64_t n, int d, int dtype, const void* queries, int nq, int k,
                         const ragarc_dense_opts_t* opts, void* workspace, size_t workspace_bytes, int path,
                         int* path_used_host, void* stream);

/* Same search for the query-owner exchange of the multi-GPU path (rag_arc_b200/sharded.py): the
 * sorted key row of query q is written straight into the inbox of the rank that owns the query,
 *   inboxes[q / nq_per_rank] + ((size_t)rank * nq_per_rank + q % nq_per_rank) * k,
 * where inboxes is a DEVICE array of n_ranks pointers, each to a [n_ranks, nq_per_rank, k] key block
 * that may live in a peer GPU's memory (stores travel over NVLink inside the merge kernel).  Every
 * rank then merges only its own nq_per_rank queries instead of all nq.
 * signal == 0: the caller orders the stores against the owner's merge itself (a device-side barrier,
 * then ragarc_merge_topk_keys on the inbox).
 * signal != 0: every inbox is followed by nq_per_rank uint32 arrival counters (zero before first use,
 *   at inbox + n_ranks*nq_per_rank*k keys); after a query's row has landed, the pushing CTA adds 1 to
 *   the owner's counter of that query (release, system scope) and the owner calls
 *   ragarc_merge_topk_inbox, whose CTA for query q waits for n_ranks arrivals, resets the counter and
 *   merges - no barrier and no second round trip between the two kernels.  Use two inboxes alternately:
 *   a rank can be at most one search ahead of the slowest one. */
int ragarc_dense_topk_keys_push(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                                int nq, int k, uint64_t id_base, uint64_t* const* inboxes, int n_ranks,
                                int rank, int nq_per_rank, int signal, void* workspace,
                                size_t workspace_bytes, int path, int* path_used_host, void* stream);
/* Owner side of the signalled exchange: inbox = this rank's [n_ranks, nq_per_rank, k_in] key block +
 * counters; nq_own <= nq_per_rank = queries of the batch this rank actually owns (the last ranks may
 * own fewer, or none).  out_*: [nq_own, k_out].  *status (device uint32, may be NULL) gets bit 0 set
 * if a wait gave up after timeout_ms (<= 0: 2000 ms) - the results of that search are then incomplete. */
int ragarc_merge_topk_inbox(uint64_t* inbox, int n_ranks, int nq_per_rank, int nq_own, int k_in, int k_out,
                            float* out_scores, int64_t* out_ids, double timeout_ms, uint32_t* status,
                            void* stream);

/* k*G-way merge after the NCCL all-gather of ragarc_dense_topk_keys outputs.
 * keys: [nlists, nq, k_in] (as all-gathered, rank-major).  Output as ragarc_dense_topk. */
int ragarc_merge_topk_keys(const uint64_t* keys, int nlists, int nq, int k_in, int k_out,
                           float* out_scores, int64_t* out_ids, void* stream);

/* Same merge without the all-gather: key_ptrs is a DEVICE array of nlists pointers, entry g
 * pointing at shard g's [nq, k_in] key block - which may live in a PEER GPU's memory mapped into
 * this process (NVLink loads inside the merge kernel; rag_arc_b200/sharded.py sets the table up
 * from symmetric memory handles and orders the accesses with a device-side barrier). */
int ragarc_merge_topk_keys_p2p(const uint64_t* const* key_ptrs, int nlists, int nq, int k_in,
                               int k_out, float* out_scores, int64_t* out_ids, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BM25 (Okapi) scoring over CSR postings.   Replaces rank_bm25.BM25Okapi.get_scores at
 *   core/retrieval/bm25.py:306,332 and np.argsort(scores)[::-1][:k] at :309,:359.
 * indptr[V+1] (int64), post_doc/post_tf[nnz] (int32; doc ids unique inside one posting list),
 * idf[V], doc_norm[n_docs] = k1*(1-b+b*dl/avgdl) (fp64, computed by the host exactly as the
 * reference expression does), k1_plus_1 = k1+1.
 * post_val[nnz] (fp64, may be NULL): the query-independent factor tf*(k1+1)/(tf+doc_norm[doc]) of
 * every posting, precomputed by the host with numpy in the reference's operation order; when given,
 * the kernels multiply it by idf[t] instead of redoing the fp64 division and the doc_norm gather
 * per query (bit-identical either way).
 * q_terms: [nq,tmax] term ids in query order, duplicates repeated, -1 = token not in the
 * vocabulary (contributes 0) ; q_len[nq].
 * score[d] += idf[t] * ( tf*(k1+1) / (tf + doc_norm[d]) ), fp64, each operation individually
 * rounded (no FMA), terms accumulated in query order => bit-identical to the numpy expression.
 * Top-k order: descending score, ties by ascending doc id (numpy's order is unspecified).
 */
size_t ragarc_bm25_workspace_bytes(int64_t n_docs, int nq);
int ragarc_bm25_scores(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                       const double* post_val, const double* idf, const double* doc_norm, double k1_plus_1,
                       const int32_t* q_terms, const int32_t* q_len, int nq, int tmax,
                       int64_t n_docs, double* out_scores /* [nq,n_docs] */, void* stream);
int ragarc_bm25_topk(const int64_t* indptr, const int32_t* post_doc, const int32_t* post_tf,
                     const double* post_val, const double* idf, const double* doc_norm, double k1_plus_1,
                     const int32_t* q_terms, const int32_t* q_len, int nq, int tmax,
                     int64_t n_docs, int k, double* out_scores /* [nq,k] */,
                     int64_t* out_ids /* [nq,k] */, void* workspace, size_t workspace_bytes,
                     void* stream);

/* Host-side query encoding for ragarc_bm25_topk / ragarc_bm25_scores: the batch form of what the
 * reference does per query in Python - preprocess_func(query) = query.split() (core/retrieval/
 * bm25.py:16-25,:302) and rank_bm25's per-token dictionary lookups (inside get_scores, :306).
 * ragarc_vocab_create copies n_tokens vocabulary strings (token i = tokens_blob[offsets[i],
 * offsets[i+1]), UTF-8, id = i; a repeated string keeps its first id).  ragarc_vocab_encode_split
 * tokenises text q = texts_blob[offsets[q], offsets[q+1]) on whitespace exactly as Python's
 * str.split() does (every character with str.isspace(), UTF-8 aware), looks every token up and
 * writes out_terms[q, 0..tmax) (-1 = out of vocabulary, and padding) and out_len[q] = min(tokens, tmax);
 * *max_len (may be NULL) = the longest query's token count - call again with a larger tmax if it
 * exceeds tmax.  All pointers are HOST memory (use pinned buffers to overlap the upload). */
typedef struct ragarc_vocab ragarc_vocab_t;
int ragarc_vocab_create(const char* tokens_blob, const int64_t* offsets, int64_t n_tokens, ragarc_vocab_t** out);
int ragarc_vocab_free(ragarc_vocab_t* vocab);
int64_t ragarc_vocab_size(const ragarc_vocab_t* vocab);
int ragarc_vocab_encode_split(const ragarc_vocab_t* vocab, const char* texts_blob, const int64_t* offsets,
                              int nq, int tmax, int32_t* out_terms_host, int32_t* out_len_host,
                              int* max_len_host);
/* Same, the nq texts given as ONE buffer of blob_bytes bytes separated by NUL bytes (what
 * "\0".join(texts).encode() produces: one allocation on the caller's side instead of nq). */
int ragarc_vocab_encode_split0(const ragarc_vocab_t* vocab, const char* texts_blob, int64_t blob_bytes,
                               int nq, int tmax, int32_t* out_terms_host, int32_t* out_len_host,
                               int* max_len_host);

/* Doc-range sharded BM25 (multi-GPU): every shard scores its own documents with the GLOBAL idf and
 * average length (the reference computes both over the whole corpus, bm25.py:218) and reports global
 * doc ids; this merges the per-shard results scores/ids [n_lists, nq, k_in] (ids -1 = padding) into
 * the global top-k_out of each query with the same order rule (n_lists*k_in <= 8192). */
int ragarc_bm25_merge_topk(const double* scores, const int64_t* ids, int n_lists, int nq, int k_in,
                           int k_out, double* out_scores, int64_t* out_ids, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reciprocal-rank fusion.   Replaces RRFusion.fuse at core/utils/Fusion.py:45-76 for a batch.
 * ids: [L, nq, kl] int32 document keys, rank = position+1, negative = padding (skipped).
 * score[key] = sum over lists (in list order) of 1.0/(rrf_k + rank) in fp64; output sorted by
 * descending score, ties by first appearance (list order, then position) - the order Python's
 * stable sorted(reverse=True) produces.  out_ids/out_scores: [nq, top_k], padded with -1 / 0;
 * out_count[nq] = number of valid entries.
 */
int ragarc_rrf_fuse(const int32_t* ids, int n_lists, int nq, int kl, double rrf_k, int top_k,
                    int32_t* out_ids, double* out_scores, int32_t* out_count, void* stream);

/* The hybrid merge of MultiPathRetriever._get_relevant_documents (core/retrieval/mutipath.py:57-93) in one
 * launch, on the retrievers' own result ROWS: list l arrives as rows[l] = device int64 [nq, kl_each[l]]
 * (corpus rows of retriever l, -1 = padding; a NULL pointer = that retriever returned nothing), the
 * content key of a row is row_to_key[l][row] (device int32, one entry per corpus row of retriever l; equal
 * contents share a key across retrievers, which is how the reference de-duplicates: Fusion.py:57-61).
 * Fusion as ragarc_rrf_fuse with kl columns per list (shorter lists are padded).  Besides the fused keys,
 * scores and counts, every fused entry gets the Document the reference returns for it - the one at the LAST
 * position holding the key, because document_map[content] is overwritten while the lists are walked in
 * order (Fusion.py:61): out_list[nq, top_k] = its list, out_row[nq, top_k] = its row in that list's corpus,
 * both -1 past out_count.  rows / kl_each / row_to_key are HOST arrays of n_lists entries (<= 8). */
int ragarc_rrf_fuse_rows(const int64_t* const* rows, const int* kl_each, const int32_t* const* row_to_key,
                         int n_lists, int nq, int kl, double rrf_k, int top_k, int32_t* out_ids, double* out_scores,
                         int32_t* out_count, int32_t* out_list, int64_t* out_row, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Pool + L2-normalise encoder outputs.   Replaces the Pooling/Normalize tail of
 *   SentenceTransformer.encode behind core/file_management/embeddings/huggingface.py:122-126.
 * x: [B,T,H] of `dtype`; mask: [B,T] int32 (0/1); out: fp32 [B,H].
 * mean: sum_t m*x / max(sum_t m, 1e-9); cls: x[:,0]; last: last unmasked token.
 * normalize!=0: out / max(||out||_2, 1e-12).
 */
int ragarc_pool_normalize(const void* x, int dtype, const int32_t* mask, int B, int T, int H,
                          int mode, int normalize, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Greedy maximal-marginal-relevance selection.   Replaces _mmr_select at
 *   VectorStore_Faiss.py:16-62 without re-embedding the candidates (:301-304): candidate rows
 *   are gathered from the resident corpus.
 * cand_rows: [nq, fetch_k] int64 corpus rows (from ragarc_dense_topk; -1 = padding).
 * out_sel: [nq, k] int32 indices INTO the candidate list (the reference returns
 * docs_and_scores[idx]), -1 padded.
 */
int ragarc_mmr_select(const void* corpus, int64_t n, int d, int dtype, const void* queries,
                      int nq, const int64_t* cand_rows, int fetch_k, int k, double lambda_mult,
                      int32_t* out_sel, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cosine distance between consecutive rows.   Replaces calculate_cosine_distances at
 *   core/file_management/chunker/spliter.py:354-372 (SemanticChunker), which calls the row-wise
 *   cosine_similarity of :307-333 once per adjacent pair in Python.
 * x: [n,d] fp32 or fp64 row-major; out[i] = 1 - <x_i,x_{i+1}> / (|x_i| |x_{i+1}|) for i < n-1,
 * evaluated in fp64 like the numpy fallback of the reference; a nan/inf similarity (zero row)
 * counts as 0, i.e. distance 1.
 */
int ragarc_adjacent_cosine_distance(const void* x, int dtype, int64_t n, int d, double* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Reranker scoring tail.   Replaces Qwen3Reranker.compute_logits after the LM forward at
 *   core/rerank/Reranker_Qwen3.py:44-49: pick the "yes"/"no" logits of the last position,
 *   log_softmax over the two, exp of the "yes" entry.
 * logits: [B, vocab] rows of the last position (row_stride elements apart), in `dtype`; the two
 * intermediate results are rounded to `dtype` as the reference's tensor library does; out: fp32 [B] (values exactly
 * representable in `dtype`).
 */
int ragarc_yes_no_score(const void* logits, int dtype, int B, int64_t row_stride, int vocab,
                        int true_id, int false_id, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAGARC_B200_H */
