/* A host that is neither Python nor PyTorch: plain C against include/ragarc_b200.h and
 * libragarc_b200.so.  Builds a small fp32 cosine index from host memory, searches it with host
 * buffers, removes rows, and checks every result against a brute-force scan in this file.
 * Exit code 0 = all checks passed (tests/test_gpu_c_abi.py compiles and runs it on the GPU box). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ragarc_b200.h"

#define N 3001
#define D 96
#define NQ 7
#define K 5

static unsigned long long state = 88172645463325252ull;
static float rnd(void) {                      /* xorshift: deterministic, no libc rand() differences */
  state ^= state << 13; state ^= state >> 7; state ^= state << 17;
  return (float)((state >> 11) & 0xFFFFF) / (float)0x100000 - 0.5f;
}

static void normalize(float* v) {
  double s = 0.0;
  for (int i = 0; i < D; ++i) s += (double)v[i] * v[i];
  if (s > 0.0) { float inv = (float)(1.0 / sqrt(s)); for (int i = 0; i < D; ++i) v[i] *= inv; }
}

#define CHECK(call) do { int rc_ = (call); if (rc_ != RAGARC_OK) { \
  fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, ragarc_last_error()); return 2; } } while (0)

static int brute_check(const float* X, const int* alive, const float* Q, const float* Ds, const int64_t* Is,
                       const int64_t* renum) {
  for (int q = 0; q < NQ; ++q) {
    for (int j = 0; j < K; ++j) {
      /* j-th best by brute force: best score strictly below the previous pick (ties by lower row) */
      double best = -1e30; int besti = -1;
      for (int r = 0; r < N; ++r) {
        if (!alive[r]) continue;
        int taken = 0;
        for (int t = 0; t < j; ++t) if (Is[q * K + t] == renum[r]) taken = 1;
        if (taken) continue;
        double s = 0.0;
        for (int i = 0; i < D; ++i) s += (double)X[(size_t)r * D + i] * Q[q * D + i];
        if (s > best + 1e-7) { best = s; besti = r; }
      }
      if (Is[q * K + j] != renum[besti] || fabs(Ds[q * K + j] - best) > 1e-5) {
        fprintf(stderr, "query %d rank %d: got row %lld score %.7f, expected row %lld score %.7f\n", q, j,
                (long long)Is[q * K + j], Ds[q * K + j], (long long)renum[besti], best);
        return 1;
      }
    }
  }
  return 0;
}

int main(void) {
  float* X = (float*)malloc(sizeof(float) * N * D);
  float* Xn = (float*)malloc(sizeof(float) * N * D);
  float Q[NQ * D], Qn[NQ * D], Ds[NQ * K];
  int64_t Is[NQ * K];
  int* alive = (int*)malloc(sizeof(int) * N);
  int64_t* renum = (int64_t*)malloc(sizeof(int64_t) * N);
  for (int i = 0; i < N * D; ++i) X[i] = rnd() * 4.0f;
  memcpy(Xn, X, sizeof(float) * N * D);
  for (int r = 0; r < N; ++r) { normalize(Xn + (size_t)r * D); alive[r] = 1; renum[r] = r; }
  for (int q = 0; q < NQ; ++q) {
    for (int i = 0; i < D; ++i) Q[q * D + i] = X[(size_t)(q * 401 + 3) * D + i] + rnd();
    memcpy(Qn + q * D, Q + q * D, sizeof(float) * D);
    normalize(Qn + q * D);
  }
  if (ragarc_abi_version() != RAGARC_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }

  ragarc_index_t* ix = NULL;
  CHECK(ragarc_index_create(D, RAGARC_F32, RAGARC_METRIC_COSINE, &ix));
  CHECK(ragarc_index_add(ix, X, 1000, 1, NULL));                 /* un-normalised rows, two batches */
  CHECK(ragarc_index_add(ix, X + (size_t)1000 * D, N - 1000, 1, NULL));
  if (ragarc_index_ntotal(ix) != N || ragarc_index_dim(ix) != D) { fprintf(stderr, "ntotal/dim wrong\n"); return 1; }
  CHECK(ragarc_index_search(ix, Q, NQ, K, Ds, Is, 1, NULL));
  if (brute_check(Xn, alive, Qn, Ds, Is, renum)) return 1;
  for (int q = 0; q < NQ; ++q)
    if (Is[q * K] != q * 401 + 3) { fprintf(stderr, "planted neighbour of query %d not first\n", q); return 1; }

  /* remove the planted neighbour of query 0 and two more rows; survivors are renumbered densely */
  int64_t drop[3] = {3, 2000, 17};
  CHECK(ragarc_index_remove(ix, drop, 3, NULL));
  for (int i = 0; i < 3; ++i) alive[drop[i]] = 0;
  { int64_t next = 0; for (int r = 0; r < N; ++r) renum[r] = alive[r] ? next++ : -1; }
  if (ragarc_index_ntotal(ix) != N - 3) { fprintf(stderr, "ntotal after remove wrong\n"); return 1; }
  CHECK(ragarc_index_search(ix, Q, NQ, K, Ds, Is, 1, NULL));
  if (brute_check(Xn, alive, Qn, Ds, Is, renum)) return 1;

  /* errors come back as codes + message, nothing aborts */
  if (ragarc_index_search(ix, Q, NQ, 5000, Ds, Is, 1, NULL) == RAGARC_OK) { fprintf(stderr, "k=5000 accepted\n"); return 1; }
  if (strlen(ragarc_last_error()) == 0) { fprintf(stderr, "no error message\n"); return 1; }
  CHECK(ragarc_index_free(ix));

  /* the same rows in a three-shard index (all shards on device 0): identical answers */
  {
    ragarc_sharded_index_t* sh = NULL;
    int devs[3] = {0, 0, 0};
    float Ds2[NQ * K]; int64_t Is2[NQ * K];
    CHECK(ragarc_sharded_create(D, RAGARC_F32, RAGARC_METRIC_COSINE, 3, devs, &sh));
    CHECK(ragarc_sharded_add(sh, X, N));
    if (ragarc_sharded_ntotal(sh) != N) { fprintf(stderr, "sharded ntotal wrong\n"); return 1; }
    CHECK(ragarc_sharded_search(sh, Q, NQ, K, Ds2, Is2));
    for (int r = 0; r < N; ++r) { alive[r] = 1; renum[r] = r; }
    if (brute_check(Xn, alive, Qn, Ds2, Is2, renum)) return 1;
    CHECK(ragarc_sharded_free(sh));
  }
  printf("index_smoke: ok (launches so far: %llu)\n", (unsigned long long)ragarc_launch_count());
  free(X); free(Xn); free(alive); free(renum);
  return 0;
}
