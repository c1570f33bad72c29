// Host-side query encoding for BM25 behind the C ABI: whitespace tokenisation (Python str.split()
// semantics on UTF-8) + vocabulary lookup + packing into the [nq, tmax] int32 layout ragarc_bm25_topk
// takes.  Replaces, for a whole batch in one call, what the reference does per query in Python:
// preprocess_func(query) (core/retrieval/bm25.py:16-25, :302) followed by rank_bm25's per-token
// dictionary lookups inside get_scores (called at :306).  No device code here; the file is part of
// the library so that a non-Python host gets the same entry points.
#include <cstring>
#include <new>
#include <string>
#include <vector>
#include "common.cuh"

// Open-addressing table over the token strings: one 8-byte slot per probe (tag + id), string compare only
// on a tag match; the strings themselves sit back to back in `blob`.  A node-based std::unordered_map costs
// two dependent cache misses per token; here the slots of all tokens of a query are prefetched while the
// query is being split, so the probes of a query overlap.
struct ragarc_vocab {
  struct Slot { uint32_t tag; int32_t id; };                     // id < 0: empty
  std::string blob;                                              // all tokens back to back
  std::vector<int64_t> off;                                      // [n+1] token i = blob[off[i], off[i+1])
  std::vector<Slot> table;                                       // power-of-two size, load factor <= 0.5
  uint64_t mask = 0;
  int64_t distinct = 0;

  static inline uint64_t hash(const unsigned char* p, size_t n) {
    uint64_t h = 0xcbf29ce484222325ull;                          // FNV-1a, finished with a multiply-shift mix
    for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
    h ^= h >> 32; h *= 0x9E3779B97F4A7C15ull; h ^= h >> 29;
    return h;
  }
  inline int32_t find(const unsigned char* p, size_t n, uint64_t h) const {
    if (table.empty()) return -1;
    const uint32_t tag = (uint32_t)(h >> 32);
    for (uint64_t i = h & mask;; i = (i + 1) & mask) {
      const Slot s = table[i];
      if (s.id < 0) return -1;
      if (s.tag == tag && (size_t)(off[s.id + 1] - off[s.id]) == n && memcmp(blob.data() + off[s.id], p, n) == 0) return s.id;
    }
  }
};

namespace ragarc {

// length in bytes of the whitespace character starting at p (0 = not whitespace): the characters for
// which Python's str.isspace() is true, in UTF-8.  One table lookup settles every byte except the four
// lead bytes a multi-byte space can start with.
struct SpaceTable {
  unsigned char cls[256];                  // 0 = not a space, 1 = single-byte space, 2 = possible multi-byte space lead
  constexpr SpaceTable() : cls{} {
    for (int c = 0; c < 256; ++c) {
      const bool one = c == ' ' || (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x1F);
      cls[c] = one ? 1 : ((c == 0xC2 || c == 0xE1 || c == 0xE2 || c == 0xE3) ? 2 : 0);
    }
  }
};
static constexpr SpaceTable kSpace{};

static inline int wide_space_len(const unsigned char* p, const unsigned char* end) {
  const unsigned char c = *p;
  if (c == 0xC2) return (p + 1 < end && (p[1] == 0x85 || p[1] == 0xA0)) ? 2 : 0;
  if (p + 2 < end) {
    if (c == 0xE1 && p[1] == 0x9A && p[2] == 0x80) return 3;                                   // U+1680
    if (c == 0xE2 && p[1] == 0x80 && ((p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF)) return 3;
    if (c == 0xE2 && p[1] == 0x81 && p[2] == 0x9F) return 3;                                   // U+205F
    if (c == 0xE3 && p[1] == 0x80 && p[2] == 0x80) return 3;                                   // U+3000
  }
  return 0;
}

static inline int space_len(const unsigned char* p, const unsigned char* end) {
  const unsigned char k = kSpace.cls[*p];
  return k < 2 ? k : wide_space_len(p, end);
}

}  // namespace ragarc

using namespace ragarc;

extern "C" {

int ragarc_vocab_create(const char* tokens_blob, const int64_t* offsets, int64_t n_tokens, ragarc_vocab_t** out) {
  RA_REQUIRE(out != nullptr, "vocab_create: out is NULL");
  *out = nullptr;
  RA_REQUIRE(n_tokens >= 0 && (n_tokens == 0 || (tokens_blob && offsets)), "vocab_create: bad arguments");
  ragarc_vocab* v = new (std::nothrow) ragarc_vocab();
  RA_REQUIRE(v != nullptr, "vocab_create: out of host memory");
  try {
    if (n_tokens > 0) v->blob.assign(tokens_blob + offsets[0], (size_t)(offsets[n_tokens] - offsets[0]));
    v->off.resize((size_t)n_tokens + 1);
    for (int64_t i = 0; i <= n_tokens; ++i) v->off[i] = n_tokens > 0 ? offsets[i] - offsets[0] : 0;
    uint64_t cap = 16;
    while (cap < (uint64_t)n_tokens * 2) cap <<= 1;
    v->table.assign(cap, ragarc_vocab::Slot{0u, -1});
    v->mask = cap - 1;
    const unsigned char* base = (const unsigned char*)v->blob.data();
    for (int64_t i = 0; i < n_tokens; ++i) {
      const unsigned char* p = base + v->off[i];
      const size_t n = (size_t)(v->off[i + 1] - v->off[i]);
      const uint64_t h = ragarc_vocab::hash(p, n);
      if (v->find(p, n, h) >= 0) continue;                       // first occurrence wins, like dict.setdefault
      uint64_t j = h & v->mask;
      while (v->table[j].id >= 0) j = (j + 1) & v->mask;
      v->table[j] = ragarc_vocab::Slot{(uint32_t)(h >> 32), (int32_t)i};
      ++v->distinct;
    }
  } catch (...) {
    delete v;
    set_error("vocab_create: out of host memory");
    return RAGARC_ERR_INVALID;
  }
  *out = v;
  return RAGARC_OK;
}

int ragarc_vocab_free(ragarc_vocab_t* v) {
  delete v;
  return RAGARC_OK;
}

int64_t ragarc_vocab_size(const ragarc_vocab_t* v) { return v ? v->distinct : -1; }

}  // extern "C"

namespace ragarc {
// one text: split, look up, pack; returns its token count.  Two passes over at most 64 tokens at a time:
// the first splits, hashes and prefetches the table slots, the second probes.
static inline int encode_one(const ragarc_vocab* v, const unsigned char* p, const unsigned char* end, int32_t* row, int tmax) {
  constexpr int CHUNK = 64;
  const unsigned char* tok[CHUNK];
  uint32_t len[CHUNK];
  uint64_t hsh[CHUNK];
  int count = 0;
  while (p < end) {
    int m = 0;
    while (p < end && m < CHUNK) {
      int sl;
      while (p < end && (sl = space_len(p, end)) > 0) p += sl;
      if (p >= end) break;
      const unsigned char* t0 = p;
      while (p < end && space_len(p, end) == 0) ++p;
      if (count + m < tmax) {
        tok[m] = t0; len[m] = (uint32_t)(p - t0);
        hsh[m] = ragarc_vocab::hash(t0, len[m]);
        if (!v->table.empty()) __builtin_prefetch(&v->table[hsh[m] & v->mask]);
      }
      ++m;
    }
    for (int i = 0; i < m && count + i < tmax; ++i)
      row[count + i] = v->find(tok[i], len[i], hsh[i]);          // -1 = out of vocabulary: contributes 0
    count += m;
  }
  for (int i = count < tmax ? count : tmax; i < tmax; ++i) row[i] = -1;
  return count;
}
}  // namespace ragarc

extern "C" {

int ragarc_vocab_encode_split0(const ragarc_vocab_t* v, const char* texts_blob, int64_t blob_bytes, int nq, int tmax,
                               int32_t* out_terms_host, int32_t* out_len_host, int* max_len_host) {
  RA_REQUIRE(v != nullptr, "vocab_encode_split0: null vocabulary");
  RA_REQUIRE(nq >= 0 && tmax > 0 && blob_bytes >= 0, "vocab_encode_split0: nq=%d tmax=%d", nq, tmax);
  RA_REQUIRE(nq == 0 || (texts_blob && out_terms_host && out_len_host), "vocab_encode_split0: null pointer");
  const unsigned char* p = (const unsigned char*)texts_blob;
  const unsigned char* blob_end = p + blob_bytes;
  int longest = 0;
  for (int q = 0; q < nq; ++q) {
    const unsigned char* end = (const unsigned char*)memchr(p, 0, (size_t)(blob_end - p));
    if (!end) end = blob_end;
    RA_REQUIRE(q == nq - 1 || end < blob_end, "vocab_encode_split0: fewer than %d NUL-separated texts in the blob", nq);
    const int count = encode_one(v, p, end, out_terms_host + (size_t)q * tmax, tmax);
    out_len_host[q] = count < tmax ? count : tmax;
    if (count > longest) longest = count;
    p = end < blob_end ? end + 1 : blob_end;
  }
  if (max_len_host) *max_len_host = longest;
  return RAGARC_OK;
}

int ragarc_vocab_encode_split(const ragarc_vocab_t* v, const char* texts_blob, const int64_t* offsets, int nq, int tmax,
                              int32_t* out_terms_host, int32_t* out_len_host, int* max_len_host) {
  RA_REQUIRE(v != nullptr, "vocab_encode_split: null vocabulary");
  RA_REQUIRE(nq >= 0 && tmax > 0, "vocab_encode_split: nq=%d tmax=%d", nq, tmax);
  RA_REQUIRE(nq == 0 || (texts_blob && offsets && out_terms_host && out_len_host), "vocab_encode_split: null pointer");
  int longest = 0;
  for (int q = 0; q < nq; ++q) {
    const unsigned char* p = (const unsigned char*)texts_blob + offsets[q];
    const unsigned char* end = (const unsigned char*)texts_blob + offsets[q + 1];
    int32_t* row = out_terms_host + (size_t)q * tmax;
    const int count0 = encode_one(v, p, end, row, tmax);
    out_len_host[q] = count0 < tmax ? count0 : tmax;
    if (count0 > longest) longest = count0;
  }
  if (max_len_host) *max_len_host = longest;
  return RAGARC_OK;
}

}  // extern "C"
