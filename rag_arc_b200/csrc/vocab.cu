// Host-side query encoding for BM25 behind the C ABI: whitespace tokenisation (Python str.split()
// semantics on UTF-8) + vocabulary lookup + packing into the [nq, tmax] int32 layout ragarc_bm25_topk
// takes.  Replaces, for a whole batch in one call, what the reference does per query in Python:
// preprocess_func(query) (core/retrieval/bm25.py:16-25, :302) followed by rank_bm25's per-token
// dictionary lookups inside get_scores (called at :306).  No device code here; the file is part of
// the library so that a non-Python host gets the same entry points.
#include <cstring>
#include <new>
#include <string>
#include <string_view>
#include <unordered_map>
#include "common.cuh"

struct ragarc_vocab {
  std::string blob;                                              // all tokens back to back
  std::unordered_map<std::string_view, int32_t> map;            // views into blob
};

namespace ragarc {

// length in bytes of the whitespace character starting at p (0 = not whitespace): the characters for
// which Python's str.isspace() is true, in UTF-8
static inline int space_len(const unsigned char* p, const unsigned char* end) {
  const unsigned char c = *p;
  if (c == ' ' || (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x1F)) return 1;
  if (c < 0x80) return 0;
  if (c == 0xC2 && p + 1 < end && (p[1] == 0x85 || p[1] == 0xA0)) return 2;
  if (p + 2 < end) {
    if (c == 0xE1 && p[1] == 0x9A && p[2] == 0x80) return 3;                                   // U+1680
    if (c == 0xE2 && p[1] == 0x80 && ((p[2] >= 0x80 && p[2] <= 0x8A) || p[2] == 0xA8 || p[2] == 0xA9 || p[2] == 0xAF)) return 3;
    if (c == 0xE2 && p[1] == 0x81 && p[2] == 0x9F) return 3;                                   // U+205F
    if (c == 0xE3 && p[1] == 0x80 && p[2] == 0x80) return 3;                                   // U+3000
  }
  return 0;
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
    v->map.reserve((size_t)n_tokens * 2);
    const char* base = v->blob.data();
    for (int64_t i = 0; i < n_tokens; ++i) {
      const std::string_view tok(base + (offsets[i] - offsets[0]), (size_t)(offsets[i + 1] - offsets[i]));
      v->map.emplace(tok, (int32_t)i);                           // first occurrence wins, like dict.setdefault
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

int64_t ragarc_vocab_size(const ragarc_vocab_t* v) { return v ? (int64_t)v->map.size() : -1; }

}  // extern "C"

namespace ragarc {
// one text: split, look up, pack; returns its token count
static inline int encode_one(const ragarc_vocab* v, const unsigned char* p, const unsigned char* end, int32_t* row, int tmax) {
  int count = 0;
  while (p < end) {
    int sl;
    while (p < end && (sl = space_len(p, end)) > 0) p += sl;
    if (p >= end) break;
    const unsigned char* t0 = p;
    while (p < end && space_len(p, end) == 0) ++p;
    if (count < tmax) {
      const auto it = v->map.find(std::string_view((const char*)t0, (size_t)(p - t0)));
      row[count] = it == v->map.end() ? -1 : it->second;         // out of vocabulary: contributes 0
    }
    ++count;
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
