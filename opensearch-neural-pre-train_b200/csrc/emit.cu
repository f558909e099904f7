// Host-side sparse emit: compact top-k buffers -> JSON objects for OpenSearch bulk indexing.
//
// The reference builds, per document, a python dict {token_string: weight} (benchmark/encoders.py:320-343) that
// benchmark/indexer.py:112-138 sends as the `sparse_embedding` rank_features field, or {"<term id>": round(w, 4)}
// for the sparse_vector field type (scripts/neural_sparse_search_aws.py:745-770).  After the GPU top-k only
// [B,k] ids and weights reach the host; this routine serialises them straight into one byte buffer (no python
// objects per term).  Pure host code (no device work); lives in the same library so that one C-ABI covers the
// path from hidden states to bulk lines.
#include "../../include/splade_b200.h"
#include "common.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace splade {
namespace {

// Append a JSON string literal (with quotes) for a UTF-8 token; returns bytes written or -1 on overflow.
long put_json_string(char* dst, size_t cap, const char* s, size_t n) {
  size_t o = 0;
  auto put = [&](char c) {
    if (o < cap) dst[o] = c;
    ++o;
  };
  put('"');
  for (size_t i = 0; i < n; ++i) {
    const unsigned char c = static_cast<unsigned char>(s[i]);
    switch (c) {
      case '"': put('\\'); put('"'); break;
      case '\\': put('\\'); put('\\'); break;
      case '\n': put('\\'); put('n'); break;
      case '\r': put('\\'); put('r'); break;
      case '\t': put('\\'); put('t'); break;
      default:
        if (c < 0x20) {
          char buf[8];
          snprintf(buf, sizeof(buf), "\\u%04x", c);
          for (int j = 0; j < 6; ++j) put(buf[j]);
        } else {
          put(static_cast<char>(c));  // UTF-8 bytes pass through
        }
    }
  }
  put('"');
  return o <= cap ? static_cast<long>(o) : -1;
}

}  // namespace
}  // namespace splade

using namespace splade;

namespace {

// Bounded byte sink that keeps counting after it is full, so one pass yields either the output or its size.
struct Sink {
  char* out;
  size_t cap;
  size_t o = 0;
  bool overflow = false;
  void raw(const char* s, size_t n) {
    if (o + n <= cap) memcpy(out + o, s, n);
    else overflow = true;
    o += n;
  }
  void lit(const char* s) { raw(s, strlen(s)); }
  void json_string(const char* s, size_t len) {
    const long wrote = splade::put_json_string(o <= cap ? out + o : out, o <= cap ? cap - o : 0, s, len);
    if (wrote < 0) {
      overflow = true;
      size_t need = 2;  // count the bytes anyway: worst case 6 per byte + the quotes
      for (size_t i = 0; i < len; ++i) {
        const unsigned char c = static_cast<unsigned char>(s[i]);
        need += (c == '"' || c == '\\' || c == '\n' || c == '\r' || c == '\t') ? 2 : (c < 0x20 ? 6 : 1);
      }
      o += need;
    } else {
      o += static_cast<size_t>(wrote);
    }
  }
};

// One document's {"term": weight, ...}.  Returns 0 or an error code (term id outside the token table).
int emit_object(Sink& w, const int32_t* ids, const float* vals, int n, const char* token_bytes,
                const int64_t* token_offsets, int n_tokens, int round_decimals, bool compact, const char* who) {
  char num[64];
  w.raw("{", 1);
  bool first = true;
  for (int j = 0; j < n; ++j) {
    const int32_t id = ids[j];
    int m;
    if (round_decimals >= 0) {
      m = snprintf(num, sizeof(num), "%.*f", round_decimals, static_cast<double>(vals[j]));
      // the rounded format drops what rounds to zero: `weight = round(w, 4); if weight <= 0.0: continue`
      // (scripts/neural_sparse_search_aws.py:753-755)
      if (strtod(num, nullptr) <= 0.0) continue;
    } else {
      m = snprintf(num, sizeof(num), "%.9g", static_cast<double>(vals[j]));  // round-trips every fp32
    }
    if (!first) w.raw(", ", compact ? 1 : 2);
    first = false;
    if (token_bytes && token_offsets) {
      if (id < 0 || id >= n_tokens)
        return splade::set_error(splade::kErrInvalidArg, "%s: term id %d outside the token table (%d)", who, id,
                                 n_tokens);
      w.json_string(token_bytes + token_offsets[id], static_cast<size_t>(token_offsets[id + 1] - token_offsets[id]));
    } else {
      char key[32];
      const int mk = snprintf(key, sizeof(key), "\"%d\"", id);
      w.raw(key, static_cast<size_t>(mk));
    }
    w.raw(": ", compact ? 1 : 2);
    w.raw(num, static_cast<size_t>(m));
  }
  w.raw("}", 1);
  return splade::kOk;
}

}  // namespace

extern "C" int splade_emit_json(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, char* out, size_t out_cap, int64_t* out_offsets,
                                size_t* bytes_needed) {
  if (!ids_host || !vals_host || !counts_host || !out_offsets || !bytes_needed || (!out && out_cap))
    return set_error(kErrInvalidArg, "splade_emit_json: null pointer argument");
  if (B < 0 || k < 1 || round_decimals > 9)
    return set_error(kErrInvalidArg, "splade_emit_json: bad arguments B=%d k=%d decimals=%d", B, k, round_decimals);
  Sink w{out, out_cap};
  for (int b = 0; b < B; ++b) {
    out_offsets[b] = static_cast<int64_t>(w.o);
    const int n = counts_host[b] < k ? (counts_host[b] > 0 ? counts_host[b] : 0) : k;
    const int rc = emit_object(w, ids_host + static_cast<size_t>(b) * k, vals_host + static_cast<size_t>(b) * k, n,
                               token_bytes, token_offsets, n_tokens, round_decimals, /*compact=*/false,
                               "splade_emit_json");
    if (rc != kOk) return rc;
  }
  out_offsets[B] = static_cast<int64_t>(w.o);
  *bytes_needed = w.o;
  if (w.overflow)
    return set_error(kErrWorkspace, "splade_emit_json: output buffer too small (%zu < %zu bytes)", out_cap, w.o);
  return kOk;
}

extern "C" int splade_emit_bulk(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, const char* index_name, const char* field_name,
                                const char* id_bytes, const int64_t* id_offsets, const char* head_bytes,
                                const int64_t* head_offsets, int skip_empty, char* out, size_t out_cap,
                                int64_t* out_offsets, size_t* bytes_needed) {
  if (!ids_host || !vals_host || !counts_host || !out_offsets || !bytes_needed || (!out && out_cap) || !index_name ||
      !field_name || !id_bytes || !id_offsets)
    return set_error(kErrInvalidArg, "splade_emit_bulk: null pointer argument");
  if (B < 0 || k < 1 || round_decimals > 9)
    return set_error(kErrInvalidArg, "splade_emit_bulk: bad arguments B=%d k=%d decimals=%d", B, k, round_decimals);
  if ((head_bytes == nullptr) != (head_offsets == nullptr))
    return set_error(kErrInvalidArg, "splade_emit_bulk: head_bytes and head_offsets go together");
  Sink w{out, out_cap};
  const size_t index_len = strlen(index_name), field_len = strlen(field_name);
  for (int b = 0; b < B; ++b) {
    out_offsets[b] = static_cast<int64_t>(w.o);
    const int n = counts_host[b] < k ? (counts_host[b] > 0 ? counts_host[b] : 0) : k;
    if (skip_empty) {  // `if not clean_vec: continue` (scripts/neural_sparse_search_aws.py:997-1001), after rounding
      int alive = 0;
      char num[64];
      for (int j = 0; j < n && !alive; ++j) {
        if (round_decimals < 0) {
          alive = 1;
        } else {
          snprintf(num, sizeof(num), "%.*f", round_decimals, static_cast<double>(vals_host[static_cast<size_t>(b) * k + j]));
          alive = strtod(num, nullptr) > 0.0;
        }
      }
      if (!alive) continue;
    }
    // action line: what opensearch-py's helpers.bulk expands {"_index", "_id", "_source"} into
    w.lit("{\"index\":{\"_index\":");
    w.json_string(index_name, index_len);
    w.lit(",\"_id\":");
    w.json_string(id_bytes + id_offsets[b], static_cast<size_t>(id_offsets[b + 1] - id_offsets[b]));
    w.lit("}}\n{");
    // source line: caller-serialised leading members, then the sparse field
    if (head_bytes && head_offsets[b + 1] > head_offsets[b]) {
      w.raw(head_bytes + head_offsets[b], static_cast<size_t>(head_offsets[b + 1] - head_offsets[b]));
      w.raw(",", 1);
    }
    w.json_string(field_name, field_len);
    w.raw(":", 1);
    const int rc = emit_object(w, ids_host + static_cast<size_t>(b) * k, vals_host + static_cast<size_t>(b) * k, n,
                               token_bytes, token_offsets, n_tokens, round_decimals, /*compact=*/true,
                               "splade_emit_bulk");
    if (rc != kOk) return rc;
    w.lit("}\n");
  }
  out_offsets[B] = static_cast<int64_t>(w.o);
  *bytes_needed = w.o;
  if (w.overflow)
    return set_error(kErrWorkspace, "splade_emit_bulk: output buffer too small (%zu < %zu bytes)", out_cap, w.o);
  return kOk;
}
