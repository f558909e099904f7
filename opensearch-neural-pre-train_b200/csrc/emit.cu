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
  // JSON string escaping of UTF-8 bytes (no quotes): ", \\ and control characters; everything else passes through
  void escaped(const char* s, size_t len) {
    size_t run = 0;  // pending bytes that need no escaping
    for (size_t i = 0; i < len; ++i) {
      const unsigned char c = static_cast<unsigned char>(s[i]);
      if (c != '"' && c != '\\' && c >= 0x20) {
        ++run;
        continue;
      }
      if (run) raw(s + i - run, run);
      run = 0;
      switch (c) {
        case '"': raw("\\\"", 2); break;
        case '\\': raw("\\\\", 2); break;
        case '\n': raw("\\n", 2); break;
        case '\r': raw("\\r", 2); break;
        case '\t': raw("\\t", 2); break;
        default: {
          char buf[8];
          snprintf(buf, sizeof(buf), "\\u%04x", c);
          raw(buf, 6);
        }
      }
    }
    if (run) raw(s + len - run, run);
  }
  void json_string(const char* s, size_t len) {
    raw("\"", 1);
    escaped(s, len);
    raw("\"", 1);
  }
  // "<prefix>.<token>" as ONE JSON string (rank_feature field paths)
  void json_field_path(const char* prefix, size_t plen, const char* tok, size_t tlen) {
    raw("\"", 1);
    escaped(prefix, plen);
    raw(".", 1);
    escaped(tok, tlen);
    raw("\"", 1);
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

extern "C" int splade_emit_search(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                  int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                  int max_clauses, int size, const char* field_name, const char* msearch_index,
                                  char* out, size_t out_cap, int64_t* out_offsets, size_t* bytes_needed) {
  if (!ids_host || !vals_host || !counts_host || !out_offsets || !bytes_needed || (!out && out_cap) || !field_name)
    return set_error(kErrInvalidArg, "splade_emit_search: null pointer argument");
  if (B < 0 || k < 1 || size < 0 || max_clauses < 1)
    return set_error(kErrInvalidArg, "splade_emit_search: bad arguments B=%d k=%d size=%d max_clauses=%d", B, k, size,
                     max_clauses);
  if (!token_bytes || !token_offsets)
    return set_error(kErrInvalidArg, "splade_emit_search: rank_feature field paths need the token table");
  Sink w{out, out_cap};
  const size_t field_len = strlen(field_name);
  char num[64];
  for (int b = 0; b < B; ++b) {
    out_offsets[b] = static_cast<int64_t>(w.o);
    int n = counts_host[b] < k ? (counts_host[b] > 0 ? counts_host[b] : 0) : k;
    if (n > max_clauses) n = max_clauses;
    if (n == 0) continue;  // `if not query_sparse: return SearchResponse(results=[], ...)`: no request at all
    if (msearch_index) {
      w.lit("{\"index\":");
      w.json_string(msearch_index, strlen(msearch_index));
      w.lit("}\n");
    }
    int m = snprintf(num, sizeof(num), "{\"size\":%d,\"query\":{\"bool\":{\"should\":[", size);
    w.raw(num, static_cast<size_t>(m));
    for (int j = 0; j < n; ++j) {
      const int32_t id = ids_host[static_cast<size_t>(b) * k + j];
      if (id < 0 || id >= n_tokens)
        return set_error(kErrInvalidArg, "splade_emit_search: term id %d outside the token table (%d)", id, n_tokens);
      if (j) w.raw(",", 1);
      w.lit("{\"rank_feature\":{\"field\":");
      w.json_field_path(field_name, field_len, token_bytes + token_offsets[id],
                        static_cast<size_t>(token_offsets[id + 1] - token_offsets[id]));
      w.lit(",\"boost\":");
      m = snprintf(num, sizeof(num), "%.9g", static_cast<double>(vals_host[static_cast<size_t>(b) * k + j]));
      w.raw(num, static_cast<size_t>(m));
      w.lit("}}");
    }
    w.lit("]}}}");
    if (msearch_index) w.raw("\n", 1);
  }
  out_offsets[B] = static_cast<int64_t>(w.o);
  *bytes_needed = w.o;
  if (w.overflow)
    return set_error(kErrWorkspace, "splade_emit_search: output buffer too small (%zu < %zu bytes)", out_cap, w.o);
  return kOk;
}
