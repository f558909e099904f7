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

extern "C" int splade_emit_json(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, char* out, size_t out_cap, int64_t* out_offsets,
                                size_t* bytes_needed) {
  if (!ids_host || !vals_host || !counts_host || !out_offsets || !bytes_needed || (!out && out_cap))
    return set_error(kErrInvalidArg, "splade_emit_json: null pointer argument");
  if (B < 0 || k < 1 || round_decimals > 9)
    return set_error(kErrInvalidArg, "splade_emit_json: bad arguments B=%d k=%d decimals=%d", B, k, round_decimals);
  size_t o = 0;
  bool overflow = false;
  char num[64];
  auto put_raw = [&](const char* s, size_t n) {
    if (o + n <= out_cap) memcpy(out + o, s, n);
    else overflow = true;
    o += n;
  };
  for (int b = 0; b < B; ++b) {
    out_offsets[b] = static_cast<int64_t>(o);
    const int n = counts_host[b] < k ? counts_host[b] : k;
    put_raw("{", 1);
    for (int j = 0; j < n; ++j) {
      const int32_t id = ids_host[static_cast<size_t>(b) * k + j];
      const float w = vals_host[static_cast<size_t>(b) * k + j];
      if (j) put_raw(", ", 2);
      if (token_bytes && token_offsets) {
        if (id < 0 || id >= n_tokens)
          return set_error(kErrInvalidArg, "splade_emit_json: term id %d outside the token table (%d)", id, n_tokens);
        const char* s = token_bytes + token_offsets[id];
        const size_t len = static_cast<size_t>(token_offsets[id + 1] - token_offsets[id]);
        const long wrote = put_json_string(o <= out_cap ? out + o : out, o <= out_cap ? out_cap - o : 0, s, len);
        if (wrote < 0) {
          overflow = true;
          // count the bytes anyway: worst case 6 per byte + 2
          size_t need = 2;
          for (size_t i = 0; i < len; ++i) {
            const unsigned char c = static_cast<unsigned char>(s[i]);
            need += (c == '"' || c == '\\' || c == '\n' || c == '\r' || c == '\t') ? 2 : (c < 0x20 ? 6 : 1);
          }
          o += need;
        } else {
          o += static_cast<size_t>(wrote);
        }
      } else {
        const int m = snprintf(num, sizeof(num), "\"%d\"", id);
        put_raw(num, static_cast<size_t>(m));
      }
      put_raw(": ", 2);
      int m;
      if (round_decimals >= 0) m = snprintf(num, sizeof(num), "%.*f", round_decimals, static_cast<double>(w));
      else m = snprintf(num, sizeof(num), "%.9g", static_cast<double>(w));  // round-trips every fp32
      put_raw(num, static_cast<size_t>(m));
    }
    put_raw("}", 1);
  }
  out_offsets[B] = static_cast<int64_t>(o);
  *bytes_needed = o;
  if (overflow) return set_error(kErrWorkspace, "splade_emit_json: output buffer too small (%zu < %zu bytes)", out_cap, o);
  return kOk;
}
