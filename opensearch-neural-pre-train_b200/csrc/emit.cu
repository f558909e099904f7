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

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

using namespace splade;

namespace {

// ------------------------------------------------------------------------------------------------ numbers
// The emitters print fp32 weights either with a fixed number of decimals (printf "%.{d}f": the AWS script's
// round(w, 4)) or as "%.9g" (round-trips every fp32).  snprintf + strtod per term was 90 % of the emit time
// (30 MB/s on one thread), so both formats are produced with exact integer arithmetic instead:
//   v = m * 2^e (m < 2^24)  =>  v * 10^P = m * 5^P * 2^(e+P),   m * 5^P < 2^24 * 5^12 < 2^52   for P <= 12
// and one right shift with round-half-even on the exact remainder gives the correctly rounded decimal digits -
// what glibc's printf computes.  Anything outside the fast range (non-finite, negative, huge, exponent notation)
// goes through snprintf, so the output bytes are identical to the printf formats in every case.
const uint64_t kPow10[13] = {1ull,          10ull,          100ull,          1000ull,         10000ull,
                             100000ull,     1000000ull,     10000000ull,     100000000ull,    1000000000ull,
                             10000000000ull, 100000000000ull, 1000000000000ull};
const uint64_t kPow5[13] = {1ull,      5ull,      25ull,      125ull,      625ull,       3125ull,     15625ull,
                            78125ull,  390625ull, 1953125ull, 9765625ull,  48828125ull,  244140625ull};

bool force_printf() {
  static const bool f = [] {
    const char* e = std::getenv("SPLADE_B200_EMIT_PRINTF");
    return e && e[0] == '1';
  }();
  return f;
}

// round_half_even(v * 10^P) for a positive finite fp32 v and 0 <= P <= 12; false when it does not fit 63 bits
inline bool scaled_round(float v, int P, uint64_t* q_out) {
  uint32_t bits;
  memcpy(&bits, &v, 4);
  const uint32_t ex = (bits >> 23) & 0xFFu;
  uint64_t m = bits & 0x7FFFFFu;
  int e;
  if (ex == 0) {
    e = -149;  // subnormal
  } else {
    m |= 0x800000u;
    e = static_cast<int>(ex) - 150;
  }
  const uint64_t N = m * kPow5[P];  // < 2^52
  const int sh = -(e + P);          // v * 10^P = N / 2^sh
  if (sh <= 0) {
    if (-sh > 11) return false;     // N << (-sh) must stay below 2^63
    *q_out = N << (-sh);
    return true;
  }
  if (sh >= 64) {
    *q_out = 0;
    return true;
  }
  uint64_t q = N >> sh;
  const uint64_t rem = N & ((1ull << sh) - 1), half = 1ull << (sh - 1);
  if (rem > half || (rem == half && (q & 1ull))) ++q;
  *q_out = q;
  return true;
}

inline int put_uint(char* dst, uint64_t x) {  // decimal digits of x, returns their count
  char tmp[24];
  int n = 0;
  do {
    tmp[n++] = static_cast<char>('0' + x % 10);
    x /= 10;
  } while (x);
  for (int i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
  return n;
}

// q with `decimals` fractional digits: "<int>.<frac zero-padded>"; strip = drop trailing zeros (and a bare '.')
inline int put_fixed(char* dst, uint64_t q, int decimals, bool strip) {
  const uint64_t ip = q / kPow10[decimals];
  uint64_t fp = q % kPow10[decimals];
  int n = put_uint(dst, ip);
  if (decimals == 0) return n;
  int nd = decimals;
  if (strip) {
    while (nd > 0 && fp % 10 == 0) {
      fp /= 10;
      --nd;
    }
    if (nd == 0) return n;
  }
  dst[n++] = '.';
  for (int i = nd - 1; i >= 0; --i) {
    dst[n + i] = static_cast<char>('0' + fp % 10);
    fp /= 10;
  }
  return n + nd;
}

// printf("%.{decimals}f", v); *is_zero = the printed value is 0 (the rounded format drops such terms)
inline int format_fixed(char* dst, size_t cap, float v, int decimals, bool* is_zero) {
  uint64_t q;
  if (!force_printf() && v > 0.f && v < 3.0e38f && decimals <= 9 && scaled_round(v, decimals, &q)) {
    *is_zero = q == 0;
    return put_fixed(dst, q, decimals, false);
  }
  const int m = snprintf(dst, cap, "%.*f", decimals, static_cast<double>(v));
  *is_zero = !(strtod(dst, nullptr) > 0.0) && !(strtod(dst, nullptr) < 0.0);
  return m;
}

// printf("%.9g", v)
inline int format_g9(char* dst, size_t cap, float v) {
  if (!force_printf() && v >= 1.0e-4f && v < 1.0e9f) {
    // decimal exponent E with 10^E <= v < 10^(E+1), then 9 significant digits = round(v * 10^(8-E))
    int E = static_cast<int>(floorf(log10f(v)));
    if (E < -4) E = -4;
    if (E > 8) E = 8;
    for (int attempt = 0; attempt < 3; ++attempt) {
      uint64_t q;
      const int P = 8 - E;
      if (P < 0 || P > 12 || !scaled_round(v, P, &q)) break;
      if (q >= kPow10[9]) {  // 10^(E+1) <= v (or the rounding carried): one digit more in front
        ++E;
        continue;
      }
      if (q < kPow10[8]) {
        --E;
        continue;
      }
      if (E < -4 || E > 8) break;  // printf switches to exponent notation
      return put_fixed(dst, q, P, true);
    }
  }
  return snprintf(dst, cap, "%.9g", static_cast<double>(v));
}

// ------------------------------------------------------------------------------------------------ byte sink
// Either a bounded caller buffer that keeps counting after it is full (one pass yields the output or its size),
// or a growing private buffer (the worker threads of the parallel emitters).
struct Sink {
  char* out = nullptr;
  size_t cap = 0;
  size_t o = 0;
  bool overflow = false;
  bool grow = false;
  Sink(char* out_, size_t cap_) : out(out_), cap(cap_) {}
  explicit Sink(size_t reserve) : grow(true) {
    cap = reserve < 4096 ? 4096 : reserve;
    out = static_cast<char*>(malloc(cap));
    if (!out) cap = 0, overflow = true;
  }
  Sink(const Sink&) = delete;
  Sink& operator=(const Sink&) = delete;
  ~Sink() {
    if (grow) free(out);
  }
  void raw(const char* s, size_t n) {
    if (o + n > cap && grow && !overflow) {
      size_t want = cap * 2;
      while (want < o + n) want *= 2;
      char* p = static_cast<char*>(realloc(out, want));
      if (p) out = p, cap = want;
      else overflow = true;
    }
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
      // the rounded format drops what rounds to zero: `weight = round(w, 4); if weight <= 0.0: continue`
      // (scripts/neural_sparse_search_aws.py:753-755)
      bool zero;
      m = format_fixed(num, sizeof(num), vals[j], round_decimals, &zero);
      if (zero || vals[j] < 0.f) continue;
    } else {
      m = format_g9(num, sizeof(num), vals[j]);  // round-trips every fp32
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
      int mk = 0;
      key[mk++] = '"';
      if (id < 0) {
        key[mk++] = '-';
        mk += put_uint(key + mk, static_cast<uint64_t>(-static_cast<int64_t>(id)));
      } else {
        mk += put_uint(key + mk, static_cast<uint64_t>(id));
      }
      key[mk++] = '"';
      w.raw(key, static_cast<size_t>(mk));
    }
    w.raw(": ", compact ? 1 : 2);
    w.raw(num, static_cast<size_t>(m));
  }
  w.raw("}", 1);
  return splade::kOk;
}

// ------------------------------------------------------------------------------------------------ threads
std::atomic<int> g_emit_threads{0};  // 0 = automatic

int emit_threads_for(int B) {
  int t = g_emit_threads.load();
  if (t <= 0) {
    const char* e = std::getenv("SPLADE_B200_EMIT_THREADS");
    t = e ? std::atoi(e) : 0;
  }
  if (t <= 0) {
    t = static_cast<int>(std::thread::hardware_concurrency());
    if (t > 16) t = 16;
  }
  const int by_work = B / 512;  // a thread is worth starting for a few hundred documents
  if (t > by_work) t = by_work;
  return t < 1 ? 1 : t;
}

// Runs doc(w, b) for b in [0,B): in the caller's sink directly, or - when the output buffer is known to be writable
// (cap > 0) and the batch is large - over contiguous document ranges on worker threads, each into a private growing
// buffer, concatenated in range order afterwards (byte-identical to the serial pass).
template <typename DocFn>
int run_documents(Sink& w, int B, int64_t* out_offsets, const DocFn& doc) {
  const int T = (w.cap > 0) ? emit_threads_for(B) : 1;
  if (T <= 1) {
    for (int b = 0; b < B; ++b) {
      out_offsets[b] = static_cast<int64_t>(w.o);
      const int rc = doc(w, b);
      if (rc != kOk) return rc;
    }
    out_offsets[B] = static_cast<int64_t>(w.o);
    return kOk;
  }
  struct Part {
    Sink* sink = nullptr;
    int rc = kOk;
    int b0 = 0, b1 = 0;
  };
  std::vector<Part> parts(static_cast<size_t>(T));
  std::vector<std::thread> pool;
  const size_t guess = w.cap / static_cast<size_t>(T) + 4096;
  for (int t = 0; t < T; ++t) {
    Part& p = parts[static_cast<size_t>(t)];
    p.b0 = static_cast<int>(static_cast<int64_t>(B) * t / T);
    p.b1 = static_cast<int>(static_cast<int64_t>(B) * (t + 1) / T);
    p.sink = new Sink(guess);
    auto work = [&p, &doc, out_offsets] {
      for (int b = p.b0; b < p.b1; ++b) {
        out_offsets[b] = static_cast<int64_t>(p.sink->o);  // relative to the part; rebased below
        p.rc = doc(*p.sink, b);
        if (p.rc != kOk) return;
      }
    };
    try {
      pool.emplace_back(work);
    } catch (...) {  // no thread to be had (std::system_error must not cross the C boundary): this range runs here
      work();
    }
  }
  for (auto& th : pool) th.join();
  int rc = kOk;
  for (int t = 0; t < T && rc == kOk; ++t) {
    Part& p = parts[static_cast<size_t>(t)];
    if (p.rc != kOk) rc = p.rc;  // (its message sits in the worker's thread-local slot: repeated below)
    else if (p.sink->overflow) rc = -1;  // a private buffer could not grow
  }
  if (rc == kOk) {
    // rebase the per-document offsets and copy the parts into place (in parallel: the copy is bandwidth-bound)
    std::vector<size_t> base(static_cast<size_t>(T) + 1, w.o);
    for (int t = 0; t < T; ++t) base[static_cast<size_t>(t) + 1] = base[static_cast<size_t>(t)] + parts[static_cast<size_t>(t)].sink->o;
    const size_t total = base[static_cast<size_t>(T)];
    const bool fits = total <= w.cap;
    std::vector<std::thread> copiers;
    for (int t = 0; t < T; ++t) {
      Part& p = parts[static_cast<size_t>(t)];
      const size_t off = base[static_cast<size_t>(t)];
      auto copy = [&p, off, fits, &w, out_offsets] {
        for (int b = p.b0; b < p.b1; ++b) out_offsets[b] += static_cast<int64_t>(off);
        if (fits && p.sink->o) memcpy(w.out + off, p.sink->out, p.sink->o);
      };
      try {
        copiers.emplace_back(copy);
      } catch (...) {
        copy();
      }
    }
    for (auto& th : copiers) th.join();
    if (!fits) w.overflow = true;
    w.o = total;
    out_offsets[B] = static_cast<int64_t>(w.o);
  }
  for (auto& p : parts) delete p.sink;
  if (rc == -1) return set_error(kErrWorkspace, "emit: out of host memory for the worker buffers");
  if (rc != kOk) {
    // a worker's set_error went to its own thread-local slot: rerun serially so the caller's thread holds the message
    Sink probe(nullptr, 0);
    for (int b = 0; b < B; ++b) {
      const int r = doc(probe, b);
      if (r != kOk) return r;
    }
  }
  return rc;
}

}  // namespace

extern "C" int splade_emit_set_threads(int n) {
  const int prev = g_emit_threads.exchange(n < 0 ? 0 : n);
  return prev;
}

static int splade_emit_json_impl(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, char* out, size_t out_cap, int64_t* out_offsets,
                                size_t* bytes_needed) {
  if (!ids_host || !vals_host || !counts_host || !out_offsets || !bytes_needed || (!out && out_cap))
    return set_error(kErrInvalidArg, "splade_emit_json: null pointer argument");
  if (B < 0 || k < 1 || round_decimals > 9)
    return set_error(kErrInvalidArg, "splade_emit_json: bad arguments B=%d k=%d decimals=%d", B, k, round_decimals);
  Sink w{out, out_cap};
  const int rc_all = run_documents(w, B, out_offsets, [&](Sink& sk, int b) {
    const int n = counts_host[b] < k ? (counts_host[b] > 0 ? counts_host[b] : 0) : k;
    return emit_object(sk, ids_host + static_cast<size_t>(b) * k, vals_host + static_cast<size_t>(b) * k, n,
                       token_bytes, token_offsets, n_tokens, round_decimals, /*compact=*/false, "splade_emit_json");
  });
  if (rc_all != kOk) return rc_all;
  *bytes_needed = w.o;
  if (w.overflow)
    return set_error(kErrWorkspace, "splade_emit_json: output buffer too small (%zu < %zu bytes)", out_cap, w.o);
  return kOk;
}

extern "C" int splade_emit_json(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, char* out, size_t out_cap, int64_t* out_offsets,
                                size_t* bytes_needed) {
  try {
    return splade_emit_json_impl(ids_host, vals_host, counts_host, B, k, token_bytes, token_offsets, n_tokens, round_decimals, out, out_cap, out_offsets, bytes_needed);
  } catch (const std::bad_alloc&) {
    return set_error(kErrWorkspace, "splade_emit_json: out of host memory");
  } catch (...) {
    return set_error(kErrInvalidArg, "splade_emit_json: unexpected C++ exception");
  }
}


static int splade_emit_bulk_impl(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
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
  const int rc_all = run_documents(w, B, out_offsets, [&](Sink& sk, int b) -> int {
    const int n = counts_host[b] < k ? (counts_host[b] > 0 ? counts_host[b] : 0) : k;
    if (skip_empty) {  // `if not clean_vec: continue` (scripts/neural_sparse_search_aws.py:997-1001), after rounding
      bool alive = false;
      char num[64];
      for (int j = 0; j < n && !alive; ++j) {
        const float v = vals_host[static_cast<size_t>(b) * k + j];
        if (round_decimals < 0) {
          alive = true;
        } else {
          bool zero;
          format_fixed(num, sizeof(num), v, round_decimals, &zero);
          alive = !zero && v > 0.f;
        }
      }
      if (!alive) return kOk;
    }
    // action line: what opensearch-py's helpers.bulk expands {"_index", "_id", "_source"} into
    sk.lit("{\"index\":{\"_index\":");
    sk.json_string(index_name, index_len);
    sk.lit(",\"_id\":");
    sk.json_string(id_bytes + id_offsets[b], static_cast<size_t>(id_offsets[b + 1] - id_offsets[b]));
    sk.lit("}}\n{");
    // source line: caller-serialised leading members, then the sparse field
    if (head_bytes && head_offsets[b + 1] > head_offsets[b]) {
      sk.raw(head_bytes + head_offsets[b], static_cast<size_t>(head_offsets[b + 1] - head_offsets[b]));
      sk.raw(",", 1);
    }
    sk.json_string(field_name, field_len);
    sk.raw(":", 1);
    const int rc = emit_object(sk, ids_host + static_cast<size_t>(b) * k, vals_host + static_cast<size_t>(b) * k, n,
                               token_bytes, token_offsets, n_tokens, round_decimals, /*compact=*/true,
                               "splade_emit_bulk");
    if (rc != kOk) return rc;
    sk.lit("}\n");
    return kOk;
  });
  if (rc_all != kOk) return rc_all;
  *bytes_needed = w.o;
  if (w.overflow)
    return set_error(kErrWorkspace, "splade_emit_bulk: output buffer too small (%zu < %zu bytes)", out_cap, w.o);
  return kOk;
}

extern "C" int splade_emit_bulk(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                int round_decimals, const char* index_name, const char* field_name,
                                const char* id_bytes, const int64_t* id_offsets, const char* head_bytes,
                                const int64_t* head_offsets, int skip_empty, char* out, size_t out_cap,
                                int64_t* out_offsets, size_t* bytes_needed) {
  try {
    return splade_emit_bulk_impl(ids_host, vals_host, counts_host, B, k, token_bytes, token_offsets, n_tokens, round_decimals, index_name, field_name, id_bytes, id_offsets, head_bytes, head_offsets, skip_empty, out, out_cap, out_offsets, bytes_needed);
  } catch (const std::bad_alloc&) {
    return set_error(kErrWorkspace, "splade_emit_bulk: out of host memory");
  } catch (...) {
    return set_error(kErrInvalidArg, "splade_emit_bulk: unexpected C++ exception");
  }
}


static int splade_emit_search_impl(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
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
      m = format_g9(num, sizeof(num), vals_host[static_cast<size_t>(b) * k + j]);
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

extern "C" int splade_emit_search(const int32_t* ids_host, const float* vals_host, const int32_t* counts_host, int B,
                                  int k, const char* token_bytes, const int64_t* token_offsets, int n_tokens,
                                  int max_clauses, int size, const char* field_name, const char* msearch_index,
                                  char* out, size_t out_cap, int64_t* out_offsets, size_t* bytes_needed) {
  try {
    return splade_emit_search_impl(ids_host, vals_host, counts_host, B, k, token_bytes, token_offsets, n_tokens, max_clauses, size, field_name, msearch_index, out, out_cap, out_offsets, bytes_needed);
  } catch (const std::bad_alloc&) {
    return set_error(kErrWorkspace, "splade_emit_search: out of host memory");
  } catch (...) {
    return set_error(kErrInvalidArg, "splade_emit_search: unexpected C++ exception");
  }
}

