// ref_driver.cpp -- C entry points around the UNMODIFIED reference library.
//
// TEST INFRASTRUCTURE ONLY (see sj_oracle.h).  Compiled by oracle/Makefile
// together with /root/reference/singleheader/simdjson.cpp (from where it lies;
// no reference source is copied into this repo) into oracle/_ref/libsj_ref.so.
// It reaches the hot path exactly the way SURVEY.md section 8(c) describes:
//   get_available_implementations()[name]->create_dom_parser_implementation()
//   parser->stage1(buf,len,mode)   (include/simdjson/internal/dom_parser_implementation.h L80)
//   impl->minify / impl->validate_utf8 (include/simdjson/implementation.h L116, L128)
#include "simdjson.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

using namespace simdjson;

#define SJR_API extern "C" __attribute__((visibility("default")))

static const implementation *find_impl(const char *name) {
  if (name == nullptr || name[0] == 0) {
    // best CPU kernel this host supports (icelake > haswell > westmere > fallback)
    for (const char *n : {"icelake", "haswell", "westmere", "fallback"}) {
      auto impl = get_available_implementations()[n];
      if (impl && impl->supported_by_runtime_system()) return impl;
    }
    return nullptr;
  }
  auto impl = get_available_implementations()[name];
  if (!impl || !impl->supported_by_runtime_system()) return nullptr;
  return impl;
}

SJR_API int sjr_supported(const char *name) { return find_impl(name) != nullptr; }

SJR_API const char *sjr_best_name() {
  static std::string s;
  auto impl = find_impl(nullptr);
  s = impl ? impl->name() : "";
  return s.c_str();
}

// Returns the error_code, or -1 if the implementation is not usable here.
// idx_out receives min(max_words, ROUNDUP(capacity,64)+9) words of the
// parser's structural_indexes; *n_inout is loaded into the parser before the
// call (the reference leaves it untouched on its early-return paths) and read
// back after.
SJR_API int sjr_stage1(const char *name, const uint8_t *buf, size_t len, size_t capacity, int mode,
                       uint32_t *idx_out, size_t max_words, uint32_t *n_inout) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  std::unique_ptr<internal::dom_parser_implementation> p;
  auto err = impl->create_dom_parser_implementation(capacity, 1024, p);
  if (err) return int(err);
  p->n_structural_indexes = *n_inout;
  err = p->stage1(buf, len, stage1_mode(mode));
  *n_inout = p->n_structural_indexes;
  size_t words = SIMDJSON_ROUNDUP_N(capacity, 64) + 9;
  if (words > max_words) words = max_words;
  if (idx_out && words) std::memcpy(idx_out, p->structural_indexes.get(), words * sizeof(uint32_t));
  return int(err);
}

SJR_API int sjr_minify(const char *name, const uint8_t *buf, size_t len, uint8_t *dst, size_t *dst_len) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  size_t n = 0;
  auto err = impl->minify(buf, len, dst, n);
  *dst_len = n;
  return int(err);
}

SJR_API int sjr_validate_utf8(const char *name, const uint8_t *buf, size_t len) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  return impl->validate_utf8(reinterpret_cast<const char *>(buf), len) ? 1 : 0;
}

// ---------------------------------------------------------------- timing ---
// The reference has no intra-call parallelism: one
// stage-1 call runs on one core.  "All host threads" therefore means T
// independent parsers, each running the whole call on the same read-only
// buffer; aggregate bytes/s = T*len / wall.  op: 0 stage1, 1 minify, 2 utf8.
// Returns seconds for the best of `iters` rounds (each round = every thread
// doing one call), after one warm-up round; negative on error.
SJR_API double sjr_time(const char *name, int op, const uint8_t *buf, size_t len, int mode, int threads,
                        int iters, int *err_out) {
  auto impl = find_impl(name);
  if (!impl) return -1.0;
  if (threads < 1) threads = 1;
  struct worker_state {
    std::unique_ptr<internal::dom_parser_implementation> parser;
    std::unique_ptr<uint8_t[]> dst;
    int err{0};
  };
  std::vector<worker_state> ws(threads);
  for (auto &w : ws) {
    if (op == 0) {
      if (impl->create_dom_parser_implementation(len, 1024, w.parser)) return -2.0;
    } else if (op == 1) {
      w.dst.reset(new uint8_t[len + SIMDJSON_PADDING]);
    }
  }
  auto one_call = [&](worker_state &w) {
    if (op == 0) {
      w.err = int(w.parser->stage1(buf, len, stage1_mode(mode)));
    } else if (op == 1) {
      size_t n = 0;
      w.err = int(impl->minify(buf, len, w.dst.get(), n));
    } else {
      w.err = impl->validate_utf8(reinterpret_cast<const char *>(buf), len) ? 0 : int(UTF8_ERROR);
    }
  };
  double best = 1e30;
  for (int it = 0; it < iters + 1; it++) {
    std::atomic<int> ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> th;
    for (int t = 1; t < threads; t++) {
      th.emplace_back([&, t] {
        ready++;
        while (!go.load(std::memory_order_acquire)) {}
        one_call(ws[t]);
      });
    }
    while (ready.load() < threads - 1) {}
    auto t0 = std::chrono::steady_clock::now();
    go.store(true, std::memory_order_release);
    one_call(ws[0]);
    for (auto &t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    double s = std::chrono::duration<double>(t1 - t0).count();
    if (it > 0 && s < best) best = s;  // round 0 is the warm-up
  }
  if (err_out) *err_out = ws[0].err;
  return best;
}

// The same measurement with the hygiene a stable number needs: the T threads are created ONCE and reused for every
// round (thread creation is outside the timed window), every thread works on its OWN copy of the document (so the
// figure is bounded by DRAM bandwidth and not by how much of one shared buffer happens to sit in the last-level cache),
// a round starts on a flag all threads spin on and ends when the last one reports.  out[0] = best round, out[1] = mean
// round (seconds; every round = every thread doing one call).  Returns 0, or a negative value on error.
SJR_API int sjr_time_rounds(const char *name, int op, const uint8_t *buf, size_t len, int mode, int threads, int warmup, int iters,
                            double *out, int *err_out) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  if (threads < 1) threads = 1;
  struct worker_state {
    std::unique_ptr<internal::dom_parser_implementation> parser;
    std::unique_ptr<uint8_t[]> src, dst;
    int err{0};
  };
  std::vector<worker_state> ws(threads);
  std::atomic<int> round{0}, done{0}, ready{0};
  std::atomic<bool> failed{false};
  const int rounds = warmup + iters;
  auto body = [&](int t) {
    worker_state &w = ws[t];
    w.src.reset(new (std::nothrow) uint8_t[len + SIMDJSON_PADDING]);
    if (!w.src) failed = true;
    else { std::memcpy(w.src.get(), buf, len); std::memset(w.src.get() + len, 0x20, SIMDJSON_PADDING); }
    if (op == 0 && impl->create_dom_parser_implementation(len, 1024, w.parser)) failed = true;
    if (op == 1) { w.dst.reset(new (std::nothrow) uint8_t[len + SIMDJSON_PADDING]); if (!w.dst) failed = true; }
    ready++;
    for (int r = 1; r <= rounds; r++) {
      while (round.load(std::memory_order_acquire) < r) {}
      if (!failed) {
        if (op == 0) w.err = int(w.parser->stage1(w.src.get(), len, stage1_mode(mode)));
        else if (op == 1) { size_t n = 0; w.err = int(impl->minify(w.src.get(), len, w.dst.get(), n)); }
        else w.err = impl->validate_utf8(reinterpret_cast<const char *>(w.src.get()), len) ? 0 : int(UTF8_ERROR);
      }
      done.fetch_add(1, std::memory_order_release);
    }
  };
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) th.emplace_back(body, t);
  while (ready.load() < threads) std::this_thread::yield();
  double best = 1e30, sum = 0;
  for (int r = 1; r <= rounds; r++) {
    const auto t0 = std::chrono::steady_clock::now();
    round.store(r, std::memory_order_release);
    while (done.load(std::memory_order_acquire) < r * threads) {}
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (r > warmup) { sum += s; if (s < best) best = s; }
  }
  for (auto &t : th) t.join();
  if (failed) return -2;
  out[0] = best;
  out[1] = sum / iters;
  if (err_out) *err_out = ws[0].err;
  return 0;
}

// ------------------------------------------------- DOM-level reference hooks
// Used by the drop-in tests to get what dom::parser::parse / parse_many /
// simdjson::minify produce through a CPU implementation, for comparison with
// the same calls routed through the b200 plug-in.

// Parses one document with implementation `name` and writes its minified
// re-serialisation (simdjson::minify(element)) to out; returns error code.
SJR_API int sjr_dom_roundtrip(const char *name, const uint8_t *buf, size_t len, char *out, size_t out_cap,
                              size_t *out_len) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  const implementation *saved = get_active_implementation();
  get_active_implementation() = impl;
  dom::parser parser;
  dom::element doc;
  auto err = parser.parse(buf, len, true).get(doc);
  int rc = int(err);
  *out_len = 0;
  if (!err) {
    std::string s = simdjson::minify(doc);
    *out_len = s.size();
    if (s.size() <= out_cap) std::memcpy(out, s.data(), s.size());
  }
  get_active_implementation() = saved;
  return rc;
}

// parse_many: returns number of documents successfully iterated, writes first
// error (0 if none) and the concatenated minified documents separated by '\n'.
SJR_API long sjr_dom_parse_many(const char *name, const uint8_t *buf, size_t len, size_t batch_size, char *out,
                                size_t out_cap, size_t *out_len, int *first_err) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  const implementation *saved = get_active_implementation();
  get_active_implementation() = impl;
  long ndocs = 0;
  *first_err = 0;
  std::string acc;
  {
    dom::parser parser;
    dom::document_stream stream;
    auto err = parser.parse_many(buf, len, batch_size).get(stream);
    if (err) {
      *first_err = int(err);
    } else {
      for (auto it = stream.begin(); it != stream.end(); ++it) {
        auto doc = *it;
        if (doc.error()) { *first_err = int(doc.error()); break; }
        acc += simdjson::minify(doc.value_unsafe());
        acc.push_back('\n');
        ndocs++;
      }
    }
  }
  *out_len = acc.size();
  if (acc.size() <= out_cap) std::memcpy(out, acc.data(), acc.size());
  get_active_implementation() = saved;
  return ndocs;
}

// ---- stage-2-lite pinning: what the reference's stage 2 makes of tokens

// dom_parser_implementation::parse_string (include/simdjson/internal/dom_parser_implementation.h L124) on a string whose
// bytes AFTER the opening quote start at src (readable for SIMDJSON_PADDING bytes past the closing quote).
// Returns the unescaped length, -1 when the reference rejects the string.
SJR_API long sjr_parse_string(const char *name, const uint8_t *src, uint8_t *dst) {
  auto impl = find_impl(name);
  if (!impl) return -2;
  std::unique_ptr<internal::dom_parser_implementation> p;
  if (impl->create_dom_parser_implementation(64, 16, p)) return -2;
  uint8_t *end = p->parse_string(src, dst, false);
  return end ? long(end - dst) : -1;
}

// DOM-parse one document and dump its tape: for every tape word that starts a value, its type char and payload
// (strings: offset into string_buf; 'l' / 'u' / 'd': the 64-bit word that follows; containers: 0), in tape order without
// the root words; string_buf bytes up to the end of the last string record.  Returns the error_code of parse().
SJR_API int sjr_dom_tape(const char *name, const uint8_t *buf, size_t len, uint8_t *types, uint64_t *payloads, size_t max_entries,
                         size_t *n_entries, uint8_t *strbuf, size_t strbuf_cap, size_t *strbuf_len) {
  auto impl = find_impl(name);
  if (!impl) return -1;
  const implementation *saved = get_active_implementation();
  get_active_implementation() = impl;
  dom::parser parser;
  dom::element doc;
  auto err = parser.parse(buf, len, true).get(doc);
  *n_entries = 0;
  *strbuf_len = 0;
  if (!err) {
    const uint64_t *tape = parser.doc.tape.get();
    const size_t tape_len = size_t(tape[0] & 0x00FFFFFFFFFFFFFFull);  // root word: index one past the closing root word
    size_t k = 0, sb_end = 0;
    for (size_t i = 1; i + 1 < tape_len; i++) {
      const uint8_t t = uint8_t(tape[i] >> 56);
      uint64_t v = 0;
      if (t == '"') {
        v = tape[i] & 0x00FFFFFFFFFFFFFFull;
        uint32_t sl;
        std::memcpy(&sl, parser.doc.string_buf.get() + v, 4);
        if (v + 5 + sl > sb_end) sb_end = size_t(v) + 5 + sl;
      } else if (t == 'l' || t == 'u' || t == 'd') {
        v = tape[++i];
      }
      if (k < max_entries) { types[k] = t; payloads[k] = v; }
      k++;
    }
    *n_entries = k;
    *strbuf_len = sb_end;
    if (sb_end <= strbuf_cap) std::memcpy(strbuf, parser.doc.string_buf.get(), sb_end);
  }
  get_active_implementation() = saved;
  return int(err);
}
