// dropin_harness.cpp -- C entry points that drive the UNMODIFIED simdjson public API (dom::parser::parse,
// parse_many, ondemand::parser::iterate, simdjson::minify, simdjson::validate_utf8) with either the "b200"
// plug-in or a CPU implementation active, so the pytest drop-in tests can compare the two.  It is the
// reference-side usage a maintainer would write, wrapped for ctypes.
#include <chrono>
#include <cstring>
#include <memory>
#include <string>

#include "b200_implementation.h"

using namespace simdjson;

#define HARNESS_API extern "C" __attribute__((visibility("default")))

namespace {
const implementation *pick(int use_b200, int device) {
  if (use_b200) return b200::get_implementation(device);
  for (const char *n : {"icelake", "haswell", "westmere", "fallback"}) {
    auto impl = get_available_implementations()[n];
    if (impl && impl->supported_by_runtime_system()) return impl;
  }
  return builtin_implementation();
}
struct scoped_impl {
  const implementation *saved;
  explicit scoped_impl(const implementation *impl) : saved(get_active_implementation()) { get_active_implementation() = impl; }
  ~scoped_impl() { get_active_implementation() = saved; }
};
void put(const std::string &s, char *out, size_t cap, size_t *out_len) {
  *out_len = s.size();
  if (s.size() <= cap) std::memcpy(out, s.data(), s.size());
}
}  // namespace

HARNESS_API const char *dropin_active_name(int use_b200) {
  static thread_local std::string s;
  scoped_impl g(pick(use_b200, 0));
  s = get_active_implementation()->name();
  return s.c_str();
}

// the implementation simdjson would use right now (no override by the harness)
HARNESS_API const char *dropin_default_active_name() {
  static thread_local std::string s;
  s = get_active_implementation()->name();
  return s.c_str();
}

// dom::parser::parse -> simdjson::minify(element).  gpu_calls reports how many stage-1 calls the parser's
// implementation sent to the GPU (0 for a CPU implementation).
HARNESS_API int dropin_dom_roundtrip(int use_b200, const uint8_t *buf, size_t len, char *out, size_t cap, size_t *out_len,
                                     unsigned long long *gpu_calls) {
  scoped_impl g(pick(use_b200, 0));
  dom::parser parser;
  dom::element doc;
  auto err = parser.parse(buf, len, true).get(doc);
  *out_len = 0;
  if (gpu_calls) {
    auto *p = dynamic_cast<b200::dom_parser_implementation *>(parser.implementation.get());
    *gpu_calls = p ? p->gpu_stage1_calls() : 0;
  }
  if (err) return int(err);
  put(simdjson::minify(doc), out, cap, out_len);
  return 0;
}

// dom::parser::parse_many (document_stream; with SIMDJSON_THREADS_ENABLED its stage-1 worker runs a second
// parser concurrently): concatenated minified documents, '\n' separated.
HARNESS_API long dropin_parse_many(int use_b200, const uint8_t *buf, size_t len, size_t batch_size, char *out, size_t cap,
                                   size_t *out_len, int *first_err, unsigned long long *gpu_calls) {
  scoped_impl g(pick(use_b200, 0));
  long ndocs = 0;
  *first_err = 0;
  std::string acc;
  dom::parser parser;
  {
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
  if (gpu_calls) {
    auto *p = dynamic_cast<b200::dom_parser_implementation *>(parser.implementation.get());
    *gpu_calls = p ? p->gpu_stage1_calls() : 0;
  }
  put(acc, out, cap, out_len);
  return ndocs;
}

// ondemand::parser::iterate -> to_json_string (On-Demand walks the GPU-produced index array lazily)
HARNESS_API int dropin_ondemand_roundtrip(int use_b200, const uint8_t *buf, size_t len, char *out, size_t cap, size_t *out_len) {
  scoped_impl g(pick(use_b200, 0));
  *out_len = 0;
  padded_string json(reinterpret_cast<const char *>(buf), len);
  ondemand::parser parser;
  ondemand::document doc;
  auto err = parser.iterate(json).get(doc);
  if (err) return int(err);
  std::string_view sv;
  err = simdjson::to_json_string(doc).get(sv);
  if (err) return int(err);
  put(std::string(sv), out, cap, out_len);
  return 0;
}

// the free functions simdjson::minify(buf,len,dst,dst_len) / simdjson::validate_utf8(buf,len)
HARNESS_API int dropin_minify(int use_b200, const uint8_t *buf, size_t len, uint8_t *dst, size_t *dst_len) {
  scoped_impl g(pick(use_b200, 0));
  size_t n = 0;
  auto err = simdjson::minify(reinterpret_cast<const char *>(buf), len, reinterpret_cast<char *>(dst), n);
  *dst_len = n;
  return int(err);
}

HARNESS_API int dropin_validate_utf8(int use_b200, const uint8_t *buf, size_t len) {
  scoped_impl g(pick(use_b200, 0));
  return simdjson::validate_utf8(reinterpret_cast<const char *>(buf), len) ? 1 : 0;
}

// Stage 1 alone through the reference's own boundary, timed inside the process: what `bench.py` reports as e2e.
// get_active_implementation()->create_dom_parser_implementation(...)->stage1(buf, len, regular) on a pageable
// padded_string (the memory dom::parser::parse hands its implementation), `iters` calls after two warm-up calls.
// Returns the error code of the last call; seconds[0] = total of the timed calls, seconds[1] = the fastest call.
// idx_out (optional, idx_cap words): the n + 3 index words of the last call, for the parity gate.
HARNESS_API int dropin_stage1_timed(int use_b200, const uint8_t *buf, size_t len, int iters, double *seconds, uint32_t *n_out,
                                    uint32_t *idx_out, size_t idx_cap, unsigned long long *gpu_calls) {
  scoped_impl g(pick(use_b200, 0));
  std::unique_ptr<internal::dom_parser_implementation> p;
  auto err = get_active_implementation()->create_dom_parser_implementation(len, 1024, p);
  if (err) return int(err);
  padded_string json(reinterpret_cast<const char *>(buf), len);
  const uint8_t *b = reinterpret_cast<const uint8_t *>(json.data());
  for (int i = 0; i < 2; i++) err = p->stage1(b, len, stage1_mode::regular);
  double total = 0, best = 1e30;
  for (int i = 0; i < iters; i++) {
    const auto t0 = std::chrono::steady_clock::now();
    err = p->stage1(b, len, stage1_mode::regular);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    total += dt;
    if (dt < best) best = dt;
  }
  seconds[0] = total;
  seconds[1] = best;
  *n_out = p->n_structural_indexes;
  if (idx_out && size_t(p->n_structural_indexes) + 3 <= idx_cap) std::memcpy(idx_out, p->structural_indexes.get(), (size_t(p->n_structural_indexes) + 3) * 4);
  if (gpu_calls) {
    auto *bp = dynamic_cast<b200::dom_parser_implementation *>(p.get());
    *gpu_calls = bp ? bp->gpu_stage1_calls() : 0;
  }
  return int(err);
}
