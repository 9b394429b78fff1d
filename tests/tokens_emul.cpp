// tokens_emul.cpp -- the per-token functions of the stage-2-lite kernels (simdjson_b200/csrc/sjb200_tokens.cuh) compiled
// for the host and driven in the kernels' decomposition (tiles of 256 structurals, one per thread, the tile's span of the document copied into a window the way the CTA stages
// it in shared memory and read through tok::FastWin when it fits: pass A
// types / payloads / tile sums, exclusive scan, pass B records at tile offset + thread prefix), sequentially.  Checked
// against the oracle by tests/test_tokens_emul.py; no GPU involved.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "sjb200_tokens.cuh"

namespace {
constexpr uint32_t kThreads = 256, kWinBytes = 12 * 1024, kWinMargin = 16, kLaneBudget = 96;  // as in sjb200_tape.cu
struct Tile {
  bool fast;
  uint64_t lo;
  std::vector<uint8_t> win;
  uint32_t limit;
};
// stage_window of sjb200_tape.cu, on the host
Tile make_tile(const uint8_t *buf, uint64_t len, const uint32_t *idx, uint32_t n, uint32_t i0) {
  Tile t;
  const uint64_t first = idx[i0];
  const uint64_t next = (uint64_t(i0) + kThreads < n) ? uint64_t(idx[i0 + kThreads]) : len;
  const uint64_t mis = (reinterpret_cast<uintptr_t>(buf) + first) & 15u;
  t.lo = first >= mis ? first - mis : first;
  uint64_t span = next + kWinMargin - t.lo;
  t.fast = span + kWinMargin <= kWinBytes;
  if (span > kWinBytes) span = kWinBytes;
  if (t.lo + span > len) span = len - t.lo;
  t.limit = uint32_t(span);
  t.win.assign(buf + t.lo, buf + t.lo + span);
  t.win.resize(span + kWinMargin, 0x20);
  return t;
}
}  // namespace

extern "C" int emu_tokens(const uint8_t *buf, uint64_t len, const uint32_t *idx, uint32_t n, uint8_t *type, uint64_t *payload, uint8_t *strbuf,
                          uint64_t cap, uint64_t *string_bytes, uint32_t *n_strings, uint32_t *first_error_index) {
  using namespace sjb200::tok;
  const uint32_t tiles = (n + kThreads - 1) / kThreads;
  const PlainSrc plain{buf, len};
  std::vector<uint64_t> tile_bytes(tiles ? tiles : 1, 0);
  std::vector<uint8_t> is_long(n ? n : 1, 0);
  unsigned long long first_error = ~0ull;
  uint32_t ns = 0;
  for (uint32_t b = 0; b < tiles; b++) {
    const Tile tile = make_tile(buf, len, idx, n, b * kThreads);
    const FastWin f{tile.win.data(), tile.limit};
    for (uint32_t t = 0; t < kThreads; t++) {
      const uint32_t i = b * kThreads + t;
      if (i >= n) break;
      unsigned long long v = 0;
      uint32_t ty;
      if (tile.fast) {
        ty = classify_token(f, f.limit, uint32_t(idx[i] - tile.lo), &v, kLaneBudget);
        if (ty == 'd') v += tile.lo;
      } else {
        ty = classify_token(plain, len, uint64_t(idx[i]), &v, uint64_t(kLaneBudget));
      }
      if (ty == kLongString) {  // (the kernels hand these to the warp: tests/tokens_warp_emul.cpp)
        is_long[i] = 1;
        const long long ul = walk_string<false>(plain, len, uint64_t(idx[i]), nullptr);
        if (ul < 0) { ty = 0; v = ul == -1 ? kStringError : kUnclosedStringError; }
        else { ty = '"'; v = (unsigned long long)ul; }
      }
      if (ty == '"') { tile_bytes[b] += v + 5; ns++; }
      type[i] = uint8_t(ty);
      payload[i] = v;
      if (ty == 0) { const unsigned long long key = ((unsigned long long)i << 8) | (v & 0xFFull); if (key < first_error) first_error = key; }
    }
  }
  uint64_t run = 0;
  for (uint32_t b = 0; b < tiles; b++) { const uint64_t v = tile_bytes[b]; tile_bytes[b] = run; run += v; }
  *string_bytes = run;
  *n_strings = ns;
  *first_error_index = first_error == ~0ull ? 0xFFFFFFFFu : uint32_t(first_error >> 8);
  if (run <= cap) {
    for (uint32_t b = 0; b < tiles; b++) {
      const Tile tile = make_tile(buf, len, idx, n, b * kThreads);
      const FastWin f{tile.win.data(), tile.limit};
      uint64_t off = tile_bytes[b];
      for (uint32_t t = 0; t < kThreads; t++) {
        const uint32_t i = b * kThreads + t;
        if (i >= n || type[i] != '"') continue;
        const uint64_t ul = payload[i];
        uint8_t *rec = strbuf + off;
        rec[0] = uint8_t(ul); rec[1] = uint8_t(ul >> 8); rec[2] = uint8_t(ul >> 16); rec[3] = uint8_t(ul >> 24);
        if (tile.fast && !is_long[i]) walk_string<true>(f, f.limit, uint32_t(idx[i] - tile.lo), rec + 4);
        else walk_string<true>(plain, len, uint64_t(idx[i]), rec + 4);
        rec[4 + ul] = 0;
        payload[i] = off;
        off += ul + 5;
      }
    }
  }
  if (first_error != ~0ull) return int(first_error & 0xFF);
  return run > cap ? 1 : 0;
}
