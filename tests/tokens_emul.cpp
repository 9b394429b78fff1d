// tokens_emul.cpp -- the per-token functions of the stage-2-lite kernels (simdjson_b200/csrc/sjb200_tokens.cuh) compiled
// for the host and driven in the kernels' decomposition (tiles of 512 structurals, one per thread: pass A
// types / payloads / tile sums, exclusive scan, pass B records at tile offset + thread prefix), sequentially.  Checked
// against the oracle by tests/test_tokens_emul.py; no GPU involved.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "sjb200_tokens.cuh"

extern "C" int emu_tokens(const uint8_t *buf, uint64_t len, const uint32_t *idx, uint32_t n, uint8_t *type, uint64_t *payload, uint8_t *strbuf,
                          uint64_t cap, uint64_t *string_bytes, uint32_t *n_strings, uint32_t *first_error_index) {
  using namespace sjb200::tok;
  const uint32_t kPer = 1, kThreads = 512, kTile = kPer * kThreads;
  const PlainSrc src{buf, len};
  const uint32_t tiles = (n + kTile - 1) / kTile;
  std::vector<uint64_t> tile_bytes(tiles ? tiles : 1, 0);
  unsigned long long first_error = ~0ull;
  uint32_t ns = 0;
  for (uint32_t b = 0; b < tiles; b++)
    for (uint32_t t = 0; t < kThreads; t++)
      for (uint32_t k = 0; k < kPer; k++) {
        const uint32_t i = b * kTile + t * kPer + k;
        if (i >= n) break;
        unsigned long long v = 0;
        const uint32_t ty = classify_token(src, len, idx[i], &v);
        if (ty == '"') { tile_bytes[b] += v + 5; ns++; }
        type[i] = uint8_t(ty);
        payload[i] = v;
        if (ty == 0) { const unsigned long long key = ((unsigned long long)i << 8) | (v & 0xFFull); if (key < first_error) first_error = key; }
      }
  uint64_t run = 0;
  for (uint32_t b = 0; b < tiles; b++) { const uint64_t v = tile_bytes[b]; tile_bytes[b] = run; run += v; }
  *string_bytes = run;
  *n_strings = ns;
  *first_error_index = first_error == ~0ull ? 0xFFFFFFFFu : uint32_t(first_error >> 8);
  if (run <= cap) {
    for (uint32_t b = 0; b < tiles; b++) {
      uint64_t off = tile_bytes[b];
      for (uint32_t t = 0; t < kThreads; t++)
        for (uint32_t k = 0; k < kPer; k++) {
          const uint32_t i = b * kTile + t * kPer + k;
          if (i >= n || type[i] != '"') continue;
          const uint64_t ul = payload[i];
          uint8_t *rec = strbuf + off;
          rec[0] = uint8_t(ul); rec[1] = uint8_t(ul >> 8); rec[2] = uint8_t(ul >> 16); rec[3] = uint8_t(ul >> 24);
          walk_string<true>(src, len, idx[i], rec + 4);
          rec[4 + ul] = 0;
          payload[i] = off;
          off += ul + 5;
        }
    }
  }
  if (first_error != ~0ull) return int(first_error & 0xFF);
  return run > cap ? 1 : 0;
}
