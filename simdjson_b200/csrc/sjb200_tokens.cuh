// sjb200_tokens.cuh -- what stage 2 decides about ONE token from its bytes (strings, numbers, atoms): the per-thread
// functions of sjb200_tape.cu.  Pure functions of (byte source, len, pos), compiled for host and device, so that they are checked
// on the CPU against the oracle (tests/tokens_emul.cpp) before the kernels around them run on a GPU.
//
// Reference: json_iterator::visit_primitive src/generic/stage2/json_iterator.h L338-360; stringparsing::parse_string
// src/generic/stage2/stringparsing.h L146-190 (handle_unicode_codepoint L55-98, escape_map L22-48);
// jsoncharutils::codepoint_to_utf8 include/simdjson/generic/jsoncharutils.h L37-62; numberparsing::parse_number
// include/simdjson/generic/numberparsing.h L860-961; atomparsing include/simdjson/generic/atomparsing.h L45-95.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SJ_TOK __host__ __device__ __forceinline__
#else
#define SJ_TOK inline
#endif

namespace sjb200 {
namespace tok {

// simdjson::error_code values this file reports (include/simdjson/error.h L19-54)
enum : uint32_t { kUnclosedStringError = 15, kTapeError = 3, kStringError = 5, kTAtomError = 6, kFAtomError = 7, kNAtomError = 8, kNumberError = 9, kBigintError = 10 };

SJ_TOK uint8_t ld_byte(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
// A byte source: byte i of the document; beyond the end the reference reads padding (numbers and atoms look one byte
// past the token).  Plain: straight from the document.  Windowed (the kernels): a span of the document staged in shared
// memory by the whole CTA, anything outside it from global memory.
// word(i, &w): the four bytes i..i+3 as a little-endian word when they can be had with ONE aligned load (false otherwise:
// the caller goes byte by byte); vec16(i, w): likewise sixteen bytes from an address that is 16-byte aligned.
SJ_TOK uint32_t ld_word(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
  return __ldg(reinterpret_cast<const uint32_t *>(p));
#else
  uint32_t v;
  __builtin_memcpy(&v, p, 4);
  return v;
#endif
}
struct PlainSrc {
  typedef uint64_t pos_t;
  const uint8_t *buf;
  uint64_t len;
  SJ_TOK uint32_t operator()(uint64_t i) const { return i < len ? uint32_t(ld_byte(buf + i)) : 0x20u; }
  SJ_TOK bool word(uint64_t i, uint32_t *w) const {
    if (i + 4 > len || ((reinterpret_cast<uintptr_t>(buf) + i) & 3u)) return false;
    *w = ld_word(buf + i);
    return true;
  }
  SJ_TOK bool vec16(uint64_t i, uint32_t w[4]) const {
    if (i + 16 > len || ((reinterpret_cast<uintptr_t>(buf) + i) & 15u)) return false;
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(buf + i));
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#else
    __builtin_memcpy(w, buf + i, 16);
#endif
    return true;
  }
};
struct WindowSrc {
  typedef uint64_t pos_t;
  const uint8_t *buf;
  uint64_t len;
  const uint8_t *win;  // win[k] = document byte lo + k, for k < span
  uint64_t lo, span;
  SJ_TOK uint32_t operator()(uint64_t i) const {
    const uint64_t k = i - lo;
    if (k < span) return win[k];
    return i < len ? uint32_t(ld_byte(buf + i)) : 0x20u;
  }
  SJ_TOK bool word(uint64_t i, uint32_t *w) const {
    const uint64_t k = i - lo;
    if (k < span) {  // (k + 4 > span: straddles the window's end)
      if (k + 4 > span || (reinterpret_cast<uintptr_t>(win + k) & 3u)) return false;
      *w = *reinterpret_cast<const uint32_t *>(win + k);
      return true;
    }
    return PlainSrc{buf, len}.word(i, w);
  }
  SJ_TOK bool vec16(uint64_t i, uint32_t w[4]) const { return PlainSrc{buf, len}.vec16(i, w); }  // long strings: straight from global memory
};
// The whole tile inside the staged window (the usual case): positions are 32-bit offsets into it and nothing is checked --
// the CTA stages the tile's span plus 16 bytes (or 0x20 padding beyond the end of the document), and no token looks
// further than 10 bytes past the closing quote / the end of its scalar, which lie before the next tile's first structural.
struct FastWin {
  typedef uint32_t pos_t;
  const uint8_t *win;  // 16-byte aligned
  uint32_t limit;      // bytes of the document in the window (the walk's "len")
  SJ_TOK uint32_t operator()(uint32_t r) const { return win[r]; }
  SJ_TOK bool word(uint32_t r, uint32_t *w) const {
    if ((r & 3u) || r + 4 > limit) return false;
    *w = *reinterpret_cast<const uint32_t *>(win + r);
    return true;
  }
};
// a byte of the word equals c
SJ_TOK bool word_has(uint32_t w, uint32_t c) {
  const uint32_t x = w ^ (c * 0x01010101u);
  return ((x - 0x01010101u) & ~x & 0x80808080u) != 0;
}
// ... a quote or a backslash
SJ_TOK bool word_has_special(uint32_t w) {
  const uint32_t x = w ^ 0x22222222u, y = w ^ 0x5C5C5C5Cu;
  return ((((x - 0x01010101u) & ~x) | ((y - 0x01010101u) & ~y)) & 0x80808080u) != 0;
}
SJ_TOK bool is_digit(uint32_t c) { return c - '0' < 10u; }
// internal::structural_or_whitespace (src/internal/jsoncharutils_tables.cpp L31-45)
SJ_TOK bool ends_scalar(uint32_t c) {
  return c == 0x20u || c == 0x09u || c == 0x0Au || c == 0x0Du || c == ',' || c == ':' || c == '[' || c == ']' || c == '{' || c == '}';
}
SJ_TOK int hex_val(uint32_t c) {
  if (c - '0' < 10u) return int(c - '0');
  const uint32_t l = c | 0x20u;
  if (l - 'a' < 6u) return int(l - 'a' + 10);
  return -1;
}
// four hex digits at i..i+3, -1 if one of them is not a hex digit (hex_to_u32_nocheck's "high bits set")
template <class S>
SJ_TOK int hex4(const S &at, typename S::pos_t i) {
  const int a = hex_val(at(i)), b = hex_val(at(i + 1)), c = hex_val(at(i + 2)), d = hex_val(at(i + 3));
  if ((a | b | c | d) < 0) return -1;
  return (a << 12) | (b << 8) | (c << 4) | d;
}

// One string: opening quote at pos.  Returns its unescaped length, -1 invalid escape, -2 the input ends first, -3 over budget.
// kWrite: the unescaped bytes go to dst.
// budget: give up (-3) once more than that many bytes have been looked at -- the caller hands such a string to the warp
// (tok::warp_string, sjb200_tokens_warp.cuh).
template <bool kWrite, class S>
SJ_TOK long long walk_string(const S &at, typename S::pos_t len, typename S::pos_t pos, uint8_t *dst, typename S::pos_t budget = ~(typename S::pos_t)0) {
  typedef typename S::pos_t P;
  P q = pos + 1;
  uint32_t out = 0;  // (a string is < 4 GiB: stage 1's limit)
  uint8_t *o = dst;  // kWrite: where the next unescaped byte goes (one pointer, bumped: the stores need no address arithmetic)
  const P stop = (budget < len && pos + 1 < len - budget) ? P(pos + 1 + budget) : len;  // the walk looks at bytes below stop
  while (q < stop) {
    uint32_t w;
    if (at.word(q, &w) && !word_has_special(w)) {  // four ordinary bytes at once
      if (kWrite) { o[0] = uint8_t(w); o[1] = uint8_t(w >> 8); o[2] = uint8_t(w >> 16); o[3] = uint8_t(w >> 24); o += 4; }
      out += 4; q += 4;
      continue;
    }
    const uint32_t b = at(q);
    if (b == '"') return out;
    if (b != '\\') {
      if (kWrite) *o++ = uint8_t(b);
      out++; q++;
      continue;
    }
    const uint32_t e = at(q + 1);
    if (e != 'u') {
      uint32_t m;
      switch (e) {  // stringparsing.h escape_map L22-48
        case '"': m = 0x22; break;
        case '\\': m = 0x5C; break;
        case '/': m = 0x2F; break;
        case 'b': m = 0x08; break;
        case 'f': m = 0x0C; break;
        case 'n': m = 0x0A; break;
        case 'r': m = 0x0D; break;
        case 't': m = 0x09; break;
        default: return -1;
      }
      if (kWrite) *o++ = uint8_t(m);
      out++; q += 2;
      continue;
    }
    int cp = hex4(at, q + 2);
    if (cp < 0) return -1;
    q += 6;
    if (cp >= 0xD800 && cp < 0xDC00) {  // a high surrogate needs its low one as the next \uXXXX (no replacement character)
      if (at(q) != '\\' || at(q + 1) != 'u') return -1;
      const int lo = hex4(at, q + 2);
      if (lo < 0xDC00 || lo > 0xDFFF) return -1;
      cp = (((cp - 0xD800) << 10) | (lo - 0xDC00)) + 0x10000;
      q += 6;
    } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
      return -1;
    }
    if (cp <= 0x7F) {
      if (kWrite) *o++ = uint8_t(cp);
      out += 1;
    } else if (cp <= 0x7FF) {
      if (kWrite) { o[0] = uint8_t(0xC0 | (cp >> 6)); o[1] = uint8_t(0x80 | (cp & 63)); o += 2; }
      out += 2;
    } else if (cp <= 0xFFFF) {
      if (kWrite) { o[0] = uint8_t(0xE0 | (cp >> 12)); o[1] = uint8_t(0x80 | ((cp >> 6) & 63)); o[2] = uint8_t(0x80 | (cp & 63)); o += 3; }
      out += 3;
    } else {
      if (kWrite) {
        o[0] = uint8_t(0xF0 | (cp >> 18)); o[1] = uint8_t(0x80 | ((cp >> 12) & 63));
        o[2] = uint8_t(0x80 | ((cp >> 6) & 63)); o[3] = uint8_t(0x80 | (cp & 63));
        o += 4;
      }
      out += 4;
    }
  }
  return q < len ? -3 : -2;  // over budget / the input ends first
}

// The number that starts at pos (parse_number, numberparsing.h L860-961).  Returns the tape type, 0 with *value = error.
template <class S>
SJ_TOK uint32_t scan_number(const S &at, typename S::pos_t pos, uint32_t first, unsigned long long *value) {
  typedef typename S::pos_t P;
  const bool neg = first == '-';
  P q = pos + (neg ? 1 : 0);
  const P start = q;
  unsigned long long i = 0;
  uint32_t c = at(q);
  const uint32_t lead = c;
  while (is_digit(c)) { i = i * 10ull + (c - '0'); c = at(++q); }
  const uint32_t digits = uint32_t(q - start);
  if (digits == 0 || (lead == '0' && digits > 1)) { *value = kNumberError; return 0; }
  bool is_float = false;
  if (c == '.') {
    is_float = true;
    const P fs = ++q;
    c = at(q);
    while (is_digit(c)) c = at(++q);
    if (q == fs) { *value = kNumberError; return 0; }  // "1." is not a number
  }
  if ((c | 0x20u) == 'e') {
    is_float = true;
    c = at(++q);
    if (c == '-' || c == '+') c = at(++q);
    const P es = q;
    while (is_digit(c)) c = at(++q);
    if (q == es) { *value = kNumberError; return 0; }
  }
  const bool dirty = !ends_scalar(c);
  if (is_float) {
    if (dirty) { *value = kNumberError; return 0; }
    *value = q;  // one past the token: the consumer converts [pos, q)
    return 'd';
  }
  const uint32_t longest = neg ? 19 : 20;  // L920-946: the 64-bit limits
  if (digits > longest) { *value = kBigintError; return 0; }
  if (digits == longest) {
    if (neg) {
      if (i > (1ull << 63)) { *value = kBigintError; return 0; }
    } else if (first != '1' || i <= 0x7FFFFFFFFFFFFFFFull) { *value = kBigintError; return 0; }  // wrapped around
  }
  if (dirty) { *value = kNumberError; return 0; }
  if (!neg && i > 0x7FFFFFFFFFFFFFFFull) { *value = i; return 'u'; }
  *value = neg ? (~i + 1ull) : i;
  return 'l';
}

template <class S>
SJ_TOK bool atom_is(const S &at, typename S::pos_t pos, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, int wl) {
  const uint32_t w[5] = {w0, w1, w2, w3, w4};
  for (int k = 1; k < wl; k++)
    if (at(pos + k) != w[k]) return false;
  return ends_scalar(at(pos + wl));
}

// type and payload of the token at structural position p (see include/sjb200.h, sjb200_tokens_dev); a string's payload is
// its unescaped length
constexpr uint32_t kLongString = 1;  // classify_token: a string longer than the budget, length not known yet
template <class S>
SJ_TOK uint32_t classify_token(const S &at, typename S::pos_t len, typename S::pos_t p, unsigned long long *value,
                                typename S::pos_t string_budget = ~(typename S::pos_t)0) {
  const uint32_t c = at(p);
  *value = 0;
  if (c == '{' || c == '}' || c == '[' || c == ']' || c == ':' || c == ',') return c;
  if (c == '"') {
    const long long ul = walk_string<false>(at, len, p, nullptr, string_budget);
    if (ul == -3) return kLongString;
    if (ul < 0) {
      *value = ul == -1 ? uint32_t(kStringError) : uint32_t(kUnclosedStringError);
      return 0;
    }
    *value = (unsigned long long)ul;
    return '"';
  }
  if (c <= '9' || c == '-') return scan_number(at, p, c, value);  // json_iterator.h L342: `(*value - '0') < 10` in int -- every byte up to '9' takes the number path
  if (c == 't') {
    if (atom_is(at, p, 't', 'r', 'u', 'e', 0, 4)) return 't';
    *value = kTAtomError;
    return 0;
  }
  if (c == 'f') {
    if (atom_is(at, p, 'f', 'a', 'l', 's', 'e', 5)) return 'f';
    *value = kFAtomError;
    return 0;
  }
  if (c == 'n') {
    if (atom_is(at, p, 'n', 'u', 'l', 'l', 0, 4)) return 'n';
    *value = kNAtomError;
    return 0;
  }
  *value = kTapeError;
  return 0;
}

}  // namespace tok
}  // namespace sjb200
