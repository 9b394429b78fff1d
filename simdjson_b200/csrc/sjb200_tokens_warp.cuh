// sjb200_tokens_warp.cuh -- one LONG string by a whole warp (stage-2-lite, sjb200_tape.cu).
//
// A thread that walks its own string byte by byte is fine for the usual 5-50 byte keys and values; a single long string
// (a padded value, an embedded document, a base64 blob) would keep one lane busy for milliseconds while 31 idle.  Such
// strings are handed to the warp: 32 bytes per step, one per lane, with the same bit algebra stage 1 uses --
//   * which backslashes start an escape: the odd-run identity of json_escape_scanner::next
//     (src/generic/stage1/json_escape_scanner.h L96-143) on a 32-bit word with a carry between steps;
//   * the closing quote is the first quote that is not escaped;
//   * every byte knows locally how many bytes it contributes to the unescaped string: an escape's backslash 0; the
//     escaped character 1 (escape_map, src/generic/stage2/stringparsing.h L22-48) or, for \uXXXX, the UTF-8 length of the
//     code point (handle_unicode_codepoint L55-98: a high surrogate takes the low one that must follow and contributes
//     4, the low one 0 -- a low surrogate is legal exactly when the escape six bytes before it is a high one); the four
//     hex digits 0; any other byte 1 -- so the output offset of a byte is a warp prefix sum;
//   * blocks of 1 KiB without a backslash or a quote (the common case in a long string) are copied without any of that.
// Same results as tok::walk_string (the sequential walk): checked against the oracle under the host SIMT emulation
// (tests/tokens_warp_emul.cpp) -- written against the primitive layer of sjb200_simt.cuh for that purpose.
#pragma once
#include "sjb200_simt.cuh"
#include "sjb200_tokens.cuh"

namespace sjb200 {
namespace tok {

// Returns (warp-uniform) the unescaped length of the string whose opening quote is at pos, -1 invalid escape, -2 the
// input ends first.  kWrite: the unescaped bytes go to dst (any address space).
template <bool kWrite, class S>
SJ_DEV long long warp_string(const S &at, uint64_t len, uint64_t pos, uint8_t *dst, unsigned lane) {
  const uint32_t ODD = 0xAAAAAAAAu;
  uint64_t q = pos + 1;       // next byte to look at
  long long out = 0;          // bytes of the unescaped string so far
  uint32_t esc_carry = 0;     // byte q is escaped (an odd run of backslashes ends at q - 1)
  uint32_t prev_uchar = 0;    // the previous step's escaped 'u' lanes (their hex digits / a low surrogate may lie in this step)
  for (;;) {
    if (q >= len) return esc_carry ? -1 : -2;  // (a backslash as the last byte: its "escaped character" is padding)
    // ---- 1 KiB at once when nothing in it needs a decision: two aligned 16-byte loads per lane, both in flight together
    // (a single warp is latency-bound: one round trip to memory per step).  The block starts at the 16-byte boundary
    // at or below q; the bytes before q (first block only) are lane 0's and ignored.
    if (!esc_carry && (prev_uchar >> 28) == 0) {
      const uint64_t mis = (reinterpret_cast<uintptr_t>(at.buf) + q) & 15u;
      if (q >= mis && q - mis + 1024 <= len) {
        const uint64_t a0 = q - mis;
        uint32_t w[8];
        const bool got = at.vec16(a0 + 16u * lane, w) && at.vec16(a0 + 512u + 16u * lane, w + 4);
        bool special = !got;
        if (got) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint32_t x = w[j];
            if (j < 4 && lane == 0 && mis > uint64_t(4 * j)) {  // bytes before q: neutral
              const uint32_t nb = mis - 4 * j >= 4 ? 4u : uint32_t(mis - 4 * j);
              const uint32_t low = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u);
              x = (x & ~low) | (0x20202020u & low);
            }
            special = special || word_has(x, '"') || word_has(x, '\\');
          }
        }
        if (!sj_any(special)) {
          if (kWrite) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              const uint64_t off = uint64_t(j < 4 ? 0 : 512) + 16u * lane + 4u * uint32_t(j & 3);  // of this word inside the block
#pragma unroll
              for (int t = 0; t < 4; t++)
                if (off + uint64_t(t) >= mis) dst[out + (long long)(off + uint64_t(t) - mis)] = uint8_t(w[j] >> (8 * t));
            }
          }
          out += (long long)(1024 - mis);
          q = a0 + 1024;
          prev_uchar = 0;
          continue;
        }
      }
    }
    // ---- 32 bytes, one per lane
    const uint64_t me = q + lane;
    const bool valid = me < len;
    const uint32_t b = valid ? at(me) : 0x20u;
    const uint32_t vmask = (len - q >= 32) ? 0xFFFFFFFFu : ((1u << uint32_t(len - q)) - 1u);
    const uint32_t bs = sj_ballot(b == '\\') & vmask;
    const uint32_t cover_in = (prev_uchar >> 31) | (prev_uchar >> 30) | (prev_uchar >> 29) | (prev_uchar >> 28);  // hex digits of an escape of the previous step
    const uint32_t quotes = sj_ballot(b == '"') & vmask;
    if (bs == 0 && !esc_carry) {
      // no escape in this step (and none reaching into it beyond hex digits, which cannot be quotes in a valid escape: if
      // one is, the escape's own lane reported the error in the previous step)
      const uint32_t qm = quotes & ~cover_in;
      const uint32_t upto = qm ? ((1u << (sj_ffs(qm) - 1)) - 1u) : vmask;  // lanes before the closing quote
      const uint32_t emit = upto & ~cover_in;
      if (kWrite && ((emit >> lane) & 1u)) dst[out + sj_popc(emit & ((1u << lane) - 1u))] = uint8_t(b);
      out += sj_popc(emit);
      if (qm) return out;
      q += 32;
      prev_uchar = 0;
      continue;
    }
    // which bytes are escaped (preceded by an odd-length run of backslashes), which backslashes start an escape
    const uint32_t potential = bs & ~esc_carry;
    const uint32_t maybe = potential << 1;
    const uint32_t eatc = ((maybe | ODD) - potential) ^ ODD;
    const uint32_t escaped = (eatc ^ (bs | esc_carry)) & 0xFFFFFFFFu;
    const uint32_t start = eatc & bs;
    const uint32_t carry_out = start >> 31;
    const uint32_t qm = quotes & ~escaped;
    const int endbit = qm ? (sj_ffs(qm) - 1) : 32;
    const uint32_t upto = endbit < 32 ? ((1u << endbit) - 1u) : 0xFFFFFFFFu;  // lanes before the closing quote (may include lanes beyond len: see below)
    const bool is_esc = ((escaped >> lane) & 1u) != 0;
    const uint32_t uchar = sj_ballot(is_esc && b == 'u') & upto;
    const uint32_t covered = ((uchar << 1) | (uchar << 2) | (uchar << 3) | (uchar << 4) | cover_in) & ~0u;
    // this lane's contribution
    uint32_t cnt = 0, ob0 = b, ob1 = 0, ob2 = 0, ob3 = 0;
    bool err = false;
    const bool active = ((upto >> lane) & 1u) != 0;
    if (active && is_esc) {
      // (a lane beyond len can be here: the escaped character of a backslash that is the last byte -- it is padding, an error)
      if (b != 'u') {
        uint32_t m = 0;
        switch (b) {
          case '"': m = 0x22; break;
          case '\\': m = 0x5C; break;
          case '/': m = 0x2F; break;
          case 'b': m = 0x08; break;
          case 'f': m = 0x0C; break;
          case 'n': m = 0x0A; break;
          case 'r': m = 0x0D; break;
          case 't': m = 0x09; break;
          default: break;
        }
        err = (m == 0) || !valid;
        cnt = 1;
        ob0 = m;
      } else {
        int cp = hex4(at, me + 1);
        if (cp < 0) {
          err = true;
        } else if (cp >= 0xD800 && cp < 0xDC00) {
          const int lo = (at(me + 5) == '\\' && at(me + 6) == 'u') ? hex4(at, me + 7) : -1;
          if (lo < 0xDC00 || lo > 0xDFFF) err = true;
          else cp = (((cp - 0xD800) << 10) | (lo - 0xDC00)) + 0x10000;
        } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
          // legal exactly as the second half of a pair: the escape six bytes back is a \u with a high surrogate
          const bool prev_is_uchar = lane >= 6 ? (((uchar >> (lane - 6)) & 1u) != 0) : (((prev_uchar >> (lane + 26)) & 1u) != 0);
          const int hi = (prev_is_uchar && me >= 5) ? hex4(at, me - 5) : -1;
          if (hi >= 0xD800 && hi < 0xDC00) cp = -2;  // consumed by the high surrogate's lane
          else err = true;
        }
        if (!err) {
          if (cp == -2) { cnt = 0; }
          else if (cp <= 0x7F) { cnt = 1; ob0 = uint32_t(cp); }
          else if (cp <= 0x7FF) { cnt = 2; ob0 = 0xC0u | uint32_t(cp >> 6); ob1 = 0x80u | uint32_t(cp & 63); }
          else if (cp <= 0xFFFF) { cnt = 3; ob0 = 0xE0u | uint32_t(cp >> 12); ob1 = 0x80u | uint32_t((cp >> 6) & 63); ob2 = 0x80u | uint32_t(cp & 63); }
          else { cnt = 4; ob0 = 0xF0u | uint32_t(cp >> 18); ob1 = 0x80u | uint32_t((cp >> 12) & 63); ob2 = 0x80u | uint32_t((cp >> 6) & 63); ob3 = 0x80u | uint32_t(cp & 63); }
        }
      }
    } else if (active && valid && !((start >> lane) & 1u) && !((covered >> lane) & 1u)) {
      cnt = 1;  // an ordinary byte
    }
    if (sj_any(err)) return -1;
    // offsets: inclusive prefix sum of cnt over the lanes
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = sj_shfl_up(incl, d);
      if (int(lane) >= d) incl += t;
    }
    const uint32_t total = sj_shfl(incl, 31);
    if (kWrite && cnt) {
      uint8_t *o = dst + out + (incl - cnt);
      o[0] = uint8_t(ob0);
      if (cnt > 1) o[1] = uint8_t(ob1);
      if (cnt > 2) o[2] = uint8_t(ob2);
      if (cnt > 3) o[3] = uint8_t(ob3);
    }
    out += total;
    if (endbit < 32) return out;
    if (q + 32 >= len) return carry_out ? -1 : -2;  // ran off the input (a trailing backslash escapes padding: an invalid escape)
    q += 32;
    esc_carry = carry_out;
    prev_uchar = uchar;
  }
}

}  // namespace tok
}  // namespace sjb200
