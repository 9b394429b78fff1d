/*
 * sj_oracle.c -- byte-at-a-time CPU restatement of simdjson's stage 1,
 * minify and validate_utf8.  TEST INFRASTRUCTURE ONLY (see sj_oracle.h).
 * Parity status: PINNED against oracle/_ref (the reference compiled from
 * /root/reference/singleheader) and tests/golden/.
 *
 * Every function cites the reference lines whose behaviour it restates.
 * Paths are relative to /root/reference/.
 */
#include "sj_oracle.h"
#include <string.h>

size_t sjo_index_capacity(size_t capacity) {
  /* include/simdjson/generic/dom_parser_implementation.h L66-82 */
  return ((capacity + 63) / 64) * 64 + 9;
}

/* ---------------------------------------------------------------- classes */

/* src/icelake.cpp L48-96 (and haswell.cpp L43-94): whitespace is
 * {20,09,0A,0D}; "op" is {2C,3A,5B,5D,7B,7D} plus 0C and 1A because the x86
 * kernels compare (byte|0x20) against a table keyed on the low nibble. */
static int is_ws(uint8_t c) { return c == 0x20 || c == 0x09 || c == 0x0A || c == 0x0D; }
static int is_op(uint8_t c) {
  switch (c) {
    case 0x2C: case 0x3A: case 0x5B: case 0x5D: case 0x7B: case 0x7D:
    case 0x0C: case 0x1A:
      return 1;
    default:
      return 0;
  }
}

/* The scanner as a per-byte automaton.
 *   esc     : json_escape_scanner::next_is_escaped (json_escape_scanner.h L50-71)
 *   instr   : json_string_scanner::prev_in_string   (json_string_scanner.h L62-85)
 *   prev_nq : json_scanner::prev_scalar             (json_scanner.h L128-157)
 */
typedef struct {
  int esc, instr, prev_nq;
  int unescaped_err; /* json_structural_indexer::unescaped_chars_error, L246 */
} scan_state;

typedef struct {
  int structural; /* json_block::structural_start, json_scanner.h L68-79 */
  int keep;       /* !(whitespace & ~in_string), json_minifier.h L37-40    */
} scan_out;

static scan_out scan_byte(scan_state *s, uint8_t c) {
  scan_out o;
  /* a byte is "escaped" iff the previous byte was an unescaped backslash */
  int escaped = s->esc;
  s->esc = (!escaped && c == '\\');
  int quote = (c == '"') && !escaped;
  if (quote) s->instr ^= 1;
  int in_string = s->instr;            /* includes opening, excludes closing quote */
  int string_tail = in_string ^ quote; /* json_string_scanner.h L30 */
  int ws = is_ws(c), op = is_op(c);
  int scalar = !(ws || op);
  int nq = scalar && !quote;           /* json_scanner.h L148 */
  int follows = s->prev_nq;            /* json_scanner.h L149 */
  s->prev_nq = nq;
  o.structural = (op || (scalar && !follows)) && !string_tail;
  if (c <= 0x1F && in_string) s->unescaped_err = 1;
  o.keep = !(ws && !in_string);
  return o;
}

/* ------------------------------------------------------------------ utf-8 */

/* Standard well-formedness (Unicode 15, table 3-7).  The reference's lookup4
 * checker (utf8_lookup4_algorithm.h L145-202) returns exactly this boolean
 * ("UTF-8 equivalence note"); the reference's own
 * tests assert the same (tests/unicode_tests.cpp L103-155). */
int sjo_validate_utf8(const uint8_t *b, size_t len) {
  size_t i = 0;
  while (i < len) {
    uint8_t c = b[i];
    if (c < 0x80) { i++; continue; }
    if (c < 0xC2) return 0; /* stray continuation, or overlong C0/C1 */
    if (c < 0xE0) {
      if (i + 1 >= len || (b[i + 1] & 0xC0) != 0x80) return 0;
      i += 2;
    } else if (c < 0xF0) {
      if (i + 2 >= len) return 0;
      uint8_t c1 = b[i + 1], c2 = b[i + 2];
      if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return 0;
      if (c == 0xE0 && c1 < 0xA0) return 0;  /* overlong */
      if (c == 0xED && c1 >= 0xA0) return 0; /* surrogates */
      i += 3;
    } else if (c < 0xF5) {
      if (i + 3 >= len) return 0;
      uint8_t c1 = b[i + 1], c2 = b[i + 2], c3 = b[i + 3];
      if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80) return 0;
      if (c == 0xF0 && c1 < 0x90) return 0;  /* overlong */
      if (c == 0xF4 && c1 >= 0x90) return 0; /* > U+10FFFF */
      i += 4;
    } else {
      return 0;
    }
  }
  return 1;
}

/* json_structural_indexer.h L156-174 */
size_t sjo_trim_partial_utf8(const uint8_t *buf, size_t len) {
  if (len >= 1 && buf[len - 1] >= 0xC0) return len - 1;
  if (len >= 2 && buf[len - 2] >= 0xE0) return len - 2;
  if (len >= 3 && buf[len - 3] >= 0xF0) return len - 3;
  return len;
}

/* ------------------------------------------------- document boundary finders */

typedef struct {
  const uint8_t *buf;
  uint32_t *idx;
  uint32_t n;
} idx_view;

/* find_next_document_index.h L39-98: walk backwards to the last place where a
 * value is followed by a value with no ',' / ':' in between. */
static uint32_t next_document_index(const idx_view *p) {
  if (p->n == 0) return 0;
  int arr = 0, obj = 0;
  for (uint32_t i = p->n - 1; i > 0; i--) {
    uint8_t b = p->buf[p->idx[i]];
    if (b == ':' || b == ',') continue;
    if (b == '}') { obj--; continue; }
    if (b == ']') { arr--; continue; }
    if (b == '{') obj++;
    else if (b == '[') arr++;
    uint8_t a = p->buf[p->idx[i - 1]];
    if (a == '{' || a == '[' || a == ':' || a == ',') continue;
    /* boundary between idx[i-1] and idx[i] */
    return (arr == 0 && obj == 0) ? p->n : i;
  }
  uint8_t f = p->buf[p->idx[0]];
  if (f == '}') obj--;
  else if (f == ']') arr--;
  else if (f == '{') obj++;
  else if (f == '[') arr++;
  return (arr == 0 && obj == 0) ? p->n : 0;
}

uint32_t sjo_find_next_document_index(const uint8_t *buf, const uint32_t *idx, uint32_t n) {
  idx_view v = {buf, (uint32_t *)idx, n};
  return next_document_index(&v);
}

#define SJO_TOO_LARGE 0xFFFFFFFFu /* find_next_document_index.h L105 */

/* find_next_document_index.h L126-267 (RFC 7464 record separators) */
static uint32_t json_sequence_filter(idx_view *p, size_t len, int is_final, uint32_t *next_start) {
  *next_start = (uint32_t)len;
  if (p->n == 0) return 0;
  uint32_t w = 0, last_rs = 0, n_rs = 0;
  for (uint32_t r = 0; r < p->n; r++) {
    uint32_t pos = p->idx[r];
    if (p->buf[pos] != 0x1E) { p->idx[w++] = pos; continue; }
    last_rs = pos; n_rs++;
    uint32_t v = pos + 1;
    for (; v < len; v++) {
      uint8_t c = p->buf[v];
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
      if (c == 0x1E) { last_rs = v; n_rs++; continue; }
      break;
    }
    while (r + 1 < p->n && p->idx[r + 1] < v) r++;
    if (v < len) {
      uint8_t c = p->buf[v];
      int oper = (c == '{' || c == '}' || c == '[' || c == ']' || c == ':' || c == ',');
      int present = (r + 1 < p->n && p->idx[r + 1] == v);
      if (!oper && !present) p->idx[w++] = v;
    }
  }
  p->n = w;
  if (w == 0) return 0;
  if (n_rs == 0) return is_final ? next_document_index(p) : 0;
  if (is_final) return p->n;
  *next_start = last_rs;
  if (n_rs < 2) return SJO_TOO_LARGE;
  for (uint32_t i = p->n; i > 0; i--)
    if (p->idx[i - 1] < last_rs) return i;
  return 0;
}

/* find_next_document_index.h L288-369 (root-level commas separate documents) */
static uint32_t comma_delimited_filter(idx_view *p, size_t len, int is_final, uint32_t *next_start) {
  *next_start = (uint32_t)len;
  if (p->n == 0) return 0;
  int depth = 0;
  uint32_t w = 0, last_comma = 0, n_comma = 0;
  for (uint32_t i = 0; i < p->n; i++) {
    uint32_t pos = p->idx[i];
    uint8_t c = p->buf[pos];
    if (c == '{' || c == '[') depth++;
    else if (c == '}' || c == ']') depth--;
    else if (c == ',' && depth == 0) { last_comma = pos; n_comma++; continue; }
    p->idx[w++] = pos;
  }
  p->n = w;
  if (w == 0) return 0;
  if (is_final) return next_document_index(p);
  if (n_comma == 0) return SJO_TOO_LARGE;
  *next_start = last_comma + 1;
  uint32_t keep = 0;
  for (uint32_t i = p->n; i > 0; i--)
    if (p->idx[i - 1] < last_comma) { keep = i; break; }
  if (keep == 0) return 0;
  p->n = keep;
  return next_document_index(p);
}

/* ---------------------------------------------------------------- stage 1 */

int sjo_stage1(const uint8_t *buf, size_t len, size_t capacity, int mode,
               uint32_t *idx, uint32_t *n_inout) {
  /* json_structural_indexer.h L193-204 */
  if (len > capacity) return SJO_CAPACITY;
  if (len == 0) return SJO_EMPTY;
  if (mode != SJO_REGULAR) {
    len = sjo_trim_partial_utf8(buf, len);
    if (len == 0) return SJO_UTF8_ERROR;
  }
  scan_state s = {0, 0, 0, 0};
  uint32_t count = 0;
  for (size_t i = 0; i < len; i++)
    if (scan_byte(&s, buf[i]).structural) idx[count++] = (uint32_t)i;
  /* the virtual 0x20 padding after len (buf_block_reader.h L98-104) can never
   * produce a structural nor an unescaped-char error, and cannot close a string */
  int unclosed = s.instr;

  /* finish(): json_structural_indexer.h L249-397 */
  if (mode == SJO_REGULAR && unclosed) return SJO_UNCLOSED_STRING;
  if (s.unescaped_err) return SJO_UNESCAPED_CHARS;
  idx_view p = {buf, idx, count};
  *n_inout = p.n;
  idx[p.n] = (uint32_t)len;
  idx[p.n + 1] = (uint32_t)len;
  idx[p.n + 2] = 0;
  if (p.n == 0) return SJO_EMPTY;
  if (idx[p.n - 1] > len) return SJO_UNEXPECTED_ERROR;

  uint32_t next_start = (uint32_t)len, m;
  switch (mode) {
    case SJO_STREAMING_PARTIAL:
      if (unclosed) { p.n--; *n_inout = p.n; if (p.n == 0) return SJO_CAPACITY; }
      m = next_document_index(&p);
      if (m == 0 && p.n > 0) {
        if (idx[0] == 0) return SJO_CAPACITY;
        *n_inout = 0;
        return SJO_EMPTY;
      }
      *n_inout = m;
      break;
    case SJO_STREAMING_FINAL:
      if (unclosed) p.n--;
      p.n = next_document_index(&p);
      *n_inout = p.n;
      idx[p.n + 1] = idx[p.n];
      idx[p.n] = (uint32_t)len;
      if (p.n == 0) return SJO_EMPTY;
      break;
    case SJO_JSON_SEQUENCE_PARTIAL:
    case SJO_COMMA_DELIMITED_PARTIAL:
      if (unclosed) { p.n--; *n_inout = p.n; if (p.n == 0) return SJO_CAPACITY; }
      m = (mode == SJO_JSON_SEQUENCE_PARTIAL) ? json_sequence_filter(&p, len, 0, &next_start)
                                              : comma_delimited_filter(&p, len, 0, &next_start);
      *n_inout = p.n;
      if (m == SJO_TOO_LARGE) return SJO_CAPACITY;
      if (m == 0) { *n_inout = 0; return SJO_EMPTY; }
      *n_inout = m;
      idx[m] = next_start;
      break;
    case SJO_JSON_SEQUENCE_FINAL:
    case SJO_COMMA_DELIMITED_FINAL:
      if (unclosed) p.n--;
      m = (mode == SJO_JSON_SEQUENCE_FINAL) ? json_sequence_filter(&p, len, 1, &next_start)
                                            : comma_delimited_filter(&p, len, 1, &next_start);
      *n_inout = m;
      idx[m + 1] = idx[m];
      idx[m] = (uint32_t)len;
      if (m == 0) return SJO_EMPTY;
      break;
    default:
      break;
  }
  return sjo_validate_utf8(buf, len) ? SJO_SUCCESS : SJO_UTF8_ERROR;
}

/* ----------------------------------------------------------------- minify */

int sjo_minify(const uint8_t *buf, size_t len, uint8_t *dst, size_t *dst_len) {
  /* json_minifier.h L68-97: same scanner; keep every byte that is not
   * whitespace outside a string; the padded tail is clamped to the bytes
   * actually consumed (L79-95), i.e. only positions < len are ever kept. */
  scan_state s = {0, 0, 0, 0};
  size_t out = 0;
  for (size_t i = 0; i < len; i++)
    if (scan_byte(&s, buf[i]).keep) dst[out++] = buf[i];
  if (s.instr) { *dst_len = 0; return SJO_UNCLOSED_STRING; } /* finish(), L42-47 */
  *dst_len = out;
  return SJO_SUCCESS;
}

/* ------------------------------------------------- helpers for the sharding tests */
/* raw scan of a shard with a given incoming scanner state (bit0 escape, bit1 in_string, bit2 prev_scalar):
 * number of structurals and the outgoing state; indexes (shard-relative) are stored when idx != NULL */
uint64_t sjo_scan_shard(const uint8_t *buf, size_t len, uint32_t state_in, uint32_t *idx, uint32_t *state_out) {
  scan_state s = {(int)(state_in & 1), (int)((state_in >> 1) & 1), (int)((state_in >> 2) & 1), 0};
  uint64_t n = 0;
  for (size_t i = 0; i < len; i++)
    if (scan_byte(&s, buf[i]).structural) { if (idx) idx[n] = (uint32_t)i; n++; }
  if (state_out) *state_out = (uint32_t)s.esc | ((uint32_t)s.instr << 1) | ((uint32_t)s.prev_nq << 2);
  return n;
}

/* the shard's carry transducer T(e) = (esc_out, quote parity, last byte is a non-quote scalar) for e = 0, 1,
 * packed like sjb200_shard_result.ttable */
uint32_t sjo_transducer(const uint8_t *buf, size_t len) {
  uint32_t T = 0;
  for (uint32_t e = 0; e < 2; e++) {
    uint32_t out = 0;
    sjo_scan_shard(buf, len, e, NULL, &out); /* in_string 0 in: bit1 of out is the parity */
    T |= (out & 7u) << (3 * e);
  }
  return T;
}
