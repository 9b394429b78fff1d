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

/* ------------------------------------------------------------ stage-2-lite
 * What the reference's stage 2 decides about ONE token from its bytes alone -- the leaves of
 * json_iterator::visit_primitive (src/generic/stage2/json_iterator.h L338-360):
 *   strings  tape_builder::visit_string (src/generic/stage2/tape_builder.h L186-205) -> stringparsing::parse_string
 *            (src/generic/stage2/stringparsing.h L146-190, handle_unicode_codepoint L55-98,
 *            jsoncharutils::codepoint_to_utf8 include/simdjson/generic/jsoncharutils.h L37-62); the record a string
 *            leaves in dom::document::string_buf is [uint32 length][bytes][0] (tape_builder.h on_start_string /
 *            on_end_string), the tape payload its offset;
 *   numbers  numberparsing::parse_number (include/simdjson/generic/numberparsing.h L860-961): grammar, int64 / uint64
 *            decision and value; a float is only recognised (type 'd'), not converted -- the reference additionally
 *            rejects floats whose VALUE is infinite (write_float L765-813), which this restatement does not model;
 *   atoms    atomparsing::is_valid_{true,false,null}_atom (include/simdjson/generic/atomparsing.h L45-95).
 * Written as plain sequential byte walks: no 32-byte chunks, no tables. */

static int sjo_hex(uint8_t c) {
  if (c >= '0' && c <= '9') return c - '0';
  if (c >= 'a' && c <= 'f') return c - 'a' + 10;
  if (c >= 'A' && c <= 'F') return c - 'A' + 10;
  return -1;
}
static uint8_t sjo_at(const uint8_t *buf, size_t len, size_t i) { return i < len ? buf[i] : (uint8_t)0x20; } /* beyond the end: padding */
static long sjo_hex4(const uint8_t *buf, size_t len, size_t i) {
  long v = 0;
  for (int k = 0; k < 4; k++) {
    int h = sjo_hex(sjo_at(buf, len, i + (size_t)k));
    if (h < 0) return -1;
    v = v * 16 + h;
  }
  return v;
}
static int sjo_is_struct_or_ws(uint8_t c) {
  return c == 0x20 || c == 0x09 || c == 0x0A || c == 0x0D || c == ',' || c == ':' || c == '[' || c == ']' || c == '{' || c == '}';
}

/* the string whose opening quote is at buf[pos]: unescaped bytes to dst (may be NULL: length only).
 * returns the unescaped length, -1 for an invalid escape (STRING_ERROR), -2 when the input ends first */
long sjo_parse_string(const uint8_t *buf, size_t len, size_t pos, uint8_t *dst) {
  size_t q = pos + 1;
  long out = 0;
  for (;;) {
    if (q >= len) return -2;
    uint8_t b = buf[q];
    if (b == '"') return out;
    if (b != '\\') {
      if (dst) dst[out] = b;
      out++; q++;
      continue;
    }
    uint8_t e = sjo_at(buf, len, q + 1);
    if (e != 'u') {
      uint8_t m = 0;
      switch (e) {
        case '"': m = 0x22; break; case '\\': m = 0x5C; break; case '/': m = 0x2F; break;
        case 'b': m = 0x08; break; case 'f': m = 0x0C; break; case 'n': m = 0x0A; break;
        case 'r': m = 0x0D; break; case 't': m = 0x09; break; default: break;
      }
      if (!m) return -1;
      if (dst) dst[out] = m;
      out++; q += 2;
      continue;
    }
    long cp = sjo_hex4(buf, len, q + 2);
    if (cp < 0) return -1;
    q += 6;
    if (cp >= 0xD800 && cp < 0xDC00) { /* high surrogate: a low one must follow as \uXXXX */
      if (sjo_at(buf, len, q) != '\\' || sjo_at(buf, len, q + 1) != 'u') return -1;
      long lo = sjo_hex4(buf, len, q + 2);
      if (lo < 0xDC00 || lo > 0xDFFF) return -1;
      cp = (((cp - 0xD800) << 10) | (lo - 0xDC00)) + 0x10000;
      q += 6;
    } else if (cp >= 0xDC00 && cp <= 0xDFFF) {
      return -1;
    }
    uint8_t u[4];
    int n;
    if (cp <= 0x7F) { u[0] = (uint8_t)cp; n = 1; }
    else if (cp <= 0x7FF) { u[0] = (uint8_t)(0xC0 | (cp >> 6)); u[1] = (uint8_t)(0x80 | (cp & 63)); n = 2; }
    else if (cp <= 0xFFFF) { u[0] = (uint8_t)(0xE0 | (cp >> 12)); u[1] = (uint8_t)(0x80 | ((cp >> 6) & 63)); u[2] = (uint8_t)(0x80 | (cp & 63)); n = 3; }
    else { u[0] = (uint8_t)(0xF0 | (cp >> 18)); u[1] = (uint8_t)(0x80 | ((cp >> 12) & 63)); u[2] = (uint8_t)(0x80 | ((cp >> 6) & 63)); u[3] = (uint8_t)(0x80 | (cp & 63)); n = 4; }
    if (dst) for (int k = 0; k < n; k++) dst[out + k] = u[k];
    out += n;
  }
}

/* the number that starts at buf[pos]; returns the tape type ('l', 'u', 'd') or 0 with *value = error_code */
static uint8_t sjo_number(const uint8_t *buf, size_t len, size_t pos, uint64_t *value) {
  const int neg = buf[pos] == '-';
  size_t q = pos + (size_t)neg;
  const size_t start = q;
  uint64_t i = 0;
  while (sjo_at(buf, len, q) >= '0' && sjo_at(buf, len, q) <= '9') { i = i * 10u + (uint64_t)(sjo_at(buf, len, q) - '0'); q++; }
  size_t digits = q - start;
  if (digits == 0 || (sjo_at(buf, len, start) == '0' && digits > 1)) { *value = SJO_NUMBER_ERROR; return 0; }
  int is_float = 0;
  if (sjo_at(buf, len, q) == '.') {
    is_float = 1;
    q++;
    const size_t fs = q;
    while (sjo_at(buf, len, q) >= '0' && sjo_at(buf, len, q) <= '9') q++;
    if (q == fs) { *value = SJO_NUMBER_ERROR; return 0; }
  }
  if (sjo_at(buf, len, q) == 'e' || sjo_at(buf, len, q) == 'E') {
    is_float = 1;
    q++;
    if (sjo_at(buf, len, q) == '-' || sjo_at(buf, len, q) == '+') q++;
    const size_t es = q;
    while (sjo_at(buf, len, q) >= '0' && sjo_at(buf, len, q) <= '9') q++;
    if (q == es) { *value = SJO_NUMBER_ERROR; return 0; }
  }
  const int dirty = !sjo_is_struct_or_ws(sjo_at(buf, len, q));
  if (is_float) {
    if (dirty) { *value = SJO_NUMBER_ERROR; return 0; }
    *value = (uint64_t)q; /* one past the token: the consumer converts [pos, q) */
    return 'd';
  }
  const size_t longest = neg ? 19 : 20;
  if (digits > longest) { *value = SJO_BIGINT_ERROR; return 0; }
  if (digits == longest) {
    if (neg) {
      if (i > ((uint64_t)1 << 63)) { *value = SJO_BIGINT_ERROR; return 0; }
    } else if (buf[pos] != '1' || i <= (uint64_t)INT64_MAX) { *value = SJO_BIGINT_ERROR; return 0; }
  }
  if (dirty) { *value = SJO_NUMBER_ERROR; return 0; }
  if (i > (uint64_t)INT64_MAX && !neg) { *value = i; return 'u'; }
  *value = neg ? (~i + 1u) : i;
  return 'l';
}

static int sjo_atom(const uint8_t *buf, size_t len, size_t pos, const char *word, size_t wl) {
  for (size_t k = 0; k < wl; k++)
    if (sjo_at(buf, len, pos + k) != (uint8_t)word[k]) return 0;
  return sjo_is_struct_or_ws(sjo_at(buf, len, pos + wl));
}

int sjo_tokens(const uint8_t *buf, size_t len, const uint32_t *idx, uint32_t n, uint8_t *type, uint64_t *payload, uint8_t *strbuf,
               size_t strbuf_cap, uint64_t *strbuf_len, uint32_t *n_strings, uint32_t *first_error_index) {
  uint64_t sb = 0;
  uint32_t ns = 0, first = 0xFFFFFFFFu;
  int err = SJO_SUCCESS;
  for (uint32_t k = 0; k < n; k++) {
    const size_t p = idx[k];
    const uint8_t c = buf[p];
    uint8_t t = 0;
    uint64_t v = 0;
    if (c == '{' || c == '}' || c == '[' || c == ']' || c == ':' || c == ',') {
      t = c;
    } else if (c == '"') {
      const long ul = sjo_parse_string(buf, len, p, NULL);
      if (ul < 0) {
        v = ul == -1 ? SJO_STRING_ERROR : SJO_UNCLOSED_STRING;
      } else {
        t = '"';
        v = (uint64_t)ul; /* becomes the record's offset below, when the buffer is large enough */
        sb += (uint64_t)ul + 5;
        ns++;
      }
    } else if (c <= '9' || c == '-') {
      /* json_iterator.h L342: `(*value - '0') < 10` is evaluated in int, so EVERY byte up to '9' takes the number path
       * (a stray '#' or 0x0C in value position is a NUMBER_ERROR, not a TAPE_ERROR) */
      t = sjo_number(buf, len, p, &v);
    } else if (c == 't') {
      if (sjo_atom(buf, len, p, "true", 4)) t = 't'; else v = SJO_T_ATOM_ERROR;
    } else if (c == 'f') {
      if (sjo_atom(buf, len, p, "false", 5)) t = 'f'; else v = SJO_F_ATOM_ERROR;
    } else if (c == 'n') {
      if (sjo_atom(buf, len, p, "null", 4)) t = 'n'; else v = SJO_N_ATOM_ERROR;
    } else {
      v = SJO_TAPE_ERROR;
    }
    type[k] = t;
    payload[k] = v;
    if (t == 0 && first == 0xFFFFFFFFu) { first = k; err = (int)v; }
  }
  *strbuf_len = sb;
  *n_strings = ns;
  *first_error_index = first;
  if (sb > strbuf_cap) return err == SJO_SUCCESS ? SJO_CAPACITY : err; /* nothing written, payloads keep the lengths */
  uint64_t off = 0;
  for (uint32_t k = 0; k < n; k++) {
    if (type[k] != '"') continue;
    const uint32_t l32 = (uint32_t)payload[k];
    for (int b = 0; b < 4; b++) strbuf[off + (uint64_t)b] = (uint8_t)(l32 >> (8 * b));
    sjo_parse_string(buf, len, idx[k], strbuf + off + 4);
    strbuf[off + 4 + l32] = 0;
    payload[k] = off;
    off += (uint64_t)l32 + 5;
  }
  return err;
}
