/*
 * sj_oracle.h -- CPU restatement of simdjson stage 1 / minify / validate_utf8.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be linked, imported or
 * executed by the product path (simdjson_b200/, include/).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker or the CPU baseline -- never as the thing
 * shipped or measured as "ours".
 *
 * Parity status: PINNED.  The restatement is checked against
 *   (1) the reference itself compiled from /root/reference/singleheader
 *       (oracle/_ref/libsj_ref.so, recipe: oracle/Makefile) on fixtures and
 *       seeded fuzz, tests/test_oracle_pinning.py;
 *   (2) golden vectors generated from the reference and committed under
 *       tests/golden/ (generator: oracle/gen_golden.py).
 *
 * It is a byte-at-a-time state machine on purpose: it shares no structure
 * with either the reference's 64-byte SIMD blocks or the CUDA kernels'
 * bit-plane formulation, so agreement between the three is meaningful.
 */
#ifndef SJ_ORACLE_H
#define SJ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numeric values of simdjson::error_code (include/simdjson/error.h L19-54) */
enum {
  SJO_SUCCESS = 0,
  SJO_CAPACITY = 1,
  SJO_MEMALLOC = 2,
  SJO_TAPE_ERROR = 3,
  SJO_STRING_ERROR = 5,
  SJO_T_ATOM_ERROR = 6,
  SJO_F_ATOM_ERROR = 7,
  SJO_N_ATOM_ERROR = 8,
  SJO_NUMBER_ERROR = 9,
  SJO_BIGINT_ERROR = 10,
  SJO_UTF8_ERROR = 11,
  SJO_EMPTY = 13,
  SJO_UNESCAPED_CHARS = 14,
  SJO_UNCLOSED_STRING = 15,
  SJO_UNEXPECTED_ERROR = 24
};

/* simdjson::stage1_mode (include/simdjson/internal/dom_parser_implementation.h L22-27) */
enum {
  SJO_REGULAR = 0,
  SJO_STREAMING_PARTIAL = 1,
  SJO_STREAMING_FINAL = 2,
  SJO_JSON_SEQUENCE_PARTIAL = 3,
  SJO_JSON_SEQUENCE_FINAL = 4,
  SJO_COMMA_DELIMITED_PARTIAL = 5,
  SJO_COMMA_DELIMITED_FINAL = 6
};

/* number of uint32 words a caller must provide for a given capacity:
 * ROUNDUP(capacity,64)+9  (include/simdjson/generic/dom_parser_implementation.h L66-82) */
size_t sjo_index_capacity(size_t capacity);

/*
 * stage 1: json_structural_indexer::index<128> + finish
 * (src/generic/stage1/json_structural_indexer.h L193-218, L249-397).
 *   idx        : >= sjo_index_capacity(capacity) words
 *   n_inout    : n_structural_indexes; left untouched on the early-return paths
 *                exactly like the reference does.
 * returns the error_code.
 */
int sjo_stage1(const uint8_t *buf, size_t len, size_t capacity, int mode,
               uint32_t *idx, uint32_t *n_inout);

/* minify: json_minifier::minify<128> (src/generic/stage1/json_minifier.h L68-97) */
int sjo_minify(const uint8_t *buf, size_t len, uint8_t *dst, size_t *dst_len);

/* validate_utf8: generic_validate_utf8 (src/generic/stage1/utf8_validator.h L18-34);
 * returns 1 for valid, 0 for invalid. */
int sjo_validate_utf8(const uint8_t *buf, size_t len);

/* stage-2-lite: per structural index what the reference's stage 2 decides from the token's
 * bytes alone.  type[k]: the tape_type char ('{' '}' '[' ']' '"' 'l' 'u' 'd' 't' 'f' 'n'; ':' and ',' for those
 * operators) or 0 for a token in error; payload[k]: string -> offset of its record ([u32 length][bytes][0], the layout of
 * dom::document::string_buf) in strbuf, 'l' / 'u' -> the value, 'd' -> byte offset one past the number, type 0 -> the
 * error_code.  Returns the error of the first token in error (document order), CAPACITY if strbuf is too small. */
int sjo_tokens(const uint8_t *buf, size_t len, const uint32_t *idx, uint32_t n, uint8_t *type, uint64_t *payload, uint8_t *strbuf,
               size_t strbuf_cap, uint64_t *strbuf_len, uint32_t *n_strings, uint32_t *first_error_index);
/* one string (opening quote at buf[pos]): unescaped length (bytes to dst unless NULL), -1 invalid escape, -2 unterminated */
long sjo_parse_string(const uint8_t *buf, size_t len, size_t pos, uint8_t *dst);

/* helpers exposed for unit tests */
uint64_t sjo_scan_shard(const uint8_t *buf, size_t len, uint32_t state_in, uint32_t *idx, uint32_t *state_out);
uint32_t sjo_transducer(const uint8_t *buf, size_t len);
size_t sjo_trim_partial_utf8(const uint8_t *buf, size_t len);
uint32_t sjo_find_next_document_index(const uint8_t *buf, const uint32_t *idx, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
