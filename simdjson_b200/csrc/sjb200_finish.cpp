// sjb200_finish.cpp -- see sjb200_finish.h
#include "sjb200_finish.h"

namespace sjb200 {

namespace {

enum Role : uint8_t { kValue = 0, kSeparator, kOpenObject, kCloseObject, kOpenArray, kCloseArray };

inline Role role_of(uint8_t c) {
  switch (c) {
    case ':': case ',': return kSeparator;
    case '{': return kOpenObject;
    case '}': return kCloseObject;
    case '[': return kOpenArray;
    case ']': return kCloseArray;
    default: return kValue;
  }
}

inline bool is_json_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }

}  // namespace

// find_next_document_index (find_next_document_index.h L39-98): scan the structurals from the end
// for the last place where one value directly follows another (no ',' / ':' between them); the
// structurals after that place form the last document, which is complete iff its brackets balance.
uint32_t complete_document_count(StructuralReader &r, uint32_t n) {
  if (n == 0) return 0;
  int net_objects = 0, net_arrays = 0;  // opens minus closes seen so far (walking backwards)
  for (uint32_t i = n - 1; i > 0; --i) {
    const Role cur = role_of(r.character(i));
    if (cur == kSeparator) continue;
    if (cur == kCloseObject) { --net_objects; continue; }
    if (cur == kCloseArray) { --net_arrays; continue; }
    if (cur == kOpenObject) ++net_objects;
    if (cur == kOpenArray) ++net_arrays;
    const Role before = role_of(r.character(i - 1));
    if (before == kOpenObject || before == kOpenArray || before == kSeparator) continue;
    // structural i starts a new document
    return (net_objects == 0 && net_arrays == 0) ? n : i;
  }
  switch (role_of(r.character(0))) {
    case kOpenObject: ++net_objects; break;
    case kCloseObject: --net_objects; break;
    case kOpenArray: ++net_arrays; break;
    case kCloseArray: --net_arrays; break;
    default: break;
  }
  return (net_objects == 0 && net_arrays == 0) ? n : 0;
}

// find_next_document_index_json_sequence (find_next_document_index.h L126-267).
// RS (0x1E) opens every record.  Stage 1 sees RS as a scalar byte, so a scalar value glued to its
// RS is not in the index array and an RS itself is; rewrite the array so that it holds the value
// starts instead of the RS bytes, then cut at the last RS for partial batches.
uint32_t filter_record_separators(const uint8_t *buf, size_t len, uint32_t *idx, uint32_t &n, bool is_final,
                                  uint32_t &next_batch_start) {
  next_batch_start = uint32_t(len);
  if (n == 0) return 0;
  uint32_t kept = 0, separators = 0, last_separator = 0;
  uint32_t i = 0;
  while (i < n) {
    const uint32_t at = idx[i];
    if (buf[at] != 0x1E) {
      idx[kept++] = at;
      ++i;
      continue;
    }
    // a run "RS (ws | RS)*": count every RS in it and find the byte the record's value starts at
    ++separators;
    last_separator = at;
    uint32_t value_at = at + 1;
    while (value_at < len && (is_json_space(buf[value_at]) || buf[value_at] == 0x1E)) {
      if (buf[value_at] == 0x1E) { ++separators; last_separator = value_at; }
      ++value_at;
    }
    ++i;
    while (i < n && idx[i] < value_at) ++i;  // structurals inside the run are separators we already counted
    if (value_at < len) {
      const Role role = role_of(buf[value_at]);
      const bool stage1_has_it = (i < n && idx[i] == value_at);
      if (role == kValue && !stage1_has_it) idx[kept++] = value_at;  // scalar glued to its RS
    }
  }
  n = kept;
  if (n == 0) return 0;
  if (separators == 0) {
    if (!is_final) return 0;
    HostStructuralReader r(buf, idx);
    return complete_document_count(r, n);
  }
  if (is_final) return n;
  next_batch_start = last_separator;
  if (separators < 2) return kDocumentTooLarge;
  uint32_t keep = n;
  while (keep > 0 && idx[keep - 1] >= last_separator) --keep;
  return keep;
}

// filter_comma_delimited (find_next_document_index.h L288-369): commas at nesting depth 0 separate
// documents; drop them so the array looks like a whitespace-separated stream.
uint32_t filter_root_commas(const uint8_t *buf, size_t len, uint32_t *idx, uint32_t &n, bool is_final,
                            uint32_t &next_batch_start) {
  next_batch_start = uint32_t(len);
  if (n == 0) return 0;
  int depth = 0;
  uint32_t kept = 0, root_commas = 0, last_root_comma = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t at = idx[i];
    const uint8_t c = buf[at];
    const Role role = role_of(c);
    if (role == kOpenObject || role == kOpenArray) ++depth;
    else if (role == kCloseObject || role == kCloseArray) --depth;
    else if (c == ',' && depth == 0) {
      ++root_commas;
      last_root_comma = at;
      continue;
    }
    idx[kept++] = at;
  }
  n = kept;
  if (n == 0) return 0;
  HostStructuralReader r(buf, idx);
  if (is_final) return complete_document_count(r, n);
  if (root_commas == 0) return kDocumentTooLarge;
  next_batch_start = last_root_comma + 1;
  uint32_t keep = n;
  while (keep > 0 && idx[keep - 1] >= last_root_comma) --keep;
  if (keep == 0) return 0;
  n = keep;
  return complete_document_count(r, n);
}

int finish_stage1(const FinishInput &in, StructuralReader &reader, IndexWriter &writer, uint32_t *n_inout,
                  const uint8_t *host_buf, uint32_t *host_idx, bool *host_idx_dirty) {
  if (in.flags & kFlagInternal) return kUnexpectedError;
  const bool unclosed = (in.state >> 1) & 1u;
  if (in.mode == kRegular && unclosed) return kUnclosedString;   // L255-259
  if (in.flags & kFlagCtl) return kUnescapedChars;                      // L261-263
  uint32_t n = uint32_t(in.count);
  const uint32_t len32 = uint32_t(in.len);
  *n_inout = n;
  if (!in.sentinels_written && !writer.set3(n, len32, len32, 0)) return kUnexpectedError;  // L284-286
  if (n == 0) return kEmpty;                                             // L289-291
  switch (in.mode) {
    case kStreamingPartial: {                                            // L295-317
      if (unclosed) { n--; *n_inout = n; if (n == 0) return kCapacity; }
      const uint32_t m = complete_document_count(reader, n);
      if (m == 0 && n > 0) {
        if (reader.position(0) == 0) return kCapacity;
        *n_inout = 0;
        return kEmpty;
      }
      *n_inout = m;
      break;
    }
    case kStreamingFinal: {                                              // L318-343
      if (unclosed) n--;
      const uint32_t m = complete_document_count(reader, n);
      *n_inout = m;
      if (!writer.final_fixup(m, len32)) return kUnexpectedError;
      if (m == 0) return kEmpty;
      break;
    }
    case kJsonSequencePartial:
    case kCommaDelimitedPartial: {                                      // L344-359, L367-384
      if (unclosed) { n--; *n_inout = n; if (n == 0) return kCapacity; }
      uint32_t next_start = len32;
      const uint32_t m = (in.mode == kJsonSequencePartial)
                             ? filter_record_separators(host_buf, in.len, host_idx, n, false, next_start)
                             : filter_root_commas(host_buf, in.len, host_idx, n, false, next_start);
      *host_idx_dirty = true;
      *n_inout = n;
      if (m == kDocumentTooLarge) return kCapacity;
      if (m == 0) { *n_inout = 0; return kEmpty; }
      *n_inout = m;
      host_idx[m] = next_start;
      break;
    }
    case kJsonSequenceFinal:
    case kCommaDelimitedFinal: {                                        // L360-366, L385-393
      if (unclosed) n--;
      uint32_t next_start = len32;
      const uint32_t m = (in.mode == kJsonSequenceFinal)
                             ? filter_record_separators(host_buf, in.len, host_idx, n, true, next_start)
                             : filter_root_commas(host_buf, in.len, host_idx, n, true, next_start);
      *host_idx_dirty = true;
      *n_inout = m;
      host_idx[m + 1] = host_idx[m];
      host_idx[m] = len32;
      if (m == 0) return kEmpty;
      break;
    }
    default:
      break;
  }
  return (in.flags & kFlagUtf8) ? kUtf8Error : kSuccess;         // L395-396
}


// trim_partial_utf8 (json_structural_indexer.h L156-174) on the last <= 3 bytes
size_t trim_partial_utf8_tail(const uint8_t *tail, size_t tail_len, size_t len) {
  // tail[tail_len-1] is the last byte of the input
  if (tail_len >= 1 && tail[tail_len - 1] >= 0xC0) return len - 1;
  if (tail_len >= 2 && tail[tail_len - 2] >= 0xE0) return len - 2;
  if (tail_len >= 3 && tail[tail_len - 3] >= 0xF0) return len - 3;
  return len;
}

}  // namespace sjb200
