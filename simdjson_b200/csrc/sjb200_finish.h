// sjb200_finish.h -- host-side epilogue of a stage-1 call: error precedence, sentinels and the
// document-boundary fix-ups of the streaming modes.
//
// Product code (not the oracle).  Restates, on top of the index array the GPU produced:
//   json_structural_indexer::finish            src/generic/stage1/json_structural_indexer.h L249-397
//   find_next_document_index                   src/generic/stage1/find_next_document_index.h L39-98
//   find_next_document_index_json_sequence     ... L126-267
//   filter_comma_delimited                     ... L288-369
// These walk the *tail* of the index array (whitespace-separated streams) or filter it serially
// (RFC 7464 / comma-delimited streams); they are O(last document) / rarely used, and SURVEY.md
// section 8(a12) keeps them on the host.  The structural characters are reached through a small
// accessor so the same code serves host arrays and windows gathered from device memory.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "sjb200_common.h"

namespace sjb200 {

// what the walk needs to know about structural i
class StructuralReader {
 public:
  virtual ~StructuralReader() {}
  virtual uint32_t position(uint32_t i) = 0;   // structural_indexes[i]
  virtual uint8_t character(uint32_t i) = 0;   // buf[structural_indexes[i]]
};

class HostStructuralReader final : public StructuralReader {
 public:
  HostStructuralReader(const uint8_t *buf, const uint32_t *idx) : buf_(buf), idx_(idx) {}
  uint32_t position(uint32_t i) override { return idx_[i]; }
  uint8_t character(uint32_t i) override { return buf_[idx_[i]]; }

 private:
  const uint8_t *buf_;
  const uint32_t *idx_;
};

// Number of structurals that belong to complete documents (whitespace-separated stream).
uint32_t complete_document_count(StructuralReader &r, uint32_t n);

constexpr uint32_t kDocumentTooLarge = 0xFFFFFFFFu;

// In-place filters for the RS-delimited and comma-delimited stream formats (host arrays).
// On return n holds the filtered count; the return value is the number of indexes to keep
// (0 = nothing usable, kDocumentTooLarge = a document started but did not fit).
uint32_t filter_record_separators(const uint8_t *buf, size_t len, uint32_t *idx, uint32_t &n, bool is_final,
                                  uint32_t &next_batch_start);
uint32_t filter_root_commas(const uint8_t *buf, size_t len, uint32_t *idx, uint32_t &n, bool is_final,
                            uint32_t &next_batch_start);

// where finish() writes index words (host array or device array)
class IndexWriter {
 public:
  virtual ~IndexWriter() {}
  virtual bool set3(uint32_t n, uint32_t a, uint32_t b, uint32_t c) = 0;  // idx[n..n+2] = a,b,c
  virtual bool final_fixup(uint32_t m, uint32_t len) = 0;                 // idx[m+1]=idx[m]; idx[m]=len
};
class HostIndexWriter final : public IndexWriter {
 public:
  explicit HostIndexWriter(uint32_t *idx) : idx_(idx) {}
  bool set3(uint32_t n, uint32_t a, uint32_t b, uint32_t c) override { idx_[n] = a; idx_[n + 1] = b; idx_[n + 2] = c; return true; }
  bool final_fixup(uint32_t m, uint32_t len) override { idx_[m + 1] = idx_[m]; idx_[m] = len; return true; }

 private:
  uint32_t *idx_;
};

// The reference's finish() (json_structural_indexer.h L249-397) after the scan.
//   reader/writer : access to the index array (host or device)
//   host_buf/host_idx : only for the RS / comma modes, which filter on host arrays (may be null -> caller staged them)
struct FinishInput {
  int mode;
  size_t len;          // trimmed length that was scanned
  uint64_t count;      // structurals found
  uint32_t state;      // final scanner state (bit1: inside a string)
  uint32_t flags;
  bool sentinels_written;  // the scan kernel already stored idx[n..n+2] (device-resident calls)
};

int finish_stage1(const FinishInput &in, StructuralReader &reader, IndexWriter &writer, uint32_t *n_inout,
                  const uint8_t *host_buf, uint32_t *host_idx, bool *host_idx_dirty);

// trim_partial_utf8: drop an unfinished trailing sequence (last3 = up to 3 last bytes, last3[k-1] is the final byte)
size_t trim_partial_utf8_tail(const uint8_t *tail, size_t tail_len, size_t len);

}  // namespace sjb200
