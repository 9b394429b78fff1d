// sjb200_docs.h -- launchers of sjb200_docs.cu (device-side epilogue of streaming scans, document boundary table)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "sjb200_params.h"

namespace sjb200 {

// what the streaming branches of finish() produce (json_structural_indexer.h L249-343)
struct StreamFinish {
  int32_t err;         // simdjson::error_code
  uint32_t n;          // n_structural_indexes (valid when n_written)
  uint32_t n_written;  // 0: the call returned before touching n (UNESCAPED_CHARS / internal error)
  uint32_t reserved;
};
// same layout as sjb200_doc_boundary (include/sjb200.h)
struct sjb200_doc_boundary_t {
  uint32_t index;  // structural index at which a document starts
  uint32_t byte;   // = structural_indexes[index]
};

// modes 1 / 2 (streaming_partial / streaming_final) behind a device-resident scan whose result block is `carry`
cudaError_t launch_stream_finish(const uint8_t *buf, uint32_t *idx, const Carry *carry, uint32_t len, int mode, StreamFinish *out_dev, StreamFinish *out_host,
                                 cudaStream_t stream);
// modes 3..6 (RS-delimited / comma-delimited streams): filter the device-resident index array in place and run the
// rest of finish(); n = structurals to consider (host-known), scratch: filter_scratch_words(n) uint32 words
size_t filter_scratch_words(uint32_t n);
cudaError_t launch_stream_filter(const uint8_t *buf, uint32_t len, uint32_t *idx, uint32_t n, int mode, uint32_t flags, uint32_t *scratch, StreamFinish *out_dev,
                                 StreamFinish *out_host, cudaStream_t stream);
// table of document starts of a whitespace-separated stream; scratch: doc_table_scratch_words(n) uint32 words
size_t doc_table_scratch_words(uint32_t n);
cudaError_t launch_doc_table(const uint8_t *buf, const uint32_t *idx, uint32_t n, uint32_t *scratch, sjb200_doc_boundary_t *table, uint32_t capacity,
                             uint32_t *ndocs_dev, cudaStream_t stream);

}  // namespace sjb200
