// sjb200_tape.h -- launcher of sjb200_tape.cu (stage-2-lite on the device)
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace sjb200 {

struct TokenTotals {
  unsigned long long string_bytes;  // bytes of string_buf the document's strings need (records of valid strings)
  unsigned long long first_error;   // (structural index << 8) | error_code of the first token in error, ~0 when none
  uint32_t n_strings;
  uint32_t reserved;
};

size_t tokens_scratch_bytes(uint32_t n);
// type[n], payload[n], strbuf[strbuf_capacity]: device memory; scratch: tokens_scratch_bytes(n) bytes, 8-byte aligned;
// stage: 1 = tiles staged through shared memory (the product path), 0 = every thread reads / writes global memory (kept as
// the A/B baseline of the staging, option tok_stage)
cudaError_t launch_tokens(const uint8_t *buf, uint64_t len, const uint32_t *idx, uint32_t n, uint8_t *type, uint64_t *payload, uint8_t *strbuf,
                          uint64_t strbuf_capacity, void *scratch, TokenTotals *tot_dev, int stage, cudaStream_t stream);

}  // namespace sjb200
