// sjb200_common.h -- constants shared by the kernels, the C ABI and the host epilogue (no CUDA types).
#pragma once
#include <stdint.h>

namespace sjb200 {

// flags a scan accumulates (ScanParams::flags, sjb200_shard_result::flags)
enum : uint32_t {
  kFlagUtf8 = 1u,      // some byte violates UTF-8 well-formedness
  kFlagCtl = 2u,       // unescaped control character inside a string
  kFlagInternal = 4u,  // a bounded spin expired (never expected; reported as UNEXPECTED_ERROR)
};

// simdjson::error_code / stage1_mode values (mirrors include/sjb200.h; kept here so host-only
// translation units do not need the C header)
enum : int { kSuccess = 0, kCapacity = 1, kMemalloc = 2, kUtf8Error = 11, kEmpty = 13, kUnescapedChars = 14, kUnclosedString = 15,
             kUnexpectedError = 24 };
enum : int { kRegular = 0, kStreamingPartial = 1, kStreamingFinal = 2, kJsonSequencePartial = 3, kJsonSequenceFinal = 4,
             kCommaDelimitedPartial = 5, kCommaDelimitedFinal = 6 };

}  // namespace sjb200
