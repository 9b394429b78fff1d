// sjb200_params.h -- geometry, scan kinds and the launch parameter block shared by the kernels, the C-ABI host
// code and the host SIMT emulation of the scan4 kernel (tests/simt_emul.cpp).  No CUDA types.
#pragma once
#include <stdint.h>

#include "sjb200_common.h"

namespace sjb200 {

// ---- geometry: launch parameters count the document in 32 KiB "tiles" (chunk launches start at tile boundaries);
// the kernels read it in 4 KiB blocks of 32 rows x 128 B
constexpr int kTileBytes = 32 * 1024;
constexpr int kTileRows = kTileBytes / 128;
constexpr int kMaxRanks = 8;                          // GPUs of one node that can share a scan (sjb200_comm)
constexpr int kXchgSteps = 64;                        // sharded passes whose exchange records an exchange window holds (two rounds each)
#ifndef SJB200_SCAN4_MIN_CTAS
#define SJB200_SCAN4_MIN_CTAS 2                       // CTAs per SM of the 8-scan-warp build of scan4
#endif

// ---- scan kinds
enum : int { kIndex = 0, kMinify = 1, kUtf8 = 2 };

// Scanner state between consecutive launches of one document (chunked streaming,
// multi-GPU shards).  state: bit0 escape, bit1 in_string, bit2 prev_scalar.
struct Carry {
  uint64_t count;     // structurals (kIndex) or kept bytes (kMinify) emitted so far
  uint32_t state;
  uint32_t ttable;    // out only: the 6-bit transducer T(e) of everything this launch scanned
  uint32_t flags;     // out only: kFlag* bits raised by this launch (the launch also clears ScanParams::flags again)
  uint32_t reserved;
};

struct ScanParams {
  const uint8_t *buf;       // device pointer to byte 0 of the document (or shard)
  uint64_t len;             // document length in bytes (<= 4 GiB - 1); bytes past it read as 0x20
  uint32_t pos_base;        // added to every emitted index (0: positions relative to buf)
  uint32_t prev_word;       // the 4 bytes that precede buf[0] (0x20202020 at start of document)
  uint32_t check_eof;       // 1: this launch scans the last tile -> flag a truncated UTF-8 sequence
  uint32_t use_tma;         // 1: full tiles arrive by cp.async.bulk.tensor (buf 16 B aligned)
  uint32_t tile_begin;      // first document tile of this launch (chunked streaming)
  uint32_t ntiles;          // tiles in this launch: document tiles [tile_begin, tile_begin+ntiles)
  uint32_t epoch;           // tags look-back descriptors so they need no per-launch reset
  uint32_t *idx_out;        // kIndex: device index array
  uint8_t *dst;             // kMinify: device output
  uint32_t write_sentinels; // kIndex: the launch that scans the last tile also stores idx[n]=idx[n+1]=len, idx[n+2]=0
  const Carry *carry_in;    // null: zero state, zero count
  Carry *carry_out;
  Carry *carry_out_host;    // scan4: optional second copy of the result in pinned host memory (saves the copy engine a trip between launches)
  uint32_t *flags;          // accumulated with atomicOr; zero between launches (the last CTA moves it to carry_out->flags)
  unsigned long long *count_desc;  // the look-back chain: one descriptor per scan4 element of the launch
  uint32_t *ticket;         // [0] next ticket, [1] CTAs finished, [2] scan4: aggregates published so far (4 words, zero between launches)
  uint32_t *park;           // scan4 with emit warps: scratch ring for parked masks, scan4_park_words(grid) words (stays in L2)
  unsigned long long *debug;  // optional [ntiles][8] timeline (globaltimer ns) for tuning; null in production
  // multi-GPU exchange fused into the scan (scan4): the launch's last CTA stores the shard record {count, state out,
  // transducer, flags} into EVERY rank's exchange window over NVLink (peer-mapped device memory), tagged with xchg_seq --
  // the path's one exchange step without a collective launch.  xchg_nranks == 0: no exchange.
  unsigned long long *xchg_peer[kMaxRanks];  // [r] = base of rank r's window: [slots][kMaxRanks][2] words
  uint32_t xchg_nranks, xchg_rank, xchg_slot, xchg_seq;
};

#if defined(__CUDACC__)
#define SJ_PARAMS_HD __host__ __device__
#else
#define SJ_PARAMS_HD
#endif
// one shard record as two independently tagged 64-bit words (8-byte stores are single transactions):
//   w0 = seq[30:0] << 33 | count[32:0]        w1 = seq << 32 | flags << 16 | ttable << 8 | state_out
SJ_PARAMS_HD inline unsigned long long xchg_word0(uint32_t seq, uint64_t count) {
  return ((unsigned long long)(seq & 0x7FFFFFFFu) << 33) | (count & 0x1FFFFFFFFull);
}
SJ_PARAMS_HD inline unsigned long long xchg_word1(uint32_t seq, uint32_t state, uint32_t ttable, uint32_t flags) {
  return ((unsigned long long)seq << 32) | ((unsigned long long)(flags & 0xFFu) << 16) | ((unsigned long long)(ttable & 0x3Fu) << 8) | (state & 7u);
}
SJ_PARAMS_HD inline bool xchg_complete(unsigned long long w0, unsigned long long w1, uint32_t seq) {
  return uint32_t(w0 >> 33) == (seq & 0x7FFFFFFFu) && uint32_t(w1 >> 32) == seq;
}
SJ_PARAMS_HD inline uint64_t xchg_count(unsigned long long w0) { return w0 & 0x1FFFFFFFFull; }

}  // namespace sjb200
