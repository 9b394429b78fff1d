// sjb200_params.h -- geometry, scan kinds and the launch parameter block shared by the kernels, the C-ABI host
// code and the host SIMT emulation of the scan4 kernel (tests/simt_emul.cpp).  No CUDA types.
#pragma once
#include <stdint.h>

#include "sjb200_common.h"

namespace sjb200 {

// ---- geometry (one "tile" is what one CTA scans per loop iteration)
constexpr int kUnitsPerLane = 4;                     // 32-byte transposition units per lane
constexpr int kLaneBytes = 32 * kUnitsPerLane;       // 128 B = one TMA 128B-swizzle row
constexpr int kWarpBytes = 32 * kLaneBytes;          // 4 KiB
constexpr int kWarps = 8;
constexpr int kThreads = 32 * kWarps;
constexpr int kTileBytes = kWarps * kWarpBytes;      // 32 KiB
constexpr int kTileRows = kTileBytes / 128;          // 256 rows of 128 B (max TMA box dim)
#ifndef SJB200_STAGES
#define SJB200_STAGES 1
#endif
#ifndef SJB200_MIN_CTAS
#define SJB200_MIN_CTAS 2
#endif
constexpr int kStages = SJB200_STAGES;      // shared-memory tile buffers per CTA
constexpr int kMinCtasPerSm = SJB200_MIN_CTAS;  // __launch_bounds__ occupancy target
// A CTA scans a "super-tile" of up to kMaxSub consecutive tiles before it consults the look-back chain once:
// the masks of every tile wait in shared memory (8 words per lane per tile) until the incoming state is known.
// (Parking them in an L2-resident global scratch instead, to fit 4 CTAs per SM, was measured slower: the SM is
// issue-bound, not latency-bound.)
#ifndef SJB200_SCAN4_MIN_CTAS
#define SJB200_SCAN4_MIN_CTAS 2
#endif
#ifndef SJB200_MAX_SUB
#define SJB200_MAX_SUB 8
#endif
constexpr int kMaxSub = SJB200_MAX_SUB;
constexpr int kMaxRanks = 8;                          // GPUs of one node that can share a scan (sjb200_comm)
constexpr int kXchgSteps = 64;                        // sharded passes whose exchange records an exchange window holds (two rounds each)
constexpr int kCtlBytes = 1024;                       // control block
constexpr int kLutBytes = 64 * 64;                    // composed-transducer table
constexpr int kEmitBytes = kWarps * 1024;             // per-warp emit scratch (128 mask words + 128 counts)
constexpr int kMaskSlotBytes = kThreads * 8 * 4;      // one tile's masks
constexpr int kSmemBytesScan = kStages * kTileBytes + 1024 /*alignment slack*/ + kCtlBytes + kLutBytes + kEmitBytes + kMaxSub * kMaskSlotBytes;
constexpr int kSmemBytesUtf8 = kStages * kTileBytes + 1024 + kCtlBytes;
constexpr int smem_bytes_for(int kind) { return kind == 2 ? kSmemBytesUtf8 : kSmemBytesScan; }

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
  uint32_t sub_per_super;   // R: tiles per super-tile (1..kMaxSub); the look-back chain has one element per super-tile
  uint32_t nsuper;          // ceil(ntiles / R)
  uint32_t full_tiles;      // document tiles that lie entirely inside floor(len/128) rows
  uint32_t epoch;           // tags look-back descriptors so they need no per-launch reset
  uint32_t *idx_out;        // kIndex: device index array
  uint8_t *dst;             // kMinify: device output
  uint32_t write_sentinels; // kIndex: the launch that scans the last tile also stores idx[n]=idx[n+1]=len, idx[n+2]=0
  const Carry *carry_in;    // null: zero state, zero count
  Carry *carry_out;
  Carry *carry_out_host;    // scan4: optional second copy of the result in pinned host memory (saves the copy engine a trip between launches)
  uint32_t *flags;          // accumulated with atomicOr; zero between launches (the last CTA moves it to carry_out->flags)
  unsigned long long *count_desc;  // [nsuper] the look-back chain
  uint32_t *ticket;         // [0] next ticket, [1] CTAs finished, [2] scan4: aggregates published so far (4 words, zero between launches)
  uint32_t *park;           // scan4, deferred mode: scratch ring for parked masks, scan4_park_words(grid) words (stays in L2)
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
