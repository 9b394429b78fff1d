// sjb200_kernels.cu -- the sm_100a scan kernels: stage-1 structural indexing,
// minify and UTF-8 validation in ONE pass over the input.
//
// What it replaces in the reference (CPU, 64-byte SIMD blocks, strictly serial carries):
//   json_structural_indexer::index<128> / step / next   src/generic/stage1/json_structural_indexer.h L193-247
//   json_scanner::next, json_string_scanner::next, json_escape_scanner::next
//                                                        json_scanner.h L134-157, json_string_scanner.h L62-85,
//                                                        json_escape_scanner.h L50-71
//   bit_indexer::write                                   json_structural_indexer.h L93-122 (src/icelake.cpp L129-160)
//   utf8_checker                                         utf8_lookup4_algorithm.h L145-202
//   json_minifier::minify<128>                           json_minifier.h L68-97
//   generic_validate_utf8                                utf8_validator.h L18-34
//
// Design (B200-first, not a port):
//   * persistent CTAs pull 32 KiB tiles from an atomic ticket counter (forward progress of the
//     chained scans does not depend on co-residency, so two parsers may run concurrently);
//   * a tile arrives in shared memory by ONE cp.async.bulk.tensor (TMA) with the 128-byte
//     swizzle, double-buffered behind an mbarrier; each lane then owns one 128-byte row and
//     reads it with conflict-free LDS.128;
//   * a lane transposes 32 bytes into 8 bit planes and evaluates every class / UTF-8 rule as
//     boolean algebra on planes (sjb200_bits.cuh); there is no per-byte code anywhere;
//   * carries: the scanner state is a 2-state transducer T(e) per chunk (6 bits).  Lanes are
//     resolved with ballots (+ one 32-bit addition for the escape chain), warps through shared
//     memory, tiles through a decoupled look-back chain on T, and output offsets through a
//     second decoupled look-back chain on counts (exact for ANY input, valid JSON or not);
//   * indexes leave the SM in global order straight from registers.
#include "sjb200_kernels.cuh"

#include "sjb200_bits.cuh"
#include "sjb200_scan4.cuh"
#include "sjb200_utf8.cuh"

namespace sjb200 {

constexpr int W = kUnitsPerLane;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kSpinLimit = 1u << 25;  // bounded spins: a stuck chain becomes kFlagInternal, never a hang

// look-back descriptor status
enum : uint32_t { kNone = 0, kAgg = 1, kInc = 2 };

struct Control {
  unsigned long long full_bar[kStages];
  uint32_t stage_super[kStages];      // super-tile (ticket) the stage's tile belongs to
  uint32_t stage_r[kStages];          // ... and its position inside the super-tile
  uint32_t stage_prev16[kStages][4];  // the 16 bytes before the tile (UTF-8 halo + boundary state)
  uint32_t next_super, next_r, next_n;  // loader state: which tile to fetch next
  uint32_t bstate_exact;
  uint32_t warpT[kWarps];
  uint32_t cntw[kMaxSub][2][kWarps];  // outputs per tile of the super-tile, per in-string polarity, per warp
  // look-back scratch, one slot per warp = per 128-element segment of the window
  uint32_t segHas[kWarps];            // the segment contains an inclusive prefix (or reaches element 0)
  uint32_t segF[kWarps];              // composed transducer of the segment's relevant elements | kValidBit
  uint32_t segIncT[kWarps];           // inclusive transducer prefix found in the segment | kValidBit if real
  uint32_t segIncC[kWarps];
  uint32_t segCount[kWarps];
};
static_assert(sizeof(Control) <= kCtlBytes, "control block");

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x4000;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void tma_load_tile(void *dst, const CUtensorMap *map, unsigned long long *bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(row)
      : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SJ_TRACE(slot)                                                                       \
  do {                                                                                       \
    if (p.debug != nullptr && tid == 0) p.debug[uint64_t(tile) * 8 + (slot)] = globaltimer_ns(); \
  } while (0)

// byte offset inside a tile -> offset in the 128B-swizzled shared-memory image
// (TMA SWIZZLE_128B: 16-byte chunk index bits [4,7) ^= row bits [7,10))
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ ((off >> 3) & 0x70u); }

// ------------------------------------------------------------- the look-back chain
// One 64-bit descriptor per chain element (= super-tile):  [63:46] epoch  [45:44] status  [43:38] transducer  [37:0] payload
//   kAgg: transducer = T of the element; payload = cnt[1]<<19 | cnt[0]: its outputs for either in-string polarity
//         at its start (the e / c bits of its incoming state are already baked in: they were read off the bytes
//         before it)
//   kInc: transducer = composition of elements [0, i]; payload = outputs of elements [0, i]
__device__ __forceinline__ unsigned long long pack_agg(uint32_t epoch, uint32_t T, uint32_t c0, uint32_t c1) {
  return ((unsigned long long)epoch << 46) | ((unsigned long long)kAgg << 44) | ((unsigned long long)(T & 63u) << 38) |
         ((unsigned long long)c1 << 19) | c0;
}
__device__ __forceinline__ unsigned long long pack_inc(uint32_t epoch, uint32_t T, uint32_t count) {
  return ((unsigned long long)epoch << 46) | ((unsigned long long)kInc << 44) | ((unsigned long long)(T & 63u) << 38) | count;
}

#ifndef SJB200_SPIN_NS
#define SJB200_SPIN_NS 100
#endif
constexpr int kLook = 4;                      // descriptors per lane: a warp covers 128 tiles of the window
constexpr int kSegTiles = 32 * kLook;
constexpr uint32_t kValidBit = 0x100u;        // marks "this composed transducer exists"

__device__ __forceinline__ uint32_t lut_compose(const uint8_t *lut, uint32_t newer, uint32_t older) {
  return lut[((newer & 63u) << 6) | (older & 63u)];
}
__device__ __forceinline__ uint32_t compose_opt(const uint8_t *lut, uint32_t newer, uint32_t older) {  // either may be absent (no kValidBit)
  if (!(newer & kValidBit)) return older;
  if (!(older & kValidBit)) return newer;
  return lut_compose(lut, newer, older) | kValidBit;
}
__device__ __forceinline__ uint32_t apply_opt(uint32_t F, uint32_t state) { return (F & kValidBit) ? tt_apply(F & 63u, state) : state; }

// Resolve tile `tile` (>= 1) of this launch: the scanner state entering it, the number of outputs before
// it, and the composed transducer of tiles [0, tile).  Called by ALL warps of the CTA (it synchronises):
// warp w looks at tiles tile-1-128w .. tile-128(w+1), so one round trip covers a window of 1024 tiles.
// The walk stops at the nearest tile that already published an inclusive prefix; everything newer only
// has an aggregate, and because an aggregate's count depends on the in-string polarity at its tile, the
// states are folded forwards (oldest to newest) with parallel suffix scans.
__device__ void resolve_tile(const ScanParams &p, Control *ctl, const uint8_t *lut, uint32_t tile, uint32_t S0, int warp, int lane,
                             uint32_t &state_in, uint32_t &base, uint32_t &Tprefix /* | kValidBit */) {
  for (;;) {
    // ---- A: every warp fetches its segment and looks for an inclusive prefix
    const int seg_newest = int(tile) - 1 - kSegTiles * warp;  // newest tile of my segment (may be < 0: empty segment)
    uint32_t status[kLook], T[kLook], pa[kLook], pb[kLook];  // pa: cnt[0] (or the inclusive count), pb: cnt[1]
    uint32_t pending = 0;
#pragma unroll
    for (int k = 0; k < kLook; k++) {
      status[k] = kNone; T[k] = 0; pa[k] = 0; pb[k] = 0;
      if (seg_newest - 32 * k - lane >= 0) pending |= 1u << k;
    }
    uint32_t spins = 0;
    while (pending) {
#pragma unroll
      for (int k = 0; k < kLook; k++) {
        if (pending & (1u << k)) {
          const unsigned long long d = ld_relaxed_u64(p.count_desc + (seg_newest - 32 * k - lane));
          const uint32_t st = uint32_t(d >> 44) & 3u;
          if (uint32_t(d >> 46) == p.epoch && st != kNone) {
            status[k] = st; T[k] = uint32_t(d >> 38) & 63u;
            if (st == kInc) { pa[k] = uint32_t(d); pb[k] = 0; }
            else { pa[k] = uint32_t(d) & 0x7FFFFu; pb[k] = uint32_t(d >> 19) & 0x7FFFFu; }
            pending &= ~(1u << k);
          }
        }
      }
      if (pending) {
        if (++spins > kSpinLimit) {
          atomicOr(p.flags, kFlagInternal);
#pragma unroll
          for (int k = 0; k < kLook; k++)
            if (pending & (1u << k)) { status[k] = kInc; T[k] = 0; pa[k] = 0; pb[k] = 0; }
          pending = 0;
        } else {
          __nanosleep(SJB200_SPIN_NS);
        }
      }
    }
    if (p.debug != nullptr && threadIdx.x == 0) p.debug[uint64_t(tile) * 8 + 6] = globaltimer_ns();  // tile = chain element
    // nearest inclusive prefix in my segment: group kinc, lane linc (groups / lanes are ordered newest first)
    int kinc = kLook, linc = 32;
#pragma unroll
    for (int k = kLook - 1; k >= 0; k--) {
      const uint32_t m = __ballot_sync(kFull, status[k] == kInc);
      if (m) { kinc = k; linc = __ffs(m) - 1; }
    }
    const bool has_inc = kinc < kLook;
    const bool reaches_start = seg_newest >= 0 && seg_newest - (kSegTiles - 1) <= 0;  // tile 0 lies in my segment
    if (lane == 0) ctl->segHas[warp] = (seg_newest >= 0 && (has_inc || reaches_start)) ? 1u : 0u;
    __syncthreads();
    int wstar = -1;
#pragma unroll
    for (int w = kWarps - 1; w >= 0; w--)
      if (ctl->segHas[w]) wstar = w;
    if (wstar < 0) {  // nothing inclusive within 1024 tiles yet: look again (predecessors are still running)
      __syncthreads();
      continue;
    }
    // ---- B: suffix scans over the relevant tiles of every segment up to wstar
    // relevant = newer than the inclusive prefix (all tiles of the segment if it has none)
    uint32_t I[kLook];       // inclusive suffix composition (| kValidBit on relevant lanes)
    int last[kLook];         // oldest relevant lane of the group (-1: none)
    uint32_t G[kLook];       // composition of the whole group (| kValidBit)
    uint32_t F = 0;          // composition of the whole segment
    if (warp <= wstar) {
#pragma unroll
      for (int k = 0; k < kLook; k++) {
        const int gnew = seg_newest - 32 * k;               // newest tile of the group
        int lst = min(31, gnew);                            // tiles below 0 do not exist
        if (has_inc && warp == wstar) {
          if (k > kinc) lst = -1;
          else if (k == kinc) lst = linc - 1;
        }
        last[k] = lst;
        uint32_t val = T[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_down_sync(kFull, val, d);
          if (lane + d <= lst) val = lut_compose(lut, val, o);   // val covers [lane, lane+d), o covers [lane+d, lane+2d)
        }
        I[k] = val | kValidBit;
        G[k] = (lst >= 0) ? (__shfl_sync(kFull, val, 0) | kValidBit) : 0u;
      }
#pragma unroll
      for (int k = kLook - 1; k >= 0; k--) F = compose_opt(lut, G[k], F);   // oldest group first
      if (lane == 0) ctl->segF[warp] = F;
      if (warp == wstar) {
        // the inclusive prefix itself: the lane that holds it publishes it (or lane 0 the virtual one before tile 0)
        if (!has_inc) {
          if (lane == 0) { ctl->segIncT[warp] = 0; ctl->segIncC[warp] = 0; }
        } else {
#pragma unroll
          for (int k = 0; k < kLook; k++)
            if (k == kinc && lane == linc) {
              ctl->segIncT[warp] = T[k] | kValidBit;
              ctl->segIncC[warp] = pa[k];
            }
        }
      }
    }
    __syncthreads();
    // ---- C: states entering every segment (oldest relevant tile first), then every tile's polarity and count
    const uint32_t incT = ctl->segIncT[wstar], incC = ctl->segIncC[wstar];
    uint32_t E = apply_opt(incT, S0);  // state entering the oldest relevant tile of segment wstar
    uint32_t myE = E;
    for (int w = wstar; w >= 0; w--) {
      if (w == warp) myE = E;
      E = apply_opt(ctl->segF[w], E);
    }
    state_in = E;  // after the newest segment
    if (warp <= wstar) {
      uint32_t Eg = myE, cnt = 0;
#pragma unroll
      for (int k = kLook - 1; k >= 0; k--) {
        if (last[k] < 0) continue;
        const uint32_t Hn = __shfl_down_sync(kFull, I[k], 1);                 // composition of the tiles older than mine
        const uint32_t S = (lane + 1 <= last[k]) ? tt_apply(Hn & 63u, Eg) : Eg;  // state entering my tile
        if (lane <= last[k]) cnt += ((S >> 1) & 1u) ? pb[k] : pa[k];
        Eg = tt_apply(G[k] & 63u, Eg);
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) cnt += __shfl_xor_sync(kFull, cnt, d);
      if (lane == 0) ctl->segCount[warp] = cnt;
    }
    __syncthreads();
    uint32_t total = incC;
    uint32_t Tp = incT;
    for (int w = wstar; w >= 0; w--) {
      total += ctl->segCount[w];
      Tp = compose_opt(lut, ctl->segF[w], Tp);
    }
    base = total;
    Tprefix = Tp;
    return;
  }
}

// ------------------------------------------------------------------ tile I/O
// All threads copy one tile global -> shared in the swizzled layout, padding with 0x20 past len.
// Used for the last (partial) tile and for buffers TMA cannot address (unaligned base).
__device__ void cooperative_fill(uint8_t *T, const ScanParams &p, uint32_t tile, int tid) {
  const uint64_t tstart = uint64_t(tile) * kTileBytes;
  const bool aligned = (reinterpret_cast<uintptr_t>(p.buf) & 15u) == 0;
  for (int c = tid; c < kTileBytes / 16; c += kThreads) {
    const uint64_t g = tstart + uint64_t(c) * 16;
    uint4 v;
    if (aligned && g + 16 <= p.len) {
      v = __ldg(reinterpret_cast<const uint4 *>(p.buf + g));
    } else {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        uint32_t x = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const uint64_t q = g + 4 * k + b;
          const uint32_t byte = (q < p.len) ? uint32_t(p.buf[q]) : 0x20u;
          x |= byte << (8 * b);
        }
        w[k] = x;
      }
      v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    *reinterpret_cast<uint4 *>(T + swz(uint32_t(c) * 16)) = v;
  }
}

__device__ __forceinline__ void load_unit(const uint8_t *T, uint32_t off, uint32_t w[8]) {
  const uint4 a = *reinterpret_cast<const uint4 *>(T + swz(off));
  const uint4 b = *reinterpret_cast<const uint4 *>(T + swz(off + 16));
  w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
  w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}

// One thread: fetch the next tile into stage s.  Tiles of a super-tile are fetched in order; when the
// current super-tile is exhausted the next ticket is taken.  For the UTF-8 scan every tile is its own element.
__device__ void refill_stage(uint8_t *tiles, Control *ctl, const CUtensorMap *tmap, const ScanParams &p, int s) {
  if (ctl->next_r >= ctl->next_n) {
    const uint32_t k = atomicAdd(p.ticket, 1u);
    ctl->next_super = k;
    ctl->next_r = 0;
    ctl->next_n = (k < p.nsuper) ? min(p.sub_per_super, p.ntiles - k * p.sub_per_super) : 1u;
  }
  const uint32_t k = ctl->next_super, r = ctl->next_r;
  ctl->next_r = r + 1;
  ctl->stage_super[s] = k;
  ctl->stage_r[s] = r;
  if (k >= p.nsuper) return;
  const uint32_t t = p.tile_begin + k * p.sub_per_super + r;  // document tile
  uint32_t w0 = 0x20202020u, w1 = 0x20202020u, w2 = 0x20202020u, w3 = p.prev_word;
  if (t > 0) {
    const uint8_t *q = p.buf + uint64_t(t) * kTileBytes - 16;
    if ((reinterpret_cast<uintptr_t>(q) & 15u) == 0) {
      const uint4 v = __ldg(reinterpret_cast<const uint4 *>(q));
      w0 = v.x; w1 = v.y; w2 = v.z; w3 = v.w;
    } else {
      uint32_t w[4];
      for (int i = 0; i < 4; i++) w[i] = uint32_t(q[4 * i]) | (uint32_t(q[4 * i + 1]) << 8) | (uint32_t(q[4 * i + 2]) << 16) | (uint32_t(q[4 * i + 3]) << 24);
      w0 = w[0]; w1 = w[1]; w2 = w[2]; w3 = w[3];
    }
  }
  ctl->stage_prev16[s][0] = w0; ctl->stage_prev16[s][1] = w1; ctl->stage_prev16[s][2] = w2; ctl->stage_prev16[s][3] = w3;
  if (p.use_tma && t < p.full_tiles) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&ctl->full_bar[s], kTileBytes);
    tma_load_tile(tiles + s * kTileBytes, tmap, &ctl->full_bar[s], 0, int(t) * kTileRows);
  }
}

// an incoming escape flips the "escaped" status of the first byte that is not a backslash;
// only matters when that byte is a quote (see SURVEY.md 8(a) carry state)
__device__ __forceinline__ void toggle_first_nonbackslash_quote(const uint32_t qu[W], uint32_t qr[W], int k) {
#pragma unroll
  for (int u = 0; u < W; u++)
    if ((k >> 5) == u) qr[u] ^= qu[u] & (1u << (k & 31));
}

// ------------------------------------------------------------------ the kernel
template <int KIND>
__global__ void __launch_bounds__(kThreads, (KIND == kUtf8) ? 4 : kMinCtasPerSm)
    scan_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KiB alignment for the 128B swizzle, computed on the shared-space address so the pointer keeps its address
  // space (a round trip through uintptr_t makes every access a generic LD/ST instead of LDS/STS)
  uint8_t *tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Control *ctl = reinterpret_cast<Control *>(tiles + kStages * kTileBytes);
  uint8_t *lut = reinterpret_cast<uint8_t *>(ctl) + kCtlBytes;                  // composed-transducer table
  uint32_t *emit_scratch = reinterpret_cast<uint32_t *>(lut + kLutBytes);       // [kWarps][256]
  uint4 *mask_slots = reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(emit_scratch) + kEmitBytes);  // [kMaxSub][2][kThreads]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  const int kRefillThread = 32;
  const uint32_t R = p.sub_per_super;

  Carry cin;
  cin.count = 0; cin.state = 0; cin.ttable = 0; cin.flags = 0; cin.reserved = 0;
  if (KIND != kUtf8 && p.carry_in != nullptr) cin = *p.carry_in;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&ctl->full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl->next_super = 0; ctl->next_r = 0; ctl->next_n = 0;
  }
  if (KIND != kUtf8) {
    for (int i = tid; i < kLutBytes; i += kThreads) lut[i] = uint8_t(tt_compose(uint32_t(i) >> 6, uint32_t(i) & 63u));
  }
  __syncthreads();
  if (tid == kRefillThread) {
    for (int s = 0; s < kStages; s++) refill_stage(tiles, ctl, &tmap, p, s);
  }
  __syncthreads();

  // state of the super-tile being scanned (identical in every thread)
  uint32_t bstate = 0;       // e / c bits of the state entering the super-tile (bit0, bit2)
  uint32_t srel = 0;         // scanner state entering the current tile, relative to super-tile polarity 0
  uint32_t Tsuper = 0;       // composed transducer of the tiles scanned so far (| kValidBit)
  bool err0 = false, err1 = false;  // unescaped control character inside a string, per super-tile polarity

  uint32_t phase_bits = 0;
  for (uint32_t it = 0;; it++) {
    const int s = it % kStages;
    const uint32_t super = ctl->stage_super[s];  // chain element (index inside this launch)
    const uint32_t r = ctl->stage_r[s];
    if (super >= p.nsuper) break;
    const uint32_t tile = super * R + r;          // tile index inside this launch
    const uint32_t dtile = p.tile_begin + tile;   // document tile (addresses, positions)
    const uint32_t nsub = min(R, p.ntiles - super * R);
    uint8_t *T = tiles + s * kTileBytes;
    if (p.debug != nullptr && tid == 0 && r == 0) {
      p.debug[uint64_t(super) * 8 + 0] = globaltimer_ns();
      p.debug[uint64_t(super) * 8 + 7] = (uint64_t(blockIdx.x) << 32) | it;
    }
    const bool via_tma = p.use_tma && dtile < p.full_tiles;
    if (via_tma) {
      const uint32_t parity = (phase_bits >> s) & 1u;
      uint32_t spins = 0;
      while (!mbar_try_wait(&ctl->full_bar[s], parity)) {
        if (++spins > kSpinLimit) {
          atomicOr(p.flags, kFlagInternal);
          break;
        }
      }
      phase_bits ^= 1u << s;
    } else {
      cooperative_fill(T, p, dtile, tid);
      __syncthreads();
    }
    const uint32_t lane_off = uint32_t(warp) * kWarpBytes + uint32_t(lane) * kLaneBytes;
    const bool last_tile = (tile == p.ntiles - 1);

    if (KIND == kUtf8) {
      // ============================ UTF-8 only: no carries beyond the 3-byte halo ============================
      uint32_t uerr = 0;
      const uint32_t pw = (lane_off == 0) ? ctl->stage_prev16[s][3] : *reinterpret_cast<const uint32_t *>(T + swz(lane_off - 4));
      utf8_carry uc = utf8_carry_from_prev_word(pw);
#pragma unroll
      for (int u = 0; u < W; u++) {
        uint32_t w8[8], pl[8];
        load_unit(T, lane_off + 32 * u, w8);
        const uint32_t hi = (w8[0] | w8[1] | w8[2] | w8[3] | w8[4] | w8[5] | w8[6] | w8[7]) & 0x80808080u;
        if (__any_sync(kFull, hi != 0 || utf8_carry_pending(uc))) {
          transpose32(w8, pl);
          uerr |= utf8_check_unit(pl, uc);
        } else {
          uc = utf8_carry_zero();
        }
      }
      if (__any_sync(kFull, uerr != 0) && lane == 0) atomicOr(p.flags, kFlagUtf8);
      if (last_tile && p.check_eof && tid == 0) {
        // the input must not end inside a multi-byte sequence (utf8_checker::check_eof, L167-171)
        uint32_t tw = 0;
        for (int d = 1; d <= 4; d++) {
          const uint32_t b = (p.len >= uint64_t(d)) ? uint32_t(p.buf[p.len - d]) : ((p.prev_word >> (8 * (4 - d + int(p.len)))) & 0xFFu);
          tw |= b << (8 * (4 - d));
        }
        if (utf8_carry_pending(utf8_carry_from_prev_word(tw))) atomicOr(p.flags, kFlagUtf8);
      }
      __syncthreads();  // every warp is done with stage s
      if (tid == kRefillThread) refill_stage(tiles, ctl, &tmap, p, s);
      __syncthreads();
      continue;
    }

    if (r == 0) {
      // ---- the e / c bits of the state entering this super-tile, from the bytes before it.  Exact unless a run of
      //      >= 16 backslashes ends right at the boundary; then (rare) wait for the predecessor's inclusive prefix.
      if (super == 0) {
        bstate = cin.state & 5u;  // the launch's carry-in is exact
      } else {
        const uint32_t b = boundary_state_from_prev16(ctl->stage_prev16[s]);
        bstate = b & 5u;
        if (b & 8u) {
          if (tid == 0) {
            uint32_t spins = 0, st_exact = 0;
            for (;;) {
              const unsigned long long d = ld_relaxed_u64(p.count_desc + (super - 1));
              if (uint32_t(d >> 46) == p.epoch && (uint32_t(d >> 44) & 3u) == kInc) {
                st_exact = tt_apply(uint32_t(d >> 38) & 63u, cin.state);
                break;
              }
              if (++spins > kSpinLimit) { atomicOr(p.flags, kFlagInternal); break; }
              __nanosleep(200);
            }
            ctl->bstate_exact = st_exact & 5u;
          }
          __syncthreads();
          bstate = ctl->bstate_exact;
        }
      }
      srel = bstate;
      Tsuper = 0;
      err0 = false; err1 = false;
    }

    // ============================ phase 1: planes, classes, UTF-8 ============================
    uint32_t bs[W], qu[W], op[W], sc[W], cl[W];
    uint32_t uerr = 0;
    {
      const uint32_t pw = (lane_off == 0) ? ctl->stage_prev16[s][3] : *reinterpret_cast<const uint32_t *>(T + swz(lane_off - 4));
      utf8_carry uc = utf8_carry_from_prev_word(pw);
#pragma unroll
      for (int u = 0; u < W; u++) {
        uint32_t w8[8], pl[8];
        load_unit(T, lane_off + 32 * u, w8);
        transpose32(w8, pl);
        const unit_classes c = classify(pl);
        bs[u] = c.bs; qu[u] = c.qu; op[u] = c.op; sc[u] = c.sc; cl[u] = c.ctl;
        if (__any_sync(kFull, pl[7] != 0 || utf8_carry_pending(uc))) {
          uerr |= utf8_check_unit(pl, uc);
        } else {
          uc = utf8_carry_zero();
        }
      }
    }
    if (KIND != kMinify) {  // minify does not validate (json_minifier.h: "does not parse or validate")
      if (__any_sync(kFull, uerr != 0) && lane == 0) atomicOr(p.flags, kFlagUtf8);
    }

    // ============================ phase 2: escapes, quotes, warp transducer ============================
    uint32_t qr[W];
    uint32_t Pmask = 0, warp_cout0 = 0;
    int nlead = 0;
    {
      const uint32_t bsany = bs[0] | bs[1] | bs[2] | bs[3];
      if (__any_sync(kFull, bsany != 0)) {
        uint32_t escaped[W];
        const uint32_t esc_out0 = escape_scan<W>(bs, escaped);
#pragma unroll
        for (int u = 0; u < W; u++) qr[u] = qu[u] & ~escaped[u];
        nlead = leading_backslashes<W>(bs);
        const uint32_t G = __ballot_sync(kFull, esc_out0 != 0);
        Pmask = __ballot_sync(kFull, nlead == 32 * W);
        const uint32_t carries = escape_carries(G, Pmask, 0u, &warp_cout0);
        if (((carries >> lane) & 1u) && nlead != 32 * W) toggle_first_nonbackslash_quote(qu, qr, nlead);
      } else {
#pragma unroll
        for (int u = 0; u < W; u++) qr[u] = qu[u];
      }
    }
    const bool warp_allbs = (Pmask == kFull);
    const int mlane = warp_allbs ? 0 : (__ffs(~Pmask) - 1);  // lane holding the first non-backslash byte
    {
      const uint32_t lp = (__popc(qr[0]) + __popc(qr[1]) + __popc(qr[2]) + __popc(qr[3])) & 1u;
      const uint32_t par0 = __popc(__ballot_sync(kFull, lp != 0)) & 1u;
      uint32_t myq = 0;
#pragma unroll
      for (int u = 0; u < W; u++)
        if ((nlead >> 5) == u) myq = (qu[u] >> (nlead & 31)) & 1u;
      uint32_t qx = __shfl_sync(kFull, myq, mlane);
      uint32_t x_is_last = __shfl_sync(kFull, uint32_t(nlead == 32 * W - 1), 31);
      if (warp_allbs) qx = 0;
      if (mlane != 31) x_is_last = 0;
      const uint32_t scal0 = __shfl_sync(kFull, (sc[W - 1] & ~qr[W - 1]) >> 31, 31);
      const uint32_t Tw = tt_make(warp_cout0, par0, scal0, warp_allbs ? 1u : warp_cout0, par0 ^ qx, scal0 ^ (qx & x_is_last));
      if (lane == 0) ctl->warpT[warp] = Tw;
    }
    __syncthreads();  // S1: all warp transducers visible; every lane holds its input in registers
    // stage s is free: fetch the next tile of this super-tile (after the last one the buffer serves as emit staging)
    if (KIND == kIndex && r + 1 < nsub && tid == kRefillThread) refill_stage(tiles, ctl, &tmap, p, s);

    // state entering this warp (relative to super-tile polarity 0): compose the warps before it
    uint32_t win = srel;
    for (int w = 0; w < warp; w++) win = tt_apply(ctl->warpT[w], win);
    const uint32_t e_w = win & 1u, s_w = (win >> 1) & 1u, c_w = (win >> 2) & 1u;
    uint32_t Ttile = ctl->warpT[0];
#pragma unroll
    for (int w = 1; w < kWarps; w++) Ttile = lut_compose(lut, ctl->warpT[w], Ttile);
    srel = tt_apply(Ttile, srel);
    Tsuper = (Tsuper & kValidBit) ? (lut_compose(lut, Ttile, Tsuper) | kValidBit) : (Ttile | kValidBit);

    // ============================ phase 3: final masks, for both polarities ============================
    if (e_w && !warp_allbs && lane == mlane) toggle_first_nonbackslash_quote(qu, qr, nlead);
    {
      uint32_t m0[W], m1[W];  // kIndex: pseudo-structurals, string tail (polarity 0).  kMinify: whitespace, in_string (polarity 0)
      uint32_t c0 = 0, c1 = 0;
      const uint32_t lp = (__popc(qr[0]) + __popc(qr[1]) + __popc(qr[2]) + __popc(qr[3])) & 1u;
      const uint32_t pb = __ballot_sync(kFull, lp != 0);
      uint32_t instr = (s_w ^ uint32_t(__popc(pb & lanemask_lt))) & 1u;
      uint32_t scal_prev = __shfl_up_sync(kFull, (sc[W - 1] & ~qr[W - 1]) >> 31, 1);
      if (lane == 0) scal_prev = c_w;
      uint32_t prev_nq = scal_prev << 31;
      uint32_t hit0 = 0, hit1 = 0;
#pragma unroll
      for (int u = 0; u < W; u++) {
        const uint32_t in_string = prefix_xor32(qr[u]) ^ (0u - instr);
        instr = in_string >> 31;
        if (KIND == kIndex) {
          const uint32_t nq = sc[u] & ~qr[u];
          const uint32_t follows = shl_in(prev_nq, nq, 1);
          prev_nq = nq;
          const uint32_t pm = op[u] | (sc[u] & ~follows);
          const uint32_t x0 = in_string ^ qr[u];
          m0[u] = pm; m1[u] = x0;
          c0 += __popc(pm & ~x0);
          c1 += __popc(pm & x0);
          hit0 |= cl[u] & in_string;
          hit1 |= cl[u] & ~in_string;
        } else {
          const uint32_t ws = ~(op[u] | sc[u]);
          uint32_t valid = kFull;
          if (last_tile) {  // the 0x20 padding past len is never output (json_minifier.h L79-95)
            const uint64_t ubase = uint64_t(dtile) * kTileBytes + lane_off + 32 * u;
            if (ubase + 32 > p.len) valid = (ubase >= p.len) ? 0u : ((1u << uint32_t(p.len - ubase)) - 1u);
          }
          m0[u] = ws; m1[u] = in_string;
          c0 += __popc(valid & ~(ws & ~in_string));
          c1 += __popc(valid & ~(ws & in_string));
        }
      }
      const bool h0 = __any_sync(kFull, hit0 != 0), h1 = __any_sync(kFull, hit1 != 0);
      err0 = err0 || h0;
      err1 = err1 || h1;
      // park the masks until the incoming polarity is known
      mask_slots[(r * 2 + 0) * kThreads + tid] = make_uint4(m0[0], m0[1], m0[2], m0[3]);
      mask_slots[(r * 2 + 1) * kThreads + tid] = make_uint4(m1[0], m1[1], m1[2], m1[3]);
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) {
        c0 += __shfl_xor_sync(kFull, c0, d);
        c1 += __shfl_xor_sync(kFull, c1, d);
      }
      if (lane == 0) { ctl->cntw[r][0][warp] = c0; ctl->cntw[r][1][warp] = c1; }
    }
    __syncthreads();  // S2: per-warp counts visible; warpT may be reused
    if (r + 1 < nsub) continue;

    // =============================================== end of the super-tile ===============================================
    if (p.debug != nullptr && tid == 0) p.debug[uint64_t(super) * 8 + 3] = globaltimer_ns();
    uint32_t tc0 = 0, tc1 = 0;
    for (uint32_t q = 0; q < nsub; q++) {
#pragma unroll
      for (int w = 0; w < kWarps; w++) { tc0 += ctl->cntw[q][0][w]; tc1 += ctl->cntw[q][1][w]; }
    }
    uint32_t state_in = cin.state, base = 0, Tprefix = 0;
    if (super > 0) {
      if (tid == 0) st_relaxed_u64(p.count_desc + super, pack_agg(p.epoch, Tsuper & 63u, tc0, tc1));
      resolve_tile(p, ctl, lut, super, cin.state, warp, lane, state_in, base, Tprefix);
      if ((state_in & 5u) != bstate) atomicOr(p.flags, kFlagInternal);  // cannot happen: bstate was exact
    }
    if (p.debug != nullptr && tid == 0) p.debug[uint64_t(super) * 8 + 4] = globaltimer_ns();
    const uint32_t pol = (state_in >> 1) & 1u;  // in_string entering the super-tile
    const uint32_t super_total = pol ? tc1 : tc0;
    const uint32_t Tincl = (Tprefix & kValidBit) ? lut_compose(lut, Tsuper, Tprefix) : (Tsuper & 63u);
    if (tid == 0) {
      st_relaxed_u64(p.count_desc + super, pack_inc(p.epoch, Tincl, base + super_total));
      if (super == p.nsuper - 1) {
        // (carry_out->flags is stored by the last CTA to leave the kernel)
        p.carry_out->count = cin.count + base + super_total;
        p.carry_out->state = tt_apply(Tincl, cin.state);
        p.carry_out->ttable = Tincl;
        if (KIND == kIndex && p.write_sentinels) {  // json_structural_indexer.h L284-286
          uint32_t *tail = p.idx_out + (cin.count + base + super_total);
          tail[0] = uint32_t(p.len);
          tail[1] = uint32_t(p.len);
          tail[2] = 0;
        }
        if (KIND == kIndex && p.check_eof) {
          uint32_t tw = 0;
          for (int d = 1; d <= 4; d++) {
            const uint32_t b = (p.len >= uint64_t(d)) ? uint32_t(p.buf[p.len - d]) : ((p.prev_word >> (8 * (4 - d + int(p.len)))) & 0xFFu);
            tw |= b << (8 * (4 - d));
          }
          if (utf8_carry_pending(utf8_carry_from_prev_word(tw))) atomicOr(p.flags, kFlagUtf8);
        }
      }
    }
    if (KIND == kIndex && lane == 0 && (pol ? err1 : err0)) atomicOr(p.flags, kFlagCtl);

    // ============================ emit every tile of the super-tile ============================
    uint32_t run_base = base;  // outputs before the tile being emitted
    for (uint32_t q = 0; q < nsub; q++) {
      uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++) {
        const uint32_t c = ctl->cntw[q][pol][w];
        if (w < warp) warp_base += c;
        tile_total += c;
      }
      const uint4 a0 = mask_slots[(q * 2 + 0) * kThreads + tid];
      const uint4 a1 = mask_slots[(q * 2 + 1) * kThreads + tid];
      if (KIND == kIndex) {
        const uint32_t flip = pol ? 0u : kFull;
        const uint32_t e0 = a0.x & (a1.x ^ flip), e1 = a0.y & (a1.y ^ flip), e2 = a0.z & (a1.z ^ flip), e3 = a0.w & (a1.w ^ flip);
        const uint32_t cnt = __popc(e0) + __popc(e1) + __popc(e2) + __popc(e3);
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += t;
        }
        // Balanced emit.  Structurals cluster (a lane inside a numeric array holds 10x more than a lane inside a long
        // string), so a loop over the lane's own bits leaves most lanes idle.  Instead the warp's 128 mask words and
        // their exclusive bit counts go to shared memory and lane L takes the contiguous run of words that holds
        // outputs [L*k, (L+1)*k), k = ceil(total/32), found by binary search.  Word w of the warp covers bytes 32w..32w+31.
        uint32_t *wmask = emit_scratch + warp * 256;
        uint32_t *wpre = wmask + 128;
        const uint32_t r0 = incl - cnt, r1 = r0 + __popc(e0), r2 = r1 + __popc(e1), r3 = r2 + __popc(e2);
        reinterpret_cast<uint4 *>(wmask)[lane] = make_uint4(e0, e1, e2, e3);
        reinterpret_cast<uint4 *>(wpre)[lane] = make_uint4(r0, r1, r2, r3);
        const uint32_t wtotal = __shfl_sync(kFull, incl, 31);
        __syncwarp();
        if (wtotal != 0) {
          const uint32_t k = (wtotal + 31u) >> 5;
          const uint32_t target = uint32_t(lane) * k;  // first word whose exclusive count is >= lane*k (128 if none)
          uint32_t lo = 0;
#pragma unroll
          for (int step = 64; step >= 1; step >>= 1)
            if (wpre[lo + step - 1] < target) lo += step;  // invariant: every word before lo has count < target
          if (wpre[lo] < target && lo == 127) lo = 128;
          uint32_t wend = __shfl_down_sync(kFull, lo, 1);
          if (lane == 31) wend = 128;
          // every lane walks its run of words with the SAME trip count (the warp maximum of words + bits) and a
          // branch-light body, so the warp stays converged: one iteration emits one index or steps to the next word.
          // Indexes go to a shared-memory staging area (this warp's 4 KiB of the idle tile buffer) and leave the SM with
          // coalesced stores: scattered 4-byte global stores cost one L1 wavefront each and were the emit bottleneck.
          const uint32_t nwords = wend > lo ? wend - lo : 0u;
          const uint32_t lim = (wend < 128u) ? wpre[wend] : wtotal;
          const uint32_t first = nwords ? wpre[lo] : 0u;
          const uint32_t nbits = nwords ? lim - first : 0u;
          uint32_t *out = p.idx_out + (cin.count + run_base + warp_base);
          uint32_t w = lo;
          uint32_t wbase = p.pos_base + (p.tile_begin + super * R + q) * uint32_t(kTileBytes) + uint32_t(warp) * kWarpBytes + 32 * lo;
          uint32_t m = nwords ? wmask[lo] : 0u;
          uint32_t left = nwords;  // words of the run not yet finished (including the current one)
          if (wtotal <= uint32_t(kWarpBytes / 4)) {
            // Indexes go to a shared-memory staging area (this warp's 4 KiB of the idle tile buffer) and leave the SM
            // with coalesced stores (scattered 4-byte global stores cost one L1 wavefront each).  The loop body is
            // branch-free and retires up to two indexes per iteration; every lane runs the same trip count, the warp
            // maximum of sum over its words of max(1, ceil(bits/2)) <= words + bits/2.
            uint32_t *stg = reinterpret_cast<uint32_t *>(T + uint32_t(warp) * kWarpBytes);
            uint32_t off = first;
            const uint32_t steps = __reduce_max_sync(kFull, nwords + ((nbits + 1u) >> 1));
            for (uint32_t i = 0; i < steps; i++) {
              const uint32_t b1 = __ffs(m) - 1;
              const bool h1 = m != 0;
              const uint32_t m1 = m & (m - 1);
              const uint32_t b2 = __ffs(m1) - 1;
              const bool h2 = m1 != 0;
              const uint32_t m2 = m1 & (m1 - 1);
              if (h1) stg[off] = wbase + b1;
              if (h2) stg[off + 1] = wbase + b2;
              off += uint32_t(h1) + uint32_t(h2);
              const bool adv = (m2 == 0) && (left > 1);
              left -= uint32_t(adv);
              w += uint32_t(adv);
              wbase += adv ? 32u : 0u;
              const uint32_t nm = wmask[w & 127u];
              m = (m2 != 0) ? m2 : (adv ? nm : 0u);
            }
            __syncwarp();
#pragma unroll 4
            for (uint32_t i = lane; i < wtotal; i += 32) out[i] = stg[i];
          } else {
            // very dense chunk (> 1 structural per 4 bytes over 4 KiB): straight to global memory
            uint32_t *dst = out + first;
            const uint32_t steps = __reduce_max_sync(kFull, nwords + nbits);
            for (uint32_t i = 0; i < steps; i++) {
              if (m != 0) {
                *dst++ = wbase + (__ffs(m) - 1);
                m &= m - 1;
              } else if (left > 1) {
                --left;
                ++w;
                wbase += 32;
                m = wmask[w];
              }
            }
          }
        }
        __syncwarp();  // the scratch is rewritten for the next tile
      } else {
        // kMinify (one tile per super-tile): the tile's bytes are still in stage s
        const uint32_t flip = pol ? kFull : 0u;
        const uint32_t k0 = ~(a0.x & ~(a1.x ^ flip)), k1 = ~(a0.y & ~(a1.y ^ flip)), k2 = ~(a0.z & ~(a1.z ^ flip)), k3 = ~(a0.w & ~(a1.w ^ flip));
        uint32_t keepm[W] = {k0, k1, k2, k3};
        uint32_t cnt = 0;
#pragma unroll
        for (int u = 0; u < W; u++) {
          if (last_tile) {
            const uint64_t ubase = uint64_t(dtile) * kTileBytes + lane_off + 32 * u;
            if (ubase + 32 > p.len) keepm[u] &= (ubase >= p.len) ? 0u : ((1u << uint32_t(p.len - ubase)) - 1u);
          }
          cnt += __popc(keepm[u]);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, d);
          if (lane >= d) incl += t;
        }
        // Kept bytes are compacted into shared memory first (byte stores there are cheap) and leave the SM as aligned
        // 16-byte vectors: scattered byte stores to global memory cost one L1 wavefront each and were ~90 % of the
        // kernel time.  The staging area (the mask slots 1..7, unused when a super-tile is one tile) starts at the
        // same offset modulo 16 as the destination, so interior 16-byte groups line up.
        const uint32_t wtotal = __shfl_sync(kFull, incl, 31);
        uint8_t *gdst = p.dst + (cin.count + run_base + warp_base);
        const uint32_t a = uint32_t(reinterpret_cast<uintptr_t>(gdst) & 15u);
        uint8_t *stg = reinterpret_cast<uint8_t *>(mask_slots) + kMaskSlotBytes + uint32_t(warp) * 6144u;
        uint8_t *sp = stg + a + (incl - cnt);
#pragma unroll
        for (int u = 0; u < W; u++) {
          uint32_t w8[8];
          load_unit(T, lane_off + 32 * u, w8);
          const uint32_t keep = keepm[u];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const uint32_t nib = (keep >> (4 * i)) & 15u;
            const uint32_t word = w8[i];
            if (nib == 15u) {
              if ((smem_u32(sp) & 3u) == 0) {
                *reinterpret_cast<uint32_t *>(sp) = word;
              } else {
                sp[0] = uint8_t(word); sp[1] = uint8_t(word >> 8); sp[2] = uint8_t(word >> 16); sp[3] = uint8_t(word >> 24);
              }
              sp += 4;
            } else {
#pragma unroll
              for (int b = 0; b < 4; b++)
                if ((nib >> b) & 1u) *sp++ = uint8_t(word >> (8 * b));
            }
          }
        }
        __syncwarp();
        {
          const uint32_t head = min(wtotal, (16u - a) & 15u);           // bytes before the first aligned 16-byte group
          const uint32_t nvec = (wtotal - head) >> 4;
          const uint32_t tail = wtotal - head - (nvec << 4);
          if (uint32_t(lane) < head) gdst[lane] = stg[a + lane];
          const uint4 *sv = reinterpret_cast<const uint4 *>(stg + a + head);  // (a + head) % 16 == 0
          uint4 *gv = reinterpret_cast<uint4 *>(gdst + head);
          for (uint32_t i = lane; i < nvec; i += 32) gv[i] = sv[i];
          if (uint32_t(lane) < tail) gdst[head + (nvec << 4) + lane] = stg[a + head + (nvec << 4) + lane];
        }
        __syncwarp();
      }
      run_base += tile_total;
    }
    if (p.debug != nullptr && tid == 0) p.debug[uint64_t(super) * 8 + 5] = globaltimer_ns();
    __syncthreads();  // every warp is done with stage s (input bytes / emit staging), the mask slots and the counts
    if (tid == kRefillThread) refill_stage(tiles, ctl, &tmap, p, s);
    __syncthreads();
  }

  // last CTA out resets the ticket for the next launch on this context
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(p.ticket + 1, 1u);
    if (done == gridDim.x - 1) {
      p.ticket[0] = 0;
      p.ticket[1] = 0;
      p.carry_out->flags = atomicExch(p.flags, 0u);  // every CTA is done raising flags; hand them over and re-arm
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------ scan4 (stage 1; see sjb200_scan4.cuh)
// two variants of one source: pipelined (masks wait in shared memory, an element is emitted two scans after it was scanned)
// and deferred (masks wait in an L2-resident scratch ring, everything is emitted after the CTA's last scan: launches
// small enough that every CTA holds all its elements at once never stall on the chain)
__global__ void __launch_bounds__(scan4::kThreads4, (SJB200_SCAN4_WARPS > 8) ? 1 : SJB200_SCAN4_MIN_CTAS)
    scan4_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw4[];
  scan4::scan4_body<0>(&tmap, p, smem_raw4, sj_smem_u32(smem_raw4));
}
__global__ void __launch_bounds__(scan4::kThreads4, (SJB200_SCAN4_WARPS > 8) ? 1 : SJB200_SCAN4_MIN_CTAS)
    scan4_deferred_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw4[];
  scan4::scan4_body<1>(&tmap, p, smem_raw4, sj_smem_u32(smem_raw4));
}
// minify on the scan4 structure (option minify_kernel=4; the block's bytes are fetched a second time, from L2, when it is emitted)
__global__ void __launch_bounds__(scan4::kThreads4, (SJB200_SCAN4_WARPS > 8) ? 1 : SJB200_SCAN4_MIN_CTAS)
    scan4_minify_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw4[];
  scan4::scan4_body<2>(&tmap, p, smem_raw4, sj_smem_u32(smem_raw4));
}

// validate_utf8, every warp on its own (sjb200_utf8.cuh)
__global__ void __launch_bounds__(utf8v2::kThreadsU, utf8v2::kCtasPerSmU)
    utf8v2_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw_u[];
  utf8v2::utf8_body(&tmap, p, smem_raw_u, sj_smem_u32(smem_raw_u));
}

// ------------------------------------------------------------------ small helpers
// last min(3, len) bytes of many device-resident documents into one small array (streaming modes trim a partial UTF-8
// tail before the scan: json_structural_indexer.h L198-204) -- one launch + one copy for a whole batch
__global__ void gather_tails_kernel(const uint8_t *const *bufs, const uint64_t *lens, uint32_t ndocs, uint8_t *out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ndocs) return;
  const uint64_t len = lens[i];
  const uint32_t k = len < 3 ? uint32_t(len) : 3u;
  for (uint32_t b = 0; b < 4; b++) out[4 * i + b] = (b < k) ? bufs[i][len - k + b] : 0;
}
__global__ void write_sentinels_kernel(uint32_t *idx, uint32_t n, uint32_t a, uint32_t b, uint32_t c) {
  idx[n] = a;
  idx[n + 1] = b;
  idx[n + 2] = c;
}

// second round of a sharded pass (after re-scans): republish this rank's record in every rank's exchange window
__global__ void xchg_post_kernel(ScanParams p, unsigned long long w0, unsigned long long w1) {
  const uint32_t r = threadIdx.x;
  if (r < p.xchg_nranks) {
    unsigned long long *rec = p.xchg_peer[r] + (size_t(p.xchg_slot) * kMaxRanks + p.xchg_rank) * 2;
    sj_st_sys_u64(rec, w0);
    sj_st_sys_u64(rec + 1, w1);
  }
}

// ------------------------------------------------------------------ launchers
template <int KIND>
static cudaError_t launch_kind(const CUtensorMap *tmap, const ScanParams &p, int grid, cudaStream_t stream) {
  static bool configured[64] = {};  // per device: the attribute lives in the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(scan_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(KIND));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  scan_kernel<KIND><<<grid, kThreads, smem_bytes_for(KIND), stream>>>(*tmap, p);
  return cudaGetLastError();
}

cudaError_t launch_scan4(const CUtensorMap *tmap, const ScanParams &p, int grid, int mode, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(scan4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scan4::kSmemBytes4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(scan4_deferred_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scan4::kSmemBytes4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(scan4_minify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scan4::kSmemBytes4);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (mode == 2) scan4_minify_kernel<<<grid, scan4::kThreads4, scan4::kSmemBytes4, stream>>>(*tmap, p);
  else if (mode == 1) scan4_deferred_kernel<<<grid, scan4::kThreads4, scan4::kSmemBytes4, stream>>>(*tmap, p);
  else scan4_kernel<<<grid, scan4::kThreads4, scan4::kSmemBytes4, stream>>>(*tmap, p);
  return cudaGetLastError();
}

cudaError_t launch_utf8v2(const CUtensorMap *tmap, const ScanParams &p, int grid, cudaStream_t stream) {
  static bool configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(utf8v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, utf8v2::kSmemBytesU);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  utf8v2_kernel<<<grid, utf8v2::kThreadsU, utf8v2::kSmemBytesU, stream>>>(*tmap, p);
  return cudaGetLastError();
}
int utf8v2_max_ctas_per_sm() {
  int n = 0;
  cudaFuncSetAttribute(utf8v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, utf8v2::kSmemBytesU);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, utf8v2_kernel, utf8v2::kThreadsU, utf8v2::kSmemBytesU);
  return (e == cudaSuccess && n > 0) ? n : 1;
}
int utf8v2_warps_per_cta() { return utf8v2::kWarpsU; }

size_t scan4_park_words(int grid) { return size_t(grid) * scan4::kParkRing * scan4::kParkSlotWords + 8; }
int scan4_parks_in_global() { return scan4::kGPark > 0 ? 1 : 0; }
int scan4_deferred_capacity() { return scan4::kParkRing; }
int scan4_tiles_per_element() { return scan4::kElemBytes / kTileBytes; }

int scan4_max_ctas_per_sm() {
  int n = 0;
  cudaFuncSetAttribute(scan4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, scan4::kSmemBytes4);
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan4_kernel, scan4::kThreads4, scan4::kSmemBytes4);
  return (e == cudaSuccess && n > 0) ? n : 1;
}

cudaError_t launch_scan(int kind, const CUtensorMap *tmap, const ScanParams &p, int grid, cudaStream_t stream) {
  switch (kind) {
    case kIndex: return launch_kind<kIndex>(tmap, p, grid, stream);
    case kMinify: return launch_kind<kMinify>(tmap, p, grid, stream);
    default: return launch_kind<kUtf8>(tmap, p, grid, stream);
  }
}

int scan_max_ctas_per_sm(int kind) {
  int n = 0;
  cudaError_t e;
  switch (kind) {
    case kIndex:
      cudaFuncSetAttribute(scan_kernel<kIndex>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(kIndex));
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kIndex>, kThreads, smem_bytes_for(kIndex));
      break;
    case kMinify:
      cudaFuncSetAttribute(scan_kernel<kMinify>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(kMinify));
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kMinify>, kThreads, smem_bytes_for(kMinify));
      break;
    default:
      cudaFuncSetAttribute(scan_kernel<kUtf8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes_for(kUtf8));
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kUtf8>, kThreads, smem_bytes_for(kUtf8));
      break;
  }
  return (e == cudaSuccess && n > 0) ? n : 1;
}

cudaError_t launch_gather_tails(const uint8_t *const *bufs, const uint64_t *lens, uint32_t ndocs, uint8_t *out, cudaStream_t stream) {
  if (ndocs == 0) return cudaSuccess;
  gather_tails_kernel<<<(ndocs + 127) / 128, 128, 0, stream>>>(bufs, lens, ndocs, out);
  return cudaGetLastError();
}

cudaError_t launch_xchg_post(const ScanParams &p, unsigned long long w0, unsigned long long w1, cudaStream_t stream) {
  xchg_post_kernel<<<1, 32, 0, stream>>>(p, w0, w1);
  return cudaGetLastError();
}

cudaError_t launch_write_sentinels(uint32_t *idx, uint32_t n, uint32_t a, uint32_t b, uint32_t c, cudaStream_t stream) {
  write_sentinels_kernel<<<1, 1, 0, stream>>>(idx, n, a, b, c);
  return cudaGetLastError();
}

}  // namespace sjb200
