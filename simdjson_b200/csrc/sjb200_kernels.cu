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

namespace sjb200 {

constexpr int W = kUnitsPerLane;
constexpr uint32_t kFull = 0xFFFFFFFFu;
constexpr uint32_t kSpinLimit = 1u << 25;  // bounded spins: a stuck chain becomes kFlagInternal, never a hang

// look-back descriptor status
enum : uint32_t { kNone = 0, kAgg = 1, kInc = 2 };

struct Control {
  unsigned long long full_bar[kStages];
  uint32_t stage_tile[kStages];
  uint32_t stage_prev16[kStages][4];  // the 16 bytes before the tile (UTF-8 halo + boundary state)
  uint32_t warpT[kWarps];
  uint32_t warpCnt[2][kWarps];        // per in-string polarity of the tile
  // look-back scratch, one slot per warp = per 128-tile segment of the window
  uint32_t segHas[kWarps];            // the segment contains an inclusive prefix (or reaches tile 0)
  uint32_t segF[kWarps];              // composed transducer of the segment's relevant tiles | 0x100 if any
  uint32_t segIncT[kWarps];           // inclusive transducer prefix found in the segment | 0x100 if real
  uint32_t segIncC[kWarps];
  uint32_t segCount[kWarps];
};
static_assert(sizeof(Control) <= 512, "control block");

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
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

// byte offset inside a tile -> offset in the 128B-swizzled shared-memory image
// (TMA SWIZZLE_128B: 16-byte chunk index bits [4,7) ^= row bits [7,10))
__device__ __forceinline__ uint32_t swz(uint32_t off) { return off ^ ((off >> 3) & 0x70u); }

// ------------------------------------------------------------- the look-back chain
// One 64-bit descriptor per tile:  [63:40] epoch  [39:38] status  [37:32] transducer  [31:0] payload
//   kAgg: transducer = T of the tile; payload = cnt[1]<<16 | cnt[0]: outputs of the tile for either
//         in-string polarity at its start (the e / c bits of its incoming state are already baked in,
//         they were read off the bytes before the tile)
//   kInc: transducer = composition of tiles [0, tile]; payload = outputs of tiles [0, tile]
__device__ __forceinline__ unsigned long long pack_desc(uint32_t epoch, uint32_t status, uint32_t T, uint32_t payload) {
  return ((unsigned long long)epoch << 40) | ((unsigned long long)status << 38) | ((unsigned long long)(T & 63u) << 32) | payload;
}

constexpr int kLook = 4;                      // descriptors per lane: a warp covers 128 tiles of the window
constexpr int kSegTiles = 32 * kLook;
constexpr uint32_t kValidBit = 0x100u;        // marks "this composed transducer exists"

__device__ __forceinline__ uint32_t compose_opt(uint32_t newer, uint32_t older) {  // either may be absent (no kValidBit)
  if (!(newer & kValidBit)) return older;
  if (!(older & kValidBit)) return newer;
  return tt_compose(newer & 63u, older & 63u) | kValidBit;
}
__device__ __forceinline__ uint32_t apply_opt(uint32_t F, uint32_t state) { return (F & kValidBit) ? tt_apply(F & 63u, state) : state; }

// Resolve tile `tile` (>= 1) of this launch: the scanner state entering it, the number of outputs before
// it, and the composed transducer of tiles [0, tile).  Called by ALL warps of the CTA (it synchronises):
// warp w looks at tiles tile-1-128w .. tile-128(w+1), so one round trip covers a window of 1024 tiles.
// The walk stops at the nearest tile that already published an inclusive prefix; everything newer only
// has an aggregate, and because an aggregate's count depends on the in-string polarity at its tile, the
// states are folded forwards (oldest to newest) with parallel suffix scans.
__device__ void resolve_tile(const ScanParams &p, Control *ctl, uint32_t tile, uint32_t S0, int warp, int lane, uint32_t &state_in,
                             uint32_t &base, uint32_t &Tprefix /* | kValidBit */) {
  for (;;) {
    // ---- A: every warp fetches its segment and looks for an inclusive prefix
    const int seg_newest = int(tile) - 1 - kSegTiles * warp;  // newest tile of my segment (may be < 0: empty segment)
    uint32_t status[kLook], T[kLook], pay[kLook];
    uint32_t pending = 0;
#pragma unroll
    for (int k = 0; k < kLook; k++) {
      status[k] = kNone; T[k] = 0; pay[k] = 0;
      if (seg_newest - 32 * k - lane >= 0) pending |= 1u << k;
    }
    uint32_t spins = 0;
    while (pending) {
#pragma unroll
      for (int k = 0; k < kLook; k++) {
        if (pending & (1u << k)) {
          const unsigned long long d = ld_relaxed_u64(p.count_desc + (seg_newest - 32 * k - lane));
          const uint32_t st = uint32_t(d >> 38) & 3u;
          if (uint32_t(d >> 40) == p.epoch && st != kNone) {
            status[k] = st; T[k] = uint32_t(d >> 32) & 63u; pay[k] = uint32_t(d);
            pending &= ~(1u << k);
          }
        }
      }
      if (pending) {
        if (++spins > kSpinLimit) {
          atomicOr(p.flags, kFlagInternal);
#pragma unroll
          for (int k = 0; k < kLook; k++)
            if (pending & (1u << k)) { status[k] = kInc; T[k] = 0; pay[k] = 0; }
          pending = 0;
        } else {
          __nanosleep(20);
        }
      }
    }
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
          if (lane + d <= lst) val = tt_compose(val, o);   // val covers [lane, lane+d), o covers [lane+d, lane+2d)
        }
        I[k] = val | kValidBit;
        G[k] = (lst >= 0) ? (__shfl_sync(kFull, val, 0) | kValidBit) : 0u;
      }
#pragma unroll
      for (int k = kLook - 1; k >= 0; k--) F = compose_opt(G[k], F);   // oldest group first
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
              ctl->segIncC[warp] = pay[k];
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
        if (lane <= last[k]) cnt += ((S >> 1) & 1u) ? (pay[k] >> 16) : (pay[k] & 0xFFFFu);
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
      Tp = compose_opt(ctl->segF[w], Tp);
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

// One thread: take the next ticket for stage s and start its load.
__device__ void refill_stage(uint8_t *tiles, Control *ctl, const CUtensorMap *tmap, const ScanParams &p, int s) {
  const uint32_t k = atomicAdd(p.ticket, 1u);
  ctl->stage_tile[s] = k;
  if (k >= p.ntiles) return;
  const uint32_t t = p.tile_begin + k;  // document tile
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
__global__ void __launch_bounds__(kThreads, kMinCtasPerSm) scan_kernel(const __grid_constant__ CUtensorMap tmap, const ScanParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *tiles = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Control *ctl = reinterpret_cast<Control *>(tiles + kStages * kTileBytes);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;
  const int kRefillThread = 32;

  Carry cin;
  cin.count = 0; cin.state = 0; cin.ttable = 0;
  if (KIND != kUtf8) cin = *p.carry_in;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStages; s++) mbar_init(&ctl->full_bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == kRefillThread) {
    for (int s = 0; s < kStages; s++) refill_stage(tiles, ctl, &tmap, p, s);
  }
  __syncthreads();

  uint32_t phase_bits = 0;
  for (uint32_t it = 0;; it++) {
    const int s = it % kStages;
    const uint32_t tile = ctl->stage_tile[s];  // index inside this launch (the look-back chain uses it)
    if (tile >= p.ntiles) break;
    const uint32_t dtile = p.tile_begin + tile;  // document tile (addresses, positions)
    uint8_t *T = tiles + s * kTileBytes;
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

    // ---- the e / c bits of the state entering this tile, from the bytes before it (exact unless a run of
    //      >= 16 backslashes ends right at the boundary: then `known` is false and the chain decides)
    uint32_t bstate;  // bit0 e, bit2 c
    bool known = true;
    if (tile == 0) {
      bstate = cin.state & 5u;  // the launch's carry-in is exact
    } else {
      const uint32_t b = boundary_state_from_prev16(ctl->stage_prev16[s]);
      bstate = b & 5u;
      known = (b & 8u) == 0;
    }

    uint32_t m0[W], m1[W];  // kIndex: pseudo-structurals, string tail (polarity 0).  kMinify: whitespace, in_string (polarity 0)
    uint32_t c0 = 0, c1 = 0;     // this lane's outputs for tile polarity 0 / 1
    bool err0 = false, err1 = false;
    uint32_t Ttile = 0, state_in = 0, base = 0, Tprefix = 0;

    for (int attempt = 0; attempt < 2; attempt++) {
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
      __syncthreads();  // S1: all warp transducers visible

      // state entering this warp, relative to the tile (tile polarity 0): compose the warps before it
      uint32_t win = bstate;
      for (int w = 0; w < warp; w++) win = tt_apply(ctl->warpT[w], win);
      const uint32_t e_w = win & 1u, s_w = (win >> 1) & 1u, c_w = (win >> 2) & 1u;
      if (attempt == 0) {
        Ttile = ctl->warpT[0];
#pragma unroll
        for (int w = 1; w < kWarps; w++) Ttile = tt_compose(ctl->warpT[w], Ttile);
      }

      // ============================ phase 3: final masks, for both polarities ============================
      if (e_w && !warp_allbs && lane == mlane) toggle_first_nonbackslash_quote(qu, qr, nlead);
      {
        const uint32_t lp = (__popc(qr[0]) + __popc(qr[1]) + __popc(qr[2]) + __popc(qr[3])) & 1u;
        const uint32_t pb = __ballot_sync(kFull, lp != 0);
        uint32_t instr = (s_w ^ uint32_t(__popc(pb & lanemask_lt))) & 1u;
        uint32_t scal_prev = __shfl_up_sync(kFull, (sc[W - 1] & ~qr[W - 1]) >> 31, 1);
        if (lane == 0) scal_prev = c_w;
        uint32_t prev_nq = scal_prev << 31;
        uint32_t hit0 = 0, hit1 = 0;
        c0 = 0; c1 = 0;
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
            uint32_t ws = ~(op[u] | sc[u]);
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
        err0 = __any_sync(kFull, hit0 != 0);
        err1 = __any_sync(kFull, hit1 != 0);
      }
      {
        uint32_t t0 = c0, t1 = c1;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
          t0 += __shfl_xor_sync(kFull, t0, d);
          t1 += __shfl_xor_sync(kFull, t1, d);
        }
        if (lane == 0) { ctl->warpCnt[0][warp] = t0; ctl->warpCnt[1][warp] = t1; }
      }
      __syncthreads();  // S2: per-warp counts visible

      if (attempt == 1) break;
      uint32_t tc0 = 0, tc1 = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++) { tc0 += ctl->warpCnt[0][w]; tc1 += ctl->warpCnt[1][w]; }
      if (tile == 0) {
        state_in = cin.state; base = 0; Tprefix = 0;
        break;  // bstate is exact
      }
      if (known && tid == 0) st_relaxed_u64(p.count_desc + tile, pack_desc(p.epoch, kAgg, Ttile, (tc1 << 16) | tc0));
      resolve_tile(p, ctl, tile, cin.state, warp, lane, state_in, base, Tprefix);
      if ((state_in & 5u) == bstate) break;
      // only possible when the boundary state could not be read off the preceding bytes: redo with the exact one
      if (known) atomicOr(p.flags, kFlagInternal);
      bstate = state_in & 5u;
      __syncthreads();
    }

    const uint32_t pol = (state_in >> 1) & 1u;  // in_string entering the tile
    uint32_t tile_total = 0, warp_base = 0;
#pragma unroll
    for (int w = 0; w < kWarps; w++) {
      const uint32_t c = ctl->warpCnt[pol][w];
      if (w < warp) warp_base += c;
      tile_total += c;
    }
    const uint32_t Tincl = (Tprefix & kValidBit) ? tt_compose(Ttile, Tprefix & 63u) : Ttile;
    if (tid == 0) {
      st_relaxed_u64(p.count_desc + tile, pack_desc(p.epoch, kInc, Tincl, base + tile_total));
      if (last_tile) {
        Carry co;
        co.count = cin.count + base + tile_total;
        co.state = tt_apply(Tincl, cin.state);
        co.ttable = Tincl;
        *p.carry_out = co;
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

    // exclusive offsets inside the warp for the polarity that turned out to be real
    const uint32_t cnt = pol ? c1 : c0;
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const uint64_t out_base = cin.count + base + warp_base + (incl - cnt);

    // ============================ emit ============================
    if (KIND == kIndex) {
      // one loop over the lane's 128 mask bits (not one per 32-bit word: the warp runs max-over-lanes iterations)
      uint32_t *dst = p.idx_out + out_base;
      uint32_t pos = p.pos_base + dtile * uint32_t(kTileBytes) + lane_off;
      const uint32_t flip = pol ? 0u : kFull;
      uint32_t m = m0[0] & (m1[0] ^ flip), n1 = m0[1] & (m1[1] ^ flip), n2 = m0[2] & (m1[2] ^ flip), n3 = m0[3] & (m1[3] ^ flip);
      for (uint32_t left = cnt; left != 0; --left) {
        while (m == 0) {  // at most three times per lane
          m = n1; n1 = n2; n2 = n3; n3 = 0;
          pos += 32;
        }
        *dst++ = pos + (__ffs(m) - 1);
        m &= m - 1;
      }
      __syncthreads();  // all warps are done with the control block
      if (tid == kRefillThread) refill_stage(tiles, ctl, &tmap, p, s);
    } else {
      uint8_t *dst = p.dst + out_base;
      const uint32_t flip = pol ? kFull : 0u;
#pragma unroll
      for (int u = 0; u < W; u++) {
        uint32_t w8[8];
        load_unit(T, lane_off + 32 * u, w8);
        uint32_t keep = ~(m0[u] & ~(m1[u] ^ flip));
        if (last_tile) {
          const uint64_t ubase = uint64_t(dtile) * kTileBytes + lane_off + 32 * u;
          if (ubase + 32 > p.len) keep &= (ubase >= p.len) ? 0u : ((1u << uint32_t(p.len - ubase)) - 1u);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint32_t nib = (keep >> (4 * i)) & 15u;
          const uint32_t word = w8[i];
          if (nib == 15u && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
            *reinterpret_cast<uint32_t *>(dst) = word;
            dst += 4;
          } else {
#pragma unroll
            for (int b = 0; b < 4; b++)
              if ((nib >> b) & 1u) *dst++ = uint8_t(word >> (8 * b));
          }
        }
      }
      __syncthreads();  // every warp is done with stage s
      if (tid == kRefillThread) refill_stage(tiles, ctl, &tmap, p, s);
    }
    __syncthreads();  // stage_tile[s] / warpCnt are reused by the next iteration
  }

  // last CTA out resets the ticket for the next launch on this context
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const uint32_t done = atomicAdd(p.ticket + 1, 1u);
    if (done == gridDim.x - 1) {
      p.ticket[0] = 0;
      p.ticket[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------ small helpers
__global__ void gather_chars_kernel(const uint8_t *buf, const uint32_t *idx, uint32_t first, uint32_t count, uint8_t *out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) out[i] = buf[idx[first + i]];
}
__global__ void write_sentinels_kernel(uint32_t *idx, uint32_t n, uint32_t a, uint32_t b, uint32_t c) {
  idx[n] = a;
  idx[n + 1] = b;
  idx[n + 2] = c;
}

// ------------------------------------------------------------------ launchers
template <int KIND>
static cudaError_t launch_kind(const CUtensorMap *tmap, const ScanParams &p, int grid, cudaStream_t stream) {
  static bool configured[64] = {};  // per device: the attribute lives in the device's context
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(scan_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  scan_kernel<KIND><<<grid, kThreads, kSmemBytes, stream>>>(*tmap, p);
  return cudaGetLastError();
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
      cudaFuncSetAttribute(scan_kernel<kIndex>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kIndex>, kThreads, kSmemBytes);
      break;
    case kMinify:
      cudaFuncSetAttribute(scan_kernel<kMinify>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kMinify>, kThreads, kSmemBytes);
      break;
    default:
      cudaFuncSetAttribute(scan_kernel<kUtf8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, scan_kernel<kUtf8>, kThreads, kSmemBytes);
      break;
  }
  return (e == cudaSuccess && n > 0) ? n : 1;
}

cudaError_t launch_gather_chars(const uint8_t *buf, const uint32_t *idx, uint32_t first, uint32_t count, uint8_t *out,
                                cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  gather_chars_kernel<<<(count + 255) / 256, 256, 0, stream>>>(buf, idx, first, count, out);
  return cudaGetLastError();
}

cudaError_t launch_write_sentinels(uint32_t *idx, uint32_t n, uint32_t a, uint32_t b, uint32_t c, cudaStream_t stream) {
  write_sentinels_kernel<<<1, 1, 0, stream>>>(idx, n, a, b, c);
  return cudaGetLastError();
}

}  // namespace sjb200
