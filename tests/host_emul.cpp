// host_emul.cpp -- CPU emulation of the CUDA stage-1 algorithm's data flow.
//
// Runs the *same* per-lane arithmetic (simdjson_b200/csrc/sjb200_bits.cuh) in
// the same tile / warp / lane decomposition the kernels use, with ballots,
// shuffles and the look-back replaced by plain loops, and checks the result
// against the byte-at-a-time oracle (oracle/sj_oracle.c).  This validates the
// carry algebra (transducer composition, escape-carry resolution, the
// one-quote toggle for an incoming escape, UTF-8 carries across units / lanes
// / tiles) on a machine without a GPU.  It is a test, not a product path.
//
// It also drives the product's host epilogue (simdjson_b200/csrc/sjb200_finish.cpp: error precedence,
// sentinels, the streaming-mode boundary walks and filters) with the emulated scan results, in all
// seven stage1 modes, against the oracle.
//
// build: see tests/test_host_emul.py
#include "sjb200_bits.cuh"
#include "sjb200_finish.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <random>
#include <type_traits>
#include <vector>

extern "C" {
#include "sj_oracle.h"
}

using namespace sjb200;

template <int W, int NWARPS>
struct Emul {
  static constexpr int LANE_BYTES = 32 * W;
  static constexpr int WARP_BYTES = 32 * LANE_BYTES;
  static constexpr int TILE_BYTES = NWARPS * WARP_BYTES;

  struct Lane {
    uint32_t bs[W], qu[W], op[W], sc[W], ctl[W], ws[W];
    uint32_t qr[W];  // real (unescaped) quotes
    uint32_t st[W];  // structurals
    uint32_t keep[W];
    int nlead;       // leading backslashes
    bool allbs;
    uint32_t esc_out0;
  };

  const uint8_t *buf;
  size_t len;
  std::vector<uint32_t> idx;
  std::vector<uint8_t> minified;
  bool utf8_err = false, ctl_err = false, unclosed = false;

  uint8_t byte_at(long pos) const { return (pos >= 0 && size_t(pos) < len) ? buf[pos] : 0x20; }
  uint32_t word_at(long pos) const {
    return uint32_t(byte_at(pos)) | (uint32_t(byte_at(pos + 1)) << 8) | (uint32_t(byte_at(pos + 2)) << 16) | (uint32_t(byte_at(pos + 3)) << 24);
  }

  static void toggle_first_nonbackslash_quote(Lane &L) {
    // an incoming escape flips the "escaped" status of the first byte that is not a backslash
    if (L.allbs) return;
    const int k = L.nlead;
    L.qr[k >> 5] ^= L.qu[k >> 5] & (1u << (k & 31));
  }

  void run() {
    const size_t ntiles = (len + TILE_BYTES - 1) / TILE_BYTES;
    uint32_t state = 0;  // bit0 esc, bit1 in_string, bit2 prev_scalar  (inclusive prefix of previous tiles)
    std::vector<Lane> lanes(NWARPS * 32);
    for (size_t t = 0; t < ntiles; t++) {
      const long tile_start = long(t) * TILE_BYTES;
      uint32_t warpT[NWARPS];
      // ---- phase 1 + 2 (independent of any carry-in)
      for (int w = 0; w < NWARPS; w++) {
        uint32_t planes[32][W][8];
        utf8_carry uc[32];
        for (int l = 0; l < 32; l++) {
          const long base = tile_start + long(w) * WARP_BYTES + long(l) * LANE_BYTES;
          Lane &L = lanes[w * 32 + l];
          for (int u = 0; u < W; u++) {
            uint32_t words[8];
            for (int i = 0; i < 8; i++) words[i] = word_at(base + 32 * u + 4 * i);
            transpose32(words, planes[l][u]);
            unit_classes c = classify(planes[l][u]);
            L.bs[u] = c.bs; L.qu[u] = c.qu; L.op[u] = c.op; L.sc[u] = c.sc; L.ctl[u] = c.ctl;
            L.ws[u] = ~(c.op | c.sc);
          }
          uc[l] = utf8_carry_from_prev_word(word_at(base - 4));
        }
        for (int u = 0; u < W; u++) {  // per-unit warp vote, like the kernel
          bool any = false;
          for (int l = 0; l < 32; l++) any |= (planes[l][u][7] != 0) || utf8_carry_pending(uc[l]);
          for (int l = 0; l < 32; l++) {
            if (any) {
              if (utf8_check_unit(planes[l][u], uc[l])) utf8_err = true;
            } else {
              uc[l] = utf8_carry_zero();
            }
          }
        }
        uint32_t G = 0, P = 0;
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          uint32_t escaped[W];
          L.esc_out0 = escape_scan<W>(L.bs, escaped);
          L.nlead = leading_backslashes<W>(L.bs);
          L.allbs = (L.nlead == 32 * W);
          for (int u = 0; u < W; u++) L.qr[u] = L.qu[u] & ~escaped[u];
          G |= (L.esc_out0 & 1u) << l;
          P |= uint32_t(L.allbs) << l;
        }
        uint32_t cout0;
        const uint32_t carries = escape_carries(G, P, 0, &cout0);
        uint32_t par0 = 0;
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          if ((carries >> l) & 1) toggle_first_nonbackslash_quote(L);
          for (int u = 0; u < W; u++) par0 ^= popc32(L.qr[u]) & 1;
        }
        // transducer of the warp chunk
        const bool warp_allbs = (P == 0xFFFFFFFFu);
        uint32_t qx = 0, x_is_last = 0;
        if (!warp_allbs) {
          const int m = ctz32(~P);
          const Lane &L = lanes[w * 32 + m];
          qx = (L.qu[L.nlead >> 5] >> (L.nlead & 31)) & 1;
          x_is_last = (m == 31 && L.nlead == 32 * W - 1);
        }
        const Lane &LL = lanes[w * 32 + 31];
        const uint32_t scal0 = ((LL.sc[W - 1] & ~LL.qr[W - 1]) >> 31) & 1;
        const uint32_t esc1 = warp_allbs ? 1u : cout0;
        warpT[w] = tt_make(cout0, par0, scal0, esc1, par0 ^ qx, scal0 ^ (qx & x_is_last));
      }
      // ---- resolve carries: tile look-back is `state`; compose the warps in order
      uint32_t warp_in[NWARPS];
      uint32_t s = state;
      for (int w = 0; w < NWARPS; w++) {
        warp_in[w] = s;
        s = tt_apply(warpT[w], s);
      }
      state = s;
      // ---- phase 3: final masks
      for (int w = 0; w < NWARPS; w++) {
        const uint32_t e_w = warp_in[w] & 1, s_w = (warp_in[w] >> 1) & 1, c_w = (warp_in[w] >> 2) & 1;
        if (e_w) {
          uint32_t P = 0;
          for (int l = 0; l < 32; l++) P |= uint32_t(lanes[w * 32 + l].allbs) << l;
          if (P != 0xFFFFFFFFu) toggle_first_nonbackslash_quote(lanes[w * 32 + ctz32(~P)]);
        }
        uint32_t instr = s_w, scal = c_w;
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          const long base = tile_start + long(w) * WARP_BYTES + long(l) * LANE_BYTES;
          uint32_t prev_nq = scal << 31;
          for (int u = 0; u < W; u++) {
            const uint32_t in_string = prefix_xor32(L.qr[u]) ^ (instr ? 0xFFFFFFFFu : 0u);
            instr = in_string >> 31;
            const uint32_t nq = L.sc[u] & ~L.qr[u];
            const uint32_t follows = shl_in(prev_nq, nq, 1);
            prev_nq = nq;
            L.st[u] = (L.op[u] | (L.sc[u] & ~follows)) & ~(in_string ^ L.qr[u]);
            if (L.ctl[u] & in_string) ctl_err = true;
            // valid mask for the padded tail
            uint32_t valid = 0xFFFFFFFFu;
            const long ubase = base + 32 * u;
            if (ubase + 32 > long(len)) valid = (ubase >= long(len)) ? 0u : ((1u << (long(len) - ubase)) - 1u);
            L.keep[u] = ~(L.ws[u] & ~in_string) & valid;
            for (uint32_t mk = L.st[u]; mk; mk &= mk - 1) idx.push_back(uint32_t(ubase + ctz32(mk)));
            for (uint32_t mk = L.keep[u]; mk; mk &= mk - 1) minified.push_back(buf[ubase + ctz32(mk)]);
          }
          scal = prev_nq >> 31;
        }
      }
    }
    unclosed = (state >> 1) & 1;
    // end-of-input: the input must not end inside a multi-byte sequence.  When len is not a
    // multiple of the tile the 0x20 padding already exposes it; check explicitly otherwise.
    if (len > 0) {
      uint32_t pw = 0;
      for (int d = 1; d <= 4; d++) pw |= uint32_t(long(len) - d >= 0 ? buf[len - d] : 0x20) << (8 * (4 - d));
      if (utf8_carry_pending(utf8_carry_from_prev_word(pw))) utf8_err = true;
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// v2 data flow: ONE look-back chain.  The tile's e / c state bits come from the 16 bytes before it
// (boundary_state_from_prev16), masks and counts are produced for BOTH in-string polarities before the
// chain is consulted, and the chain element is {T, cnt[0], cnt[1]}.  A tile whose boundary state is
// unknown (>= 16 backslashes right before it) publishes no aggregate, waits for exact state and redoes
// its final-mask phase.  `window` controls how many predecessors are presented as un-finished
// (aggregate only) to the look-back, like tiles in flight on the GPU.
template <int W, int NWARPS>
struct Emul2 {
  static constexpr int LANE_BYTES = 32 * W;
  static constexpr int WARP_BYTES = 32 * LANE_BYTES;
  static constexpr int TILE_BYTES = NWARPS * WARP_BYTES;
  struct Lane {
    uint32_t bs[W], qu[W], op[W], sc[W], ctl[W];
    uint32_t qr[W], pm[W], x0[W], in0[W];
    int nlead; bool allbs;
  };
  struct Desc { bool has_agg = false, inc = false; uint32_t T = 0, cnt[2] = {0, 0}; uint32_t Tp = 0; uint64_t C = 0; bool first = false; };

  const uint8_t *buf; size_t len;
  uint32_t state0 = 0;  // scanner state entering the buffer
  int window = 0;
  std::vector<uint32_t> idx;
  std::vector<uint8_t> minified;
  bool utf8_err = false, ctl_err = false, unclosed = false;
  int redo_count = 0;
  std::mt19937_64 rng{12345};

  uint8_t byte_at(long pos) const { return (pos >= 0 && size_t(pos) < len) ? buf[pos] : 0x20; }
  uint32_t word_at(long pos) const {
    return uint32_t(byte_at(pos)) | (uint32_t(byte_at(pos + 1)) << 8) | (uint32_t(byte_at(pos + 2)) << 16) | (uint32_t(byte_at(pos + 3)) << 24);
  }
  static void toggle(Lane &L) {
    if (L.allbs) return;
    const int k = L.nlead;
    L.qr[k >> 5] ^= L.qu[k >> 5] & (1u << (k & 31));
  }

  void run() {
    const size_t ntiles = (len + TILE_BYTES - 1) / TILE_BYTES;
    std::vector<Desc> desc(ntiles);
    std::vector<Lane> lanes(NWARPS * 32);
    for (size_t t = 0; t < ntiles; t++) {
      const long tile_start = long(t) * TILE_BYTES;
      uint32_t warpT[NWARPS], Pmask[NWARPS];
      // ---- phases 1+2 (as v1)
      for (int w = 0; w < NWARPS; w++) {
        uint32_t planes[32][W][8];
        utf8_carry uc[32];
        for (int l = 0; l < 32; l++) {
          const long base = tile_start + long(w) * WARP_BYTES + long(l) * LANE_BYTES;
          Lane &L = lanes[w * 32 + l];
          for (int u = 0; u < W; u++) {
            uint32_t words[8];
            for (int i = 0; i < 8; i++) words[i] = word_at(base + 32 * u + 4 * i);
            transpose32(words, planes[l][u]);
            unit_classes c = classify(planes[l][u]);
            L.bs[u] = c.bs; L.qu[u] = c.qu; L.op[u] = c.op; L.sc[u] = c.sc; L.ctl[u] = c.ctl;
          }
          uc[l] = utf8_carry_from_prev_word(word_at(base - 4));
        }
        for (int u = 0; u < W; u++) {
          bool any = false;
          for (int l = 0; l < 32; l++) any |= (planes[l][u][7] != 0) || utf8_carry_pending(uc[l]);
          for (int l = 0; l < 32; l++) {
            if (any) { if (utf8_check_unit(planes[l][u], uc[l])) utf8_err = true; } else uc[l] = utf8_carry_zero();
          }
        }
        uint32_t G = 0, P = 0;
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          uint32_t escaped[W];
          uint32_t eo = escape_scan<W>(L.bs, escaped);
          L.nlead = leading_backslashes<W>(L.bs);
          L.allbs = (L.nlead == 32 * W);
          for (int u = 0; u < W; u++) L.qr[u] = L.qu[u] & ~escaped[u];
          G |= (eo & 1u) << l; P |= uint32_t(L.allbs) << l;
        }
        uint32_t cout0;
        const uint32_t carries = escape_carries(G, P, 0, &cout0);
        uint32_t par0 = 0;
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          if ((carries >> l) & 1) toggle(L);
          for (int u = 0; u < W; u++) par0 ^= popc32(L.qr[u]) & 1;
        }
        const bool warp_allbs = (P == 0xFFFFFFFFu);
        uint32_t qx = 0, x_is_last = 0;
        if (!warp_allbs) {
          const int m = ctz32(~P);
          const Lane &L = lanes[w * 32 + m];
          qx = (L.qu[L.nlead >> 5] >> (L.nlead & 31)) & 1;
          x_is_last = (m == 31 && L.nlead == 32 * W - 1);
        }
        const Lane &LL = lanes[w * 32 + 31];
        const uint32_t scal0 = ((LL.sc[W - 1] & ~LL.qr[W - 1]) >> 31) & 1;
        warpT[w] = tt_make(cout0, par0, scal0, warp_allbs ? 1u : cout0, par0 ^ qx, scal0 ^ (qx & x_is_last));
        Pmask[w] = P;
      }
      uint32_t Ttile = warpT[0];
      for (int w = 1; w < NWARPS; w++) Ttile = tt_compose(warpT[w], Ttile);

      // ---- boundary state of the tile from the 16 bytes before it (tile 0: the carry-in state is exact)
      uint32_t guess;
      bool known;
      if (t == 0) { guess = state0 & 5u; known = true; }
      else {
        uint32_t p16[4];
        for (int i = 0; i < 4; i++) p16[i] = word_at(tile_start - 16 + 4 * i);
        const uint32_t b = boundary_state_from_prev16(p16);
        guess = b & 5u; known = !(b & 8u);
      }
      uint32_t cnt[2] = {0, 0};
      bool err_pol[2] = {false, false};
      uint32_t toggled_lane[NWARPS];
      auto phase3 = [&](uint32_t et, uint32_t ct) {
        cnt[0] = cnt[1] = 0; err_pol[0] = err_pol[1] = false;
        uint32_t s = et | (ct << 2);  // tile-relative: in_string bit = 0
        for (int w = 0; w < NWARPS; w++) {
          const uint32_t e_w = s & 1, s_w = (s >> 1) & 1, c_w = (s >> 2) & 1;
          s = tt_apply(warpT[w], s);
          toggled_lane[w] = 0xFFFFFFFFu;
          if (e_w && Pmask[w] != 0xFFFFFFFFu) { toggled_lane[w] = ctz32(~Pmask[w]); toggle(lanes[w * 32 + toggled_lane[w]]); }
          uint32_t instr = s_w, scal = c_w;
          for (int l = 0; l < 32; l++) {
            Lane &L = lanes[w * 32 + l];
            const long base = tile_start + long(w) * WARP_BYTES + long(l) * LANE_BYTES;
            uint32_t prev_nq = scal << 31;
            for (int u = 0; u < W; u++) {
              const uint32_t in_string = prefix_xor32(L.qr[u]) ^ (instr ? 0xFFFFFFFFu : 0u);
              instr = in_string >> 31;
              const uint32_t nq = L.sc[u] & ~L.qr[u];
              const uint32_t follows = shl_in(prev_nq, nq, 1);
              prev_nq = nq;
              L.pm[u] = L.op[u] | (L.sc[u] & ~follows);
              L.x0[u] = in_string ^ L.qr[u];
              L.in0[u] = in_string;
              cnt[0] += popc32(L.pm[u] & ~L.x0[u]);
              cnt[1] += popc32(L.pm[u] & L.x0[u]);
              if (L.ctl[u] & in_string) err_pol[0] = true;
              if (L.ctl[u] & ~in_string) err_pol[1] = true;
              (void)base;
            }
            scal = prev_nq >> 31;
          }
        }
      };
      auto undo_toggles = [&]() {
        for (int w = 0; w < NWARPS; w++) if (toggled_lane[w] != 0xFFFFFFFFu) toggle(lanes[w * 32 + toggled_lane[w]]);
      };
      phase3(guess & 1, (guess >> 2) & 1);
      Desc &D = desc[t];
      D.T = Ttile;
      if (known) { D.has_agg = true; D.cnt[0] = cnt[0]; D.cnt[1] = cnt[1]; }
      // ---- the look-back: nearest predecessor presented as inclusive, then a forward fold
      long i = long(t) - 1;
      {
        const long oldest_unfinished = std::max<long>(0, long(t) - long(window ? rng() % (window + 1) : 0));
        while (i >= 0 && i >= oldest_unfinished && desc[i].has_agg) i--;  // these are seen as aggregates
      }
      uint32_t S = (i < 0) ? state0 : tt_apply(desc[i].Tp, state0);
      uint64_t C = (i < 0) ? 0 : desc[i].C;
      uint32_t Tp = (i < 0) ? 0 : desc[i].Tp;
      bool haveTp = i >= 0;
      for (long j = i + 1; j < long(t); j++) {
        C += desc[j].cnt[(S >> 1) & 1];
        S = tt_apply(desc[j].T, S);
        Tp = haveTp ? tt_compose(desc[j].T, Tp) : desc[j].T;
        haveTp = true;
      }
      if ((S & 5u) != guess) {
        if (known) { fprintf(stderr, "BUG: boundary state mismatch at tile %zu (guess %u exact %u)\n", t, guess, S & 5u); exit(2); }
        undo_toggles();
        phase3(S & 1, (S >> 2) & 1);
        redo_count++;
      }
      const uint32_t pol = (S >> 1) & 1;
      D.Tp = haveTp ? tt_compose(Ttile, Tp) : Ttile;
      D.C = C + cnt[pol];
      D.inc = true;
      if (err_pol[pol]) ctl_err = true;
      // ---- emit
      for (int w = 0; w < NWARPS; w++)
        for (int l = 0; l < 32; l++) {
          Lane &L = lanes[w * 32 + l];
          const long base = tile_start + long(w) * WARP_BYTES + long(l) * LANE_BYTES;
          for (int u = 0; u < W; u++) {
            const long ubase = base + 32 * u;
            uint32_t valid = 0xFFFFFFFFu;
            if (ubase + 32 > long(len)) valid = (ubase >= long(len)) ? 0u : ((1u << (long(len) - ubase)) - 1u);
            const uint32_t st = pol ? (L.pm[u] & L.x0[u]) : (L.pm[u] & ~L.x0[u]);
            const uint32_t ws = ~(L.op[u] | L.sc[u]);
            const uint32_t ins = pol ? ~L.in0[u] : L.in0[u];
            const uint32_t keep = ~(ws & ~ins) & valid;
            for (uint32_t mk = st; mk; mk &= mk - 1) idx.push_back(uint32_t(ubase + ctz32(mk)));
            for (uint32_t mk = keep; mk; mk &= mk - 1) minified.push_back(buf[ubase + ctz32(mk)]);
          }
        }
      if (idx.size() != D.C) { fprintf(stderr, "BUG: count chain %zu vs %llu at tile %zu\n", idx.size(), (unsigned long long)D.C, t); exit(2); }
      if (t + 1 == ntiles) unclosed = (tt_apply(D.Tp, state0) >> 1) & 1;
    }
    if (ntiles == 0) unclosed = (state0 >> 1) & 1;
    if (len > 0) {
      uint32_t pw = 0;
      for (int d = 1; d <= 4; d++) pw |= uint32_t(long(len) - d >= 0 ? buf[len - d] : 0x20) << (8 * (4 - d));
      if (utf8_carry_pending(utf8_carry_from_prev_word(pw))) utf8_err = true;
    }
  }
};

static long g_redos = 0;

template <int W, int NWARPS, int V2 = 0>
static int check(const std::vector<uint8_t> &in, const char *what) {
  typename std::conditional<V2 != 0, Emul2<W, NWARPS>, Emul<W, NWARPS>>::type e;
  e.buf = in.data();
  e.len = in.size();
  if constexpr (V2 != 0) e.window = (V2 == 1) ? 0 : 7;
  e.run();
  if constexpr (V2 != 0) g_redos += e.redo_count;
  // oracle, raw pieces
  std::vector<uint32_t> oidx(sjo_index_capacity(in.size()) + 16);
  uint32_t on = 0xDEADBEEF;
  int oerr = sjo_stage1(in.data(), in.size(), in.size(), SJO_STREAMING_FINAL /*tolerates unclosed*/, oidx.data(), &on);
  (void)oerr;
  // raw structural list: recompute with regular mode to learn the error class
  std::vector<uint32_t> ridx(sjo_index_capacity(in.size()) + 16);
  uint32_t rn = 0xDEADBEEF;
  int rerr = sjo_stage1(in.data(), in.size(), in.size(), SJO_REGULAR, ridx.data(), &rn);
  bool o_utf8 = sjo_validate_utf8(in.data(), in.size());
  std::vector<uint8_t> mout(in.size() + 1);
  size_t mlen = 0;
  int merr = sjo_minify(in.data(), in.size(), mout.data(), &mlen);
  int bad = 0;
  if (e.utf8_err == o_utf8) { bad = 1; }
  if (e.unclosed != (merr == SJO_UNCLOSED_STRING)) { bad = 2; }
  if (!e.unclosed) {
    if (e.minified.size() != mlen || memcmp(e.minified.data(), mout.data(), mlen) != 0) bad = 3;
  }
  if (in.size() > 0) {
    if (e.unclosed) {
      if (rerr != SJO_UNCLOSED_STRING) bad = 4;
    } else if (e.ctl_err) {
      if (rerr != SJO_UNESCAPED_CHARS) bad = 5;
    } else {
      if (rerr == SJO_UNCLOSED_STRING || rerr == SJO_UNESCAPED_CHARS) bad = 6;
      else if (rn != e.idx.size() || memcmp(ridx.data(), e.idx.data(), rn * 4) != 0) bad = 7;
    }
  }
  if (bad) {
    fprintf(stderr, "MISMATCH kind=%d (%s) W=%d NWARPS=%d len=%zu\n", bad, what, W, NWARPS, in.size());
    fprintf(stderr, "  emul: n=%zu utf8_err=%d ctl=%d unclosed=%d | oracle: err=%d n=%u utf8ok=%d merr=%d\n", e.idx.size(), e.utf8_err,
            e.ctl_err, e.unclosed, rerr, rn, o_utf8, merr);
    fprintf(stderr, "  hex:");
    for (size_t i = 0; i < in.size() && i < 400; i++) fprintf(stderr, "%02x", in[i]);
    fprintf(stderr, "\n");
  }
  return bad;
}

// full stage1 semantics in every mode: emulated scan + product finish_stage1 vs oracle
static int check_modes(const std::vector<uint8_t> &in, int mode) {
  const size_t cap = in.size();
  std::vector<uint32_t> oidx(sjo_index_capacity(cap) + 16, 0xABABABABu);
  uint32_t on = 0xDEADBEEF;
  const int oerr = sjo_stage1(in.data(), in.size(), cap, mode, oidx.data(), &on);
  // product path, host flavour (mirrors sjb200_stage1 in sjb200_capi.cu)
  std::vector<uint32_t> pidx(sjo_index_capacity(cap) + 16, 0xABABABABu);
  uint32_t pn = 0xDEADBEEF;
  int perr;
  size_t len = in.size();
  do {
    if (len > cap) { perr = kCapacity; break; }
    if (len == 0) { perr = kEmpty; break; }
    if (mode != kRegular) {
      const size_t k = len < 3 ? len : 3;
      len = trim_partial_utf8_tail(in.data() + len - k, k, len);
      if (len == 0) { perr = kUtf8Error; break; }
    }
    Emul<4, 2> e;
    e.buf = in.data();
    e.len = len;
    e.run();
    FinishInput fi;
    fi.mode = mode; fi.len = len; fi.count = e.idx.size();
    fi.state = e.unclosed ? 2u : 0u;
    fi.flags = (e.utf8_err ? kFlagUtf8 : 0u) | (e.ctl_err ? kFlagCtl : 0u);
    fi.sentinels_written = false;
    const bool early = (mode == kRegular && e.unclosed) || e.ctl_err;
    if (!early) memcpy(pidx.data(), e.idx.data(), e.idx.size() * 4);
    HostStructuralReader reader(in.data(), pidx.data());
    HostIndexWriter writer(pidx.data());
    bool dirty = false;
    perr = finish_stage1(fi, reader, writer, &pn, in.data(), pidx.data(), &dirty);
  } while (0);
  int bad = 0;
  if (perr != oerr || pn != on) bad = 1;
  else if (on != 0xDEADBEEF && memcmp(pidx.data(), oidx.data(), (size_t(on) + 3) * 4) != 0) bad = 2;
  if (bad) {
    fprintf(stderr, "MODE MISMATCH kind=%d mode=%d len=%zu: product err=%d n=%u | oracle err=%d n=%u\n  hex:", bad, mode, in.size(), perr, pn, oerr, on);
    for (size_t i = 0; i < in.size() && i < 400; i++) fprintf(stderr, "%02x", in[i]);
    fprintf(stderr, "\n");
  }
  return bad;
}

static std::vector<uint8_t> multi_document(std::mt19937_64 &rng, const char *sep) {
  static const char *frag[] = {"{\"a\":1}", "[1,2,3]", "{\"k\":[true,false,null]}", "\"str\"", "123", "true", "{\"x\":\"y\\\"z\"}", "[[],{}]",
                               "{\"u\":\"\xc3\xa9\xe2\x82\xac\"}", "{\"deep\":{\"a\":[1,{\"b\":2}]}}", "null", "-1.5e3", "{\"a\":\"\\\\\"}"};
  static const char *broken[] = {"{\"a\":", "[1,2", "{\"k\":[tr", "\"unterminated", "{\"x\":\"y\\", "[[", "{\"u\":\"\xe2\x82", "{\"u\":\"\xf0\x9f", "]", "}", ","};
  std::vector<uint8_t> out;
  auto put = [&](const char *t) { out.insert(out.end(), t, t + strlen(t)); };
  auto putsep = [&]() {
    int k = rng() % 5;
    if (k == 0) put(sep); else if (k == 1) { put(sep); put(" "); } else if (k == 2) { put(" "); put(sep); } else if (k == 3) { put(sep); put("\n"); } else if (strcmp(sep, " ") == 0) {} else put(sep);
  };
  if (rng() % 3 == 0) putsep();
  int nd = rng() % 12;
  for (int i = 0; i < nd; i++) { put(frag[rng() % 13]); putsep(); }
  if (rng() % 5 < 3) { put(broken[rng() % 11]); if (rng() % 2) putsep(); }
  if (rng() % 2) while (!out.empty() && (out.back() == ' ' || out.back() == '\n')) out.pop_back();
  return out;
}

int main(int argc, char **argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 20000;
  std::mt19937_64 rng(0x5eed1234);
  const char *alphabets[] = {"\\\\\\\"\" {}[],: \n\tabc1\x01\x0c\x1a\x1e", "\\\"", "\\\\\\\\\\\\\\\"a ", "\"{}[],:0 ", " \n\r\t\"a\\", ",{}[] 1 \"a\":\n"};
  const char *utf8bits[] = {"\xc3\xa9", "\xe2\x82\xac", "\xf0\x9f\x98\x80", "\xff", "\xc3", "\xe2\x82", "\xf0\x9f\x98", "\x80", "\xed\xa0\x80", "\xc0\xaf", "\xf4\x90\x80\x80", "\xe0\x9f\xbf", "\xf0\x8f\xbf\xbf", "\xf5\x80\x80\x80", "\xed\x9f\xbf", "\xf4\x8f\xbf\xbf", "\xe0\xa0\x80", "\xf0\x90\x80\x80", "\xc2\x80", "\xdf\xbf"};
  int fails = 0;
  for (int it = 0; it < iters && fails < 5; it++) {
    std::vector<uint8_t> in;
    int kind = rng() % 8;
    size_t n = rng() % 1200;
    const char *a = alphabets[rng() % 6];
    size_t alen = strlen(a);
    if (kind == 0) {  // long backslash runs around lane / warp / tile boundaries
      size_t pre = rng() % 3 ? (rng() % 5) * 128 + (rng() % 9) - 4 + 4096 * (rng() % 3) : rng() % 300;
      if (pre > 20000) pre = 0;
      for (size_t i = 0; i < pre; i++) in.push_back(a[rng() % alen]);
      size_t run = (size_t[]){1, 2, 3, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 4095, 4096, 4097, 8191, 8192, 8193}[rng() % 21] + rng() % 2;
      for (size_t i = 0; i < run; i++) in.push_back('\\');
      size_t post = rng() % 300;
      for (size_t i = 0; i < post; i++) in.push_back(a[rng() % alen]);
    } else if (kind == 1) {  // UTF-8 fragments at arbitrary offsets
      for (size_t i = 0; i < n; i++) in.push_back(rng() % 4 ? 'a' + rng() % 26 : a[rng() % alen]);
      int k = 1 + rng() % 8;
      for (int j = 0; j < k; j++) {
        const char *f = utf8bits[rng() % 20];
        size_t pos = in.empty() ? 0 : rng() % (in.size() + 1);
        if (rng() % 3 == 0 && in.size() > 140) pos = (rng() % (in.size() / 128)) * 128 + (rng() % 7) - 3;
        if (pos > in.size()) pos = in.size();
        in.insert(in.begin() + pos, f, f + strlen(f));
      }
    } else if (kind == 2) {  // mostly valid multi-byte text
      while (in.size() < n) {
        const char *f = utf8bits[(size_t[]){0, 1, 2, 14, 15, 16, 17, 18, 19}[rng() % 9]];
        if (rng() % 3) in.push_back(' ' + rng() % 90); else in.insert(in.end(), f, f + strlen(f));
      }
      if (rng() % 2 && !in.empty()) in[rng() % in.size()] ^= 1u << (rng() % 8);
      if (rng() % 4 == 0 && !in.empty()) in.resize(in.size() - rng() % std::min<size_t>(in.size(), 4));
    } else if (kind == 3) {  // big-ish: several tiles
      n = 3000 + rng() % 40000;
      for (size_t i = 0; i < n; i++) in.push_back(rng() % 3 ? 'a' + rng() % 26 : a[rng() % alen]);
    } else {
      for (size_t i = 0; i < n; i++) in.push_back(a[rng() % alen]);
    }
    fails += check<4, 2>(in, "W4x2") != 0;
    fails += check<2, 1>(in, "W2x1") != 0;
    if (it % 8 == 0) fails += check<4, 8>(in, "W4x8") != 0;
    if (it % 8 == 1) fails += check<1, 3>(in, "W1x3") != 0;
    fails += check<1, 1, 2>(in, "v2 W1x1 windowed") != 0;   // 1 KiB tiles: many tile boundaries, aggregates in the chain
    fails += check<2, 1, 1>(in, "v2 W2x1") != 0;
    if (it % 4 == 0) fails += check<4, 2, 2>(in, "v2 W4x2 windowed") != 0;
  }
  // exact tile multiples (no padding anywhere) incl. a truncated sequence at the very end
  for (int rep = 0; rep < 50 && fails < 5; rep++) {
    std::vector<uint8_t> in(8192, 'a');
    for (int j = 0; j < 40; j++) in[rng() % in.size()] = "\"\\ {}:,\n"[rng() % 8];
    const char *tails[] = {"\xc3", "\xe2\x82", "\xf0\x9f\x98", "\xf0\x9f\x98\x80", "ab", "\xe2\x82\xac"};
    const char *t = tails[rep % 6];
    memcpy(in.data() + in.size() - strlen(t), t, strlen(t));
    fails += check<4, 2>(in, "exact") != 0;
    fails += check<2, 1>(in, "exact") != 0;
    fails += check<1, 1, 2>(in, "v2 exact") != 0;
    fails += check<4, 2, 2>(in, "v2 exact") != 0;
  }
  // all seven stage1 modes through the product's finish logic
  for (int it = 0; it < iters && fails < 5; it++) {
    int family = it % 4;
    std::vector<uint8_t> in;
    int mode;
    if (family == 0) {
      const char *a = alphabets[rng() % 6];
      size_t n = rng() % 500, alen = strlen(a);
      for (size_t i = 0; i < n; i++) in.push_back(a[rng() % alen]);
      mode = rng() % 7;
    } else if (family == 1) { in = multi_document(rng, " "); mode = 1 + rng() % 2; }
    else if (family == 2) { in = multi_document(rng, "\x1e"); in.insert(in.begin(), 0x1e); mode = 3 + rng() % 2; }
    else { in = multi_document(rng, ","); mode = 5 + rng() % 2; }
    fails += check_modes(in, mode) != 0;
  }
  if (fails) { printf("FAILED\n"); return 1; }
  printf("host emulation OK (%d cases, %ld v2 redo tiles)\n", iters, g_redos);
  return 0;
}
