// tokens_warp_emul.cpp -- tok::warp_string (simdjson_b200/csrc/sjb200_tokens_warp.cuh: a long string unescaped by a whole
// warp) under the host SIMT emulation: 32 OS threads are the lanes, the warp collectives are rendezvous
// (sjb200_simt.cuh, SJB200_HOST_EMU).  Driven by tests/test_tokens_emul.py against the oracle; no GPU involved.
#define SJB200_HOST_EMU 1
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "sjb200_tokens_warp.cuh"

using namespace sjb200;
thread_local simt::ThreadCtx simt::tctx;

namespace {
struct Job {
  const uint8_t *buf;
  uint64_t len;
  const uint64_t *pos;
  uint32_t npos;
  long long *out_len;   // [npos]
  uint8_t *out;         // [npos][stride]
  uint64_t stride;
  uint64_t win_lo, win_span;  // > 0: bytes [win_lo, win_lo + win_span) come from a staged copy (WindowSrc)
  const uint8_t *win;
  simt::WarpShared *warp;
  simt::CtaShared *cta;
};
struct LaneArg { Job *job; unsigned lane; };

void *lane_main(void *vp) {
  LaneArg *a = static_cast<LaneArg *>(vp);
  Job &j = *a->job;
  simt::tctx = simt::ThreadCtx();
  simt::tctx.tid = a->lane;
  simt::tctx.nctas = 1;
  simt::tctx.warp = j.warp;
  simt::tctx.ctas = j.cta;
  for (uint32_t k = 0; k < j.npos; k++) {
    long long r0, r1;
    if (j.win_span) {
      tok::WindowSrc src{j.buf, j.len, j.win, j.win_lo, j.win_span};
      r0 = tok::warp_string<false>(src, j.len, j.pos[k], nullptr, a->lane);
      r1 = tok::warp_string<true>(src, j.len, j.pos[k], j.out + k * j.stride, a->lane);
    } else {
      tok::PlainSrc src{j.buf, j.len};
      r0 = tok::warp_string<false>(src, j.len, j.pos[k], nullptr, a->lane);
      r1 = tok::warp_string<true>(src, j.len, j.pos[k], j.out + k * j.stride, a->lane);
    }
    if (a->lane == 0) j.out_len[k] = (r0 == r1) ? r0 : -99;  // the two passes must agree
  }
  return nullptr;
}
}  // namespace

extern "C" int emu_warp_strings(const uint8_t *buf, uint64_t len, const uint64_t *pos, uint32_t npos, long long *out_len, uint8_t *out, uint64_t stride,
                                uint64_t win_lo, uint64_t win_span) {
  simt::WarpShared warp;
  simt::CtaShared cta;
  pthread_barrier_init(&warp.bar, nullptr, 32);
  pthread_barrier_init(&cta.bar, nullptr, 32);
  cta.smem = nullptr;
  std::vector<uint8_t> win(win_span ? win_span : 1);
  if (win_span) memcpy(win.data(), buf + win_lo, win_span);
  Job job{buf, len, pos, npos, out_len, out, stride, win_lo, win_span, win.data(), &warp, &cta};
  LaneArg args[32];
  pthread_t th[32];
  for (unsigned l = 0; l < 32; l++) {
    args[l] = LaneArg{&job, l};
    if (pthread_create(&th[l], nullptr, lane_main, &args[l]) != 0) return -1;
  }
  for (unsigned l = 0; l < 32; l++) pthread_join(th[l], nullptr);
  pthread_barrier_destroy(&warp.bar);
  pthread_barrier_destroy(&cta.bar);
  return 0;
}
