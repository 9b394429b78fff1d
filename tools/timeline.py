"""Per-chain-element (super-tile) phase timeline of one stage-1 launch (tuning aid).
Slots: 0 first tile starts, 3 all tiles scanned (aggregate ready), 6 look-back: descriptors valid, 4 resolved, 5 emitted;
slot 7 = blockIdx<<32 | loop iteration."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simdjson_b200 as sj  # noqa: E402
from simdjson_b200 import corpus  # noqa: E402

size = int(os.environ.get("PROBE_BYTES", 64 << 20))
doc = corpus.random_json(size).copy()
rc, p = sj.get_active_implementation().create_dom_parser_implementation(len(doc))
p.set_option("debug_timeline", 1)
p.set_option("time_kernel", 1)
d = torch.from_numpy(doc).cuda()
for _ in range(3):
    p.stage1_device(d, 0)
print("kernel_ms", p.get_stat("kernel_ms"))
nt = (size + 32767) // 32768
buf = np.zeros((nt, 8), dtype=np.uint64)
n = sj.lib().sjb200_get_debug_timeline(p._ctx, buf.ctypes.data, nt)
t = buf[:n].astype(np.int64)
t = t[t[:, 0] > 0]
n = len(t)
t0 = t[:, 0].min()
print("elements", n, "span_us", (t[:, 5].max() - t0) / 1e3)
ctas = (t[:, 7] >> 32) & 0xFFFF
smid = t[:, 7] >> 48
iters = t[:, 7] & 0xFFFFFFFF
print("ctas", len(np.unique(ctas)), "elements per cta: max", np.bincount(ctas.astype(np.int64)).max())
print("resolve lag (resolved - scanned)  mean %.2f p90 %.2f max %.2f us" % (((t[:, 4] - t[:, 3]) / 1e3).mean(), np.percentile((t[:, 4] - t[:, 3]) / 1e3, 90), ((t[:, 4] - t[:, 3]) / 1e3).max()))
print("emit done after resolved          mean %.2f p90 %.2f max %.2f us" % (((t[:, 5] - t[:, 4]) / 1e3).mean(), np.percentile((t[:, 5] - t[:, 4]) / 1e3, 90), ((t[:, 5] - t[:, 4]) / 1e3).max()))
if os.environ.get("PROBE_KERNEL", "4") == "4":
    # scan4: slots 0 scan start (warp 0), 3 all 8 blocks scanned (last scan warp; aggregate published), 6 the chain warp
    # picks the element up, 4 resolved, 5 emitted (warp 0)
    o = np.argsort(t[:, 0])  # tickets are handed out in order; rows are indexed by element already
    scanned = t[:, 3].astype(np.float64)
    ready = np.maximum.accumulate(scanned)  # every predecessor's aggregate is out
    us = lambda a: (a.mean() / 1e3, np.median(a) / 1e3, np.percentile(a, 90) / 1e3, a.max() / 1e3)
    for nm, a in (("scan (start -> all blocks scanned)", t[:, 3] - t[:, 0]), ("resolved - scanned", t[:, 4] - t[:, 3]), ("chain pick-up - scanned", t[:, 6] - t[:, 3]),
                  ("resolved - (all predecessors scanned)", t[:, 4] - ready), ("emitted - resolved", t[:, 5] - t[:, 4]),
                  ("emitted - scan start", t[:, 5] - t[:, 0])):
        print("%-40s mean %7.2f us  p50 %7.2f  p90 %7.2f  max %7.2f" % ((nm,) + us(a.astype(np.float64))))
    v = (t[:, 1] > 0) & (t[:, 2] > 0)
    if v.any():
        for nm, a in (("warp 0: its own block scan (slot1 - slot0)", (t[:, 1] - t[:, 0])[v]), ("warp 0: emit (slot5 - slot2: resolved seen -> stored)", (t[:, 5] - t[:, 2])[v]),
                      ("warp 0: emit start - resolved (slot2 - slot4)", (t[:, 2] - t[:, 4])[v])):
            print("%-52s mean %7.2f us  p50 %7.2f  p90 %7.2f  max %7.2f" % ((nm,) + us(a.astype(np.float64))))
    dur = (t[:, 3] - t[:, 0]) / 1e3
    print("slowest scans (element, cta, sm, iteration, start us, duration us):")
    for i in np.argsort(-dur)[:14]:
        print("   #%d cta %d sm %d it %d start %.1f dur %.1f" % (i, ctas[i], smid[i], iters[i], (t[i, 0] - t0) / 1e3, dur[i]))
    gate = np.maximum.accumulate(scanned)
    jumps = np.argsort(-(gate[1:] - gate[:-1]))[:8] + 1
    print("elements that gate their successors longest (element, scanned us, gate jump us):", [(int(i), round((scanned[i] - t0) / 1e3, 1), round((gate[i] - gate[i - 1]) / 1e3, 1)) for i in sorted(jumps)])
    for i in sorted(jumps)[:6]:
        rows = np.where(ctas == ctas[i])[0]
        rows = rows[np.argsort(t[rows, 0])]
        print("gate #%d: cta %d sm %d it %d | its CTA:" % (i, ctas[i], smid[i], iters[i]), " ".join("[#%d it%d: %.1f %.1f %.1f %.1f %.1f]" % (r, iters[r], (t[r, 0] - t0) / 1e3, (t[r, 3] - t0) / 1e3, (t[r, 6] - t0) / 1e3, (t[r, 4] - t0) / 1e3, (t[r, 5] - t0) / 1e3) for r in rows))
    for c in np.unique(ctas)[:3]:
        rows = t[ctas == c]
        rows = rows[np.argsort(rows[:, 0])]
        print("cta", int(c), " ".join("[#%d: %.1f %.1f %.1f %.1f %.1f]" % (int(np.where((t == r).all(axis=1))[0][0]), (r[0] - t0) / 1e3, (r[3] - t0) / 1e3, (r[6] - t0) / 1e3, (r[4] - t0) / 1e3, (r[5] - t0) / 1e3) for i, r in enumerate(rows)))
        print("   warp 0 (scan start, scan end, emit start, emit end):", " ".join("[%.1f %.1f | %.1f %.1f]" % ((r[0] - t0) / 1e3, (r[1] - t0) / 1e3, (r[2] - t0) / 1e3, (r[5] - t0) / 1e3) for r in rows))
    sys.exit(0)
front = (t[:, 3] - t[:, 0]) / 1e3
v6 = t[:, 6] > 0
wait = (t[v6, 6] - t[v6, 3]) / 1e3
walk = (t[v6, 4] - t[v6, 6]) / 1e3
emit = (t[:, 5] - t[:, 4]) / 1e3
for nm, a in (("front (scan all tiles)", front), ("look-back: wait for predecessors", wait), ("look-back: fold", walk), ("emit", emit)):
    print(f"{nm:34s} mean {a.mean():7.2f} us  p50 {np.median(a):7.2f}  p90 {np.percentile(a, 90):7.2f}  max {a.max():7.2f}")
order = np.argsort(t[:, 0])
for q in np.linspace(0, n - 1, 8).astype(int):
    i = order[q]
    r = (t[i] - t0) / 1e3
    print(f"element started #{q:4d}: start {r[0]:7.2f} scanned {r[3]:7.2f} desc-valid {r[6] if t[i,6] else float('nan'):7.2f} resolved {r[4]:7.2f} emitted {r[5]:7.2f}")
