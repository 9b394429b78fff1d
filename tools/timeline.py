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
if os.environ.get("PROBE_KERNEL"):
    p.set_option("kernel", int(os.environ["PROBE_KERNEL"]))
if os.environ.get("PROBE_R"):
    p.set_option("sub_per_super", int(os.environ["PROBE_R"]))
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
ctas = t[:, 7] >> 32
print("ctas", len(np.unique(ctas)), "elements per cta: max", np.bincount(ctas.astype(np.int64)).max())
print("resolve lag (resolved - scanned)  mean %.2f p90 %.2f max %.2f us" % (((t[:, 4] - t[:, 3]) / 1e3).mean(), np.percentile((t[:, 4] - t[:, 3]) / 1e3, 90), ((t[:, 4] - t[:, 3]) / 1e3).max()))
print("emit done after resolved          mean %.2f p90 %.2f max %.2f us" % (((t[:, 5] - t[:, 4]) / 1e3).mean(), np.percentile((t[:, 5] - t[:, 4]) / 1e3, 90), ((t[:, 5] - t[:, 4]) / 1e3).max()))
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
