"""GPU probe: kernel timings of the three scans on the 64 MiB bench document (diagnostic, not the bench).
SJB200_LIB selects a build variant; PROBE_BYTES the size; PROBE_TMA=0 the plain-load path."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
import simdjson_b200 as sj  # noqa: E402
from simdjson_b200 import corpus  # noqa: E402

tag = os.path.basename(os.environ.get("SJB200_LIB", "default"))
port = O.Port()
impl = sj.get_active_implementation()
size = int(os.environ.get("PROBE_BYTES", 64 << 20))
doc = corpus.random_json(size).copy()
want = port.stage1(doc, 0)
rc, p = impl.create_dom_parser_implementation(len(doc))
assert rc == 0, rc
p.set_option("time_kernel", 1)
p.set_option("use_tma", int(os.environ.get("PROBE_TMA", 1)))
if os.environ.get("PROBE_GRID"):
    p.set_option("grid", int(os.environ["PROBE_GRID"]))
d = torch.from_numpy(doc).cuda()
dst = torch.empty(len(doc), dtype=torch.uint8, device="cuda")
rc = p.stage1_device(d, 0)
got = p.device_index_buffer().cpu().numpy().view(np.uint32)
okk = rc == want.err and p.n_structural_indexes == want.n and np.array_equal(got[: want.n + 3], want.words())
res = {"tag": tag, "parity": bool(okk), "grid": p.get_stat("grid_index")}
kinds = os.environ.get("PROBE_KINDS", "stage1,minify,utf8").split(",")
for kind in kinds:
    ts = []
    for it in range(10):
        if kind == "stage1":
            rc = p.stage1_device(d, 0)
        elif kind == "minify":
            rc, dl = p.minify_device(d, dst)
        else:
            rc = p.validate_utf8_device(d)
        ts.append(p.get_stat("kernel_ms"))
    res[kind] = {"best_ms": min(ts[2:]), "median_ms": sorted(ts[2:])[len(ts[2:]) // 2], "gbs": len(doc) / min(ts[2:]) / 1e6}
u = torch.from_numpy(corpus.random_utf8(size)).cuda()
ts = []
for it in range(6):
    r = p.validate_utf8_device(u)
    ts.append(p.get_stat("kernel_ms"))
res["utf8_dense"] = {"best_ms": min(ts[1:]), "gbs": size / min(ts[1:]) / 1e6}
pin = torch.from_numpy(doc).pin_memory()
hb = pin.numpy()
hs = []
for it in range(4):
    t0 = time.time(); rc = p.stage1(hb, 0); hs.append(time.time() - t0)
res["host_path_gbs"] = len(doc) / min(hs) / 1e9
print(json.dumps(res), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "probe.jsonl"), "a") as f:
    f.write(json.dumps(res) + "\n")
