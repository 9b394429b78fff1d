"""First-contact GPU probe: correctness spot checks + kernel timings (diagnostic, not the bench)."""
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

out = {}
port = O.Port()
impl = sj.get_active_implementation()
size = int(os.environ.get("PROBE_BYTES", 64 << 20))
doc = corpus.random_json(size)
want = port.stage1(doc, 0)
rc, p = impl.create_dom_parser_implementation(len(doc))
assert rc == 0, rc
p.set_option("time_kernel", 1)
d = torch.from_numpy(doc).cuda()
dst = torch.empty(len(doc), dtype=torch.uint8, device="cuda")
for tma in (1, 0):
    p.set_option("use_tma", tma)
    t0 = time.time()
    rc = p.stage1_device(d, 0)
    dt = time.time() - t0
    got = p.device_index_buffer().cpu().numpy().view(np.uint32)
    okk = rc == want.err and p.n_structural_indexes == want.n and np.array_equal(got[: want.n + 3], want.words())
    print(f"stage1 tma={tma}: rc={rc} n={p.n_structural_indexes} want={want.n} parity={okk} first_call={dt*1e3:.2f} ms err='{p.last_cuda_error()}'", flush=True)
    best = {}
    for kind in ("stage1", "minify", "utf8"):
        ts = []
        for it in range(8):
            if kind == "stage1":
                rc = p.stage1_device(d, 0)
            elif kind == "minify":
                rc, dl = p.minify_device(d, dst)
            else:
                rc = p.validate_utf8_device(d)
            ts.append(p.get_stat("kernel_ms"))
        best[kind] = min(ts[2:])
        print(f"  {kind:7s} tma={tma} kernel_ms best={best[kind]:.4f} median={sorted(ts)[len(ts)//2]:.4f}  -> {len(doc)/best[kind]/1e6:.1f} GB/s in (rc={rc})", flush=True)
    out[f"tma{tma}"] = best
werr, wout = port.minify(doc)
rc, dl = p.minify_device(d, dst)
print("minify parity:", rc == werr and bytes(dst[:dl].cpu().numpy()) == wout)
print("utf8:", p.validate_utf8_device(d), port.validate_utf8(doc))
u = corpus.random_utf8(size)
du = torch.from_numpy(u).cuda()
ts = []
for it in range(6):
    r = p.validate_utf8_device(du)
    ts.append(p.get_stat("kernel_ms"))
print(f"utf8 (53% non-ascii) valid={r} kernel_ms best={min(ts[1:]):.4f} -> {len(u)/min(ts[1:])/1e6:.1f} GB/s")
# host path
pin = torch.from_numpy(doc).pin_memory()
hb = pin.numpy()
for it in range(3):
    t0 = time.time(); rc = p.stage1(hb, 0); dt = time.time() - t0
    print(f"host-path stage1: rc={rc} n={p.n_structural_indexes} {dt*1e3:.2f} ms -> {len(doc)/dt/1e9:.2f} GB/s e2e")
print("grid:", p.get_stat("grid_index"), "sms:", p.get_stat("sm_count"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w"))
