"""configs[2]/[3]/[4] of BASELINE.json at full size on one GPU: 1 GiB NDJSON stage1 (streaming_final), 256 MiB utf8 + minify.
Parity: full comparison with the CPU oracle port (seconds) + size-independent properties."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O
import simdjson_b200 as sj
from simdjson_b200 import corpus
port = O.Port()
res = {}
size = int(os.environ.get("BIG_BYTES", 1 << 30))
t0 = time.time(); nd = corpus.ndjson_rows(size); print("gen ndjson", time.time() - t0, flush=True)
rc, p = sj.get_active_implementation().create_dom_parser_implementation(len(nd)); assert rc == 0
p.set_option("time_kernel", 1)
d = torch.from_numpy(nd.copy()).cuda()
ts = []
for _ in range(4):
    t0 = time.time(); rc = p.stage1_device(d, sj.STREAMING_FINAL); dt = time.time() - t0
    ts.append(p.get_stat("kernel_ms"))
print(f"1 GiB ndjson streaming_final: rc={rc} n={p.n_structural_indexes} kernel_ms best {min(ts):.3f} -> {len(nd)/min(ts)/1e6:.1f} GB/s ; whole call {dt*1e3:.2f} ms", flush=True)
t0 = time.time(); want = port.stage1(nd, sj.STREAMING_FINAL); print("oracle port", time.time() - t0, want.err, want.n, flush=True)
got = p.device_index_buffer().cpu().numpy().view(np.uint32)
same = rc == want.err and p.n_structural_indexes == want.n and np.array_equal(got[: want.n + 3], want.words())
print("parity 1 GiB ndjson:", same, flush=True)
res["ndjson_1g"] = {"kernel_ms": min(ts), "gbs": len(nd) / min(ts) / 1e6, "parity": bool(same), "n": int(want.n)}
del d, got
# 256 MiB utf8 + minify
u = corpus.random_utf8(256 << 20); du = torch.from_numpy(u).cuda()
ts = []
for _ in range(4):
    r = p.validate_utf8_device(du); ts.append(p.get_stat("kernel_ms"))
print(f"256 MiB utf8: valid={r} kernel_ms best {min(ts):.3f} -> {len(u)/min(ts)/1e6:.1f} GB/s")
res["utf8_256m"] = {"kernel_ms": min(ts), "gbs": len(u) / min(ts) / 1e6, "valid": int(r)}
del du
j = corpus.random_json(256 << 20, pretty_bias=0.8, utf8_rate=0.15)
dj = torch.from_numpy(j.copy()).cuda(); dst = torch.empty(len(j), dtype=torch.uint8, device="cuda")
rc2, p2 = sj.get_active_implementation().create_dom_parser_implementation(len(j))
p2.set_option("time_kernel", 1)
ts = []
for _ in range(4):
    rcm, dl = p2.minify_device(dj, dst); ts.append(p2.get_stat("kernel_ms"))
werr, wout = port.minify(j)
same = rcm == werr and dl == len(wout) and bytes(dst[:dl].cpu().numpy()) == wout
print(f"256 MiB minify: rc={rcm} kept={dl} ({dl/len(j):.3f}) kernel_ms best {min(ts):.3f} -> {len(j)/min(ts)/1e6:.1f} GB/s parity={same}")
res["minify_256m"] = {"kernel_ms": min(ts), "gbs": len(j) / min(ts) / 1e6, "parity": bool(same), "kept_frac": dl / len(j)}
ts = []
for _ in range(4):
    rc = p2.stage1_device(dj, 0); ts.append(p2.get_stat("kernel_ms"))
print(f"256 MiB stage1 (pretty, utf8-heavy): rc={rc} n={p2.n_structural_indexes} kernel_ms best {min(ts):.3f} -> {len(j)/min(ts)/1e6:.1f} GB/s")
res["stage1_256m"] = {"kernel_ms": min(ts), "gbs": len(j) / min(ts) / 1e6}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "big_probe.json"), "w"), indent=1)
