"""host-pointer path throughput vs chunk size (tuning aid)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simdjson_b200 as sj
from simdjson_b200 import corpus
doc = corpus.random_json(64 << 20)
pin = torch.from_numpy(doc.copy()).pin_memory(); hb = pin.numpy()
rc, p = sj.get_active_implementation().create_dom_parser_implementation(len(doc))
# raw PCIe reference: H2D and D2H of the same sizes with torch
d = torch.empty(len(doc), dtype=torch.uint8, device="cuda")
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(pin, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw H2D 64 MiB pinned: {len(doc)/dt/1e9:.1f} GB/s")
for chunk in (1 << 20, 4 << 20, 16 << 20, 64 << 20):
    p.set_option("chunk_bytes", chunk)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); rc = p.stage1(hb, 0); ts.append(time.perf_counter() - t0)
    print(f"chunk {chunk>>20:3d} MiB: best {len(doc)/min(ts)/1e9:6.2f} GB/s  median {len(doc)/sorted(ts)[3]/1e9:6.2f} GB/s  rc={rc} n={p.n_structural_indexes}")
