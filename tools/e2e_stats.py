"""where a host-pointer stage-1 call spends its time (Python mirror of the C ABI, pageable input): waiting for staged
chunks / inside CUDA calls / final synchronise"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import simdjson_b200 as sj
from simdjson_b200 import corpus
doc = corpus.random_json(64 << 20).copy()
rc, p = sj.get_active_implementation().create_dom_parser_implementation(len(doc))
for threads, chunk, skip, zc in ((6, 4 << 20, 0, 1), (6, 4 << 20, 1, 1), (6, 4 << 20, 0, 0), (8, 8 << 20, 0, 1), (8, 8 << 20, 1, 1), (6, 1 << 20, 0, 1), (4, 4 << 20, 0, 1)):
    p.set_option("copy_threads", threads); p.set_option("chunk_bytes", chunk); p.set_option("host_skip_scan", skip); p.set_option("zero_copy_out", zc)
    for _ in range(3):
        p.stage1(doc, 0)
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); rc = p.stage1(doc, 0); ts.append(time.perf_counter() - t0)
    print(f"threads {threads} chunk {chunk>>20}M skip_scan {skip} zero_copy {zc}: best {min(ts)*1e3:.3f} ms ({len(doc)/min(ts)/1e9:.1f} GB/s)  wait {p.get_stat('host_wait_ms'):.3f}  issue {p.get_stat('host_issue_ms'):.3f}  sync {p.get_stat('host_sync_ms'):.3f} ms  paths in/out {p.get_stat('input_path')}/{p.get_stat('output_path')}")
