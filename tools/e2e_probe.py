"""e2e probe: stage 1 through the plug-in (libsimdjson_b200.so, unmodified reference API, pageable padded_string) and
through the Python mirror of the C ABI; sweeps the host-pipeline knobs via the SJB200_* environment.
  python tools/e2e_probe.py            one line per configuration"""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "simdjson_b200", "plugin", "libsimdjson_b200.so")


def one():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from simdjson_b200 import corpus
    size = int(os.environ.get("PROBE_BYTES", 64 << 20))
    cache = f"/tmp/probe_cache_{size}.npz"
    z = np.load(cache) if os.path.exists(cache) else None
    doc = z["doc"] if z is not None else corpus.random_json(size).copy()
    L = C.CDLL(PLUGIN)
    L.dropin_stage1_timed.restype = C.c_int
    L.dropin_stage1_timed.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_ulonglong)]
    secs = (C.c_double * 2)()
    n = C.c_uint32(0)
    calls = C.c_ulonglong(0)
    idx = np.zeros(len(doc) // 4 + 16, dtype=np.uint32)
    use = int(os.environ.get("PROBE_USE_B200", 1))
    iters = int(os.environ.get("PROBE_ITERS", 10))
    rc = L.dropin_stage1_timed(use, doc.ctypes.data, len(doc), iters, secs, C.byref(n), idx.ctypes.data, len(idx), C.byref(calls))
    res = {"tag": os.environ.get("PROBE_TAG", ""), "bytes": size, "rc": rc, "n": int(n.value), "gpu_calls": int(calls.value),
           "gbs_mean": size * iters / secs[0] / 1e9, "gbs_best": size / secs[1] / 1e9, "ms_best": secs[1] * 1e3}
    if z is not None:
        res["parity"] = bool(rc == int(z["err"]) and int(n.value) == int(z["n"]) and np.array_equal(idx[: int(z["n"]) + 3], z["words"]))
    print(json.dumps(res), flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "e2e_probe.jsonl"), "a") as f:
        f.write(json.dumps(res) + "\n")


if __name__ == "__main__":
    if os.environ.get("PROBE_CHILD"):
        one()
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    configs = [("threads4 chunk4M first512K (default)", {}),
               ("threads6", {"SJB200_COPY_THREADS": "6"}),
               ("threads6 chunk2M", {"SJB200_COPY_THREADS": "6", "SJB200_CHUNK_BYTES": str(2 << 20)}),
               ("threads4 chunk2M first256K", {"SJB200_CHUNK_BYTES": str(2 << 20), "SJB200_FIRST_CHUNK_BYTES": str(256 << 10)}),
               ("threads4 no ramp", {"SJB200_FIRST_CHUNK_BYTES": str(4 << 20)}),
               ("threads6 chunk2M slots16", {"SJB200_COPY_THREADS": "6", "SJB200_CHUNK_BYTES": str(2 << 20), "SJB200_RING_SLOTS": "16"})]
    for tag, env in configs:
        e = dict(os.environ, PROBE_CHILD="1", PROBE_TAG=tag, **env)
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=e)
