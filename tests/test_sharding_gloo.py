"""N>1 host logic on CPU: two gloo ranks run the product's sharding protocol (simdjson_b200/sharding.py:
all-gather of {transducer,count,flags}, fold, re-scan on wrong speculation) with the shard scans done by the CPU
oracle, and together reproduce a single scan of the whole buffer."""
import ctypes as C
import os
import random
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, docs, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch

    import oracle_lib as O
    from simdjson_b200 import sharding
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = O.Port().L
    L.sjo_scan_shard.restype = C.c_uint64
    L.sjo_scan_shard.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    L.sjo_transducer.restype = C.c_uint32
    L.sjo_transducer.argtypes = [C.c_void_p, C.c_size_t]
    results = []
    for doc in docs:
        a = np.frombuffer(doc, dtype=np.uint8)
        cuts = sharding.shard_cuts(a, world)
        shard = np.ascontiguousarray(a[cuts[rank]: cuts[rank + 1]])
        idx = np.zeros(len(shard) + 1, dtype=np.uint32)

        def scan(state_in):
            so = C.c_uint32(0)
            n = L.sjo_scan_shard(shard.ctypes.data, len(shard), state_in, idx.ctypes.data, C.byref(so))
            utf8_bad = 0 if O.Port().validate_utf8(shard) else 1
            return int(L.sjo_transducer(shard.ctypes.data, len(shard))), int(n), utf8_bad

        def all_gather(v):
            t = torch.from_numpy(v.copy())
            outl = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(outl, t)
            return np.stack([o.numpy() for o in outl])

        r = sharding.exchange(scan, rank, world, all_gather)
        glob = idx[: r["count"]].astype(np.int64) + cuts[rank]
        results.append((r["base"], r["count"], r["state_in"], r["rescanned"], r["flags"], glob.tolist()))
    out.put((rank, results))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_protocol_gloo():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    from simdjson_b200 import corpus
    rng = random.Random(11)
    docs = [bytes(corpus.random_json(200000, seed=5)), bytes(corpus.ndjson_rows(150000, seed=6))]
    # adversarial: the cut lands inside a string / after a backslash, so rank 1's speculation is wrong
    docs.append(b'{"k":"' + b"x" * 5000 + b'\\\\' + b"y" * 4999 + b'","z":[1,2,3]}')
    docs.append(b'["' + b"a\\\"" * 3000 + b'", 1, 2]')
    for _ in range(6):
        docs.append(b"".join(corpus.adversarial(rng, 300) for _ in range(40)))
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, docs, q)) for r in range(world)]
    [p.start() for p in procs]
    got = dict(q.get(timeout=150) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    port_lib = O.Port()
    L = port_lib.L
    L.sjo_scan_shard.restype = C.c_uint64
    L.sjo_scan_shard.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32)]
    rescans = 0
    for k, doc in enumerate(docs):
        a = np.frombuffer(doc, dtype=np.uint8)
        idx = np.zeros(len(a) + 1, dtype=np.uint32)
        n = L.sjo_scan_shard(a.ctypes.data, len(a), 0, idx.ctypes.data, None)
        whole = idx[:n].astype(np.int64).tolist()
        r0, r1 = got[0][k], got[1][k]
        assert r0[0] == 0 and r1[0] == r0[1], "index base of shard 1 = count of shard 0"
        assert r0[5] + r1[5] == whole, k
        assert (r0[4] & 1) == (0 if port_lib.validate_utf8(a) else 1)
        rescans += int(r1[3])
    assert rescans >= 2  # the adversarial documents really exercised the wrong-speculation path
