"""Multi-GPU sharding of one stage-1 scan: the host-side protocol around
`sjb200_stage1_shard_dev`.

Every rank scans its byte range with a *speculated* incoming scanner state (0 = outside a string, no pending
escape, previous byte not a scalar).  A shard's 6-bit carry transducer does not depend on the incoming state, so
one all-gather of {transducer, count, flags} per rank tells every rank everybody's true incoming state; a rank
whose speculation was wrong scans again with the true state and a second all-gather republishes the counts.
Indexes stay shard-relative (uint32) plus a 64-bit base, like document_stream's batch_start + structural_indexes[i]
(include/simdjson/dom/document_stream-inl.h L250).
"""
import ctypes as C

import numpy as np

from .implementation import lib as _lib


def fold_states(ttables):
    """true incoming state of every shard from the shards' transducers (document starts in state 0)"""
    L = _lib()
    arr = (C.c_uint32 * len(ttables))(*[int(t) for t in ttables])
    return [int(L.sjb200_fold_state(arr, r)) for r in range(len(ttables))]


def shard_cuts(buf, nshards):
    """cut a host buffer into nshards byte ranges at UTF-8 character boundaries"""
    L = _lib()
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    cuts = [0]
    for k in range(1, nshards):
        cuts.append(int(L.sjb200_shard_cut(a.ctypes.data, len(a), (len(a) * k) // nshards)))
    cuts.append(len(a))
    return cuts


def exchange(scan, rank, world, all_gather):
    """Run the protocol on one rank.
      scan(state_in) -> (ttable, count, flags)   scans this rank's shard (GPU: sjb200_stage1_shard_dev)
      all_gather(int64[4]) -> int64[world][4]     the collective (NCCL / gloo all_gather of 32 bytes per rank)
    Returns dict(state_in, base, count, flags, rescanned, ttables)."""
    tt, count, flags = scan(0)
    g = all_gather(np.array([tt, count, flags, 0], dtype=np.int64))
    states = fold_states([int(x) for x in g[:, 0]])
    rescanned = False
    if states[rank] != 0:
        tt2, count, flags = scan(states[rank])
        assert tt2 == tt, "a shard's transducer cannot depend on its incoming state"
        rescanned = True
    if any(s != 0 for s in states):  # somebody's count changed: publish the corrected counts
        g = all_gather(np.array([tt, count, flags, 0], dtype=np.int64))
    return {"state_in": states[rank], "base": int(g[:rank, 1].sum()), "count": int(count), "flags": int(np.bitwise_or.reduce(g[:, 2])),
            "rescanned": rescanned, "ttables": [int(x) for x in g[:, 0]], "total": int(g[:, 1].sum())}


def unpack_result(row):
    """int64[3] written by sjb200_stage1_shard_dev_enqueue -> (ttable, count, flags, state_out)"""
    count, st_tt, fl = int(row[0]), int(row[1]), int(row[2])
    return (st_tt >> 32) & 0xFFFFFFFF, count, fl & 0xFFFFFFFF, st_tt & 0xFFFFFFFF


def verify_speculation(gathered, rank):
    """gathered: int64[world][3] from the all-gather of one speculative pass.  Returns (ok_for_everyone, my_state_in, my_base)."""
    rows = [unpack_result(r) for r in gathered]
    states = fold_states([r[0] for r in rows])
    base = sum(r[1] for r in rows[:rank])
    return all(s == 0 for s in states), states[rank], base
