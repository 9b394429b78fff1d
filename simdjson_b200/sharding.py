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


def shard_cuts_at_lines(buf, nshards, window=1 << 20):
    """the same, preferring cuts right after a raw line feed (a raw 0x0A cannot lie inside a JSON string, so for valid
    input every shard starts in state 0 and no rank ever scans twice)"""
    L = _lib()
    a = np.ascontiguousarray(buf, dtype=np.uint8)
    cuts = [0]
    for k in range(1, nshards):
        cuts.append(max(cuts[-1], int(L.sjb200_shard_cut_line(a.ctypes.data, len(a), (len(a) * k) // nshards, window))))
    cuts.append(len(a))
    return cuts


class Comm:
    """sjb200_comm: the per-rank object of the sharded scan with the exchange fused into the scan kernel
    (include/sjb200.h).  connect() maps the peers' exchange windows: pass `all_gather_bytes`, a callable that all-gathers
    a 64-byte uint8 array over the job's process group (torch.distributed over NCCL or gloo) -- the only collective of
    the path, once per job -- or, for ranks living in one process, connect_local()."""

    def __init__(self, parser, rank, world):
        from . import capi
        self._capi = capi
        self.parser, self.rank, self.world = parser, rank, world
        self._h = C.c_void_p()
        rc = _lib().sjb200_comm_create(parser._ctx, rank, world, C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"sjb200_comm_create failed ({rc}): " + parser.last_cuda_error())

    def handle(self):
        h = np.zeros(self._capi.COMM_HANDLE_BYTES, dtype=np.uint8)
        rc = _lib().sjb200_comm_get_handle(self._h, h.ctypes.data)
        if rc != 0:
            raise RuntimeError("sjb200_comm_get_handle failed: " + self.parser.last_cuda_error())
        return h

    def connect(self, all_gather_bytes):
        if self.world == 1:
            return
        handles = np.ascontiguousarray(all_gather_bytes(self.handle()), dtype=np.uint8).reshape(self.world, self._capi.COMM_HANDLE_BYTES)
        rc = _lib().sjb200_comm_connect(self._h, handles.ctypes.data)
        if rc != 0:
            raise RuntimeError("sjb200_comm_connect failed: " + self.parser.last_cuda_error())

    @staticmethod
    def connect_local(comms):
        arr = (C.c_void_p * len(comms))(*[c._h for c in comms])
        for c in comms:
            rc = _lib().sjb200_comm_connect_local(c._h, arr)
            if rc != 0:
                raise RuntimeError("sjb200_comm_connect_local failed: " + c.parser.last_cuda_error())

    def enqueue(self, d_shard, d_idx, last_shard, stream=None):
        from .implementation import _stream_ptr
        return _lib().sjb200_stage1_sharded_enqueue(self._h, d_shard.data_ptr(), d_shard.numel(), int(last_shard), d_idx.data_ptr(), _stream_ptr(stream))

    def finish(self):
        res = self._capi.ShardedResult()
        rc = _lib().sjb200_stage1_sharded_finish(self._h, C.byref(res))
        return rc, res

    def scan(self, d_shard, d_idx, last_shard, stream=None):
        rc = self.enqueue(d_shard, d_idx, last_shard, stream)
        if rc != 0:
            return rc, None
        return self.finish()

    def close(self):
        if self._h:
            _lib().sjb200_comm_destroy(self._h)
            self._h = C.c_void_p()


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
