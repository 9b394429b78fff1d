"""bench.py builds a rank's shard of the N x 64 MiB document from the 8 MiB pieces around it only (make_shard); it must be
byte-identical to cutting the whole document (make_stream + sharding.shard_cuts_at_lines).  CPU only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_windowed_shards_equal_the_cut_document():
    import bench as B
    from simdjson_b200 import sharding
    world, k = 2, 1
    doc = B.make_stream(world, k)
    cuts = sharding.shard_cuts_at_lines(doc, world)
    assert cuts[0] == 0 and cuts[-1] == len(doc) == world * B.DOC_BYTES
    for r in range(world):
        assert np.array_equal(B.make_shard(world, k, r), doc[cuts[r]: cuts[r + 1]]), r
    # arbitrary ranges, incl. the separators, the head and the padded tail
    n, end_pieces, total = B._stream_layout(world)
    for lo, hi in ((0, 5), (B.STREAM_PIECE - 3, B.STREAM_PIECE + 9), (end_pieces - 7, end_pieces + 11), (total - 9, total), (12345, 12345)):
        assert np.array_equal(B.stream_range(world, k, lo, hi), doc[lo:hi]), (lo, hi)
