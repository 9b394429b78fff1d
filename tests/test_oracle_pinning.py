"""Pins the CPU restatement (oracle/sj_oracle.c) before anything trusts it.

 * against the committed golden vectors produced by the unmodified reference
   (tests/golden/*.json, generator oracle/gen_golden.py) -- runs everywhere;
 * against the live reference (oracle/_ref/libsj_ref.so: icelake AND haswell)
   on seeded adversarial inputs, all seven stage1 modes -- runs wherever the
   prebuilt reference .so is present (it ships to the GPU box too).
"""
import hashlib
import json
import os
import random

import numpy as np
import pytest

import oracle_lib as O
from simdjson_b200 import corpus

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def port():
    return O.Port()


def _load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def test_port_matches_golden_stage1(port):
    cases = _load("stage1.json")["cases"]
    assert len(cases) > 500
    for c in cases:
        r = port.stage1(bytes.fromhex(c["hex"]), c["mode"])
        assert r.err == c["err"], c
        if c["n"] is None:
            assert not r.wrote, c
        else:
            assert r.n == c["n"], c
            assert [int(x) for x in r.words()] == c["words"], c


def test_port_matches_golden_minify(port):
    cases = _load("minify.json")["cases"]
    assert len(cases) > 400
    for c in cases:
        err, out = port.minify(bytes.fromhex(c["hex"]))
        assert (err, out.hex()) == (c["err"], c["out"]), c


def test_port_matches_golden_utf8(port):
    cases = _load("utf8.json")["cases"]
    assert len(cases) > 300
    for c in cases:
        assert port.validate_utf8(bytes.fromhex(c["hex"])) == c["valid"], c


@pytest.mark.skipif(not os.path.isdir(O.JSONEXAMPLES), reason="reference corpora not staged (make -C oracle)")
def test_port_matches_golden_corpora(port):
    for f in _load("corpora.json")["files"]:
        data = np.fromfile(os.path.join(O.JSONEXAMPLES, f["file"]), dtype=np.uint8)
        assert len(data) == f["len"]
        r = port.stage1(data, f["mode"])
        assert (r.err, r.n) == (f["err"], f["n"])
        assert hashlib.sha256(r.words().tobytes()).hexdigest() == f["idx_sha256"]
        err, out = port.minify(data)
        assert (err, len(out), hashlib.sha256(out).hexdigest()) == (f["minify_err"], f["minify_len"], f["minify_sha256"])
        assert port.validate_utf8(data) == f["utf8"]


def test_amazon_is_793_documents(port):
    """tests/dom/basictests.cpp L36, L692-708: parse_many over amazon_cellphones.ndjson yields 793 docs;
    at stage-1 level: 793 top-level '[' ... ']' rows in streaming_final mode."""
    path = os.path.join(O.JSONEXAMPLES, "amazon_cellphones.ndjson")
    if not os.path.exists(path):
        pytest.skip("reference corpora not staged")
    data = np.fromfile(path, dtype=np.uint8)
    r = port.stage1(data, O.STREAMING_FINAL)
    assert r.err == 0 and r.n == 15067
    # count row starts: a '[' structural that is the first structural on its line
    starts = 0
    idx = r.idx[: r.n]
    for k, pos in enumerate(idx):
        if data[pos] == ord("[") and (k == 0 or data[idx[k - 1]] == ord("]")):
            starts += 1
    assert starts == 793


needs_ref = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref/libsj_ref.so not built")


@needs_ref
@pytest.mark.parametrize("impl", ["icelake", "haswell"])
def test_port_matches_live_reference_fuzz(port, impl):
    if impl not in O.ref_impls():
        pytest.skip(f"{impl} not supported by this host")
    ref = O.Ref(impl)
    rng = random.Random(corpus.SEED + (1 if impl == "haswell" else 0))
    n_cases = 4000
    for i in range(n_cases):
        kind = i % 4
        if kind == 0:
            b, mode = corpus.adversarial(rng), rng.choice(O.ALL_MODES)
        elif kind == 1:
            b, mode = corpus.multi_document(rng), rng.choice([1, 2])
        elif kind == 2:
            b, mode = b"\x1e" + corpus.multi_document(rng, sep=b"\x1e"), rng.choice([3, 4])
        else:
            b, mode = corpus.multi_document(rng, sep=b","), rng.choice([5, 6])
        a, r = port.stage1(b, mode), ref.stage1(b, mode)
        assert O.same_stage1(a, r), (b, mode, a.err, r.err, a.n, r.n)
        if kind == 0:
            assert port.minify(b) == ref.minify(b), b
            assert port.validate_utf8(b) == ref.validate_utf8(b), b


@needs_ref
def test_port_matches_live_reference_capacity_and_sizes(port):
    ref = O.Ref("")
    doc = b'{"a":[1,2,3],"b":"xyz"}'
    # len > capacity -> CAPACITY before anything is written (json_structural_indexer.h L195)
    a, r = port.stage1(doc, 0, capacity=8), ref.stage1(doc, 0, capacity=8)
    assert a.err == r.err == O.CAPACITY and not a.wrote and not r.wrote
    # block-boundary lengths
    rng = random.Random(5)
    for n in list(range(0, 200)) + [255, 256, 257, 4095, 4096, 4097]:
        b = bytes(corpus.random_json(max(n, 64)))[:n]
        for mode in (0, 1, 2):
            assert O.same_stage1(port.stage1(b, mode), ref.stage1(b, mode)), (n, mode)
        assert port.minify(b) == ref.minify(b)
        u = bytes(corpus.random_utf8(max(n, 1), seed=n))[:n]
        assert port.validate_utf8(u) == ref.validate_utf8(u), n
        _ = rng


@needs_ref
def test_port_matches_live_reference_large(port):
    ref = O.Ref("")
    d = corpus.random_json(4 << 20)
    assert O.same_stage1(port.stage1(d), ref.stage1(d))
    assert port.minify(d) == ref.minify(d)
    nd = corpus.ndjson_rows(2 << 20)
    for mode in (1, 2):
        assert O.same_stage1(port.stage1(nd, mode), ref.stage1(nd, mode))
    # cut an NDJSON buffer mid-row: streaming_partial must point at the last complete row
    cut = nd[: (1 << 20) + 123]
    a, r = port.stage1(cut, 1), ref.stage1(cut, 1)
    assert O.same_stage1(a, r) and a.err == 0 and a.idx[a.n] < len(cut)
    u = corpus.random_utf8(1 << 20)
    assert port.validate_utf8(u) and ref.validate_utf8(u)
    for pos in (0, len(u) // 2, len(u) - 2):
        v = u.copy()
        v[pos] = 0xFF
        assert not port.validate_utf8(v) and not ref.validate_utf8(v)
