"""CPU check of the stage-2-lite per-token functions (simdjson_b200/csrc/sjb200_tokens.cuh, compiled for the host by
tests/tokens_emul.cpp in the kernels' tile decomposition) against the oracle.  The GPU run of the kernels themselves is
tests/test_gpu_parity.py::test_tokens_*."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import token_fuzz as TF
from simdjson_b200 import corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("tokemu") / "libtokemu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "simdjson_b200", "csrc"),
                           os.path.join(ROOT, "tests", "tokens_emul.cpp"), "-o", so])
    L = C.CDLL(so)
    L.emu_tokens.restype = C.c_int
    L.emu_tokens.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64),
                             C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    return L


def run_emu(L, doc, idx, n, cap=None):
    a = np.frombuffer(bytes(doc), dtype=np.uint8)
    ix = np.ascontiguousarray(idx[: max(n, 1)], dtype=np.uint32)
    types = np.zeros(max(n, 1), dtype=np.uint8)
    pay = np.zeros(max(n, 1), dtype=np.uint64)
    cap = (len(a) + 5 * n + 64) if cap is None else cap
    sb = np.zeros(max(cap, 1), dtype=np.uint8)
    sl, ns, fe = C.c_uint64(0), C.c_uint32(0), C.c_uint32(0)
    err = L.emu_tokens(a.ctypes.data, len(a), ix.ctypes.data, n, types.ctypes.data, pay.ctypes.data, sb.ctypes.data, cap, C.byref(sl), C.byref(ns), C.byref(fe))
    return err, types[:n], pay[:n], sb[: min(sl.value, cap)], sl.value, ns.value, fe.value


def same(a, b):
    return a[0] == b[0] and bytes(a[1]) == bytes(b[1]) and np.array_equal(a[2], b[2]) and bytes(a[3]) == bytes(b[3]) and a[4:] == b[4:]


def test_token_functions_match_oracle(emu):
    port = O.Port()
    rng = random.Random(99)
    docs = [bytes(corpus.random_json(rng.randrange(300, 200000), seed=4000 + i)) for i in range(10)]
    docs += [open(os.path.join(O.JSONEXAMPLES, f), "rb").read() for f in ("twitter.json", "citm_catalog.json")] if os.path.isdir(O.JSONEXAMPLES) else []
    # documents made of adversarial tokens: every scalar kind and string body the fuzzers produce, errors included
    for _ in range(60):
        parts = []
        for _ in range(rng.randrange(1, 400)):
            if rng.random() < 0.5:
                body, _bad = TF.string_body(rng)
                parts.append(b'"' + body + b'"')
            else:
                tok = TF.scalar_token(rng)
                if b'"' in tok or b"\\" in tok:
                    continue
                parts.append(tok)
        docs.append(b"[" + rng.choice([b",", b" ,\n ", b", "]).join(parts) + b"]")
    for d in docs:
        r = port.stage1(d)
        assert r.err == 0
        want = port.tokens(d, r.idx, r.n)
        got = run_emu(emu, d, r.idx, r.n)
        assert same(got, want), d[:100]
    # tokens cut off by the end of the input, capacity
    for d in (b'["abc', b'[12', b'[tru', b'["\\u12', b'["\\ud800\\u', b"[-", b"[1e"):
        r = port.stage1(d, mode=O.STREAMING_PARTIAL)
        ix = np.array([1], dtype=np.uint32)
        assert same(run_emu(emu, d, ix, 1), port.tokens(d, ix, 1)), d
    d = docs[0]
    r = port.stage1(d)
    assert same(run_emu(emu, d, r.idx, r.n, cap=16), port.tokens(d, r.idx, r.n, strbuf_cap=16))
