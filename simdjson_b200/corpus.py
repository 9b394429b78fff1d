"""Seeded synthetic corpora for the parity tests and bench.py.

Nothing here touches the GPU or the oracle.  Every generator is deterministic
in its seed; the default seed is the reference's own fuzz seed 0x5eed1234
(tests/document_stream_fuzz_test_common.h L13).
"""
import random

import numpy as np

SEED = 0x5EED1234

# ---------------------------------------------------------------------------
# adversarial byte soups for stage-1 / minify parity (mirrors the alphabet the
# survey used to validate the scalar specification, Appendix A)
# ---------------------------------------------------------------------------
_ALPHABETS = [
    b'\\\\\\"" {}[],: \n\tabc1\x01\x0c\x1a\x1e',
    b'\\"',
    b'\\\\\\\\\\\\\\"a ',
    b'"{}[],:0 ',
    b' \n\r\t"a\\',
    b'\x1e{"a":1} \n\x1e[1,2]"x"',
    b',{}[] 1 "a":\n',
]


def adversarial(rng, max_len=700):
    """one short byte string rich in backslashes, quotes, operators and controls"""
    n = rng.randint(0, max_len)
    alpha = rng.choice(_ALPHABETS)
    kind = rng.random()
    if kind < 0.15:  # long backslash runs crossing 32/64/128-byte boundaries
        run = rng.choice([31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257])
        s = bytearray(rng.choice(alpha) for _ in range(rng.randint(0, 80)))
        s += b"\\" * (run + rng.randint(0, 1))
        s += bytes(rng.choice(alpha) for _ in range(rng.randint(0, 200)))
        return bytes(s)
    if kind < 0.25:  # sprinkle non-ASCII / truncated UTF-8
        s = bytearray(rng.choice(alpha) for _ in range(n))
        for _ in range(rng.randint(1, 6)):
            if not s:
                break
            pos = rng.randrange(len(s))
            s[pos:pos] = rng.choice([b"\xc3\xa9", b"\xe2\x82\xac", b"\xf0\x9f\x98\x80", b"\xff", b"\xc3", b"\xe2\x82", b"\xf0\x9f\x98", b"\x80", b"\xed\xa0\x80", b"\xc0\xaf", b"\xf4\x90\x80\x80"])
        return bytes(s)
    return bytes(rng.choice(alpha) for _ in range(n))


_FRAGMENTS = [
    b'{"a":1}', b'[1,2,3]', b'{"k":[true,false,null]}', b'"str"', b"123", b"true", b'{"x":"y\\"z"}', b'[[],{}]',
    b'{"u":"\xc3\xa9\xe2\x82\xac"}', b'{"deep":{"a":[1,{"b":2}]}}', b"null", b"-1.5e3", b'{"a":"\\\\"}',
]
_BROKEN = [b'{"a":', b"[1,2", b'{"k":[tr', b'"unterminated', b'{"x":"y\\', b"[[", b'{"u":"\xe2\x82', b'{"u":"\xf0\x9f', b"]", b"}", b","]


def multi_document(rng, sep=b" ", allow_broken=True):
    """a stream of complete documents with an optionally truncated last one"""
    parts = [rng.choice(_FRAGMENTS) for _ in range(rng.randint(0, 12))]
    if allow_broken and rng.random() < 0.6:
        parts.append(rng.choice(_BROKEN))
    seps = [sep, sep * 2, b"\n", b" \n ", b""] if sep == b" " else [sep, sep + b" ", b" " + sep, sep + b"\n"]
    out = bytearray()
    if rng.random() < 0.3:
        out += rng.choice(seps)
    for p in parts:
        out += p
        out += rng.choice(seps)
    if rng.random() < 0.5:
        out = out.rstrip()
    return bytes(out)


# ---------------------------------------------------------------------------
# RandomUTF8-style text (reference: tests/unicode_tests.cpp L7-101)
# ---------------------------------------------------------------------------
def random_utf8(nbytes, seed=SEED, weights=(70, 15, 10, 5)):
    """valid UTF-8 of exactly nbytes bytes; code-point lengths drawn 70/15/10/5 %"""
    rng = np.random.default_rng(seed)
    n_cp = nbytes  # upper bound on code points
    kinds = rng.choice(4, size=n_cp, p=np.array(weights) / 100.0)
    lens = kinds + 1
    csum = np.cumsum(lens)
    k = int(np.searchsorted(csum, nbytes, side="right"))
    kinds, lens, csum = kinds[:k], lens[:k], csum[:k]
    total = int(csum[-1]) if k else 0
    out = np.full(nbytes, 0x20, dtype=np.uint8)  # tail padded with spaces
    starts = csum - lens
    r = rng.integers(0, 1 << 30, size=k)
    # 1 byte: 0x20..0x7E
    m = kinds == 0
    out[starts[m]] = 0x20 + (r[m] % 0x5F)
    # 2 bytes: U+0080..U+07FF
    m = kinds == 1
    cp = 0x80 + (r[m] % (0x800 - 0x80))
    out[starts[m]] = 0xC0 | (cp >> 6)
    out[starts[m] + 1] = 0x80 | (cp & 0x3F)
    # 3 bytes: U+0800..U+FFFF minus surrogates
    m = kinds == 2
    cp = 0x800 + (r[m] % (0x10000 - 0x800 - 0x800))
    cp = np.where(cp >= 0xD800, cp + 0x800, cp)
    out[starts[m]] = 0xE0 | (cp >> 12)
    out[starts[m] + 1] = 0x80 | ((cp >> 6) & 0x3F)
    out[starts[m] + 2] = 0x80 | (cp & 0x3F)
    # 4 bytes: U+10000..U+10FFFF
    m = kinds == 3
    cp = 0x10000 + (r[m] % (0x110000 - 0x10000))
    out[starts[m]] = 0xF0 | (cp >> 18)
    out[starts[m] + 1] = 0x80 | ((cp >> 12) & 0x3F)
    out[starts[m] + 2] = 0x80 | ((cp >> 6) & 0x3F)
    out[starts[m] + 3] = 0x80 | (cp & 0x3F)
    assert total <= nbytes
    return out


# ---------------------------------------------------------------------------
# random-structure JSON (config 2 of BASELINE.json)
# ---------------------------------------------------------------------------
_KEYCH = "abcdefghijklmnopqrstuvwxyz_ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789"
_STRCH = _KEYCH + "      .,;:!?-+/()#@*&%$"
_MB = ["é", "ü", "α", "€", "中", "文", "ツ", "\U0001f600", "\U0001f4a9", "ß"]
_ESC = ['\\"', "\\\\", "\\n", "\\/", "\\t", "\\u00e9"]


def _rand_string(rng, utf8_rate, max_len=64):
    n = rng.randint(0, max_len)
    out = []
    for _ in range(n):
        x = rng.random()
        if x < 0.02:
            out.append(rng.choice(_ESC))
        elif x < 0.02 + utf8_rate:
            out.append(rng.choice(_MB))
        else:
            out.append(rng.choice(_STRCH))
    return '"' + "".join(out) + '"'


def _rand_number(rng):
    k = rng.random()
    if k < 0.5:
        return str(rng.randint(-100000, 1000000))
    if k < 0.85:
        return "%.*f" % (rng.randint(1, 8), rng.uniform(-1e4, 1e4))
    return "%.*e" % (rng.randint(1, 6), rng.uniform(-1e4, 1e4))


def _rand_value(rng, depth, pretty, indent, utf8_rate, out, budget):
    """append one random JSON value to the list `out`; budget = [remaining node count]
    bounds the subtree (depth 8 x fan-out 16 would otherwise explode)"""
    budget[0] -= 1
    if depth <= 0 or budget[0] <= 0 or rng.random() < 0.35:
        k = rng.random()
        if k < 0.45:
            out.append(_rand_string(rng, utf8_rate))
        elif k < 0.8:
            out.append(_rand_number(rng))
        else:
            out.append(rng.choice(["true", "false", "null"]))
        return
    # whitespace style switches per subtree between minified and 2-space pretty-print
    if rng.random() < 0.25:
        pretty = not pretty
    fan = rng.randint(1, 16)
    is_obj = rng.random() < 0.6
    nl = ("\n" + " " * (indent + 2)) if pretty else ""
    nl_end = ("\n" + " " * indent) if pretty else ""
    out.append("{" if is_obj else "[")
    for i in range(fan):
        if i:
            out.append(",")
        out.append(nl)
        if is_obj:
            out.append('"' + "".join(rng.choice(_KEYCH) for _ in range(rng.randint(3, 12))) + '"')
            out.append(": " if pretty else ":")
        _rand_value(rng, depth - 1, pretty, indent + 2, utf8_rate, out, budget)
    out.append(nl_end)
    out.append("}" if is_obj else "]")


def random_json_piece(seed, target_bytes, utf8_rate=0.05, pretty_bias=0.5):
    """a comma-joined run of random values, about target_bytes long (no enclosing brackets)"""
    rng = random.Random(seed)
    out, size = [], 0
    first = True
    while size < target_bytes:
        chunk = []
        if not first:
            chunk.append(",")
        first = False
        pretty = rng.random() < pretty_bias
        if pretty:
            chunk.append("\n  ")
        _rand_value(rng, 8, pretty, 2, utf8_rate, chunk, [rng.randint(20, 600)])
        s = "".join(chunk)
        out.append(s)
        size += len(s.encode("utf-8"))
    return "".join(out).encode("utf-8")


def random_json(nbytes, seed=SEED, pool=48, piece_bytes=96 * 1024, utf8_rate=0.05, pretty_bias=0.5):
    """one valid JSON document of exactly nbytes bytes: '[' + random subtrees + ',"pad…"]'.

    A pool of distinct random subtree runs is generated (python recursion, a
    few MB) and the document is assembled from a seeded random sequence of
    pool entries, so 64 MiB+ documents take seconds rather than minutes while
    the structure still varies along the whole buffer."""
    assert nbytes >= 64
    small = nbytes < pool * piece_bytes
    if small:
        piece_bytes = max(256, nbytes // 8)
        pool = 6
    pieces = [random_json_piece(seed + 1000003 * i, piece_bytes, utf8_rate, pretty_bias) for i in range(pool)]
    rng = random.Random(seed ^ 0xABCDEF)
    parts, size = [b"["], 1
    tail_min = len(b',"pad":""]') + 0  # we close with ,"<padding>"]
    while True:
        p = pieces[rng.randrange(pool)]
        extra = len(p) + (1 if len(parts) > 1 else 0)
        if size + extra + 4 > nbytes:
            break
        if len(parts) > 1:
            parts.append(b",")
        parts.append(p)
        size += extra
    # close with a string padded to the exact length:  ,"xxxx"]   (or  "xxxx"] if nothing fit)
    lead = b"," if len(parts) > 1 else b""
    pad = nbytes - size - len(lead) - 3
    assert pad >= 0, (nbytes, size)
    parts.append(lead + b'"' + b"x" * pad + b'"]')
    doc = b"".join(parts)
    assert len(doc) == nbytes, (len(doc), nbytes)
    return np.frombuffer(doc, dtype=np.uint8)


# ---------------------------------------------------------------------------
# amazon_cellphones-style NDJSON (config 3)
# ---------------------------------------------------------------------------
def ndjson_rows(nbytes, seed=SEED, base_rows=None):
    """NDJSON of exactly nbytes bytes.  Rows are array-of-9 records in the style of
    jsonexamples/amazon_cellphones.ndjson; when base_rows (list of bytes, no newline)
    is given they are the seed rows, otherwise rows are synthesised.  A pool of
    deterministically mutated rows is tiled to the requested size; the final row is
    padded inside a string so the buffer ends with '\\n' at exactly nbytes."""
    rng = random.Random(seed)
    if not base_rows:
        base_rows = []
        for _ in range(793):
            asin = "B%09d" % rng.randrange(10**9)
            brand = rng.choice(["Motorola", "Samsung", "Nokia", "Sony", "Apple", "Google", "HUAWEI", "ASUS", "OnePlus", "Xiaomi"])
            title = "".join(rng.choice(_STRCH) for _ in range(rng.randint(20, 120))).strip()
            url = "https://www.amazon.com/%s/dp/%s" % (title.replace(" ", "-")[:40].replace("/", ""), asin)
            img = "https://m.media-amazon.com/images/I/%s._AC_UY218_.jpg" % "".join(rng.choice(_KEYCH) for _ in range(11))
            row = '["%s","%s","%s","%s","%s",%.1f,"%s",%d,%.2f,%.1f]' % (
                asin, brand, title.replace("\\", "").replace('"', '\\"'), url, img, rng.uniform(1, 5),
                "https://www.amazon.com/product-reviews/" + asin, rng.randint(1, 3000), rng.uniform(0, 999), rng.uniform(0, 999))
            base_rows.append(row.encode())
    pool = []
    for i in range(4096):
        r = bytearray(base_rows[i % len(base_rows)])
        # deterministic mutation: rewrite digits in place (keeps the row valid JSON)
        for _ in range(6):
            pos = rng.randrange(len(r))
            if 0x30 <= r[pos] <= 0x39:
                r[pos] = 0x30 + rng.randrange(10)
        pool.append(bytes(r) + b"\n")
    order = np.random.default_rng(seed).integers(0, len(pool), size=nbytes // 64 + 16)
    parts, size = [], 0
    for j in order:
        p = pool[int(j)]
        if size + len(p) + 16 > nbytes:
            break
        parts.append(p)
        size += len(p)
    pad = nbytes - size - len(b'["",0]\n')
    assert pad >= 0
    parts.append(b'["' + b"p" * pad + b'",0]\n')
    doc = b"".join(parts)
    assert len(doc) == nbytes
    return np.frombuffer(doc, dtype=np.uint8)


def tile_documents(docs, nbytes, sep=b"\n"):
    """config 5 style: the given documents repeated (each followed by sep) up to at most nbytes;
    returns the buffer (length <= nbytes, cut at a document boundary)"""
    unit = b"".join(d + sep for d in docs)
    reps = max(1, nbytes // len(unit))
    return np.frombuffer(unit * reps, dtype=np.uint8)
