"""Seeded generators of stage-2-lite test inputs (strings with escapes, scalar tokens, small documents): shared by
oracle/gen_golden_tokens.py (golden vectors from the reference), tests/test_oracle_pinning.py and the GPU parity tests."""
import random

ESC_OK = [b'\\"', b"\\\\", b"\\/", b"\\b", b"\\f", b"\\n", b"\\r", b"\\t"]
ESC_BAD = [b"\\a", b"\\x", b"\\U", b"\\0", b"\\ ", b"\\'", b"\\u12", b"\\u12G4", b"\\uD800", b"\\uD800x", b"\\uD800\\n", b"\\uD800\\u0041",
           b"\\uDC00", b"\\uDFFF", b"\\uD800\\uD800", b"\\u+123", b"\\u 123"]
HEX = b"0123456789abcdefABCDEF"


def _u4(rng, lo, hi):
    v = rng.randrange(lo, hi + 1)
    s = "%04x" % v
    return ("\\u" + "".join(c.upper() if rng.random() < 0.5 else c for c in s)).encode()


def string_body(rng, bad_rate=0.15, maxlen=80):
    """bytes between the quotes of a JSON string (no raw quote, backslashes only as part of escapes); returns (body, may_be_bad)"""
    out = bytearray()
    bad = False
    n = rng.choice([0, 1, 2, 3, 5, 8, 13, 21, 31, 32, 33, 40, 63, 64, 65, maxlen])
    while len(out) < n:
        r = rng.random()
        if r < 0.55:
            out.append(rng.choice(b"abcxyz 0123456789_-.,:;{}[]/'"))
        elif r < 0.62:
            out += rng.choice(["é", "€", "😀", "ü", "中"]).encode()
        elif r < 0.80:
            out += rng.choice(ESC_OK)
        elif r < 0.88:
            out += _u4(rng, 0, 0xD7FF)
        elif r < 0.91:
            out += _u4(rng, 0xE000, 0xFFFF)
        elif r < 0.95:
            out += _u4(rng, 0xD800, 0xDBFF) + _u4(rng, 0xDC00, 0xDFFF)
        elif rng.random() < bad_rate * 4:
            out += rng.choice(ESC_BAD)
            bad = True
        else:
            out.append(0x20)
    return bytes(out), bad


NUM_PARTS_INT = ["0", "1", "7", "10", "42", "123", "9007199254740993", "9223372036854775807", "9223372036854775808", "9223372036854775809",
                 "18446744073709551615", "18446744073709551616", "12345678901234567890", "99999999999999999999", "10000000000000000000",
                 "19999999999999999999", "123456789012345678901", "00", "01", "007"]


def scalar_token(rng):
    """one scalar token (valid or not) as bytes: numbers, atoms, near misses"""
    r = rng.random()
    if r < 0.45:
        s = ("-" if rng.random() < 0.4 else "") + (rng.choice(NUM_PARTS_INT) if rng.random() < 0.6 else str(rng.randrange(0, 10 ** rng.randrange(1, 21))))
        if rng.random() < 0.35:
            s += "." + "".join(rng.choice("0123456789") for _ in range(rng.choice([0, 1, 2, 5, 17])))
        if rng.random() < 0.3:
            s += rng.choice("eE") + rng.choice(["", "+", "-"]) + "".join(rng.choice("0123456789") for _ in range(rng.choice([0, 1, 2, 3])))
        if rng.random() < 0.1:
            s += rng.choice(["x", "-", "+", ".", "e", "a", "\"", "\x0c", "0x1"])
        return s.encode()
    if r < 0.75:
        w = rng.choice(["true", "false", "null"])
        m = rng.random()
        if m < 0.5:
            return w.encode()
        if m < 0.6:
            return (w + rng.choice(["x", "1", "e", "\"", "\x0c", "_"])).encode()
        if m < 0.7:
            return w[:-1].encode()
        if m < 0.8:
            return (w[:-1] + rng.choice("xyzEUL")).encode()
        if m < 0.9:
            return w.upper().encode()
        return (w[0] + "".join(rng.choice("aelrstu") for _ in range(len(w) - 1))).encode()
    if r < 0.85:
        return rng.choice([b"-", b"--1", b"-a", b"+1", b".5", b"1.", b"1e", b"1e+", b"-0", b"-0.0", b"0e0", b"1E5", b"1.5e-3", b"0.1e1x"])
    return rng.choice([b"x", b"abc", b"nan", b"NaN", b"Infinity", b"-Infinity", b"tru", b"nul", b"fals", b"'a'", b"\x0c", b"#", b"@1"])


def wrap_scalar(tok, rng):
    """a document with the token in value position and the structural index of the token"""
    style = rng.randrange(4)
    if style == 0:
        return b"[" + tok + b"]", 1
    if style == 1:
        return b"[ " + tok + b" ,1]", 1
    if style == 2:
        return b'{"a":' + tok + b"}", 3
    return b"[0,\n" + tok + b"\n]", 3


def long_body(rng, n, esc_rate):
    """a string body of about n bytes: plain runs (so that 512-byte blocks without specials occur) mixed with escapes at esc_rate"""
    out = bytearray()
    while len(out) < n:
        r = rng.random()
        if r < esc_rate:
            b, _ = string_body(rng, bad_rate=0.0, maxlen=40)
            out += b
        elif r < esc_rate + 0.02:
            out += rng.choice([b"\\\\" * rng.randrange(1, 40), b"\\\\\\\"", b"\\ud83d\\ude00" * rng.randrange(1, 8), b"\\u0041", b"\\n\\t\\\\u0041"])
        else:
            out += bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyz 0123456789,.:;{}[]") for _ in range(rng.choice([1, 7, 31, 32, 33, 100, 511, 512, 513, 700, 1500])))
    return bytes(out)
