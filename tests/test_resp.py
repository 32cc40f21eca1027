"""RESP ingest mirror (throttlecrab_b200/resp.py) against the reference's RESP tests: parser/serializer
(redis/resp.rs:234-317, transport/redis_test.rs:166-260,505-554) and security limits
(transport/redis_security_test.rs), plus the command semantics of process_command (redis_test.rs) driven through
the BATCHED pipeline.  On the CPU the engine is replaced by the oracle (checker's stand-in, tests only); the GPU
test at the bottom runs the same pipeline against the real engine."""
import numpy as np
import pytest

import oracle
import throttlecrab_b200 as tc
from throttlecrab_b200 import resp
from throttlecrab_b200.resp import (Array, BulkString, Error, Integer, RespError, RespParser, RespSerializer,
                                    SimpleString, process_pipeline)

NOW = 1_700_000_000 * 10**9


def test_parse_basic_values():                          # resp.rs:238-301
    p = RespParser()
    assert p.parse(b"+OK\r\n") == (SimpleString("OK"), 5)
    assert p.parse(b"-ERR unknown command\r\n") == (Error("ERR unknown command"), 22)
    assert p.parse(b":42\r\n") == (Integer(42), 5)
    assert p.parse(b"$6\r\nfoobar\r\n") == (BulkString("foobar"), 12)
    assert p.parse(b"$-1\r\n") == (BulkString(None), 5)
    assert p.parse(b"*2\r\n$3\r\nfoo\r\n$3\r\nbar\r\n") == (Array([BulkString("foo"), BulkString("bar")]), 22)


def test_serialize():                                    # resp.rs:303-316
    assert RespSerializer.serialize(SimpleString("OK")) == b"+OK\r\n"
    assert RespSerializer.serialize(Array([BulkString("foo"), Integer(42)])) == b"*2\r\n$3\r\nfoo\r\n:42\r\n"


def test_partial_data():                                 # redis_test.rs:166-190
    p = RespParser()
    assert p.parse(b"+OK") is None
    assert p.parse(b"+OK\r\n") == (SimpleString("OK"), 5)
    assert p.parse(b"$6\r\nfoo") is None
    assert p.parse(b"$6\r\nfoobar\r\n") == (BulkString("foobar"), 12)


def test_roundtrip_and_edge_cases():                     # redis_test.rs:193-240
    p = RespParser()
    values = [SimpleString("OK"), Error("ERR something"), Integer(42), BulkString("hello world"), BulkString(None),
              Array([BulkString("foo"), Integer(123), SimpleString("bar")])]
    for v in values:
        ser = RespSerializer.serialize(v)
        assert p.parse(ser) == (v, len(ser))
    assert p.parse(b"*0\r\n") == (Array([]), 4)
    nested = Array([Array([Integer(1), Integer(2)]), BulkString("test")])
    assert p.parse(RespSerializer.serialize(nested))[0] == nested


def test_multiple_commands_and_integer_args():           # redis_test.rs:505-554
    p = RespParser()
    data = b"*1\r\n$4\r\nPING\r\n" + b"*2\r\n$4\r\nPING\r\n$5\r\nhello\r\n"
    cmd1, c1 = p.parse(data)
    assert cmd1 == Array([BulkString("PING")])
    cmd2, c2 = p.parse(data[c1:])
    assert cmd2 == Array([BulkString("PING"), BulkString("hello")]) and c2 == len(data) - c1
    cmd, _ = p.parse(b"*5\r\n$8\r\nTHROTTLE\r\n$8\r\ntest_key\r\n:10\r\n:100\r\n:60\r\n")
    assert cmd == Array([BulkString("THROTTLE"), BulkString("test_key"), Integer(10), Integer(100), Integer(60)])


@pytest.mark.parametrize("bad", [b"$999999999999999999999\r\n", b"*999999999999999999999\r\n", b"$-999999999\r\n",
                                 b"*-999999999\r\n", b"*1\r\n" * 200 + b":42\r\n", b"*%d\r\n" % (2**63 - 1),
                                 b"$4\r\n\xff\xfe\xfd\xfc\r\n", b"*3\r\n:42\r\n$999999999999\r\ntest\r\n:100\r\n",
                                 b"*10000000\r\n", b"?what\r\n"])
def test_security_limits(bad):                           # redis_security_test.rs:8-160
    with pytest.raises(RespError):
        RespParser().parse(bad)


def test_security_tolerated_inputs():                    # redis_security_test.rs:82-128
    p = RespParser()
    try:
        p.parse(b"$%d\r\n" % (2**63 - 1))                 # may need more data or be rejected, must not blow up
    except RespError:
        pass
    v, _ = p.parse(b"$5\r\nhel\x00lo\r\n")
    assert v[0] == "bulk" and len(v[1]) == 5 and v[1][3] == "\x00"   # like the reference: 5 bytes, CRLF not verified


# ---- command semantics through the batched pipeline -------------------------------------------------------
def _cmd(*args):
    return RespSerializer.serialize(Array([BulkString(a) if isinstance(a, str) or a is None else Integer(a)
                                           for a in args]))


def _throttle(key, b, c, p, q=None):
    a = ["THROTTLE", key, str(b), str(c), str(p)] + ([str(q)] if q is not None else [])
    return _cmd(*a)


class OracleEngine:
    """tests only: the oracle standing in for RateLimiter.rate_limit_batch (key column = key hash -> string)"""

    def __init__(self):
        self.st = oracle.OracleStore(oracle.PERIODIC, capacity=1000, created_ns=NOW, p0=10**9)

    def __call__(self, req):
        out = np.zeros(len(req), tc.RES_DTYPE)
        for i, r in enumerate(req):
            s, a, rem, reset, retry = self.st.rate_limit("h%d" % int(r["key_hash"]), int(r["max_burst"]),
                                                         int(r["count_per_period"]), int(r["period"]),
                                                         int(r["quantity"]), int(r["now_ns"]))
            out[i] = (rem, reset, retry, s, a, [0, 0, 0])
        return out


def _replies(buf, engine=None):
    out, consumed, _, _ = process_pipeline(buf, engine or OracleEngine(), NOW)
    p, vals, pos = RespParser(), [], 0
    while pos < len(out):
        v, c = p.parse(out[pos:])
        vals.append(v)
        pos += c
    return vals, consumed


def test_commands_known_answers():
    buf = (_cmd("PING") + _cmd("PING", "hello") + _throttle("test_key", 10, 100, 60)       # redis_test.rs:99-130
           + _throttle("test_key2", 10, 100, 60, 5) + _cmd("UNKNOWN") + _cmd("THROTTLE", "test_key")
           + _cmd("THROTTLE", "test_key", "not_a_number", "100", "60") + _cmd("THROTTLE", "test_key", "-5", "100", "60")
           + _cmd("THROTTLE", None, "10", "100", "60") + _throttle("", 10, 100, 60)
           + _throttle("large_quantity_key", 10, 100, 60, 15) + _throttle("zero_quantity_key", 10, 100, 60, 0)
           + _cmd("THROTTLE", "ik", 10, 100, 60) + _cmd("throttle", "ik", "10", "100", "60")
           + _cmd("THROTTLE", "neg", "10", "100", "60", "-1") + _cmd("QUIT"))
    v, consumed = _replies(buf)
    assert consumed == len(buf)
    assert v[0] == SimpleString("PONG") and v[1] == BulkString("hello")
    assert v[2] == Array([Integer(1), Integer(10), Integer(9), Integer(5), Integer(0)])        # :116-130
    assert v[3] == Array([Integer(1), Integer(10), Integer(5), Integer(7), Integer(0)])        # :132-144
    assert v[4][0] == "error" and "unknown command" in v[4][1]                                 # :147-153
    assert "wrong number of arguments" in v[5][1]                                              # :156-163
    assert "invalid max_burst" in v[6][1]                                                      # :476-484
    assert v[7][0] == "error" and v[7][1].startswith("ERR")                                    # :486-489 (InvalidRateLimit)
    assert "invalid key" in v[8][1]                                                            # :658-676
    assert v[9][1][:3] == [Integer(1), Integer(10), Integer(9)]                                # :633-656 empty key
    assert v[10][1][0] == Integer(0) and v[10][1][2] == Integer(10)                            # :384-395
    assert v[11][1][0] == Integer(1) and v[11][1][2] == Integer(10)                            # :491-502
    assert v[12][1][2] == Integer(9) and v[13][1][2] == Integer(8)                             # integer args, lower case
    assert v[14] == Error("ERR Rate limit check failed: negative quantity: -1")
    assert v[15] == SimpleString("OK")                                                         # QUIT, last


def test_pipeline_keeps_arrival_order_and_partial_tail():
    """exhaustion inside ONE buffer (redis_test.rs:271-304) + an incomplete trailing frame is left unconsumed"""
    one = _throttle("exhaustion_test", 3, 100, 60)
    buf = one * 4 + one[:-7]
    v, consumed = _replies(buf)
    assert consumed == 4 * len(one)
    assert [x[1][0] for x in v] == [Integer(1)] * 3 + [Integer(0)]
    assert [x[1][2] for x in v] == [Integer(2), Integer(1), Integer(0), Integer(0)]
    eng = OracleEngine()
    calls = []
    process_pipeline(buf, lambda r: (calls.append(len(r)), eng(r))[1], NOW)
    assert calls == [4]                                   # ONE engine batch for the whole buffer


@pytest.mark.gpu
def test_pipeline_against_the_engine():
    """redis_test.rs:422-473 (mixed commands) and :720-762 (case-insensitive) through the real engine"""
    lim = tc.RateLimiter(tc.PeriodicStore(capacity=1000, created_ns=NOW, max_batch=4096))
    buf = (_cmd("PING") + _throttle("mixed_key", 10, 100, 60) + _cmd("PING", "test message")
           + _cmd("throttle", "mixed_key", "10", "100", "60") + _cmd("Throttle", "mixed_key", "10", "100", "60")
           + _throttle("boundary_key", 2**63 - 1, 2**63 - 1, 2**63 - 1) + _throttle("tiny_key", 1, 1, 1))
    v, consumed = _replies(buf, lim.rate_limit_batch)
    assert consumed == len(buf)
    assert v[0] == SimpleString("PONG") and v[2] == BulkString("test message")
    assert [v[i][1][2] for i in (1, 3, 4)] == [Integer(9), Integer(8), Integer(7)]
    assert v[5][1][:2] == [Integer(1), Integer(2**63 - 1)]                                     # :677-697
    assert v[6][1][:3] == [Integer(1), Integer(1), Integer(0)]                                 # :699-716


def test_pipeline_error_and_quit_like_the_connection_loop():
    """redis/mod.rs:128-149: commands before a protocol error are applied and answered, then the connection closes;
    nothing after QUIT is executed."""
    one = _throttle("order_key", 2, 100, 60)
    eng = OracleEngine()
    out, consumed, n, close = process_pipeline(one + one + b"?garbage\r\n" + one, eng, NOW)
    assert close == "error" and n == 2 and consumed == 2 * len(one)
    vals, pos, p = [], 0, RespParser()
    while pos < len(out):
        v, c = p.parse(out[pos:])
        vals.append(v)
        pos += c
    assert [x[1][0] for x in vals] == [Integer(1), Integer(1)]
    out, consumed, n, close = process_pipeline(one + _cmd("QUIT") + one, eng, NOW)
    assert close == "quit" and n == 1 and consumed == len(one) + len(_cmd("QUIT"))
    assert out.endswith(b"+OK\r\n")
    v0, _ = RespParser().parse(out)
    assert v0[1][0] == Integer(0)                      # the key was exhausted by the two requests above; the third
    # command (after QUIT) was never applied: a fresh engine would have allowed it


class _FakeStore:
    """tests only (no GPU): what process_pipeline_native needs from a limiter besides the two C helpers"""

    def __init__(self):
        self.eng = OracleEngine()

    def _check(self, rc):
        assert rc == 0

    def hash_seed(self):
        return (0, 0)


class _FakeLimiter:
    def __init__(self):
        from throttlecrab_b200 import _native
        self._L, self._h, self.store = _native.lib(), None, _FakeStore()

    def rate_limit_batch(self, req):
        return self.store.eng(req)


def test_native_batch_parser_matches_the_general_parser():
    """gcra_resp_parse_throttle + gcra_resp_format_replies (csrc/gcra_resp.inc) against the mirror of resp.rs on a
    pipeline that mixes plain THROTTLE frames with everything the fast path must hand back: other commands, RESP
    integer arguments, lower case, quantity, negative numbers, non-numeric arguments, a split tail."""
    from throttlecrab_b200.resp import process_pipeline_native
    frames = [_throttle("k%d" % (i % 7), 3, 100, 60) for i in range(50)]
    frames[5] = _cmd("PING")
    frames[9] = _cmd("throttle", "k1", 3, 100, 60)                     # RESP integers: general parser
    frames[12] = _throttle("k2", 3, 100, 60, 2)
    frames[13] = _throttle("k2", 3, 100, 60, -1)                       # negative quantity -> error reply
    frames[20] = _cmd("THROTTLE", "k3", "abc", "100", "60")            # not a number -> ERR invalid max_burst
    frames[21] = _cmd("tHrOtTlE", "k3", "3", "100", "60")
    frames[30] = _cmd("THROTTLE", "k4", "0", "100", "60")              # invalid rate limit (engine error)
    frames[31] = _cmd("UNKNOWN")
    frames[40] = _throttle("binary\r\nkey", 3, 100, 60)                # CRLF inside a bulk string
    buf = b"".join(frames) + _throttle("tail", 3, 100, 60)[:-5]
    want = process_pipeline(buf, OracleEngine(), NOW)
    got = process_pipeline_native(buf, _FakeLimiter(), NOW)
    assert got == want
    assert got[1] == len(b"".join(frames)) and got[2] == 47 and got[3] is None
    # errors and QUIT exactly as the general pipeline
    bad = frames[0] + b"!x\r\n" + frames[1]
    assert process_pipeline_native(bad, _FakeLimiter(), NOW) == process_pipeline(bad, OracleEngine(), NOW)
    q = frames[0] + _cmd("quit") + frames[1]
    assert process_pipeline_native(q, _FakeLimiter(), NOW) == process_pipeline(q, OracleEngine(), NOW)


def test_native_batch_parser_limits():
    """frames the fast path must not swallow: huge declared lengths, 20-digit numbers, wrong arity"""
    import ctypes as C
    from throttlecrab_b200 import _native
    L = _native.lib()
    req = np.zeros(8, tc.REQ_DTYPE)

    def run(buf):
        used, cnt, stop = C.c_uint64(), C.c_uint32(), C.c_int32()
        assert L.gcra_resp_parse_throttle(None, buf, len(buf), NOW, 8, req.ctypes.data, C.byref(used), C.byref(cnt),
                                          C.byref(stop)) == 0
        return used.value, cnt.value, stop.value
    ok = _throttle("k", 5, 10, 60)
    assert run(ok) == (len(ok), 1, 0)
    assert int(req[0]["key_hash"]) == tc.hash_key("k") and int(req[0]["quantity"]) == 1 and int(req[0]["now_ns"]) == NOW
    assert run(ok + ok[:10]) == (len(ok), 1, 1)                                        # incomplete frame: read more
    assert run(_cmd("THROTTLE", "k", "99999999999999999999", "10", "60"))[1:] == (0, 2)   # 20 digits: general parser
    assert run(b"*5\r\n$8\r\nTHROTTLE\r\n$999999999999\r\n")[1:] == (0, 2)              # over MAX_BULK_STRING_SIZE
    assert run(_cmd("THROTTLE", "k", "5", "10"))[1:] == (0, 2)                           # wrong arity
    assert run(_throttle("k", 9223372036854775807, 1, 1))[1:] == (1, 0)                # i64::MAX fits
    assert run(ok * 9)[1:] == (8, 3)                                                   # max_frames


@pytest.mark.gpu
def test_native_pipeline_through_the_pinned_ring():
    """A pipelined read buffer -> request rows written straight into a pinned ring slot -> one engine batch -> reply
    bytes; equal to the general pipeline over the blocking API on a twin engine."""
    from throttlecrab_b200.resp import process_pipeline_native
    frames = [_throttle("user:%d" % (i % 97), 5, 10, 60) for i in range(3000)]
    frames[100] = _cmd("PING")
    frames[2000] = _cmd("throttle", "user:1", 5, 10, 60)
    buf = b"".join(frames)
    a = tc.RateLimiter(tc.PeriodicStore(capacity=1000, created_ns=NOW, max_batch=4096))
    b = tc.RateLimiter(tc.PeriodicStore(capacity=1000, created_ns=NOW, max_batch=4096))
    ring = tc.Ring(b, slots=2, slot_capacity=4096)
    want = process_pipeline(buf, a.rate_limit_batch, NOW)
    got = process_pipeline_native(buf, b, NOW, ring=ring, slot=1)
    assert got == want and got[2] == 2999


def _raw_throttle(key_bytes, *nums):
    parts = [b"THROTTLE", key_bytes] + [str(x).encode() for x in nums]
    return b"*%d\r\n" % len(parts) + b"".join(b"$%d\r\n%s\r\n" % (len(p), p) for p in parts)


def test_native_batch_parser_rejects_what_from_utf8_rejects():
    """resp.rs:110 turns every bulk string into a String: a key that is not well-formed UTF-8 fails the frame.  The
    fast path hands exactly those frames to the general parser (stop 2) and accepts every well-formed key."""
    import ctypes as C
    from throttlecrab_b200 import _native
    L = _native.lib()
    req = np.zeros(4, tc.REQ_DTYPE)

    def accepted(key):
        buf = _raw_throttle(key, 3, 100, 60)
        used, cnt, stop = C.c_uint64(), C.c_uint32(), C.c_int32()
        assert L.gcra_resp_parse_throttle(None, buf, len(buf), NOW, 4, req.ctypes.data, C.byref(used), C.byref(cnt),
                                          C.byref(stop)) == 0
        assert (cnt.value, stop.value) in ((1, 0), (0, 2))
        return cnt.value == 1
    rng = np.random.default_rng(11)
    cases = [b"", b"plain", "kéy".encode(), "€".encode(), "\U0001f980".encode(), b"\xc0\x80", b"\xc1\xbf",
             b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xe0\xa0\x80", b"\xed\x9f\xbf", b"\xed\xa0\x80", b"\xef\xbf\xbf",
             b"\xf0\x8f\xbf\xbf", b"\xf0\x90\x80\x80", b"\xf4\x8f\xbf\xbf", b"\xf4\x90\x80\x80", b"\xf5\x80\x80\x80",
             b"\x80", b"\xbf", b"ab\xc3", b"ab\xe2\x82", b"\xf0\x9f\xa6", b"\xff", b"a\xc3\xa9b\xe2\x82\xacc"]
    for _ in range(3000):
        n = int(rng.integers(1, 6))
        cases.append(bytes(rng.choice([0x41, 0x7f, 0x80, 0x9f, 0xa0, 0xbf, 0xc1, 0xc2, 0xdf, 0xe0, 0xe1, 0xec, 0xed,
                                       0xee, 0xef, 0xf0, 0xf1, 0xf3, 0xf4, 0xf5, 0x8f, 0x90], n).astype(np.uint8)))
    for key in cases:
        try:
            key.decode("utf-8")
            ok = True
        except UnicodeDecodeError:
            ok = False
        assert accepted(key) == ok, key
    # and the pipeline as a whole answers such a frame like the general path (a parse error closes the connection)
    from throttlecrab_b200.resp import process_pipeline_native
    buf = _raw_throttle(b"good", 3, 100, 60) + _raw_throttle(b"bad\xff", 3, 100, 60) + _raw_throttle(b"after", 3, 100, 60)
    assert process_pipeline_native(buf, _FakeLimiter(), NOW) == process_pipeline(buf, OracleEngine(), NOW)


def test_native_pipeline_fuzz_against_the_general_pipeline():
    """Seeded fuzz: pipelines of well-formed frames, frames with flipped/dropped/inserted bytes and raw garbage, cut
    at a random byte; the native fast path + fallback must answer byte for byte like the general pipeline (replies,
    bytes consumed, commands applied, close reason)."""
    from throttlecrab_b200.resp import process_pipeline_native
    rng = np.random.default_rng(20260923)
    keys = [b"a", b"user:1", "kéy".encode(), b"", b"x" * 70, b"\xff\xfe", b"with\r\ncrlf", b"-1", b"*5", b"$3"]
    nums = [0, 1, 2, 3, 5, 60, 100, -1, -5, 2**31, 2**63 - 1, -2**63, 2**63, "abc", "", "+7", "007", " 5", "5 ", "1e3"]

    def frame():
        kind = int(rng.integers(0, 10))
        if kind <= 5:
            args = [nums[int(rng.integers(0, len(nums)))] if rng.random() < 0.15 else int(rng.integers(1, 200))
                    for _ in range(int(rng.choice([3, 3, 3, 4, 4, 2, 5])))]
            f = _raw_throttle(keys[int(rng.integers(0, len(keys)))], *args)
            if rng.random() < 0.1:
                f = f.replace(b"THROTTLE", rng.choice([b"throttle", b"Throttle", b"THROTTLF", b"THROTTL"]).item(), 1)
            return f
        if kind == 6:
            return _cmd(*[["PING"], ["PING", "hello"], ["QUIT"], ["UNKNOWN"], ["throttle", "k", 3, 100, 60],
                          ["THROTTLE", "k", 3, 100, 60, 2]][int(rng.integers(0, 6))])
        if kind == 7:                                               # a damaged throttle frame
            f = bytearray(_raw_throttle(b"dmg", 3, 100, 60))
            for _ in range(int(rng.integers(1, 3))):
                op, pos = int(rng.integers(0, 3)), int(rng.integers(0, len(f)))
                if op == 0:
                    f[pos] = int(rng.integers(0, 256))
                elif op == 1:
                    del f[pos]
                else:
                    f.insert(pos, int(rng.choice([0x0d, 0x0a, 0x24, 0x2a, 0x3a, 0x2d, 0x30, 0x39, 0xff])))
            return bytes(f)
        if kind == 8:
            return [b"+OK\r\n", b":5\r\n", b"$-1\r\n", b"*0\r\n", b"*-1\r\n", b"$5\r\nab\r\n", b"*5\r\n$8\r\nTHROTTLE\r\n$99999999999\r\n",
                    b"*5\r\n$8\r\nTHROTTLE\r\n$-3\r\n", b"*99999999999\r\n", b"\r\n"][int(rng.integers(0, 10))]
        return bytes(rng.integers(0, 256, int(rng.integers(1, 12))).astype(np.uint8))

    for case in range(400):
        buf = b"".join(frame() for _ in range(int(rng.integers(1, 12))))
        if rng.random() < 0.4 and len(buf) > 1:
            buf = buf[:int(rng.integers(1, len(buf)))]
        want = process_pipeline(buf, OracleEngine(), NOW)
        got = process_pipeline_native(buf, _FakeLimiter(), NOW)
        assert got == want, (case, buf)
