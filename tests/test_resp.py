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
    out, consumed, _ = process_pipeline(buf, engine or OracleEngine(), NOW)
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
           + _cmd("THROTTLE", None, "10", "100", "60") + _throttle("", 10, 100, 60) + _cmd("QUIT")
           + _throttle("large_quantity_key", 10, 100, 60, 15) + _throttle("zero_quantity_key", 10, 100, 60, 0)
           + _cmd("THROTTLE", "ik", 10, 100, 60) + _cmd("throttle", "ik", "10", "100", "60")
           + _cmd("THROTTLE", "neg", "10", "100", "60", "-1"))
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
    assert v[10] == SimpleString("OK")
    assert v[11][1][0] == Integer(0) and v[11][1][2] == Integer(10)                            # :384-395
    assert v[12][1][0] == Integer(1) and v[12][1][2] == Integer(10)                            # :491-502
    assert v[13][1][2] == Integer(9) and v[14][1][2] == Integer(8)                             # integer args, lower case
    assert v[15] == Error("ERR Rate limit check failed: negative quantity: -1")


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
