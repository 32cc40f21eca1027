"""The CPU oracle against the reference's own known-answer tests (transcribed as data in
tests/golden/).  This is the parity pin for the oracle: the reference (Rust) cannot run here."""
import pytest

import oracle
from golden import ns_vectors, ref_scripts

T0 = ns_vectors.T0
KINDS = [oracle.PERIODIC, oracle.PROBABILISTIC, oracle.ADAPTIVE]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", sorted(ref_scripts.SCENARIOS))
def test_reference_scripts(name, kind):
    st = oracle.OracleStore(kind, capacity=1000, created_ns=T0)
    for (key, b, c, p, q, t, expect) in ref_scripts.SCENARIOS[name]:
        out = st.rate_limit(key, b, c, p, q, T0 + t)
        ref_scripts.check_step(expect, *out)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", sorted(ns_vectors.VECTORS))
def test_ns_vectors(name, kind):
    st = oracle.OracleStore(kind, capacity=1000, created_ns=T0)
    for row in ns_vectors.VECTORS[name]:
        key, b, c, p, q, t, status, allowed, rem, reset, retry, tat_rel, exp_rel = row
        out = st.rate_limit(key, b, c, p, q, T0 + t)
        assert out[0] == status
        if status == 0:
            assert out[1:] == (bool(allowed), rem, reset, retry), (name, row, out)
        ent = st.entry(key)
        if tat_rel is None:
            assert ent is None
        else:
            assert ent == (T0 + tat_rel, ns_vectors.sat_expiry(exp_rel)), (name, row, ent)


def test_rate_vectors():
    # throttlecrab/src/core/rate/tests.rs:41-48
    for (count, period), ei in ref_scripts.RATE_VECTORS:
        st, e, _ = oracle.derive(1, count, period)
        assert st == 0 and e == ei
    # (7,60) -> 8571428571 ns ; (i64::MAX/1000, 60) -> 0
    assert oracle.derive(3, 7, 60)[1] == 8571428571
    assert oracle.derive(10, 9223372036854775, 60)[1] == 0
    # burst truncated to its low 32 bits (rate_limiter.rs:122)
    st, ei, dvt = oracle.derive(9223372036854775, 100, 60)
    assert (st, ei, dvt) == (0, 600000000, 600000000 * 2783138806)
    # Duration * u32 overflow -> the reference panics -> Internal
    assert oracle.derive(2**32, 1, 9223372036854775807)[0] == 3
