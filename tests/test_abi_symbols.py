"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports every
symbol include/gcra_b200.h declares; host-only helpers agree with the oracle.  No device work."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
import throttlecrab_b200 as tc
from throttlecrab_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gcra_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcra_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    _native.build()
    L = ctypes.CDLL(_native.SO_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libgcra_b200.so does not export %s" % n
    assert sorted(_native.SYMBOLS) == names, set(_native.SYMBOLS) ^ set(names)


def test_struct_layouts():
    assert tc.REQ_DTYPE.itemsize == 48 and tc.RES_DTYPE.itemsize == 32 and tc.REQ16_DTYPE.itemsize == 16
    assert tc.RES_DTYPE.fields["status"][1] == 24 and tc.RES_DTYPE.fields["allowed"][1] == 28
    assert ctypes.sizeof(_native.Config) == 72      # ... + hash_seed[2]


def test_derive_params_matches_oracle():
    rng = np.random.default_rng(1)
    cases = [(10, 100, 60), (3, 7, 60), (1, 1, 1), (2**63 - 1,) * 3, (9223372036854775, 100, 60),
             (10, 9223372036854775, 60), (2**32, 1, 2**63 - 1), (2**32 + 5, 3, 17), (5, 1, 2**63 - 1)]
    for _ in range(2000):
        e = rng.integers(0, 63, 3)
        cases.append(tuple(int(rng.integers(1, 2**int(x) + 1)) for x in e))
    for c in cases:
        assert tc.derive_params(*c) == oracle.derive(*c), c


def test_hash_key_is_stable_and_spread():
    h = tc.hash_key_ids(np.arange(200_000, dtype=np.uint64))
    assert len(np.unique(h)) == len(h)
    assert tc.hash_key("k:17") == int(h[17])
    assert tc.hash_key("") != tc.hash_key("\0")
    # owner sharding is balanced
    L = _native.lib()
    own = np.array([L.gcra_owner_of(int(x), 8) for x in h[:20000]])
    cnt = np.bincount(own, minlength=8)
    assert cnt.min() > 2000 and cnt.max() < 3000


def test_engine_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        tc.AdaptiveStore(capacity=10)


def test_rate_mirror_matches_reference_tests():
    """throttlecrab/src/core/rate/tests.rs:5-63"""
    S = 10**9
    assert tc.Rate.per_second(10).period() == 100_000_000
    assert tc.Rate.per_second(1).period() == S
    assert tc.Rate.per_minute(60).period() == S
    assert tc.Rate.per_minute(1).period() == 60 * S
    assert tc.Rate.per_hour(3600).period() == S
    assert tc.Rate.per_hour(1).period() == 3600 * S
    assert tc.Rate.per_day(86400).period() == S
    assert tc.Rate.per_day(1).period() == 86400 * S
    assert tc.Rate.from_count_and_period(10, 60).period() == 6 * S
    assert tc.Rate.from_count_and_period(30, 60).period() == 2 * S
    assert tc.Rate.from_count_and_period(0, 60).period() == (2**64 - 1) * S
    assert tc.Rate.from_count_and_period(10, 0).period() == (2**64 - 1) * S
    assert tc.Rate.new(250_000_000).period() == 250_000_000


def test_seeded_key_hash_is_siphash_2_4():
    """gcra_hash_key_seeded = SipHash-2-4 (reference vectors of the SipHash paper's test program: key 00..0f, input
    00, 01, ... of increasing length)."""
    L = _native.lib()
    k0, k1 = 0x0706050403020100, 0x0f0e0d0c0b0a0908
    want = {0: 0x726fdb47dd0e0e31, 1: 0x74f839c593dc67fd, 8: 0x93f5f5799a932462, 15: 0xa129ca6149be45e5}
    for n, w in want.items():
        data = bytes(range(n))
        assert L.gcra_hash_key_seeded(data, n, k0, k1) == w, n
    assert L.gcra_hash_key_seeded(b"abc", 3, 1, 2) != L.gcra_hash_key_seeded(b"abc", 3, 1, 3)
