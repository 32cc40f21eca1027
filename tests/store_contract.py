"""The reference's Store-trait contract (throttlecrab/src/core/store/store_test_suite.rs:22-539,
store/tests.rs) restated once against a duck-typed store with
    get(key, now_ns) -> value | None
    compare_and_swap_with_ttl(key, old, new, ttl_ns, now_ns) -> bool
    set_if_not_exists_with_ttl(key, value, ttl_ns, now_ns) -> bool
so the same checks run on the CPU oracle and on the CUDA table through the C ABI."""
S = 1_000_000_000
MS = 1_000_000
I64_MAX = 9223372036854775807
I64_MIN = -9223372036854775808
NOW = 1_700_000_000 * S


def basic_operations(st):                      # store_test_suite.rs:22-59
    ttl = 60 * S
    assert st.set_if_not_exists_with_ttl("key1", 100, ttl, NOW)
    assert st.get("key1", NOW) == 100
    assert not st.set_if_not_exists_with_ttl("key1", 200, ttl, NOW)
    assert st.get("key1", NOW) == 100


def compare_and_swap(st):                      # :63-109
    ttl = 60 * S
    st.set_if_not_exists_with_ttl("key1", 100, ttl, NOW)
    assert st.compare_and_swap_with_ttl("key1", 100, 200, ttl, NOW)
    assert st.get("key1", NOW) == 200
    assert not st.compare_and_swap_with_ttl("key1", 100, 300, ttl, NOW)
    assert st.get("key1", NOW) == 200
    assert not st.compare_and_swap_with_ttl("key2", 0, 100, ttl, NOW)


def ttl_expiration(st):                        # :113-170
    ttl = 60 * S
    st.set_if_not_exists_with_ttl("key1", 100, ttl, NOW)
    assert st.get("key1", NOW) == 100
    assert st.get("key1", NOW + 59 * S) == 100
    expired = NOW + 61 * S
    assert st.get("key1", expired) is None
    assert not st.compare_and_swap_with_ttl("key1", 100, 200, ttl, expired)
    assert st.set_if_not_exists_with_ttl("key1", 300, ttl, expired)
    assert st.get("key1", expired) == 300


def expiry_boundary(st):                       # adaptive_cleanup.rs:232,248,264: expired AT now == expiry
    st.set_if_not_exists_with_ttl("b", 7, 10 * S, NOW)
    assert st.get("b", NOW + 10 * S - 1) == 7
    assert st.get("b", NOW + 10 * S) is None
    assert not st.compare_and_swap_with_ttl("b", 7, 8, 10 * S, NOW + 10 * S)
    assert st.set_if_not_exists_with_ttl("b", 9, 10 * S, NOW + 10 * S)


def negative_tat(st):                          # :174-209
    ttl = 60 * S
    assert st.set_if_not_exists_with_ttl("key1", -1000, ttl, NOW)
    assert st.get("key1", NOW) == -1000
    assert st.compare_and_swap_with_ttl("key1", -1000, -500, ttl, NOW)
    assert st.get("key1", NOW) == -500


def short_ttl(st):                             # :213-247
    st.set_if_not_exists_with_ttl("key1", 100, 1 * MS, NOW)
    assert st.get("key1", NOW) == 100
    assert st.get("key1", NOW + 2 * MS) is None


def extreme_values(st):                        # :251-286
    ttl = 60 * S
    st.set_if_not_exists_with_ttl("max", I64_MAX, ttl, NOW)
    assert st.get("max", NOW) == I64_MAX
    st.set_if_not_exists_with_ttl("min", I64_MIN, ttl, NOW)
    assert st.get("min", NOW) == I64_MIN
    assert st.compare_and_swap_with_ttl("max", I64_MAX, I64_MAX - 1, ttl, NOW)
    assert st.get("max", NOW) == I64_MAX - 1


def special_keys(st):                          # :290-338
    ttl = 60 * S
    for k, v in (("", 100), ("a" * 1000, 200), ("\U0001F980\U0001F525\U0001F4BB", 300),
                 ("key:with:colons/and/slashes\\and\\backslashes", 400)):
        st.set_if_not_exists_with_ttl(k, v, ttl, NOW)
        assert st.get(k, NOW) == v


def concurrent_operations(st):                 # :342-375
    ttl = 60 * S
    st.set_if_not_exists_with_ttl("counter", 0, ttl, NOW)
    cur = 0
    for _ in range(10):
        v = st.get("counter", NOW)
        if st.compare_and_swap_with_ttl("counter", v, v + 1, ttl, NOW):
            cur += 1
    assert st.get("counter", NOW) == cur == 10


def cleanup_behavior(st):                      # :379-419
    for i in range(100):
        st.set_if_not_exists_with_ttl("key%d" % i, i, 1 * S, NOW)
    for i in range(100):
        assert st.get("key%d" % i, NOW) is not None
    for i in range(100):
        assert st.get("key%d" % i, NOW + 2 * S) is None


def ttl_update_on_cas(st):                     # :423-461
    st.set_if_not_exists_with_ttl("key1", 100, 10 * S, NOW)
    assert st.compare_and_swap_with_ttl("key1", 100, 200, 100 * S, NOW)
    assert st.get("key1", NOW + 11 * S) == 200
    assert st.get("key1", NOW + 101 * S) is None


def zero_ttl(st):                              # :465-487
    st.set_if_not_exists_with_ttl("key1", 100, 0, NOW)
    assert st.get("key1", NOW + 1) is None


def many_keys(st):                             # :491-539
    ttl = 3600 * S
    for i in range(500):
        assert st.set_if_not_exists_with_ttl("key_%d" % i, i, ttl, NOW)
    for i in range(500):
        assert st.get("key_%d" % i, NOW) == i
    for i in range(0, 500, 7):
        assert st.compare_and_swap_with_ttl("key_%d" % i, i, i + 1000, ttl, NOW)
    for i in range(0, 500, 7):
        assert st.get("key_%d" % i, NOW) == i + 1000


def huge_ttl(st):                              # rate_limiter.rs:179-183: negative ttl wraps to ~2^64 ns
    st.set_if_not_exists_with_ttl("h", 5, 2**64 - 1, NOW)
    assert st.get("h", I64_MAX - 1) == 5


CONTRACT = [basic_operations, compare_and_swap, ttl_expiration, expiry_boundary, negative_tat,
            short_ttl, extreme_values, special_keys, concurrent_operations, cleanup_behavior,
            ttl_update_on_cas, zero_ttl, many_keys, huge_ttl]
