"""CPU model of the two exactness arguments the decide kernels rest on, checked against
plain sequential application on random runs of requests on ONE key:

  speculate-and-commit   every pending request is evaluated against the run's current state; everything up to
                         and including the first state-changing request is final; the rest re-evaluates.
  finite-state rounds    the last K distinct states are candidates; each request is a map candidate -> candidate
                         (or NEW); an inclusive scan of the composed maps gives every request its true input
                         state; a round is repeated only when a request creates a state outside the candidates.

This is a model of the ALGORITHM (the CUDA kernels are tested against the oracle on the GPU); it documents why
the parallel evaluation reproduces the sequential semantics bit for bit, including keys that toggle between
states (max_burst = 1 with zero-quantity requests)."""
from hypothesis import given, settings, strategies as st

from test_oracle_vs_python_model import I64_MAX, U64, clamp, derive, wrap64, _tdiv

NEW = "new"


def decide(state, req):
    """state = (tat, expiry) or None; req = (now, ei, dvt, q) -> (allowed, new_state, outputs)"""
    now, ei, dvt, q = req
    live = state is not None and state[1] > now
    tat = max(state[0], clamp(now - dvt)) if live else clamp(now - ei)
    new_tat = clamp(tat + clamp(ei * q))
    allow_at = clamp(new_tat - dvt)
    allowed = now >= allow_at
    new_state = state
    if allowed:
        ttl = clamp(clamp(new_tat - now) + dvt) % U64
        new_state = (new_tat, min(now + ttl, I64_MAX))
    cur = new_tat if allowed else tat
    room = clamp(wrap64(now + dvt) - cur)
    out = (allowed, max(_tdiv(room, ei), 0) if ei > 0 else 0, max(clamp(clamp(cur - now) + dvt), 0),
           0 if allowed else max(clamp(allow_at - now), 0))
    return allowed, new_state, out


def sequential(run, s0):
    s, outs = s0, []
    for r in run:
        _, s, o = decide(s, r)
        outs.append(o)
    return outs, s


def speculate_commit(run, s0, width=32):
    outs, s = [None] * len(run), s0
    for base in range(0, len(run), width):
        pending = list(range(base, min(base + width, len(run))))
        while pending:
            ev = {i: decide(s, run[i]) for i in pending}
            first = next((i for i in pending if ev[i][1] != s), None)      # first state-changing request
            final = [i for i in pending if first is None or i <= first]
            for i in final:
                outs[i] = ev[i][2]
            if first is not None:
                s = ev[first][1]
            pending = [i for i in pending if i not in final]
    return outs, s


def fsm_rounds(run, s0, k=4, stride=16):
    outs = [None] * len(run)
    cands, cur, victim = [s0], 0, 0
    for base in range(0, len(run), stride):
        pending = list(range(base, min(base + stride, len(run))))
        while pending:
            maps = {}
            for i in pending:                                    # request -> map over the candidate indices
                m = []
                for c, cs in enumerate(cands):
                    ns = decide(cs, run[i])[1]
                    m.append(c if ns == cs else (cands.index(ns) if ns in cands else NEW))
                maps[i] = m
            idx, true_in = cur, {}
            for i in pending:                                    # what the inclusive prefix scan computes
                true_in[i] = idx
                if idx != NEW:
                    idx = maps[i][idx]
            done = [i for i in pending if true_in[i] != NEW]
            creator = None
            for i in done:
                _, ns, o = decide(cands[true_in[i]], run[i])
                outs[i] = o
                if maps[i][true_in[i]] == NEW:
                    creator = (i, ns)
            if creator is None:
                cur = idx
            else:                                                # a state outside the candidate set was created
                if len(cands) < k:
                    cands.append(creator[1])
                    cur = len(cands) - 1
                else:
                    cands[victim] = creator[1]
                    cur, victim = victim, (victim + 1) % k
            pending = [i for i in pending if i not in done]
    return outs, cands[cur]


T0 = 1_700_000_000 * 10**9
policies = st.sampled_from([(1, 1, 1), (1, 3, 3600), (2, 120, 60), (5, 10, 60), (100, 1000, 60), (3, 7, 60)])
req = st.tuples(policies, st.sampled_from([0, 0, 1, 1, 1, 1, 2, 5]), st.sampled_from([0, 0, 0, 1000, 10**9, 7 * 10**9]))


@settings(max_examples=300, deadline=None)
@given(st.lists(req, min_size=1, max_size=120), st.booleans())
def test_parallel_schemes_equal_sequential(reqs, start_with_entry):
    now, run = T0, []
    for (b, c, p), q, d in reqs:
        now += d
        _, ei, dvt = derive(b, c, p)
        run.append((now, ei, dvt, q))
    s0 = (T0 - 10**9, T0 + 5 * 10**9) if start_with_entry else None
    want = sequential(run, s0)
    assert speculate_commit(run, s0) == want
    assert fsm_rounds(run, s0) == want


def test_toggling_key_needs_few_finite_state_rounds():
    """max_burst = 1 with zero-quantity requests: the entry flips between two states."""
    _, ei, dvt = derive(1, 3, 3600)
    run = [(T0, ei, dvt, q) for q in [1, 0, 1, 1, 0, 0, 1, 2, 0, 1] * 30]
    want = sequential(run, None)
    assert fsm_rounds(run, None, stride=64) == want
    assert speculate_commit(run, None) == want
    assert len({o for o in want[0]}) > 2          # allowed and denied outcomes really alternate


# ---------------------------------------------------------------------------------------------------------
# the index-order pipeline (csrc/gcra_index_path.cuh): probe -> decide against the batch-start state -> resolve
# -> sequential residue, over a batch that mixes several keys; `entries` models the hashed batch bitmap (two
# keys may share an entry: both are then treated as shared, which must not change any result)
# ---------------------------------------------------------------------------------------------------------
def index_order_pipeline(batch, table, entries=4):
    """batch = [(key, req)], table = {key: state}.  Returns (outputs per row, table after the batch)."""
    table = dict(table)
    seen, shared = set(), set()
    for key, _ in batch:                                         # pass A: the batch bitmap
        e = key % entries
        (shared if e in seen else seen).add(e)
    outs, first_change, verdict = [None] * len(batch), {}, {}
    for i, (key, r) in enumerate(batch):                         # pass B (any order: reads batch-start states only)
        s0 = table.get(key)
        _, ns, o = decide(s0, r)
        outs[i] = o
        if key % entries not in shared:
            table[key] = ns                                      # alone on its key: commit
        else:
            verdict[i] = ns != s0
    for i in sorted(verdict, reverse=True):                      # atomicMin(mark[slot]), any order
        if verdict[i]:
            first_change[batch[i][0]] = i
    residue = []
    commits = {}
    for i in verdict:                                            # pass C
        key = batch[i][0]
        fm = first_change.get(key)
        if fm is not None and i > fm:
            residue.append(i)
        elif fm == i:
            commits[key] = decide(table.get(key), batch[i][1])[1]
    table.update(commits)
    for i in sorted(residue):                                    # residue: one after another, in batch order
        key, r = batch[i]
        _, table[key], outs[i] = decide(table.get(key), r)
    return outs, table


@settings(max_examples=300, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 9), req), min_size=1, max_size=150), st.integers(1, 7))
def test_index_order_pipeline_equals_sequential(rows, entries):
    now, batch = T0, []
    for key, ((b, c, p), q, d) in rows:
        now += d
        _, ei, dvt = derive(b, c, p)
        batch.append((key, (now, ei, dvt, q)))
    table0 = {k: (T0 - 10**9, T0 + 5 * 10**9) for k in range(0, 10, 3)}
    want_outs, want_table = [], dict(table0)
    for key, r in batch:
        _, want_table[key], o = decide(want_table.get(key), r)
        want_outs.append(o)
    outs, table = index_order_pipeline(batch, table0, entries)
    assert outs == want_outs
    assert {k: v for k, v in table.items() if v is not None} == {k: v for k, v in want_table.items() if v is not None}
