"""CPU tests (no GPU): the oracle against the hand-derived KATs, the two oracle
implementations against each other, and properties of the spec.

PARITY UNPINNED for best-fit: the reference has no such loop; the
KATs pin the builder-defined spec."""
import hashlib
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "bestfit_kat.json")))
SYN = json.load(open(os.path.join(HERE, "golden", "bestfit_synth.json")))


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", KAT["snapshot"], ids=lambda c: c["name"])
@pytest.mark.parametrize("impl", ["c", "np"])
def test_snapshot_kat(case, impl, oracle_c, oracle_np):
    o = oracle_c if impl == "c" else oracle_np
    idx, dc, dm, tab = o.snapshot(case["free_core"], case["free_mem"], case["req_core"], case["req_mem"])
    D = len(case["free_core"])
    assert idx.tolist() == case["idx"]
    assert dc.tolist() == case["delta_core"]
    assert dm.tolist() == case["delta_mem"]
    assert tab[:D].tolist() == case["table_core"]
    assert tab[D:2 * D].tolist() == case["table_mem"]
    assert tab[2 * D:].tolist() == case["oversub"]


@pytest.mark.parametrize("case", KAT["sequential"], ids=lambda c: c["name"])
@pytest.mark.parametrize("impl", ["c", "np"])
def test_sequential_kat(case, impl, oracle_c, oracle_np):
    o = oracle_c if impl == "c" else oracle_np
    idx, fc, fm = o.replay(case["free_core"], case["free_mem"], case["kind"], case["a"], case["b"])
    assert idx.tolist() == case["idx"]
    assert fc.tolist() == case["table_core"]
    assert fm.tolist() == case["table_mem"]


@pytest.mark.parametrize("case", KAT["prefix_commit"], ids=lambda c: c["name"])
@pytest.mark.parametrize("impl", ["c", "np"])
def test_prefix_commit_kat(case, impl, oracle_c, oracle_np):
    o = oracle_c if impl == "c" else oracle_np
    idx, dc, dm, tab = o.prefix_commit(case["free_core"], case["free_mem"], case["req_core"], case["req_mem"])
    D = len(case["free_core"])
    assert idx.tolist() == case["idx"]
    assert dc.tolist() == case["delta_core"] and dm.tolist() == case["delta_mem"]
    assert tab[:D].tolist() == case["table_core"] and tab[D:2 * D].tolist() == case["table_mem"]
    assert not tab[2 * D:].any()


@pytest.mark.parametrize("case", KAT["rounds"], ids=lambda c: c["name"])
@pytest.mark.parametrize("impl", ["c", "np"])
def test_rounds_kat(case, impl, oracle_c, oracle_np):
    o = oracle_c if impl == "c" else oracle_np
    idx, dc, dm, fc, fm, rounds, left = o.rounds(case["free_core"], case["free_mem"], case["req_core"], case["req_mem"],
                                                 case["max_rounds"])
    assert idx.tolist() == case["idx"]
    assert dc.tolist() == case["delta_core"] and dm.tolist() == case["delta_mem"]
    assert fc.tolist() == case["table_core"] and fm.tolist() == case["table_mem"]
    assert (rounds, left) == (case["rounds"], case["left"])


@pytest.mark.parametrize("seed", range(6))
def test_rounds_oracles_agree_and_reach_a_fixed_point(seed, oracle_c, oracle_np, egpu):
    """The C loop and the numpy construction (built on prefix_commit) agree; at the fixed point no
    row is deferred, the table is never negative, and every row reported infeasible really does
    not fit the final table."""
    rng = np.random.default_rng(seed)
    D = int(rng.choice([1, 3, 8, 17, 64]))
    fc = rng.integers(0, 101, D).astype(np.int32)
    fm = rng.integers(0, 4096, D).astype(np.int32)
    R = int(rng.integers(1, 400))
    rc = rng.integers(0, 40, R).astype(np.int32)
    rm = rng.integers(0, 600, R).astype(np.int32)
    a = oracle_c.rounds(fc, fm, rc, rm)
    b = oracle_np.rounds(fc, fm, rc, rm)
    assert all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and a[5:] == b[5:]
    idx, dc, dm, tfc, tfm, rounds, left = a
    assert left == 0 and not (idx == -2).any() and (tfc >= 0).all() and (tfm >= 0).all()
    assert np.array_equal(tfc, fc - dc) and np.array_equal(tfm, fm - dm)
    for r in np.flatnonzero(idx == -1):
        # infeasible when it was last scored; the table only shrinks afterwards
        assert not ((tfc >= rc[r]) & (tfm >= rm[r])).any()
    capped = oracle_c.rounds(fc, fm, rc, rm, 1)
    one = oracle_c.prefix_commit(fc, fm, rc, rm)
    assert np.array_equal(capped[0], one[0]) and capped[5] == 1 and capped[6] == int((one[0] == -2).sum())


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg3_1m", "cfg4"])
def test_c_oracle_matches_golden(name, oracle_c, egpu):
    g = SYN["snapshot"][name]
    w = egpu.synth.workload(name)
    assert w["free_core"].tolist() == g["free_core"] and w["free_mem"].tolist() == g["free_mem"]
    rc, rm = egpu.synth.requests(w["dist"], w["seed"], w["R"])
    assert digest(rc) == g["req_core_sha256"] and digest(rm) == g["req_mem_sha256"]
    for nthreads in (1, 4):
        idx, dc, dm, tab = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm, nthreads)
        assert digest(idx) == g["idx_sha256"]
        assert dc.tolist() == g["delta_core"] and dm.tolist() == g["delta_mem"]
        assert tab.tolist() == g["table_out"]


def test_c_oracle_matches_golden_sequential(oracle_c, egpu):
    g = SYN["sequential"]["cfg5_head20000"]
    w = egpu.synth.workload("cfg5")
    kind, a, b = egpu.synth.churn_events(w["seed"], w["R"])
    ev = SYN["sequential"]["cfg5_events_sha256"]
    assert (digest(kind), digest(a), digest(b)) == (ev["kind"], ev["a"], ev["b"])
    n = g["E"]
    idx, fc, fm = oracle_c.replay(w["free_core"], w["free_mem"], kind[:n], a[:n], b[:n])
    assert digest(idx) == g["idx_sha256"]
    assert fc.tolist() == g["free_core"] and fm.tolist() == g["free_mem"]


def test_cfg3_forced_infeasible_rows(oracle_c, egpu):
    w = egpu.synth.workload("cfg3")
    rc, rm = egpu.synth.requests(w["dist"], w["seed"], w["R"])
    idx, *_ = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm)
    assert (idx[15::16] == -1).all()
    assert (rc[15::32] == 101).all() and (rm[31::32] == egpu.synth.CAP_MEM + 1).all()


tables = st.integers(1, 64).flatmap(lambda D: st.tuples(
    st.lists(st.integers(0, 100), min_size=D, max_size=D),
    st.lists(st.integers(0, (1 << 18) - 1), min_size=D, max_size=D)))
reqs = st.lists(st.tuples(st.integers(-2, 130), st.one_of(st.integers(-2, 300), st.integers(0, (1 << 18) + 5))),
                min_size=0, max_size=40)


@settings(max_examples=150, deadline=None)
@given(tables, reqs)
def test_two_oracles_agree_and_properties(oracle_c, oracle_np, table, rq):
    fc, fm = table
    rc = np.array([r[0] for r in rq], dtype=np.int32)
    rm = np.array([r[1] for r in rq], dtype=np.int32)
    i1, dc1, dm1, t1 = oracle_c.snapshot(fc, fm, rc, rm)
    i2, dc2, dm2, t2 = oracle_np.snapshot(fc, fm, rc, rm)
    assert i1.tolist() == i2.tolist()
    assert dc1.tolist() == dc2.tolist() and dm1.tolist() == dm2.tolist() and t1.tolist() == t2.tolist()
    D = len(fc)
    for r, d in enumerate(i1.tolist()):
        c, m = int(rc[r]), int(rm[r])
        feas = [k for k in range(D) if 0 <= c <= fc[k] and 0 <= m <= fm[k]]
        if d < 0:
            assert not feas
        else:
            # brute-force optimality with lowest-index tie-break
            assert d == min(feas, key=lambda k: (fc[k] - c, fm[k] - m, k))
    # demand sums are what the indices say
    for d in range(D):
        sel = i1 == d
        assert dc1[d] == int(rc[sel].astype(np.int64).sum()) and dm1[d] == int(rm[sel].astype(np.int64).sum())


@settings(max_examples=100, deadline=None)
@given(tables, reqs)
def test_prefix_commit_oracles_agree_and_never_oversubscribe(oracle_c, oracle_np, table, rq):
    fc, fm = table
    rc = np.array([r[0] for r in rq], dtype=np.int32)
    rm = np.array([r[1] for r in rq], dtype=np.int32)
    a = oracle_c.prefix_commit(fc, fm, rc, rm)
    b = oracle_np.prefix_commit(fc, fm, rc, rm)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    idx, dc, dm, tab = a
    D = len(fc)
    assert (tab[:2 * D] >= 0).all() and not tab[2 * D:].any()
    snap, *_ = oracle_c.snapshot(fc, fm, rc, rm)
    # deferred rows are exactly snapshot choices that were cut; committed ones keep their choice
    assert ((idx == snap) | ((idx == -2) & (snap >= 0))).all()
    for d in range(D):
        sel = idx == d
        assert dc[d] == int(rc[sel].astype(np.int64).sum()) <= fc[d]
        assert dm[d] == int(rm[sel].astype(np.int64).sum()) <= fm[d]



events = st.lists(st.tuples(st.integers(0, 1), st.integers(-1, 60), st.integers(0, 2000)), min_size=0, max_size=60)


@settings(max_examples=100, deadline=None)
@given(tables, events)
def test_replay_oracles_agree(oracle_c, oracle_np, table, ev):
    fc, fm = table
    kind = np.array([e[0] for e in ev], dtype=np.int32)
    a = np.array([e[1] for e in ev], dtype=np.int32)
    b = np.array([e[2] for e in ev], dtype=np.int32)
    i1, c1, m1 = oracle_c.replay(fc, fm, kind, a, b)
    i2, c2, m2 = oracle_np.replay(fc, fm, kind, a, b)
    assert i1.tolist() == i2.tolist() and c1.tolist() == c2.tolist() and m1.tolist() == m2.tolist()
    # conservation: free + live demand == initial
    live_c = np.zeros(len(fc), dtype=np.int64)
    held = {}
    for i in range(len(ev)):
        if kind[i] == 0 and i1[i] >= 0:
            held[i] = int(i1[i])
        elif kind[i] == 1 and i1[i] >= 0:
            assert held.pop(int(a[i])) == int(i1[i])
    for i, d in held.items():
        live_c[d] += int(a[i])
    assert (np.array(fc) - live_c == c1).all()


def test_sequential_equals_snapshot_for_single_requests(oracle_c, egpu):
    """With one request per batch the two modes are the same function."""
    w = egpu.synth.workload("cfg3")
    rc, rm = egpu.synth.requests(3, 11, 200)
    for r in range(200):
        i_s, *_ = oracle_c.snapshot(w["free_core"], w["free_mem"], rc[r:r + 1], rm[r:r + 1])
        i_q, _, _ = oracle_c.replay(w["free_core"], w["free_mem"], [0], rc[r:r + 1], rm[r:r + 1])
        assert i_s[0] == i_q[0]
