"""oracle/bestfit_np.py — second, independent CPU restatement of the best-fit spec.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's CPU-baseline legs; never by the product package.

PARITY UNPINNED: the reference (elastic-ai/elastic-gpu-agent @ 2609107) has no
best-fit loop (GetPreferredAllocation is a stub, pkg/plugins/base.go:94-96; the
GPU index is read from an annotation, pkg/plugins/gpushare.go:107-125).  This
file follows DESIGN.md §2 / SURVEY.md Appendix A.  It is deliberately written a
different way from oracle/bestfit_oracle.c — the packed-key argmin of Appendix
A.3 over a dense (request x device) grid — so that agreement between the two is
evidence the spec is unambiguous.

Units pinned by the reference: core percent, 100 per card
(pkg/common/const.go:4); memory MiB (pkg/plugins/gpushare.go:161); device index
= NVML order (pkg/operator/base.go:29-33).
"""
from __future__ import annotations

import numpy as np

MAX_DEVICES = 64
CORE_MAX = 100
MEM_MAX = (1 << 18) - 1
INT32_MAX = np.int64(2**31 - 1)


def table_valid(free_core, free_mem) -> bool:
    fc = np.asarray(free_core)
    fm = np.asarray(free_mem)
    if fc.ndim != 1 or fc.shape != fm.shape or not (1 <= fc.size <= MAX_DEVICES):
        return False
    return bool(((fc >= 0) & (fc <= CORE_MAX) & (fm >= 0) & (fm <= MEM_MAX)).all())


def pick_grid(free_core, free_mem, req_core, req_mem) -> np.ndarray:
    """Appendix A.3: key = (lc << 24) | (lm << 6) | d, infeasible -> INT32_MAX,
    idx = argmin key (its low 6 bits), -1 when the minimum is INT32_MAX."""
    fc = np.asarray(free_core, dtype=np.int64)[None, :]
    fm = np.asarray(free_mem, dtype=np.int64)[None, :]
    rc = np.asarray(req_core, dtype=np.int64)[:, None]
    rm = np.asarray(req_mem, dtype=np.int64)[:, None]
    d = np.arange(fc.shape[1], dtype=np.int64)[None, :]
    lc = fc - rc
    lm = fm - rm
    feasible = (rc >= 0) & (rm >= 0) & (lc >= 0) & (lm >= 0)
    key = np.where(feasible, (lc << 24) | (lm << 6) | d, INT32_MAX)
    if key.shape[0] == 0:
        return np.zeros(0, dtype=np.int32)
    kmin = key.min(axis=1)
    return np.where(kmin == INT32_MAX, -1, kmin & 63).astype(np.int32)


def snapshot(free_core, free_mem, req_core, req_mem, chunk: int = 1 << 18):
    """Spec §2.4.  Returns (idx int32[R], delta_core int64[D], delta_mem int64[D],
    table_out int32[3*D] = free_core', free_mem', oversub)."""
    assert table_valid(free_core, free_mem)
    fc = np.asarray(free_core, dtype=np.int64)
    fm = np.asarray(free_mem, dtype=np.int64)
    rc = np.asarray(req_core, dtype=np.int32)
    rm = np.asarray(req_mem, dtype=np.int32)
    D = fc.size
    R = rc.size
    idx = np.empty(R, dtype=np.int32)
    for s in range(0, R, chunk):
        idx[s:s + chunk] = pick_grid(fc, fm, rc[s:s + chunk], rm[s:s + chunk])
    ok = idx >= 0
    dc = np.zeros(D, dtype=np.int64)
    dm = np.zeros(D, dtype=np.int64)
    np.add.at(dc, idx[ok], rc[ok].astype(np.int64))
    np.add.at(dm, idx[ok], rm[ok].astype(np.int64))
    return idx, dc, dm, apply_delta(fc, fm, dc, dm)


def prefix_commit(free_core, free_mem, req_core, req_mem):
    """Spec §2.5 written with per-device cumulative sums (independent of the C loop)."""
    idx, _, _, _ = snapshot(free_core, free_mem, req_core, req_mem)
    fc = np.asarray(free_core, dtype=np.int64)
    fm = np.asarray(free_mem, dtype=np.int64)
    rc = np.asarray(req_core, dtype=np.int64)
    rm = np.asarray(req_mem, dtype=np.int64)
    D = fc.size
    out = idx.copy()
    dc = np.zeros(D, dtype=np.int64)
    dm = np.zeros(D, dtype=np.int64)
    for d in range(D):
        sel = idx == d
        pc = np.cumsum(np.where(sel, rc, 0))
        pm = np.cumsum(np.where(sel, rm, 0))
        ok = sel & (pc <= fc[d]) & (pm <= fm[d])
        out[sel & ~ok] = -2
        dc[d] = rc[ok].sum()
        dm[d] = rm[ok].sum()
    return out, dc, dm, apply_delta(fc, fm, dc, dm)


def rounds(free_core, free_mem, req_core, req_mem, max_rounds: int = 1 << 20):
    """Spec §2.5 "rounds", built from prefix_commit above (independent of the C loop): returns
    (idx, delta_core, delta_mem, free_core', free_mem', rounds, still_deferred)."""
    fc = np.asarray(free_core, dtype=np.int64).copy()
    fm = np.asarray(free_mem, dtype=np.int64).copy()
    rc = np.asarray(req_core, dtype=np.int32)
    rm = np.asarray(req_mem, dtype=np.int32)
    out = np.empty(rc.size, dtype=np.int32)
    rows = np.arange(rc.size)
    tc = np.zeros(fc.size, dtype=np.int64)
    tm = np.zeros(fc.size, dtype=np.int64)
    n_rounds = 0
    while rows.size and n_rounds < max_rounds:
        idx, dc, dm, tab = prefix_commit(fc, fm, rc[rows], rm[rows])
        out[rows] = idx
        fc, fm = tab[:fc.size].astype(np.int64), tab[fc.size:2 * fc.size].astype(np.int64)
        tc += dc
        tm += dm
        rows = rows[idx == -2]
        n_rounds += 1
    return out, tc, tm, fc.astype(np.int32), fm.astype(np.int32), n_rounds, int(rows.size)


def apply_delta(free_core, free_mem, delta_core, delta_mem) -> np.ndarray:
    c = np.asarray(free_core, dtype=np.int64) - np.asarray(delta_core, dtype=np.int64)
    m = np.asarray(free_mem, dtype=np.int64) - np.asarray(delta_mem, dtype=np.int64)
    lo, hi = -(2**31), 2**31 - 1
    over = ((c < 0) | (m < 0)).astype(np.int64)
    return np.concatenate([np.clip(c, lo, hi), np.clip(m, lo, hi), over]).astype(np.int32)


def replay(free_core, free_mem, kind, a, b):
    """Spec §2.6, pure-Python loop (small cases only).  Returns
    (idx int32[E], free_core', free_mem')."""
    assert table_valid(free_core, free_mem)
    fc = [int(x) for x in free_core]
    fm = [int(x) for x in free_mem]
    D = len(fc)
    E = len(kind)
    out = np.empty(E, dtype=np.int32)
    live = {}
    for i in range(E):
        if int(kind[i]) == 0:
            c, m = int(a[i]), int(b[i])
            best = None
            if c >= 0 and m >= 0:
                for d in range(D):
                    if fc[d] >= c and fm[d] >= m:
                        k = (fc[d] - c, fm[d] - m, d)
                        if best is None or k < best:
                            best = k
            if best is None:
                out[i] = -1
            else:
                d = best[2]
                fc[d] -= c
                fm[d] -= m
                live[i] = d
                out[i] = d
        else:
            t = int(a[i])
            if int(kind[i]) == 1 and 0 <= t < i and t in live:
                d = live.pop(t)
                fc[d] += int(a[t])
                fm[d] += int(b[t])
                out[i] = d
            else:
                out[i] = -1
    return out, np.array(fc, dtype=np.int32), np.array(fm, dtype=np.int32)
