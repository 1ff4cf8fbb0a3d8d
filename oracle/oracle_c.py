"""ctypes loader for oracle/libbestfit_oracle.so (built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY — see oracle/bestfit_oracle.c.  PARITY UNPINNED for the
best-fit functions (the reference has no such loop).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbestfit_oracle.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        lib.oracle_table_valid.restype = C.c_int
        lib.oracle_table_valid.argtypes = [vp, vp, C.c_int32]
        lib.oracle_pick_one.restype = C.c_int32
        lib.oracle_pick_one.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32]
        lib.oracle_bestfit_snapshot.restype = C.c_int
        lib.oracle_bestfit_snapshot.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int64, vp, vp, vp, vp, C.c_int]
        lib.oracle_bestfit_prefix_commit.restype = C.c_int
        lib.oracle_bestfit_prefix_commit.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int64, vp, vp, vp, vp]
        lib.oracle_bestfit_rounds.restype = C.c_int
        lib.oracle_bestfit_rounds.argtypes = [vp, vp, C.c_int32, vp, vp, C.c_int64, vp, vp, vp, C.c_int32, vp]
        lib.oracle_replay.restype = C.c_int
        lib.oracle_replay.argtypes = [vp, vp, C.c_int32, vp, vp, vp, C.c_int64, vp]
        lib.oracle_device_hash.restype = C.c_int
        lib.oracle_device_hash.argtypes = [vp, C.c_int64, vp, vp]
        lib.oracle_sha256.restype = None
        lib.oracle_sha256.argtypes = [vp, C.c_uint64, vp]
        lib.oracle_max_threads.restype = C.c_int
        lib.oracle_max_threads.argtypes = []
        _lib = lib
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _cgroup_cpus() -> float:
    """CPU quota of this container (cgroup v2 cpu.max or v1 cfs quota), inf if unlimited."""
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            return float(q) / float(p)
    except Exception:
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
        p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        if q > 0:
            return q / p
    except Exception:
        pass
    return float("inf")


def max_threads() -> int:
    """Host threads the oracle can really use: the smaller of the CPU affinity mask and the
    container's CPU quota.  (On the round-1 GPU box 128 cores are visible but cpu.max grants
    16: 16 threads run at 950 M decisions/s, 128 at 10 M.)  OMP_NUM_THREADS is deliberately
    ignored - torchrun sets it to 1 for every rank, which would turn the multi-threaded CPU
    baseline into a single-threaded one; the count is passed to OpenMP explicitly."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    q = _cgroup_cpus()
    if q != float("inf"):
        n = min(n, max(1, int(q)))
    return max(1, n)


def snapshot(free_core, free_mem, req_core, req_mem, nthreads: int = 1):
    """(idx, delta_core, delta_mem, table_out[3D]) per spec §2.4."""
    fc, fm, rc, rm = _i32(free_core), _i32(free_mem), _i32(req_core), _i32(req_mem)
    D, R = fc.size, rc.size
    idx = np.empty(R, dtype=np.int32)
    dc = np.zeros(D, dtype=np.int64)
    dm = np.zeros(D, dtype=np.int64)
    tab = np.zeros(3 * D, dtype=np.int32)
    r = load().oracle_bestfit_snapshot(_p(fc), _p(fm), D, _p(rc), _p(rm), R, _p(idx), _p(dc), _p(dm), _p(tab),
                                       int(nthreads))
    if r != 0:
        raise ValueError(f"oracle_bestfit_snapshot failed: {r}")
    return idx, dc, dm, tab


def prefix_commit(free_core, free_mem, req_core, req_mem):
    """(idx with -2 = DEFERRED, delta_core, delta_mem, table_out[3D]) per spec §2.5."""
    fc, fm, rc, rm = _i32(free_core), _i32(free_mem), _i32(req_core), _i32(req_mem)
    D, R = fc.size, rc.size
    idx = np.empty(R, dtype=np.int32)
    dc = np.zeros(D, dtype=np.int64)
    dm = np.zeros(D, dtype=np.int64)
    tab = np.zeros(3 * D, dtype=np.int32)
    r = load().oracle_bestfit_prefix_commit(_p(fc), _p(fm), D, _p(rc), _p(rm), R, _p(idx), _p(dc), _p(dm), _p(tab))
    if r != 0:
        raise ValueError(f"oracle_bestfit_prefix_commit failed: {r}")
    return idx, dc, dm, tab


def rounds(free_core, free_mem, req_core, req_mem, max_rounds: int = 1 << 20):
    """(idx, delta_core, delta_mem, free_core', free_mem', rounds, still_deferred) per spec §2.5 "rounds"."""
    fc, fm = _i32(free_core).copy(), _i32(free_mem).copy()
    rc, rm = _i32(req_core), _i32(req_mem)
    D, R = fc.size, rc.size
    idx = np.empty(R, dtype=np.int32)
    dc = np.zeros(D, dtype=np.int64)
    dm = np.zeros(D, dtype=np.int64)
    left = np.zeros(1, dtype=np.int64)
    r = load().oracle_bestfit_rounds(_p(fc), _p(fm), D, _p(rc), _p(rm), R, _p(idx), _p(dc), _p(dm), int(max_rounds), _p(left))
    if r < 0:
        raise ValueError(f"oracle_bestfit_rounds failed: {r}")
    return idx, dc, dm, fc, fm, int(r), int(left[0])


def snapshot_into(fc, fm, rc, rm, idx, nthreads: int):
    """Timing helper: preallocated int32 arrays, no result marshalling."""
    r = load().oracle_bestfit_snapshot(_p(fc), _p(fm), fc.size, _p(rc), _p(rm), rc.size, _p(idx), None, None, None,
                                       int(nthreads))
    if r != 0:
        raise ValueError(f"oracle_bestfit_snapshot failed: {r}")


def replay(free_core, free_mem, kind, a, b):
    """(idx, free_core', free_mem') per spec §2.6."""
    fc, fm = _i32(free_core).copy(), _i32(free_mem).copy()
    k, a_, b_ = _i32(kind), _i32(a), _i32(b)
    out = np.empty(k.size, dtype=np.int32)
    r = load().oracle_replay(_p(fc), _p(fm), fc.size, _p(k), _p(a_), _p(b_), k.size, _p(out))
    if r != 0:
        raise ValueError(f"oracle_replay failed: {r}")
    return out, fc, fm


def device_hash(ids) -> str:
    """types.NewDevice(...).Hash per pkg/types/device.go:17-25,49-54 (C restatement)."""
    arr = (C.c_char_p * max(1, len(ids)))(*[s.encode() for s in ids])
    out = C.create_string_buffer(9)
    r = load().oracle_device_hash(arr, len(ids), out, None)
    if r != 0:
        raise ValueError(f"oracle_device_hash failed: {r}")
    return out.value.decode()


def sha256(data: bytes) -> bytes:
    out = C.create_string_buffer(32)
    load().oracle_sha256(C.c_char_p(data), len(data), out)
    return out.raw


def device_hash_prepared(ids):
    """(callable, keepalive): the C call alone, for timing without Python marshalling."""
    enc = [x.encode() for x in ids]
    arr = (C.c_char_p * max(1, len(enc)))(*enc)
    out = C.create_string_buffer(9)
    fn = load().oracle_device_hash
    n = len(enc)

    def call():
        fn(arr, n, out, None)
        return out.value.decode()
    return call, (enc, arr, out)
