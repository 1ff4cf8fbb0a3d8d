"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, the
all-gather of per-rank demand vectors and the combine step must reproduce the
unsharded oracle.  The per-shard scan is done by the oracle here (no GPU); on the GPU
box the same flow runs through the CUDA library (test_gpu_parity / bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_rows, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import elastic_gpu_agent_b200 as e
        from elastic_gpu_agent_b200 import sharding
        from oracle import oracle_c
        w = e.synth.workload("cfg4")
        lo, hi = sharding.shard_bounds(total_rows, world, rank)
        rc, rm = e.synth.requests(w["dist"], w["seed"], hi - lo, first_row=lo)
        idx, dc, dm, _ = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm)
        delta = torch.from_numpy(np.concatenate([dc, dm]))
        gathered = sharding.gather_demands(delta, world)
        tab = sharding.combine_demands(w["free_core"], w["free_mem"], gathered.numpy())
        # every rank must hold the same table'
        tabs = [torch.zeros(tab.size, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(tabs, torch.from_numpy(tab))
        same = all(torch.equal(tabs[0], t) for t in tabs)
        q.put((rank, lo, hi, idx.tolist() if total_rows <= 4096 else int(idx.astype(np.int64).sum()), tab.tolist(), same))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total_rows", [1001, 50_000])
def test_row_sharding_world2_matches_unsharded_oracle(total_rows, egpu, oracle_c):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = egpu.synth.workload("cfg4")
    rc, rm = egpu.synth.requests(w["dist"], w["seed"], total_rows)
    idx, dc, dm, tab = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total_rows
    for r in res:
        assert r[5], "ranks disagree on table'"
        assert r[4] == tab.tolist()
    if total_rows <= 4096:
        assert res[0][3] + res[1][3] == idx.tolist()
    else:
        assert res[0][3] + res[1][3] == int(idx.astype(np.int64).sum())


def test_shard_bounds_cover_and_balance(egpu):
    from elastic_gpu_agent_b200 import sharding
    for total in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 3, 8):
            b = [sharding.shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def test_combine_demands_matches_oracle_apply(egpu, oracle_np):
    from elastic_gpu_agent_b200 import sharding
    rng = np.random.default_rng(5)
    fc = rng.integers(0, 101, 16)
    fm = rng.integers(0, 1 << 18, 16)
    g = rng.integers(0, 1 << 40, (4, 32))
    g[:, :16] = rng.integers(0, 60, (4, 16))
    tot = g.sum(axis=0)
    assert np.array_equal(sharding.combine_demands(fc, fm, g), oracle_np.apply_delta(fc, fm, tot[:16], tot[16:]))


# -- prefix-commit over shards: rank-major base offsets (host logic; the CUDA path is
#    tests/test_gpu_peer_exchange.py::test_world2_prefix_commit_rank_major).  The two helpers
#    below restate prefix_base_kernel / prefix_cut_kernel with numpy: checker code, kept out of
#    the product package on purpose -----------------------------------------------------------
def prefix_base(gathered_uncapped, rank: int) -> np.ndarray:
    """Demand (int64[2*D]) already on the running sums when rank `rank`'s first row is
    considered: the UNCAPPED demand vectors of the lower ranks, summed.  Test-side restatement of prefix_base_kernel."""
    g = np.asarray(gathered_uncapped, dtype=np.int64)
    return g[:rank].sum(axis=0) if rank > 0 else np.zeros(g.shape[1], dtype=np.int64)


def prefix_cut_shard(free_core, free_mem, idx, req_core, req_mem, base):
    """Test-side restatement of prefix_cut/apply_kernel for one shard: rows whose device's running
    demand (base + the rows of this shard up to and including it, committed or not) exceeds
    free[d] become -2.  Returns (idx', committed int64[2*D])."""
    fc = np.asarray(free_core, dtype=np.int64)
    fm = np.asarray(free_mem, dtype=np.int64)
    D = fc.size
    idx = np.asarray(idx, dtype=np.int32)
    rc = np.asarray(req_core, dtype=np.int64)
    rm = np.asarray(req_mem, dtype=np.int64)
    base = np.asarray(base, dtype=np.int64)
    out = idx.copy()
    committed = np.zeros(2 * D, dtype=np.int64)
    for d in range(D):
        sel = idx == d
        pc = base[d] + np.cumsum(np.where(sel, rc, 0))
        pm = base[D + d] + np.cumsum(np.where(sel, rm, 0))
        ok = sel & (pc <= fc[d]) & (pm <= fm[d])
        out[sel & ~ok] = -2
        committed[d] = rc[ok].sum()
        committed[D + d] = rm[ok].sum()
    return out, committed


def _prefix_case(total_rows):
    rng = np.random.default_rng(total_rows)
    fc = np.array([100, 100, 70, 30, 100, 50, 100, 100], dtype=np.int32)
    fm = np.array([183359, 183359, 183359, 60, 183359, 183359, 183359, 183359], dtype=np.int32)
    cls = rng.integers(0, 4, total_rows)
    core = np.select([cls == 0, cls == 1, cls == 2], [rng.integers(1, 5, total_rows), rng.integers(31, 35, total_rows),
                                                      rng.integers(51, 55, total_rows)], 0).astype(np.int32)
    mem = rng.integers(1, 9, total_rows).astype(np.int32)
    return fc, fm, core, mem


def _prefix_worker(rank, world, port, total_rows, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from elastic_gpu_agent_b200 import sharding
        from oracle import oracle_c
        fc, fm, core, mem = _prefix_case(total_rows)
        lo, hi = sharding.shard_bounds(total_rows, world, rank)
        rc, rm = core[lo:hi], mem[lo:hi]
        idx, dc, dm, _ = oracle_c.snapshot(fc, fm, rc, rm)                       # the scan of this shard
        uncapped = sharding.gather_demands(torch.from_numpy(np.concatenate([dc, dm])), world)   # exchange step 1
        base = prefix_base(uncapped.numpy(), rank)
        idx2, committed = prefix_cut_shard(fc, fm, idx, rc, rm, base)
        gathered = sharding.gather_demands(torch.from_numpy(committed), world)   # exchange step 2
        tab = sharding.combine_demands(fc, fm, gathered.numpy())
        q.put((rank, idx2.tolist(), committed.tolist(), tab.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total_rows", [(2, 40), (2, 2001), (3, 95)])
def test_prefix_commit_rank_major_matches_unsharded_oracle(world, total_rows, oracle_c):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_prefix_worker, args=(r, world, port, total_rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    fc, fm, core, mem = _prefix_case(total_rows)
    o_idx, o_dc, o_dm, o_tab = oracle_c.prefix_commit(fc, fm, core, mem)
    assert sum((r[1] for r in res), []) == o_idx.tolist()
    assert np.array_equal(np.sum([r[2] for r in res], axis=0), np.concatenate([o_dc, o_dm]))
    for r in res:
        assert r[3] == o_tab.tolist()
    assert (o_idx == -2).any() and (o_idx >= 0).any()
