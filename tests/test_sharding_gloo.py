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
