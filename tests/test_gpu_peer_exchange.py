"""GPU tests of the multi-GPU path that exchanges demand vectors through peer memory
(CUDA IPC) instead of a collective: world = 1 in-process, and world = 2 as two processes
that share cuda:0 (IPC works between processes on one device; the round-end multi-GPU
bench runs the same code across GPUs).  Expected values come from the CPU oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world1_shard_step_equals_snapshot(alloc, oracle_c, egpu):
    import torch
    w = egpu.synth.workload("cfg4")
    alloc.set_table(w["free_core"], w["free_mem"])
    alloc.peer_attach(0, 1, [alloc.peer_export()])
    s = torch.cuda.current_stream().cuda_stream
    D, R = 64, 50_003
    cur_c, cur_m = w["free_core"].copy(), w["free_mem"].copy()
    for step in range(80):  # more steps than exchange slots (64)
        rc, rm = egpu.synth.requests(4, 900 + step, R)
        c, m = torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda()
        idx = torch.empty(R + 1, dtype=torch.int32, device="cuda")
        dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
        tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
        alloc.bestfit_shard_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), step, s)
        commit = step % 5 == 4
        alloc.apply_peers_dev(step, tab.data_ptr(), commit, s)
        torch.cuda.synchronize()
        o_idx, o_dc, o_dm, o_tab = oracle_c.snapshot(cur_c, cur_m, rc, rm, 4)
        assert np.array_equal(idx[:R].cpu().numpy(), o_idx)
        assert np.array_equal(dl.cpu().numpy(), np.concatenate([o_dc, o_dm]))
        assert np.array_equal(tab.cpu().numpy(), o_tab)
        if commit:
            cur_c, cur_m = np.maximum(o_tab[:D], 0), np.maximum(o_tab[D:2 * D], 0)
    assert alloc.peer_last_timeout == 0
    g_c, g_m, _ = alloc.table()
    assert np.array_equal(g_c, cur_c) and np.array_equal(g_m, cur_m)
    alloc.peer_detach()


def test_world1_fused_lagged_apply(alloc, oracle_c, egpu):
    """egpu_bestfit_batch_shard_lag_dev: the scan of step k also applies step k - lag; the tail
    is flushed by one apply launch.  More steps than exchange slots, world = 1."""
    import torch
    w = egpu.synth.workload("cfg3")
    alloc.set_table(w["free_core"], w["free_mem"])
    alloc.peer_attach(0, 1, [alloc.peer_export()])
    st = torch.cuda.Stream()
    D, R, LAG, N = 8, 30_001, 3, 150
    with torch.cuda.stream(st):
        bufs = []
        for step in range(N):
            rc, rm = egpu.synth.requests(3, 500 + step, R)
            bufs.append((rc, rm, torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda(),
                         torch.empty(R, dtype=torch.int32, device="cuda"), torch.zeros(2 * D, dtype=torch.int64, device="cuda"),
                         torch.full((3 * D,), -7, dtype=torch.int32, device="cuda")))
    torch.cuda.synchronize()
    for step in range(N):
        _, _, c, m, idx, dl, _ = bufs[step]
        lagged = bufs[step - LAG][6].data_ptr() if step >= LAG else 0
        alloc.bestfit_shard_lag_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), step, LAG, lagged, st.cuda_stream,
                                    inputs_ready=True)
    alloc.apply_peers_multi_dev(N - LAG, [bufs[j][6].data_ptr() for j in range(N - LAG, N)], False, st.cuda_stream)
    torch.cuda.synchronize()
    assert alloc.peer_last_timeout == 0
    for step in range(N):
        rc, rm, _, _, idx, dl, tab = bufs[step]
        o_idx, o_dc, o_dm, o_tab = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm, 4)
        assert np.array_equal(idx.cpu().numpy(), o_idx), step
        assert np.array_equal(dl.cpu().numpy(), np.concatenate([o_dc, o_dm])), step
        assert np.array_equal(tab.cpu().numpy(), o_tab), step
    alloc.peer_detach()


def test_shard_calls_need_attach(alloc, egpu):
    alloc.set_table([1], [1])
    with pytest.raises(egpu.EgpuError) as ei:
        alloc.apply_peers_dev(0)
    assert ei.value.code == -6


def _rank_main(rank, world, conn, peer_conn, steps, R, lag=0):
    sys.path.insert(0, ROOT)
    import torch
    import elastic_gpu_agent_b200 as e
    torch.cuda.set_device(0)
    w = e.synth.workload("cfg3")
    a = e.BestFitAllocator(0)
    a.set_table(w["free_core"], w["free_mem"])
    mine = a.peer_export()
    peer_conn.send(mine)
    other = peer_conn.recv()
    handles = [mine, other] if rank == 0 else [other, mine]
    a.peer_attach(rank, world, handles)
    peer_conn.send("attached")
    assert peer_conn.recv() == "attached"
    s = torch.cuda.current_stream().cuda_stream
    D = 8
    out = []
    keep = []
    for step in range(steps):
        rc, rm = e.synth.requests(3, 70 + step, R, first_row=rank * R)
        c, m = torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda()
        idx = torch.empty(R, dtype=torch.int32, device="cuda")
        dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
        tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
        if lag:
            lagged = keep[step - lag][4].data_ptr() if step >= lag else 0
            a.bestfit_shard_lag_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), step, lag, lagged, s)
        else:
            a.bestfit_shard_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), step, s)
            a.apply_peers_dev(step, tab.data_ptr(), step % 3 == 2, s)
        keep.append((c, m, idx, dl, tab))
    if lag:
        a.apply_peers_multi_dev(steps - lag, [keep[j][4].data_ptr() for j in range(steps - lag, steps)], False, s)
    torch.cuda.synchronize()
    for c, m, idx, dl, tab in keep:
        out.append((idx.cpu().numpy(), dl.cpu().numpy(), tab.cpu().numpy()))
    fc, fm, ov = a.table()
    conn.send((rank, out, fc, fm, a.peer_last_timeout))
    peer_conn.send("done")
    peer_conn.recv()
    a.peer_detach()
    a.close()


@pytest.mark.parametrize("lag", [0, 2])
def test_world2_two_processes_one_gpu(lag, oracle_c, egpu):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    steps, R, world, D = 7, 20_001, 2, 8
    a_conn, b_conn = ctx.Pipe()          # rank0 <-> rank1
    res0_r, res0_w = ctx.Pipe(False)
    res1_r, res1_w = ctx.Pipe(False)
    p0 = ctx.Process(target=_rank_main, args=(0, world, res0_w, a_conn, steps, R, lag))
    p1 = ctx.Process(target=_rank_main, args=(1, world, res1_w, b_conn, steps, R, lag))
    p0.start()
    p1.start()
    assert res0_r.poll(180) and res1_r.poll(180), "ranks did not finish"
    r0, r1 = res0_r.recv(), res1_r.recv()
    p0.join(60)
    p1.join(60)
    assert p0.exitcode == 0 and p1.exitcode == 0
    assert r0[4] == 0 and r1[4] == 0, "an apply kernel timed out waiting for its peer"
    w = egpu.synth.workload("cfg3")
    cur_c, cur_m = w["free_core"].copy(), w["free_mem"].copy()
    for step in range(steps):
        tot = np.zeros(2 * D, dtype=np.int64)
        for rank, res in ((0, r0), (1, r1)):
            rc, rm = egpu.synth.requests(3, 70 + step, R, first_row=rank * R)
            o_idx, o_dc, o_dm, _ = oracle_c.snapshot(cur_c, cur_m, rc, rm)
            idx, dl, tab = res[1][step]
            assert np.array_equal(idx, o_idx), f"indices rank {rank} step {step}"
            assert np.array_equal(dl, np.concatenate([o_dc, o_dm]))
            tot += np.concatenate([o_dc, o_dm])
        from elastic_gpu_agent_b200 import sharding
        etab = sharding.combine_demands(cur_c, cur_m, tot[None, :])
        assert np.array_equal(r0[1][step][2], etab) and np.array_equal(r1[1][step][2], etab)
        if step % 3 == 2 and not lag:  # the fused lagged apply never commits
            cur_c, cur_m = np.maximum(etab[:D], 0), np.maximum(etab[D:2 * D], 0)
    for res in (r0, r1):
        assert np.array_equal(res[2], cur_c) and np.array_equal(res[3], cur_m)


def _multi_rank_main(rank, world, conn, peer_conn, K, R, replays, fused):
    """cfg4 (D = 64) row-sharded: every replay is gate -> one multi-batch launch of K sharded steps
    -> apply launches on a second stream, the shape bench.py --gpus N captures in a CUDA graph."""
    sys.path.insert(0, ROOT)
    import torch
    import elastic_gpu_agent_b200 as e
    torch.cuda.set_device(0)
    w = e.synth.workload("cfg4")
    D = 64
    a = e.BestFitAllocator(0)
    a.set_table(w["free_core"], w["free_mem"])
    mine = a.peer_export()
    peer_conn.send(mine)
    other = peer_conn.recv()
    a.peer_attach(rank, world, [mine, other] if rank == 0 else [other, mine])
    peer_conn.send("attached")
    assert peer_conn.recv() == "attached"
    st, ap = torch.cuda.Stream(), torch.cuda.Stream()
    with torch.cuda.stream(st):
        tens, tup = [], []
        for k in range(K):
            rc, rm = e.synth.requests(4, 40 + k, R, first_row=rank * R)
            c, m = torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda()
            idx = torch.empty(R, dtype=torch.int32, device="cuda")
            dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
            tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
            tens.append((c, m, idx, dl, tab))
            tup.append((c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), tab.data_ptr()))
    torch.cuda.synchronize()
    batches = a.make_batches(tup)
    for rep in range(replays):
        # step numbers restart at 0 on every replay (as a replayed CUDA graph does): the apply
        # kernels consumed the flags, and the ranks are synchronised between replays
        a.gate_dev(st.cuda_stream)
        a.bestfit_batches_shard_dev(batches, 0, st.cuda_stream, inputs_ready=True, apply=fused)
        if not fused:
            a.apply_peers_multi_dev(0, [t[4].data_ptr() for t in tens], False, ap.cuda_stream)
        a.gate_open()
        torch.cuda.synchronize()
        peer_conn.send("replayed")
        assert peer_conn.recv() == "replayed"
    out = [(idx.cpu().numpy(), dl.cpu().numpy(), tab.cpu().numpy()) for c, m, idx, dl, tab in tens]
    conn.send((rank, out, a.peer_last_timeout))
    peer_conn.send("done")
    peer_conn.recv()
    a.peer_detach()
    a.close()


@pytest.mark.parametrize("fused", [False, True])
def test_world2_cfg4_sharded_multi_batch_with_gate(fused, oracle_c, egpu):
    """BASELINE config 4's shape (64 devices, request rows sharded over the ranks) at world = 2:
    lookup scan + fused peer push out of a multi-batch launch, start gate; table' from a batched
    apply launch or (fused = EGPU_F_APPLY) from the last CTA of every batch."""
    import torch.multiprocessing as mp
    from elastic_gpu_agent_b200 import sharding
    ctx = mp.get_context("spawn")
    K, R, D = 12, 30_001, 64
    a_conn, b_conn = ctx.Pipe()
    res0_r, res0_w = ctx.Pipe(False)
    res1_r, res1_w = ctx.Pipe(False)
    p0 = ctx.Process(target=_multi_rank_main, args=(0, 2, res0_w, a_conn, K, R, 3, fused))
    p1 = ctx.Process(target=_multi_rank_main, args=(1, 2, res1_w, b_conn, K, R, 3, fused))
    p0.start()
    p1.start()
    assert res0_r.poll(180) and res1_r.poll(180), "ranks did not finish"
    r0, r1 = res0_r.recv(), res1_r.recv()
    p0.join(60)
    p1.join(60)
    assert p0.exitcode == 0 and p1.exitcode == 0
    assert r0[2] == 0 and r1[2] == 0, "a gate or an apply kernel timed out waiting for its peer"
    w = egpu.synth.workload("cfg4")
    for k in range(K):
        tot = np.zeros(2 * D, dtype=np.int64)
        for rank, res in ((0, r0), (1, r1)):
            rc, rm = egpu.synth.requests(4, 40 + k, R, first_row=rank * R)
            o_idx, o_dc, o_dm, _ = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm)
            idx, dl, _ = res[1][k]
            assert np.array_equal(idx, o_idx), f"indices rank {rank} batch {k}"
            assert np.array_equal(dl, np.concatenate([o_dc, o_dm]))
            tot += np.concatenate([o_dc, o_dm])
        etab = sharding.combine_demands(w["free_core"], w["free_mem"], tot[None, :])
        assert np.array_equal(r0[1][k][2], etab) and np.array_equal(r1[1][k][2], etab)


# ------------------------------------------------------------------ prefix-commit over shards
def _rank_partial(idx, rc, rm, D):
    d = np.zeros(2 * D, dtype=np.int64)
    ok = idx >= 0
    np.add.at(d, idx[ok], rc[ok].astype(np.int64))
    np.add.at(d, D + idx[ok], rm[ok].astype(np.int64))
    return d


def test_world1_shard_prefix_equals_single_gpu_prefix_commit(alloc, oracle_c, egpu):
    """world = 1: the sharded entry point must be the plain prefix-commit (base offset 0)."""
    import torch
    D = 8
    fc, fm = egpu.synth.table_full(D)
    alloc.set_table(fc, fm)
    alloc.peer_attach(0, 1, [alloc.peer_export()])
    s = torch.cuda.current_stream().cuda_stream
    cur_c, cur_m = fc.copy(), fm.copy()
    for k, R in enumerate([37, 20_003, 1, 0, 4096]):
        rc, rm = egpu.synth.requests(2, 300 + k, R)
        c = torch.from_numpy(rc).cuda() if R else torch.empty(4, dtype=torch.int32, device="cuda")
        m = torch.from_numpy(rm).cuda() if R else torch.empty(4, dtype=torch.int32, device="cuda")
        idx = torch.empty(max(R, 4), dtype=torch.int32, device="cuda")
        dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
        tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
        alloc.bestfit_shard_prefix_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), tab.data_ptr(), 2 * k,
                                       commit=True, stream=s)
        torch.cuda.synchronize()
        o_idx, o_dc, o_dm, o_tab = oracle_c.prefix_commit(cur_c, cur_m, rc, rm)
        assert np.array_equal(idx[:R].cpu().numpy(), o_idx), k
        assert np.array_equal(dl.cpu().numpy(), np.concatenate([o_dc, o_dm])), k
        assert np.array_equal(tab.cpu().numpy(), o_tab), k
        cur_c, cur_m = o_tab[:D].copy(), o_tab[D:2 * D].copy()
        if k == 1:  # start again from a roomy table so the later steps still place something
            cur_c, cur_m = fc.copy(), fm.copy()
            alloc.set_table(fc, fm)
    assert alloc.peer_last_timeout == 0
    g_c, g_m, _ = alloc.table()
    assert np.array_equal(g_c, cur_c) and np.array_equal(g_m, cur_m)
    alloc.peer_detach()


PREFIX_STEPS = [(7, 20_001, True), (3, 50, True), (0, 64, True), (1_000, 1_000, True), (5, 5, False), (2, 3, True),
                (30_000, 30_000, True)]  # (rows of rank 0, rows of rank 1, fresh table?)
PREFIX_FC = np.array([100, 100, 70, 30, 100, 50, 100, 100], dtype=np.int32)
PREFIX_FM = np.array([183359, 183359, 183359, 60, 183359, 183359, 183359, 183359], dtype=np.int32)


def _prefix_requests(k, n):
    """Four request classes that the best-fit rule sends to different devices of PREFIX_FC/FM, so
    several devices fill up at different rows: core 1..4 and core 0 -> device 3 (30 %, 60 MiB: the
    cut may come from either resource), core 31..34 -> device 5, core 51..54 -> device 2."""
    rng = np.random.default_rng(1000 + k)
    cls = rng.integers(0, 4, n)
    core = np.select([cls == 0, cls == 1, cls == 2], [rng.integers(1, 5, n), rng.integers(31, 35, n), rng.integers(51, 55, n)], 0)
    mem = rng.integers(1, 9, n)
    return core.astype(np.int32), mem.astype(np.int32)


def _prefix_rank_main(rank, world, conn, peer_conn):
    sys.path.insert(0, ROOT)
    import torch
    import elastic_gpu_agent_b200 as e
    torch.cuda.set_device(0)
    a = e.BestFitAllocator(0)
    a.set_table(PREFIX_FC, PREFIX_FM)
    mine = a.peer_export()
    peer_conn.send(mine)
    other = peer_conn.recv()
    a.peer_attach(rank, world, [mine, other] if rank == 0 else [other, mine])
    peer_conn.send("attached")
    assert peer_conn.recv() == "attached"
    s = torch.cuda.current_stream().cuda_stream
    D = 8
    out = []
    for k, (r0, r1, fresh) in enumerate(PREFIX_STEPS):
        R = r0 if rank == 0 else r1
        if fresh:
            torch.cuda.synchronize()
            a.set_table(PREFIX_FC, PREFIX_FM)
        rc, rm = _prefix_requests(k, r0 + r1)
        rc, rm = (rc[:r0], rm[:r0]) if rank == 0 else (rc[r0:], rm[r0:])
        rc, rm = np.ascontiguousarray(rc), np.ascontiguousarray(rm)
        c = torch.from_numpy(rc).cuda() if R else torch.empty(4, dtype=torch.int32, device="cuda")
        m = torch.from_numpy(rm).cuda() if R else torch.empty(4, dtype=torch.int32, device="cuda")
        idx = torch.empty(max(R, 4), dtype=torch.int32, device="cuda")
        dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
        tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
        a.bestfit_shard_prefix_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), tab.data_ptr(), 2 * k,
                                   commit=True, stream=s)
        torch.cuda.synchronize()
        out.append((idx[:R].cpu().numpy(), dl.cpu().numpy(), tab.cpu().numpy()))
    fc, fm, ov = a.table()
    conn.send((rank, out, fc, fm, a.peer_last_timeout))
    peer_conn.send("done")
    peer_conn.recv()
    a.peer_detach()
    a.close()


def test_world2_prefix_commit_rank_major(oracle_c, egpu):
    """Two ranks (two processes on cuda:0): the concatenation of the shards in rank order must be
    exactly the single-batch prefix-commit of the oracle; both ranks end with the same table."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    a_conn, b_conn = ctx.Pipe()
    res0_r, res0_w = ctx.Pipe(False)
    res1_r, res1_w = ctx.Pipe(False)
    p0 = ctx.Process(target=_prefix_rank_main, args=(0, 2, res0_w, a_conn))
    p1 = ctx.Process(target=_prefix_rank_main, args=(1, 2, res1_w, b_conn))
    p0.start()
    p1.start()
    assert res0_r.poll(180) and res1_r.poll(180), "ranks did not finish"
    r0, r1 = res0_r.recv(), res1_r.recv()
    p0.join(60)
    p1.join(60)
    assert p0.exitcode == 0 and p1.exitcode == 0
    assert r0[4] == 0 and r1[4] == 0, "a kernel timed out waiting for its peer"
    D = 8
    cur_c, cur_m = PREFIX_FC, PREFIX_FM
    cuts_in_rank1 = 0
    for k, (n0, n1, fresh) in enumerate(PREFIX_STEPS):
        if fresh:
            cur_c, cur_m = PREFIX_FC, PREFIX_FM
        rc, rm = _prefix_requests(k, n0 + n1)
        o_idx, o_dc, o_dm, o_tab = oracle_c.prefix_commit(cur_c, cur_m, rc, rm)
        i0, d0, t0 = r0[1][k]
        i1, d1, t1 = r1[1][k]
        assert np.array_equal(i0, o_idx[:n0]), f"rank 0 step {k}"
        assert np.array_equal(i1, o_idx[n0:]), f"rank 1 step {k}"
        assert np.array_equal(d0, _rank_partial(o_idx[:n0], rc[:n0], rm[:n0], D)), k
        assert np.array_equal(d1, _rank_partial(o_idx[n0:], rc[n0:], rm[n0:], D)), k
        assert np.array_equal(d0 + d1, np.concatenate([o_dc, o_dm])), k
        assert np.array_equal(t0, o_tab) and np.array_equal(t1, o_tab), k
        assert not o_tab[2 * D:].any()
        cuts_in_rank1 += int((o_idx[n0:] >= 0).any() and (o_idx[n0:] == -2).any())
        cur_c, cur_m = o_tab[:D].copy(), o_tab[D:2 * D].copy()
    assert cuts_in_rank1 >= 3, "the cases must put some cuts inside rank 1's shard"
    for res in (r0, r1):
        assert np.array_equal(res[2], cur_c) and np.array_equal(res[3], cur_m)


# ------------------------------------------------------------------ more than four ranks (second pull group)
def _rank_main_n(rank, world, conn, steps, R, fused):
    """One of `world` processes on GPU 0; the parent relays the IPC handles.  fused = False: one sharded scan
    + one apply launch per step; True: the steps as one multi-batch launch with EGPU_F_APPLY (the word
    finishers of the scan pull the peers' words themselves)."""
    sys.path.insert(0, ROOT)
    import torch
    import elastic_gpu_agent_b200 as e
    torch.cuda.set_device(0)
    w = e.synth.workload("cfg3")
    a = e.BestFitAllocator(0)
    a.set_table(w["free_core"], w["free_mem"])
    conn.send(a.peer_export())
    a.peer_attach(rank, world, conn.recv())
    conn.send("attached")
    assert conn.recv() == "go"
    s = torch.cuda.current_stream().cuda_stream
    D = 8
    keep, tup = [], []
    for step in range(steps):
        rc, rm = e.synth.requests(3, 90 + step, R, first_row=rank * R)
        c, m = torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda()
        idx = torch.empty(R, dtype=torch.int32, device="cuda")
        dl = torch.zeros(2 * D, dtype=torch.int64, device="cuda")
        tab = torch.zeros(3 * D, dtype=torch.int32, device="cuda")
        keep.append((c, m, idx, dl, tab))
        tup.append((c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), tab.data_ptr()))
    torch.cuda.synchronize()
    if fused:
        a.bestfit_batches_shard_dev(a.make_batches(tup), 0, s, inputs_ready=True, apply=True)
    else:
        for step, (c, m, idx, dl, tab) in enumerate(keep):
            a.bestfit_shard_dev(c.data_ptr(), m.data_ptr(), R, idx.data_ptr(), dl.data_ptr(), step, s)
            a.apply_peers_dev(step, tab.data_ptr(), False, s)
    torch.cuda.synchronize()
    conn.send((rank, [(idx.cpu().numpy(), dl.cpu().numpy(), tab.cpu().numpy()) for c, m, idx, dl, tab in keep], a.peer_last_timeout))
    assert conn.recv() == "done"
    a.peer_detach()
    a.close()


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [False, True])
def test_world6_six_processes_one_gpu(fused, oracle_c, egpu):
    """Six ranks: the waits collect the peers' words four ranks at a time, so ranks 4 and 5 come from the
    second group - in apply_peers_kernel and in the scan's own epilogue (EGPU_F_APPLY)."""
    import torch.multiprocessing as mp
    from elastic_gpu_agent_b200 import sharding
    ctx = mp.get_context("spawn")
    world, steps, R, D = 6, 3, 4_001, 8
    conns, procs = [], []
    for rank in range(world):
        parent, child = ctx.Pipe()
        p = ctx.Process(target=_rank_main_n, args=(rank, world, child, steps, R, fused))
        p.start()
        conns.append(parent)
        procs.append(p)
    try:
        assert all(c.poll(240) for c in conns), "ranks did not come up"
        handles = [c.recv() for c in conns]
        for c in conns:
            c.send(handles)
        assert all(c.poll(120) and c.recv() == "attached" for c in conns)
        for c in conns:
            c.send("go")
        assert all(c.poll(240) for c in conns), "ranks did not finish"
        res = sorted((c.recv() for c in conns), key=lambda r: r[0])
        for c in conns:
            c.send("done")
        for p in procs:
            p.join(60)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert all(r[2] == 0 for r in res), "a wait for a peer timed out"
    w = egpu.synth.workload("cfg3")
    for step in range(steps):
        tot = np.zeros(2 * D, dtype=np.int64)
        for rank in range(world):
            rc, rm = egpu.synth.requests(3, 90 + step, R, first_row=rank * R)
            o_idx, o_dc, o_dm, _ = oracle_c.snapshot(w["free_core"], w["free_mem"], rc, rm)
            idx, dl, _ = res[rank][1][step]
            assert np.array_equal(idx, o_idx), f"indices rank {rank} step {step}"
            assert np.array_equal(dl, np.concatenate([o_dc, o_dm]))
            tot += np.concatenate([o_dc, o_dm])
        etab = sharding.combine_demands(w["free_core"], w["free_mem"], tot[None, :])
        for rank in range(world):
            assert np.array_equal(res[rank][1][step][2], etab), f"table' rank {rank} step {step}"
