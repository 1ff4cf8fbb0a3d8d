"""Builds the agent's persisted state (Bolt records + symlinks, in the reference's formats) for
a set of placements — what PreStartContainer leaves behind (pkg/plugins/gpushare.go:114-147,
:239-263).  Test helper; uses the oracle's marshal_record."""
import numpy as np

from oracle import restore_py as R


def persisted_state(placements, D, mem_per_gpu, seed=0):
    """placements: list of (gpu, core_percent, mem_mib).  Each becomes one pod with a "core" and a
    "mem" container.  The IDs are what kubelet would hand out: arbitrary unused IDs of the
    advertised pools (pkg/plugins/gpushare.go:24-33,159-168), NOT tied to the chosen GPU.
    Returns (records, links)."""
    rng = np.random.default_rng(seed)
    core_pool = rng.permutation(D * 100)
    mem_pool = rng.permutation(D * mem_per_gpu)
    cp = mp = 0
    records, links = [], []
    for n, (gpu, core, mem) in enumerate(placements):
        containers = {}
        if core > 0:
            take = core_pool[cp:cp + core]
            cp += core
            assert len(take) == core, "core pool exhausted: placements oversubscribe the node"
            ids = ["%d-%02d" % (int(t) // 100, int(t) % 100) for t in take]
            containers["core"] = (ids, R.CORE)
            links.append(("elastic-gpu-%s-0" % R.device_hash(ids), "/dev/nvidia%d" % gpu))
        if mem > 0:
            take = mem_pool[mp:mp + mem]
            mp += mem
            assert len(take) == mem, "memory pool exhausted"
            ids = ["%d-%02d" % (int(t) // mem_per_gpu, int(t) % mem_per_gpu) for t in take]
            containers["mem"] = (ids, R.MEM)
            links.append(("elastic-gpu-%s-0" % R.device_hash(ids), "/dev/nvidia%d" % gpu))
        records.append(R.marshal_record("default", "pod-%d" % n, containers))
    return records, links


def live_placements(kind, a, b, out_idx):
    """(gpu, core, mem) of every ALLOC event that was placed and not freed, from a replay's result."""
    live = {}
    for i in range(len(kind)):
        if kind[i] == 0:
            if out_idx[i] >= 0:
                live[i] = (int(out_idx[i]), int(a[i]), int(b[i]))
        elif out_idx[i] >= 0:
            live.pop(int(a[i]), None)
    return [live[k] for k in sorted(live)]
