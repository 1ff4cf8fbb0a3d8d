"""Diagnostic: how many host threads can the CPU oracle really use on this box?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--torch-first" in sys.argv:
    import torch  # noqa
import numpy as np
import elastic_gpu_agent_b200 as e
from oracle import oracle_c
print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "omp max", oracle_c.max_threads(),
      "OMP_WAIT_POLICY", os.environ.get("OMP_WAIT_POLICY"), "torch first", "--torch-first" in sys.argv)
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
w = e.synth.workload("cfg3_1m")
rc, rm = e.synth.requests(3, 6, 1_000_000)
fc = np.ascontiguousarray(w["free_core"]); fm = np.ascontiguousarray(w["free_mem"]); idx = np.empty(rc.size, np.int32)
for t in (1, 4, 8, 16, 32, 64, 128):
    if t > oracle_c.max_threads() * 2:
        break
    oracle_c.snapshot_into(fc, fm, rc, rm, idx, t)
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 0.5:
        oracle_c.snapshot_into(fc, fm, rc, rm, idx, t); n += 1
    dt = time.perf_counter() - t0
    print(f"threads {t:4d}: {n * rc.size / dt / 1e6:9.1f} M decisions/s  ({1e3 * dt / n:.2f} ms/pass)")
