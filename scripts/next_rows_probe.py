"""Runs the rows next to the scan once each (replay of cfg5, device-identity batch at node
scale) so that `ncu --metrics gpu__time_duration.sum` can list their kernels."""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import elastic_gpu_agent_b200 as e
from elastic_gpu_agent_b200 import devhash
a = e.BestFitAllocator(0)
w5 = e.synth.workload("cfg5")
kind, ea, eb = e.synth.churn_events(w5["seed"], w5["R"])
for _ in range(2):
    a.set_table(w5["free_core"], w5["free_mem"])
    t0 = time.perf_counter(); a.replay(kind, ea, eb); print("replay e2e ms", 1e3 * (time.perf_counter() - t0))
rng = random.Random(11)
sets = [["%d-%02d" % (c % 8, j) for j in rng.sample(range(183359), rng.choice([4096, 8192, 16384]))] for c in range(96)]
flat, id_off, set_off = devhash.flatten(sets)
for _ in range(3):
    t0 = time.perf_counter(); hs = devhash.device_hashes_flat(a, flat, id_off, set_off); print("hash batch ms", 1e3 * (time.perf_counter() - t0))
