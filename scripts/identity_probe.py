"""The node-scale device-set identity batch of bench.py's next_rows (96 sets, 934 k IDs), alone: for an ncu
launch list of its kernels (`ncu --metrics gpu__time_duration.sum ... python scripts/identity_probe.py`)
and a host-clock time per call with pageable buffers."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402,F401
import elastic_gpu_agent_b200 as e  # noqa: E402
from elastic_gpu_agent_b200 import devhash  # noqa: E402

alloc = e.BestFitAllocator(0)
rng = random.Random(11)
sets = [["%d-%02d" % (c % 8, j) for j in rng.sample(range(183359), rng.choice([4096, 8192, 16384]))] for c in range(96)]
flat, id_off, set_off = devhash.flatten(sets)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
for k in range(n):
    t0 = time.perf_counter()
    hs = devhash.device_hashes_flat(alloc, flat, id_off, set_off)
    print("call %d: %.3f ms" % (k, 1e3 * (time.perf_counter() - t0)))
print(hs[:3], len(flat), id_off.size)
