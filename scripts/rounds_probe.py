"""Where does egpu_bestfit_batch_rounds spend its time?  Wall clock per call (repeated, so
one-time allocations show) for the host-buffer and the device-buffer forms; run it under
`ncu --metrics gpu__time_duration.sum` for the per-kernel durations."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import elastic_gpu_agent_b200 as e  # noqa: E402

w3 = e.synth.workload("cfg3")
rc, rm = e.synth.requests(3, 5, 1 << 20)
rc, rm = np.minimum(rc, 5).astype(np.int32), np.minimum(rm, 2048).astype(np.int32)
with e.BestFitAllocator(0) as a:
    for rep in range(4):
        a.set_table(w3["free_core"], w3["free_mem"])
        t0 = time.perf_counter()
        idx, dc, dm, rounds, left = a.bestfit_rounds(rc, rm)
        print("host form   rep", rep, "ms", round(1e3 * (time.perf_counter() - t0), 3), "rounds", rounds, flush=True)
    c, m = torch.from_numpy(rc).cuda(), torch.from_numpy(rm).cuda()
    out = torch.empty(rc.size, dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(4):
        a.set_table(w3["free_core"], w3["free_mem"])
        t0 = time.perf_counter()
        delta, rounds, left = a.bestfit_rounds_dev(c.data_ptr(), m.data_ptr(), rc.size, out.data_ptr(), stream=st.cuda_stream)
        print("device form rep", rep, "ms", round(1e3 * (time.perf_counter() - t0), 3), "rounds", rounds, flush=True)
    for cap in (1, 2, 3):
        a.set_table(w3["free_core"], w3["free_mem"])
        t0 = time.perf_counter()
        delta, rounds, left = a.bestfit_rounds_dev(c.data_ptr(), m.data_ptr(), rc.size, out.data_ptr(), cap, stream=st.cuda_stream)
        print("device form max_rounds", cap, "ms", round(1e3 * (time.perf_counter() - t0), 3), "left", left, flush=True)
    a.set_table(w3["free_core"], w3["free_mem"])
    t0 = time.perf_counter()
    a.bestfit(rc, rm, commit=True, prefix_commit=True)
    print("single prefix-commit, host form ms", round(1e3 * (time.perf_counter() - t0), 3), flush=True)
