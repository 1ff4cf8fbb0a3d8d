"""Is a CUDA event recorded after PDL-attributed launches ordered after their completion?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import elastic_gpu_agent_b200 as e
a = e.BestFitAllocator(0)
w = e.synth.workload("cfg3")
a.set_table(w["free_core"], w["free_mem"])
R = 64 << 20
st = torch.cuda.Stream(); torch.cuda.set_stream(st); sh = st.cuda_stream
c = torch.empty(R, dtype=torch.int32, device="cuda"); m = torch.empty_like(c)
a.synth_requests_dev(3, 7, 0, R, c.data_ptr(), m.data_ptr(), sh)
outs = [torch.full((R,), -9, dtype=torch.int32, device="cuda") for _ in range(3)]
dl = [torch.zeros(16, dtype=torch.int64, device="cuda") for _ in range(3)]
torch.cuda.synchronize()
for ready in (False, True):
    for o in outs: o.fill_(-9)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st)
    n = 9
    for i in range(n):
        a.bestfit_dev(c.data_ptr(), m.data_ptr(), R, outs[i % 3].data_ptr(), dl[i % 3].data_ptr(), 0, False, sh, inputs_ready=ready)
    e1.record(st)
    t_launch = time.perf_counter() - t0
    e1.synchronize()
    t_ev = time.perf_counter() - t0
    unfinished = int((outs[(n - 1) % 3] == -9).sum())   # rows not yet written when the event fired
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    print(f"inputs_ready={ready}: events {e0.elapsed_time(e1)*1e3/n:.1f} us/launch; host: launch {t_launch*1e6/n:.1f}, event sync {t_ev*1e6/n:.1f}, device sync {t_all*1e6/n:.1f} us/launch; rows unwritten at event: {unfinished}")
