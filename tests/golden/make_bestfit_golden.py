"""Generates tests/golden/bestfit_synth.json: outputs of the numpy oracle
(oracle/bestfit_np.py) on the BASELINE.json configs, so that the C oracle, the
CUDA path and any future reimplementation can be checked without re-deriving
them.  Run from the repo root:  python tests/golden/make_bestfit_golden.py

The reference cannot generate these (it has no best-fit loop);
they pin the builder-defined spec.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import elastic_gpu_agent_b200 as e  # noqa: E402  (synthetic inputs only)
from oracle import bestfit_np  # noqa: E402


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


out = {"_generator": "tests/golden/make_bestfit_golden.py (oracle/bestfit_np.py)", "snapshot": {}, "sequential": {}}
for name in ["cfg2", "cfg3", "cfg3_1m", "cfg4"]:
    w = e.synth.workload(name)
    rc, rm = e.synth.requests(w["dist"], w["seed"], w["R"])
    idx, dc, dm, tab = bestfit_np.snapshot(w["free_core"], w["free_mem"], rc, rm)
    rec = {
        "D": int(w["D"]), "R": int(w["R"]),
        "free_core": w["free_core"].tolist(), "free_mem": w["free_mem"].tolist(),
        "req_core_sha256": digest(rc), "req_mem_sha256": digest(rm),
        "idx_sha256": digest(idx), "idx_histogram": np.bincount(idx + 1, minlength=w["D"] + 1).tolist(),
        "delta_core": dc.tolist(), "delta_mem": dm.tolist(), "table_out": tab.tolist(),
    }
    if w["R"] <= 1000:
        rec["req_core"] = rc.tolist()
        rec["req_mem"] = rm.tolist()
        rec["idx"] = idx.tolist()
    else:
        rec["idx_head"] = idx[:64].tolist()
    out["snapshot"][name] = rec

# sequential: cfg5 head (pure-Python oracle is slow; 20000 events) --------------
w = e.synth.workload("cfg5")
kind, a, b = e.synth.churn_events(w["seed"], w["R"])
n = 20000
idx, fc, fm = bestfit_np.replay(w["free_core"], w["free_mem"], kind[:n], a[:n], b[:n])
out["sequential"]["cfg5_head20000"] = {
    "E": n, "kind_sha256": digest(kind[:n]), "a_sha256": digest(a[:n]), "b_sha256": digest(b[:n]),
    "idx_sha256": digest(idx), "idx_head": idx[:64].tolist(), "free_core": fc.tolist(), "free_mem": fm.tolist(),
}
out["sequential"]["cfg5_events_sha256"] = {"E": int(w["R"]), "kind": digest(kind), "a": digest(a), "b": digest(b)}

with open(os.path.join(ROOT, "tests", "golden", "bestfit_synth.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote bestfit_synth.json")
