# A/B of the lookup scan's demand sums (run under gpurun): 32-bit shared-memory atomics
# (default) vs lane-private phased sums (EGPU_LUT_ACC=lane).  cfg4 = 64 devices x 1M requests.
run() { env "$@" python bench.py --workload cfg4 --steps 480 --warmup 20 --cpu-budget 0.1 --no-sweep > gpurun_out/b.json 2> gpurun_out/b.err; python - "$*" <<PY
import json,sys
try:
    j=json.load(open("gpurun_out/b.json")); print(sys.argv[1], "us/step", round(1e3*j["ms_per_step"],3), "frac", round(j["roofline"]["frac"],3), "parity", j["parity_vs_oracle"])
except Exception as ex: print(sys.argv[1], "FAILED", ex, open("gpurun_out/b.err").read()[-800:])
PY
}
run EGPU_LUT_ACC=atomic
run EGPU_LUT_ACC=lane
run EGPU_LUT_ACC=atomic EGPU_CTAS_PER_SM=2
run EGPU_LUT_ACC=atomic EGPU_CTAS_PER_SM=3
run EGPU_LUT_ACC=atomic
