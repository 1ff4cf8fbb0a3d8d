# tuning sweep used in round 1 (run under gpurun): prints us/step for the headline workload
run() { env "$@" python bench.py --steps 480 --warmup 20 --cpu-budget 0.2 --no-sweep > gpurun_out/b.json 2> gpurun_out/b.err; python - "$*" <<PY
import json,sys
try:
    j=json.load(open("gpurun_out/b.json")); print(sys.argv[1], "us/step", round(1e3*j["ms_per_step"],3), "frac", round(j["roofline"]["frac"],3), "e2e us", round(1e3*j["e2e"]["ms_per_step"],1), "parity", j["parity_vs_oracle"])
except Exception as ex: print(sys.argv[1], "FAILED", ex, open("gpurun_out/b.err").read()[-500:])
PY
}
run EGPU_X=0
run EGPU_ROWS_PER_THREAD=32
run EGPU_ROWS_PER_THREAD=64
run EGPU_ROWS_PER_THREAD=128
run EGPU_ROWS_PER_THREAD=256
