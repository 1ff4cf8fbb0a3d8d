# tuning sweep (run under gpurun): prints us/step for the headline workload
run() { env "$@" python bench.py --steps 480 --warmup 20 --cpu-budget 0.1 --no-sweep > gpurun_out/b.json 2> gpurun_out/b.err; python - "$*" <<PY
import json,sys
try:
    j=json.load(open("gpurun_out/b.json")); print(sys.argv[1], "us/step", round(1e3*j["ms_per_step"],3), "frac", round(j["roofline"]["frac"],3), "parity", j["parity_vs_oracle"])
except Exception as ex: print(sys.argv[1], "FAILED", ex, open("gpurun_out/b.err").read()[-500:])
PY
}
for v in 2 4; do for t in 32 48 64 96; do run EGPU_VEC=$v EGPU_ROWS_PER_THREAD=$t; done; done
run EGPU_VEC=4 EGPU_ROWS_PER_THREAD=64 EGPU_PIPE_GROUP=24
run EGPU_VEC=2 EGPU_ROWS_PER_THREAD=48 EGPU_PIPE_GROUP=24
