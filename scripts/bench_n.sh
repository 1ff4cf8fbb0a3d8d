#!/bin/bash
# bench.py on N GPUs the way the driver launches it; JSON line -> gpurun_out/r2_bench_n$N[_$EX].json
N=$1; EX=${2:-peer}; O=gpurun_out; mkdir -p $O
SUF=""; [ "$EX" != "peer" ] && SUF="_$EX"
if [ "$N" = "1" ]; then
  timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err
else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 20 --warmup 5 --exchange $EX > $O/r2_bench_n$N$SUF.json 2> $O/r2_bench_n$N.err
fi
echo "rc=$?"
python - <<PY
import json
d=json.load(open("$O/r2_bench_n$N$SUF.json"))
print("N=$N $EX", "value", d["value"], "us/step", 1e3*d["ms_per_step"], "min/max", 1e3*d["timing"]["ms_per_step_min"], 1e3*d["timing"]["ms_per_step_max"], "parity", d["parity_vs_oracle"], "launches", d["gpu_launches"])
c=d.get("cfg4_sharded")
if c: print("  cfg4_sharded us/step", c["us_per_step"], "frac", c["frac"], "parity", c["parity_vs_oracle"])
a=d.get("allgather_path")
if a: print("  allgather", {k:(v["parity_vs_oracle"], round(v["us_per_step_eager"],1)) for k,v in a["cases"].items()})
p=d.get("prefix_commit_over_shards")
if p: print("  prefix over shards", p["bit_exact_all_ranks"])
print("  e2e", d["e2e"]["value"], "packed", d["e2e_packed"]["value"], "clocks", d["clocks"])
PY
