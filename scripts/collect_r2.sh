#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): GPU tests, bench, ncu launch list, ncu --set full of the
# three scan kernels that matter, sanitizer subset.  Everything lands in gpurun_out/; the summaries worth
# keeping are copied to profiles/ by hand (profiles/README.md says which).
set -u
O=gpurun_out
mkdir -p $O
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4) > $O/r2_gputests.txt
(timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err); echo "bench rc=$?" >> $O/r2_gputests.txt
(timeout 200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_ref.json 2>> $O/r2_bench_n1.err)
# launch list of the same command (cold, serialised: shares, not absolutes)
EGPU_BENCH_NO_GATE=1 EGPU_BENCH_REPLAYS=5 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $O/r2_launches.csv python bench.py --gpus 1 --steps 20 --warmup 5 --no-sweep > $O/r2_bench_under_ncu.log 2>&1
# full captures: the multi-batch launches of the headline (8 devices), of cfg4 (64 devices) and one 64 Mi-row launch
for spec in "cfg3_1m 20 sorted_multi" "cfg4 20 lut_multi" "cfg3_64mi 2 sorted_64mi"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:multi -c 3 -f -o $O/r2_$3 \
      python scripts/multi_probe.py $1 $2 --only-multi > $O/r2_ncu_$3.log 2>&1
done
bash scripts/sanitize.sh $O > $O/r2_sanitize.log 2>&1
cat $O/r2_gputests.txt
