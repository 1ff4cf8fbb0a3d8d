#!/bin/bash
# compute-sanitizer over a reduced -m gpu subset (run under gpurun; summaries land in gpurun_out/).
# memcheck: every kernel family once; racecheck / synccheck: the shared-memory-heavy scans and epilogues.
set -u
export EGPU_UNDER_SANITIZER=1  # launches are blocking under the tool: tests skip the start gate (it waits for the host)
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
SUB_A='test_multi_batch_launch_equals_separate_calls and (K3 or 3-) or test_multi_batch_launches_pipelined or test_query_leaves or test_commit_after or test_prefix_commit_cut or test_lookup_scan_with_clustered or test_world1_sharded_multi_batch_and_gate or test_apply_deltas_on_gathered_vectors and G2'
SUB_B='test_empty_and_ragged_batches or test_snapshot_kat or test_sequential_kat or test_prefix_commit_kat or test_pipelined_launches_overlap_safely'
run() {  # tool, label, extra args..., then pytest selection
  local tool=$1 label=$2; shift 2
  timeout 900 compute-sanitizer --tool "$tool" --target-processes all --error-exitcode 86 --print-limit 30 "$@" \
      > "$OUT/r2_sanitizer_${tool}_${label}.txt" 2>&1
  echo "exit=$? tool=$tool label=$label" >> "$OUT/r2_sanitizer_${tool}_${label}.txt"
  tail -n 6 "$OUT/r2_sanitizer_${tool}_${label}.txt"
}
for tool in memcheck racecheck synccheck; do
  run $tool multi python -m pytest tests/test_gpu_multi.py -q -x -m gpu -k "$SUB_A"
  run $tool parity python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "$SUB_B"
done
[ -n "${SKIP_DEVHASH:-}" ] && { run memcheck world2 python -m pytest tests/test_gpu_peer_exchange.py -q -x -m gpu -k "world2_two_processes_one_gpu or world1"; exit 0; }
run memcheck devhash python -m pytest tests/test_devhash.py tests/test_restore.py -q -x -m gpu
run memcheck world2 python -m pytest tests/test_gpu_peer_exchange.py -q -x -m gpu -k "world2_two_processes_one_gpu or world1"
run racecheck devhash python -m pytest tests/test_devhash.py -q -x -m gpu -k "long_messages or golden_batch"
