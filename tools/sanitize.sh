#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels with shared-memory staging, warp shuffles, mbarrier /
# TMA rings and batch strides, at test sizes.  Run on a GPU box:   bash tools/sanitize.sh [outdir]
# Writes <outdir>/sanitizer_{memcheck,racecheck}.log; the last lines carry the error summaries.
out=${1:-gpurun_out}
mkdir -p "$out"
sel='low_order_target_pass[9 or low_order_target_pass[12 or diagonal_block or default_schedule_ends or batched_handle_random_programs[9 or batched_handle_random_programs[12 or gather_block_on_virtual_peers[variant0-1 or gather_block_on_virtual_peers[variant1-2 or beta_sweep_through_one_batched_handle[double-release] or expansion_fast_path or rotated'
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
      python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$sel" > "$out/sanitizer_$tool.log" 2>&1
  echo "$tool exit code: $?" >> "$out/sanitizer_$tool.log"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$out/sanitizer_$tool.log" | tail -5
done
