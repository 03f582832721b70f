#!/bin/bash
# Workload sweep of SURVEY.md 8d on one B200: configs 2 and 4 at size multipliers m = 1, 2, 4.
# Output: gpurun_out/sweep.jsonl (one bench.py line per workload).
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for P in 500000 2000000; do
  for m in 1 2 4; do
    python bench.py --P $P --m $m --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 >> gpurun_out/sweep.jsonl
  done
done
cat gpurun_out/sweep.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config']['workload'][:40], 'R', d['config']['tile_instances'], 'ms', round(d['ms_per_step'], 3), 'GGf/s', round(d['value'] / 1e6, 1), d['kernel_ms'])
"
