#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun; one GPU).  Outputs under gpurun_out/:
#   r2_launches.csv                 launch list of the bench step (config 2), `--metrics gpu__time_duration.sum`
#   r2_gen10.ncu-rep (+ summary)    `ncu --set full` of the last of 3 steps at config 2
#   r2_config4_densify.ncu-rep      `ncu --set full` of a mapping iteration at config-4 size AFTER a densify_and_prune
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python tools/one_step.py 500000 2 3 > gpurun_out/r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 18 -c 9 -f -o gpurun_out/r2_gen10 \
    python tools/one_step.py 500000 2 3 > gpurun_out/r2_gen10.log 2>&1
python tools/ncu_summarize.py gpurun_out/r2_gen10.ncu-rep gpurun_out/r2_ncu_gen10_summary.json gpurun_out/r2_traffic.json
ncu --set full --clock-control none -k regex:^k_ --launch-skip 24 -c 12 -f -o gpurun_out/r2_config4_densify \
    python tools/config4_densify_loop.py 4 3 > gpurun_out/r2_config4_densify.log 2>&1
python tools/ncu_summarize.py gpurun_out/r2_config4_densify.ncu-rep gpurun_out/r2_ncu_config4_densify_summary.json
tail -5 gpurun_out/r2_config4_densify.log
