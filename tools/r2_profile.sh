#!/bin/bash
# Round-2 ncu evidence, final code of the round (run on the GPU box through gpurun; one GPU).  The reports themselves go
# to $OUT (default /tmp/r2_prof: together they exceed what gpurun copies back); under gpurun_out/ land
#   r2_launches_final.csv                      launch list of the bench step (config 2), `--metrics gpu__time_duration.sum`
#   r2_ncu_final_summary.json, r2_traffic.json `ncu --set full` of the last of 3 steps at config 2 (+ the report itself)
#   r2_final_{bwd,fwd}_source.csv              per-SASS-instruction counts of the two compositors (tools/sass_mix.py)
#   r2_ncu_final_track_summary.json            a tracking step against a frozen model (frozen-model forward, pose-only backward)
#   r2_ncu_config4_densify_final_summary.json  a mapping loop at config-4 size with a densify_and_prune inside
set -x
OUT=${OUT:-/tmp/r2_prof}
mkdir -p gpurun_out $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_final.csv \
    python tools/one_step.py 500000 2 3 > $OUT/launches.log 2>&1
# step 1 launches 9 library kernels (no hints yet: scatter path), steps 2 and 3 launch 8
ncu --set full --clock-control none --import-source on -k regex:^k_ --launch-skip 17 -c 8 -f -o $OUT/r2_final \
    python tools/one_step.py 500000 2 3 > $OUT/final.log 2>&1
python tools/ncu_summarize.py $OUT/r2_final.ncu-rep gpurun_out/r2_ncu_final_summary.json gpurun_out/r2_traffic.json
ncu -i $OUT/r2_final.ncu-rep --page source --csv --kernel-name regex:k_composite_bwd > gpurun_out/r2_final_bwd_source.csv
ncu -i $OUT/r2_final.ncu-rep --page source --csv --kernel-name regex:k_composite_fwd > gpurun_out/r2_final_fwd_source.csv
cp $OUT/r2_final.ncu-rep gpurun_out/r2_final.ncu-rep
ncu --set full --clock-control none -k regex:^k_ --launch-skip 17 -c 8 -f -o $OUT/r2_final_track \
    python tools/one_step.py 500000 2 3 track_frozen > $OUT/track.log 2>&1
python tools/ncu_summarize.py $OUT/r2_final_track.ncu-rep gpurun_out/r2_ncu_final_track_summary.json
ncu --set full --clock-control none -k regex:^k_ -c 60 -f -o $OUT/r2_config4_densify_final \
    python tools/config4_densify_loop.py 4 3 > gpurun_out/r2_config4_densify_final.log 2>&1
python tools/ncu_summarize.py $OUT/r2_config4_densify_final.ncu-rep gpurun_out/r2_ncu_config4_densify_final_summary.json
du -sh gpurun_out
