#!/bin/bash
# ncu captures for profiles/ (run on the GPU box through gpurun; one GPU).
#   bash tools/profile_round.sh r01
set -u
R=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
# 1. launch list of the default bench command (cold-cache, serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $OUT/launches_${R}_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_${R}_cfg3.log 2>&1
# 2. the sampler kernel (dominant) with full sections + source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nuts -s 1 -c 1 \
    -o $OUT/prof_${R}_nuts_cfg4 python tools/sampler_workload.py cfg4 148 40 > $OUT/prof_${R}_nuts.log 2>&1
# 3. moment / update / cavity kernels at the config-4 state shapes (K=1024, d=50, n=800)
timeout 900 ncu --set full --clock-control none -k regex:'k_moments|k_update_partial|k_cavity|k_sum_chunks' -c 8 \
    -o $OUT/prof_${R}_linalg_cfg4 python tools/linalg_workload.py 1024 50 800 > $OUT/prof_${R}_linalg.log 2>&1
python tools/linalg_workload.py 1024 50 800 time > $OUT/linalg_timing_${R}.txt 2>&1
python tools/linalg_workload.py 64 20 800 time >> $OUT/linalg_timing_${R}.txt 2>&1
python tools/linalg_workload.py 256 200 3200 time >> $OUT/linalg_timing_${R}.txt 2>&1
ls -la $OUT
