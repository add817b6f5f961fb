#!/bin/bash
# ncu captures for profiles/ (run on the GPU box through gpurun; one GPU).
#   bash tools/profile_round.sh r02
set -u
R=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
# 1. launch list of the default bench command (cold-cache, serialised: compare SHARES)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/launches_${R}_cfg4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/launches_${R}_cfg4.log 2>&1
# 2. the sampler kernel (dominant) of the SAME command: the 4th launch = the timed EP iteration; full sections
#    + source; dram__bytes of this capture are roofline.traffic of the bench line (profiles/traffic.json)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_nuts -s 3 -c 1 -f \
    -o $OUT/prof_${R}_nuts_cfg4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/prof_${R}_nuts.log 2>&1
# 3. moment / update / cavity kernels at the config-4 state shapes (K=1024, d=50, n=800)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gram_mma|k_moments_tail|k_update_partial|k_cavity|k_sum_chunks' -c 10 -f \
    -o $OUT/prof_${R}_linalg_cfg4 python tools/linalg_workload.py 1024 50 800 > $OUT/prof_${R}_linalg.log 2>&1
ls -la $OUT
