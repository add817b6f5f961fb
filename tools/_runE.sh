set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_cavity_col|k_gram_mma|k_moments_tail_col" -c 4 -f -o gpurun_out/r2_linalg_prof python tools/linalg_workload.py 1024 50 800 > gpurun_out/r2_ncu_linalg.log 2>&1
tail -5 gpurun_out/r2_ncu_linalg.log
ls -la gpurun_out/*.ncu-rep
