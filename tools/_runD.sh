set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py tests/test_damping.py -m gpu -x -q > gpurun_out/r2_pytestD.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestD.log
tail -15 gpurun_out/r2_pytestD.log
rm -f gpurun_out/r2_linalg_row.txt
for shape in "1024 50 800" "4096 20 800" "64 20 800" "512 100 800"; do
  python tools/linalg_workload.py $shape time >> gpurun_out/r2_linalg_row.txt 2>&1
done
cat gpurun_out/r2_linalg_row.txt
ncu --metrics gpu__time_duration.sum,smsp__pipe_tensor_subpipe_dmma_cycles_active.avg,sm__cycles_active.avg --clock-control none --csv --log-file gpurun_out/r2_linalg_launches.csv python tools/linalg_workload.py 1024 50 800 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2_linalg_launches.csv | cut -d, -f5,13-15 | tail -40
