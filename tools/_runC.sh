set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_linalg.py -m gpu -x -q > gpurun_out/r2_pytestC.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestC.log
tail -15 gpurun_out/r2_pytestC.log
for shape in "1024 50 800" "4096 20 800" "64 20 800" "512 100 800"; do
  python tools/linalg_workload.py $shape time >> gpurun_out/r2_linalg_mma.txt 2>&1
  EPGPU_MOMENTS_SIMT=1 python tools/linalg_workload.py $shape time >> gpurun_out/r2_linalg_simt.txt 2>&1
done
cat gpurun_out/r2_linalg_mma.txt; cat gpurun_out/r2_linalg_simt.txt
timeout 900 python tools/diag_sites2.py cfg4 4 > gpurun_out/r2_diag2.txt 2>&1; tail -40 gpurun_out/r2_diag2.txt
