mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_linalg.py tests/test_damping.py -m gpu -x -q 2>&1 | tail -3
rm -f gpurun_out/r2_linalg_timing.txt
for shape in "1024 50 800" "4096 20 800" "64 20 800" "512 100 800" "256 200 3200"; do python tools/linalg_workload.py $shape time >> gpurun_out/r2_linalg_timing.txt 2>&1; done
cat gpurun_out/r2_linalg_timing.txt
