set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytestG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestG.log
tail -12 gpurun_out/r2_pytestG.log | cut -c1-300
