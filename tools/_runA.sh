set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_linalg.py tests/test_gpu_experiment.py -m gpu -x -q -k "mix_phi or oracle_posterior or consensus" > gpurun_out/r2_pytestI.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestI.log
tail -30 gpurun_out/r2_pytestI.log | cut -c1-300
EPGPU_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "sampler" | tail -8
EPGPU_TRACE=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sites 148 2>&1 >/dev/null | grep "sampler" | tail -4
