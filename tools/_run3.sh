set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sampler.py tests/test_damping.py -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -15 gpurun_out/r2_pytest3.log
python __graft_entry__.py smoke 2>&1 | tail -3
python tools/diag_sites.py cfg4 2>&1 | tail -4
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_cfg4_n1_short.json 2> gpurun_out/r2_cfg4_n1_short.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_cfg4_n1_short.err
