set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytestA.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestA.log
tail -8 gpurun_out/r2_pytestA.log
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_cfg4_n1_A.json 2> gpurun_out/r2_cfg4_n1_A.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_cfg4_n1_A.err
cut -c1-3000 gpurun_out/r2_cfg4_n1_A.json
