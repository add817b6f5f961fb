set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
for damp in schedule auto auto1; do
  timeout 600 python bench.py --sites 256 --steps 10 --warmup 3 --no-cpu-baseline --damp $damp > gpurun_out/r2_k256_$damp.json 2> gpurun_out/r2_k256_$damp.err; echo "rc=$?"
  tail -c 1500 gpurun_out/r2_k256_$damp.err
done
