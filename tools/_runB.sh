set -x
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q > gpurun_out/r2_pytestB.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestB.log
tail -12 gpurun_out/r2_pytestB.log | cut -c1-400
for wl in cfg3 cfg4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 4 --workload $wl > gpurun_out/r2_${wl}_n$N.json 2> gpurun_out/r2_${wl}_n$N.err; echo "rc=$?"
tail -c 600 gpurun_out/r2_${wl}_n$N.err
grep "^{" gpurun_out/r2_${wl}_n$N.json | cut -c1-300
done
timeout 600 python bench.py --steps 6 --warmup 4 --workload cfg3 > gpurun_out/r2_cfg3_n1.json 2> gpurun_out/r2_cfg3_n1.err; echo "rc=$?"
grep "^{" gpurun_out/r2_cfg3_n1.json | cut -c1-300
