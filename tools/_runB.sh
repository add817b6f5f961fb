set -x
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_sampler.py -k "multirank or nccl or reinit" -m gpu -x -q > gpurun_out/r2_pytestB.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestB.log
tail -25 gpurun_out/r2_pytestB.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo "rc=$?"
tail -c 2500 gpurun_out/r2_scale_n$N.err
cut -c1-1500 gpurun_out/r2_scale_n$N.json
