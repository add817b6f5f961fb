set -x
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2_scale_n$N.err
cut -c1-700 gpurun_out/r2_scale_n$N.json
