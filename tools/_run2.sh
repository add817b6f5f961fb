set -x
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo "rc=$?"
tail -c 2500 gpurun_out/r2_scale_n$N.err
cat gpurun_out/r2_scale_n$N.json | cut -c1-600
