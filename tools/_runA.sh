set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_experiment.py tests/test_gpu_linalg.py -m gpu -x -q > gpurun_out/r2_pytestK.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestK.log
tail -30 gpurun_out/r2_pytestK.log | cut -c1-300
EPGPU_TRACE=1 timeout 900 python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/r2_cfg4_n1_F.json 2> gpurun_out/r2_cfg4_n1_F.err; echo "rc=$?"
grep "sampler" gpurun_out/r2_cfg4_n1_F.err | tail -4
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_cfg4_n1_F.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['grad_evals_per_s'], d['roofline'])
PY
