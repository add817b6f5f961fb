set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sampler.py tests/test_damping.py -m gpu -x -q > gpurun_out/r2_pytestJ.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytestJ.log
tail -30 gpurun_out/r2_pytestJ.log | cut -c1-300
EPGPU_TRACE=1 timeout 900 python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/r2_cfg4_n1_E.json 2> gpurun_out/r2_cfg4_n1_E.err; echo "rc=$?"
grep "sampler" gpurun_out/r2_cfg4_n1_E.err | tail -12
tail -c 600 gpurun_out/r2_cfg4_n1_E.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_cfg4_n1_E.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['grad_evals_per_s'], d['roofline']['frac'])
h=d['health']
for k in ('df_used','update_attempts','sites_skipped','sites_rejected_rank0','max_rhat','mean_stepsize','sampling_s','kl_step','rhat_sites_median_p90_frac_gt_1p1'): print(k,h[k])
PY
