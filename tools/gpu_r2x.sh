# round 2, call X (1 GPU): packed-fp32 UnCL forward + batched row-kernel prologue -- UnCL / step-loss / FeCL parity, bench
set -x
timeout 600 python -m pytest tests/test_gpu_uncl.py tests/test_gpu_module.py tests/test_gpu_fecl.py -m gpu -x -q -k "not subprocess" > gpurun_out/pytest_r2x.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2x.log | cut -c1-300
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r2x.json 2> gpurun_out/bench_r2x.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2x.json'))
print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['repetitions_ms_per_step'], 'allocs', d['e2e']['device_allocs_in_timed_regions'])
for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
PY
