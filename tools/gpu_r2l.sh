# round 2, call L (1 GPU): stored-pairs backward -- FeCL parity tests, smoke, bench A/B against the recomputing backward
set -x
timeout 600 python -m pytest tests/test_gpu_fecl.py tests/test_gpu_module.py -m gpu -x -q > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2l.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2l.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2l.log | cut -c1-300
for mode in stored recompute; do
  [ $mode = recompute ] && export DYCON_FECL_BWD=recompute
  timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r2l_$mode.json 2> gpurun_out/bench_r2l_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2l_$mode.json'))
    print('$mode value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_r2l_$mode.err').read()[-3000:])
PY
done
