# round 2, call T (1 GPU): polynomial pass 2 of the row kernel -- quick parity, bench, in-stream kernel times
set -x
timeout 400 python -m pytest tests/test_gpu_fecl.py -m gpu -x -q -k "golden or seeded or ragged or work_split or single_class or near_identical or unnormalised" > gpurun_out/pytest_r2t.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2t.log | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2t.json'))
    print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_r2t.err').read()[-3000:])
PY
timeout 300 python tools/kernel_times.py --steps 30 > gpurun_out/kernel_times_r2t.md 2> gpurun_out/kernel_times_r2t.err; echo rc=$?; head -12 gpurun_out/kernel_times_r2t.md | cut -c1-150
