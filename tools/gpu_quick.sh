# quick GPU check: parity tests, then the bench line (usage: bash tools/gpu_quick.sh <tag> [pytest -k expr])
tag=${1:-q}; shift
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$tag.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$tag.json'))
    print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
    print('e2e', d['e2e']['value']/1e9, d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value']/1e6, 'Mvox/s')
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_$tag.err').read()[-2000:])
PY
