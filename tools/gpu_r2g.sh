# round 2, call G (1 GPU): full GPU suite (f1/f2/f4 + sorted subprocess), train-step harness (small, then config 3), bench
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_r2g.log | cut -c1-300
python tools/train_step.py --patch 32,32,32 --batch 4 --labeled-bs 2 --steps 6 --timed 3 > gpurun_out/train_small.json 2> gpurun_out/train_small.err; echo "small rc=$?"; tail -3 gpurun_out/train_small.err; cut -c1-1500 gpurun_out/train_small.json
python tools/train_step.py --shape pancreas --batch 8 --steps 50 --timed 15 --json gpurun_out/train_step_pancreas.json > /dev/null 2> gpurun_out/train_pancreas.err; echo "pancreas rc=$?"; tail -3 gpurun_out/train_pancreas.err
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2g.json'))
print('ms/step', round(d['ms_per_step']*1e3,1), {k:(round(v['avg_ms']*1e3,1), round(v['frac'],3)) for k,v in d['roofline_all'].items()}, 'e2e', d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value']/1e6)
try:
    t=json.load(open('gpurun_out/train_step_pancreas.json')); print(json.dumps(t['timing'])); print(json.dumps(t['trajectory']))
except Exception as e: print('train json', e)
PY
