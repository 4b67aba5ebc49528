# round 2, call C (1 GPU): label-sorted FeCL -- parity suite, then bench A/B (sorted+classes | sorted only | unsorted)
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2c.log
B="python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e"
$B > gpurun_out/bench_r2c_sorted.json 2> gpurun_out/bench_r2c_sorted.err; echo "bench rc=$?"
DYCON_FECL_CLASSES=0 $B > gpurun_out/bench_r2c_noclasses.json 2> /dev/null
DYCON_FECL_SORT=0 $B > gpurun_out/bench_r2c_unsorted.json 2> /dev/null
python - <<PY
import json
for tag in ("sorted","noclasses","unsorted"):
    try:
        d=json.load(open(f'gpurun_out/bench_r2c_{tag}.json'))
        print(tag, 'ms/step', round(d['ms_per_step']*1e3,1), {k:round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()}, 'launches', d['gpu_launches'], 'loss', d['config']['loss_check'])
    except Exception as e: print(tag, 'failed', e)
PY
tail -5 gpurun_out/bench_r2c_sorted.err
