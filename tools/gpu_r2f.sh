# round 2, call F (1 GPU): elect.sync issue loops + new rank kernel: FeCL tests, bench A/B, timelines
set -x
python -m pytest tests/test_gpu_fecl.py tests/test_gpu_fecl_global.py tests/test_gpu_module.py -x -q > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2f.log
B="python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e"
$B > gpurun_out/bench_r2f_sorted.json 2> gpurun_out/bench_r2f_sorted.err; echo "bench rc=$?"
DYCON_FECL_CLASSES=0 $B > gpurun_out/bench_r2f_noclasses.json 2> /dev/null
DYCON_FECL_SORT=0 $B > gpurun_out/bench_r2f_unsorted.json 2> /dev/null
python - <<PY
import json
for tag in ("sorted","noclasses","unsorted"):
    try:
        d=json.load(open(f'gpurun_out/bench_r2f_{tag}.json'))
        print(tag, 'ms/step', round(d['ms_per_step']*1e3,1), {k:round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()}, 'launches', d['gpu_launches'], 'loss', d['config']['loss_check'])
    except Exception as e: print(tag, 'failed', e)
PY
DYCON_SO_VARIANT=timeline python tools/timeline.py > gpurun_out/timeline2_sorted.md 2> gpurun_out/timeline2_sorted.err; echo "rc=$?"
DYCON_SO_VARIANT=timeline python tools/timeline.py --unsorted > gpurun_out/timeline2_unsorted.md 2> /dev/null; echo "rc=$?"
