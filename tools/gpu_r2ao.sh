# round 2, call AO (1 GPU): the working tree's product library once more (golden vectors, bench), and a DIAGNOSTIC build of the
# similarity sweep without its Gc stores (wrong results; how much of the sweep is the L2 write path?)
set -x
timeout 120 python -m pytest tests/test_gpu_fecl.py tests/test_gpu_uncl.py -x -q -m gpu -k "test_golden" 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ao.json 2> gpurun_out/ao.err; echo "bench rc=$?"
DYCON_SO_VARIANT=nogc timeout 100 python tools/spans.py > gpurun_out/spans_r2ao_nogc.md 2> gpurun_out/spans_r2ao.err; echo rc=$?
grep "sweep" gpurun_out/spans_r2ao_nogc.md
python - <<'PY'
import json
d=json.load(open('gpurun_out/ao.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
