# round 2, call AE (1 GPU): 128-row tiles for the similarity sweep (4-stage ring) -- DYCON_FECL_RT=1
set -x
export DYCON_FECL_RT=1
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py > gpurun_out/spans_r2ae_rt1.md 2> gpurun_out/spans_r2ae.err; echo rc=$?
grep "sweep\|row kernel" gpurun_out/spans_r2ae_rt1.md
timeout 300 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "golden or seeded" 2>&1 | tail -3
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ae_rt1.json 2> gpurun_out/ae_rt1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ae_rt1.json'))
print('rt1 ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
