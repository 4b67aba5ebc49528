# round 2, call AM (1 GPU): similarity sweep with its A tiles in tensor memory (TS-form MMAs, five-stage ring)
set -x
timeout 200 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "(test_golden or seeded or ragged) and fp16" 2>&1 | tail -8
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py > gpurun_out/spans_r2am.md 2> gpurun_out/spans_r2am.err; echo rc=$?
grep "sweep\|row kernel" gpurun_out/spans_r2am.md
DYCON_SO_VARIANT=timeline timeout 100 python tools/timeline.py > gpurun_out/timeline_r2am.md 2> gpurun_out/timeline_r2am.err; echo rc=$?
head -18 gpurun_out/timeline_r2am.md
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/am.json 2> gpurun_out/am.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/am.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
