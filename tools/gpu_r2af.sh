# round 2, call AF (1 GPU): K-chunk-granular operand ring of the similarity sweep
set -x
timeout 200 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "test_golden and fp16" 2>&1 | tail -3
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py > gpurun_out/spans_r2af.md 2> gpurun_out/spans_r2af.err; echo rc=$?
grep "sweep\|row kernel" gpurun_out/spans_r2af.md
DYCON_SO_VARIANT=timeline timeout 100 python tools/timeline.py > gpurun_out/timeline_r2af.md 2> gpurun_out/timeline_r2af.err; echo rc=$?
head -20 gpurun_out/timeline_r2af.md
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/af.json 2> gpurun_out/af.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/af.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
