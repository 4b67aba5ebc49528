# round 2, call AG (1 GPU): row kernel with two warps per row (16 warps per SM)
set -x
export DYCON_SO_VARIANT=rows2
timeout 200 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "test_golden and fp16" 2>&1 | tail -3
DYCON_SO_VARIANT=rows2tl timeout 100 python tools/spans.py > gpurun_out/spans_r2ag.md 2> gpurun_out/spans_r2ag.err; echo rc=$?
grep "row kernel" gpurun_out/spans_r2ag.md
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ag.json 2> gpurun_out/ag.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ag.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
