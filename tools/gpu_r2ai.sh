# round 2, call AI (1 GPU): UnCL logits fetched with an L2 evict-first policy (FeCL's pair matrices stay in the L2 for the backward)
set -x
timeout 200 python -m pytest tests/test_gpu_uncl.py -x -q -m gpu 2>&1 | tail -3
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py > gpurun_out/spans_r2ai.md 2> gpurun_out/spans_r2ai.err; echo rc=$?
grep "GEMM" gpurun_out/spans_r2ai.md; tail -14 gpurun_out/spans_r2ai.md | head -11
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ai.json 2> gpurun_out/ai.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ai.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
