# round 2, call AJ (1 GPU): backward GEMM -- the two column splits of a row block as a cluster of two CTAs, reduction through DSMEM
set -x
timeout 120 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "(test_golden and fp16) or reproducible" 2>&1 | tail -5
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py > gpurun_out/spans_r2aj.md 2> gpurun_out/spans_r2aj.err; echo rc=$?
grep "GEMM" gpurun_out/spans_r2aj.md
timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/aj.json 2> gpurun_out/aj.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/aj.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()}, 'launches', d['gpu_launches'])
PY
tail -5 gpurun_out/aj.err
