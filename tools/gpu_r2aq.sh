# round 2, call AQ (1 GPU): backward GEMM read-out with 16-byte vector reductions (warp-tile transpose)
set -x
timeout 150 python -m pytest tests/test_gpu_fecl.py -x -q -m gpu -k "((test_golden or seeded or ragged) and fp16) or reproducible or isles22_sample" 2>&1 | tail -4
DYCON_SO_VARIANT=timeline timeout 60 python tools/spans.py > gpurun_out/spans_r2aq.md 2> gpurun_out/spans_r2aq.err; echo rc=$?
grep "GEMM" gpurun_out/spans_r2aq.md | head -6; grep -A 12 "tiles | with Gc" gpurun_out/spans_r2aq.md | tail -4
timeout 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/aq.json 2> gpurun_out/aq.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/aq.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
