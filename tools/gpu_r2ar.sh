# round 2, call AR (1 GPU): pack kernel with a swizzled tile, 16-byte shared-memory stores and 16-byte global stores
set -x
timeout 100 python -m pytest tests/test_gpu_fecl.py tests/test_gpu_prep.py tests/test_gpu_fecl_global.py -x -q -m gpu -k "test_golden or prep or virtual or ragged or label_sorted" 2>&1 | tail -4
DYCON_SO_VARIANT=timeline timeout 60 python tools/spans.py > gpurun_out/spans_r2ar.md 2> gpurun_out/spans_r2ar.err; echo rc=$?
grep "pack16\|sweep | 140 | 140 | main loop reached" gpurun_out/spans_r2ar.md | head -5
timeout 120 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/ar.json 2> gpurun_out/ar.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ar.json'))
print('ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
PY
