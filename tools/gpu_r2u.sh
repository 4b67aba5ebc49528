# round 2, call U (1 GPU): row-kernel launch-bound variants (threads x resident CTAs): default 256x2, v1 128x3, v2 256x1, v3 128x4
set -x
timeout 400 python -m pytest tests/test_gpu_fecl.py -m gpu -x -q -k "golden or seeded or ragged or work_split or single_class or near_identical or unnormalised" > gpurun_out/pytest_r2u.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2u.log | cut -c1-400
for v in "" v1 v2 v3; do
  export DYCON_SO_VARIANT=$v
  timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2u_$v.json 2> gpurun_out/bench_r2u_$v.err; echo "bench [$v] rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2u_$v.json'))
    print('[$v] ms/step', d['ms_per_step'], ' '.join(f"{k}={v['avg_ms']*1e3:.1f}" for k,v in d['roofline_all'].items()))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_r2u_$v.err').read()[-2000:])
PY
done
