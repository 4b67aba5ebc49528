# round 2, call O (1 GPU): packed-fp32 row kernel -- quick parity, bench, launch list, timelines, --set full of the three FeCL kernels
set -x
timeout 400 python -m pytest tests/test_gpu_fecl.py -m gpu -x -q -k "golden or seeded or ragged or work_split or single_class or near_identical or unnormalised" > gpurun_out/pytest_r2o.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2o.log | cut -c1-400
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2o.json'))
    print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_r2o.err').read()[-3000:])
PY
DYCON_SO_VARIANT=timeline timeout 300 python tools/timeline.py > gpurun_out/timeline_r2o.md 2> gpurun_out/timeline_r2o.err; echo "timeline rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2o.csv $B > gpurun_out/ncu_l_r2o.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_r2o.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.defaultdict(list)
for r in rows[1:]:
    try: agg[r[ki][:70]].append(float(r[vi].replace(',', '')))
    except Exception: pass
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])): print(f'{k:72s} n={len(v):3d} avg={sum(v)/len(v)/1e3:8.2f} us')
PY
DYCON_NO_PDL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_sweep|fecl_row_pairs|fecl_tc_bwd_gemm" -s 12 -c 3 -o gpurun_out/prof_r2o_fecl $B > gpurun_out/ncu_f_r2o.log 2>&1; tail -2 gpurun_out/ncu_f_r2o.log
