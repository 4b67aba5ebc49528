# round 2, call Q (1 GPU): carve-out hint A/B (bench), --set full of the similarity sweep and the row kernel
set -x
for mode in off on; do
  [ $mode = on ] && export DYCON_CARVEOUT=1
  timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e > gpurun_out/bench_r2q_$mode.json 2> gpurun_out/bench_r2q_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2q_$mode.json'))
    print('$mode value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
    for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_r2q_$mode.err').read()[-3000:])
PY
done
unset DYCON_CARVEOUT
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
DYCON_NO_PDL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_sweep|fecl_row_pairs" -s 8 -c 2 -o gpurun_out/prof_r2q_fwd $B > gpurun_out/ncu_f_r2q.log 2>&1; tail -2 gpurun_out/ncu_f_r2q.log
