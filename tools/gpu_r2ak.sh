# round 2, call AK (1 GPU): evidence of the final build (UnCL logits with an L2 evict-first policy) -- the whole GPU suite, smoke,
# bench (both arms), in-stream kernel durations, per-CTA spans, ncu launch list, --set full of the UnCL kernels
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2ak.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2ak.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2ak.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2ak.log | cut -c1-300
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2ak.json 2> gpurun_out/bench_r2ak.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2ak_reference.json 2> gpurun_out/bench_r2ak_reference.err; echo "reference rc=$?"
timeout 300 python tools/kernel_times.py --steps 30 > gpurun_out/kernel_times_r2ak.md 2> gpurun_out/kernel_times_r2ak.err; echo "ktimes rc=$?"
DYCON_SO_VARIANT=timeline timeout 100 python tools/spans.py --row-detail > gpurun_out/spans_r2ak.md 2> gpurun_out/spans_r2ak.err; echo "spans rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2ak.csv $B > gpurun_out/ncu_l_r2ak.log 2>&1; echo "ncu list rc=$?"
DYCON_NO_PDL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"uncl_" -s 6 -c 2 -o gpurun_out/prof_r2ak_uncl $B > gpurun_out/ncu_f_r2ak.log 2>&1; tail -2 gpurun_out/ncu_f_r2ak.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2ak.json'))
print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('repetitions_ms_per_step'), 'cpu', d['cpu_baseline']['value']/1e6)
for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3), 'traffic', v.get('traffic'))
PY
