# round 2, call V (1 GPU): evidence of the final build -- the whole GPU suite, smoke, bench (both arms), ncu launch list,
# --set full of the step kernels, DRAM traffic by range replay, timelines, in-stream kernel durations
set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2v.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2v.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2v.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2v.log | cut -c1-300
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2v.json 2> gpurun_out/bench_r2v.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2v_reference.json 2> gpurun_out/bench_r2v_reference.err; echo "reference rc=$?"
timeout 300 python tools/kernel_times.py --steps 30 > gpurun_out/kernel_times_r2v.md 2> gpurun_out/kernel_times_r2v.err; echo "ktimes rc=$?"
DYCON_SO_VARIANT=timeline timeout 300 python tools/timeline.py > gpurun_out/timeline_r2v.md 2> gpurun_out/timeline_r2v.err; echo "timeline rc=$?"
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2v.csv $B > gpurun_out/ncu_l_r2v.log 2>&1
DYCON_NO_PDL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_|fecl_row|pack16|uncl_|zero_fill" -s 24 -c 8 -o gpurun_out/prof_r2v_step $B > gpurun_out/ncu_f_r2v.log 2>&1; tail -2 gpurun_out/ncu_f_r2v.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ema_" -s 3 -c 1 -o gpurun_out/prof_r2v_ema $B > gpurun_out/ncu_e_r2v.log 2>&1
P="python bench.py --profile-ranges"
DYCON_NO_PDL=1 $P > gpurun_out/ranges_plain_r2v.log 2>&1 && DYCON_NO_PDL=1 timeout 900 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ranges_r2v.csv $P > gpurun_out/ncu_r_r2v.log 2>&1
tail -3 gpurun_out/ncu_r_r2v.log; head -12 gpurun_out/ranges_r2v.csv | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2v.json'))
print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'e2e', d.get('e2e',{}).get('value',0)/1e9, 'cpu', d.get('cpu_baseline',{}).get('value',0)/1e6)
for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3))
PY
