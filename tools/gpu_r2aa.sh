# round 2, call AA (1 GPU): records of the final build -- bench line (both arms), timelines, in-stream kernel durations
set -x
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2aa_reference.json 2> gpurun_out/bench_r2aa_reference.err; echo "reference rc=$?"
timeout 300 python tools/kernel_times.py --steps 30 > gpurun_out/kernel_times_r2aa.md 2> gpurun_out/kernel_times_r2aa.err; echo "ktimes rc=$?"
DYCON_SO_VARIANT=timeline timeout 300 python tools/timeline.py > gpurun_out/timeline_r2aa.md 2> gpurun_out/timeline_r2aa.err; echo "timeline rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2aa.json'))
print('value', d['value']/1e9, 'Gvox/s  ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['repetitions_ms_per_step'], 'allocs', d['e2e']['device_allocs_in_timed_regions'], 'cpu', d['cpu_baseline']['value']/1e6)
for k,v in d['roofline_all'].items(): print(' ', k, round(v['avg_ms']*1e3,1),'us frac', round(v['frac'],3), 'traffic', v.get('traffic'))
PY
