# round 2, call AB (1 GPU): stream-overlap experiment (UnCL branch on a second stream), row kernel at 128 registers
set -x
B="timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-e2e"
$B > gpurun_out/ab_base.json 2> gpurun_out/ab_base.err; echo "base rc=$?"
$B --overlap > gpurun_out/ab_overlap.json 2> gpurun_out/ab_overlap.err; echo "overlap rc=$?"
DYCON_SO_VARIANT=row2 $B > gpurun_out/ab_row2.json 2> gpurun_out/ab_row2.err; echo "row2 rc=$?"
DYCON_SO_VARIANT=row2 $B --overlap > gpurun_out/ab_row2_overlap.json 2> gpurun_out/ab_row2_overlap.err; echo "row2 overlap rc=$?"
python - <<'PY'
import json
for n in ('base','overlap','row2','row2_overlap'):
    try:
        d=json.load(open(f'gpurun_out/ab_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step']*1e3,2), 'us; loss', d['config']['loss_check'], {k: round(v['avg_ms']*1e3,1) for k,v in d['roofline_all'].items()})
    except Exception as e:
        print(n, 'failed', e)
PY
tail -3 gpurun_out/ab_overlap.err
