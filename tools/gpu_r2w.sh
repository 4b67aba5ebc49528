# round 2, call W (1 GPU): is the end-to-end number box noise or path dependent?  default / three-sweep forward / default again
set -x
for tag in a_default b_sweeps c_default d_recompute; do
  unset DYCON_FECL_FWD DYCON_FECL_BWD
  [ $tag = b_sweeps ] && export DYCON_FECL_FWD=sweeps
  [ $tag = d_recompute ] && export DYCON_FECL_BWD=recompute
  timeout 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/bench_r2w_$tag.json 2> gpurun_out/bench_r2w_$tag.err; echo "bench $tag rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2w_$tag.json'))
print('$tag step us', round(d['ms_per_step']*1e3,1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'h2d', round(d['e2e']['h2d_gbs_plain_copy'],1))
PY
done
nvidia-smi topo -m 2>/dev/null | head -8; lscpu | grep -E "Model name|NUMA node\(s\)|Socket" 
