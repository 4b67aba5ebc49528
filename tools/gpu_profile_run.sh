# Full profile pass on one B200 (usage: gpurun -- 'bash tools/gpu_profile_run.sh <tag>'):
#   bench line, ncu launch list, --set full captures of every kernel of the step and of the EMA kernel.
# Summaries for profiles/ are then made locally with tools/ncu_summary.py (launches / full).
tag=${1:-r1}
set -x
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
$B > gpurun_out/plain_$tag.log 2>&1 || { tail -20 gpurun_out/plain_$tag.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1
# DYCON_NO_PDL=1: ncu serialises kernels anyway, and its kernel replay is happier without programmatic launches
DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_sweep" -s 15 -c 3 -o gpurun_out/prof_${tag}_sweeps $B > gpurun_out/ncu_f_${tag}_sweeps.log 2>&1
DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_bwd|uncl_|pack16|zero_fill" -s 15 -c 5 -o gpurun_out/prof_${tag}_rest $B > gpurun_out/ncu_f_${tag}_rest.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ema_" -s 3 -c 1 -o gpurun_out/prof_${tag}_ema $B > gpurun_out/ncu_e_$tag.log 2>&1
grep -h "ERROR\|passes" gpurun_out/ncu_f_${tag}_*.log gpurun_out/ncu_e_$tag.log | head
tail -c 400 gpurun_out/bench_$tag.json
