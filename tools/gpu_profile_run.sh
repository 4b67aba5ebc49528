# full profile pass (usage: bash tools/gpu_profile_run.sh <tag>): bench line, launch list, --set full captures
tag=${1:-r1}
set -x
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
$B > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1
true
ncu --set full --clock-control none --import-source on -k regex:"ema_" -s 3 -c 1 -o gpurun_out/prof_${tag}_ema $B > gpurun_out/ncu_e_$tag.log 2>&1
tail -c 600 gpurun_out/bench_$tag.json
