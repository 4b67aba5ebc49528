# launch list (per-kernel durations) of a short eager bench run (usage: bash tools/gpu_ncu_list.sh <tag>)
tag=${1:-x}
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_$tag.csv
