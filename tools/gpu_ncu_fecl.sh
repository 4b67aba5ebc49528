# --set full capture of the FeCL kernels only (usage: bash tools/gpu_ncu_fecl.sh <tag> [kernel regex] [count])
tag=${1:-x}; rx=${2:-fecl_tc}; cnt=${3:-4}
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_$tag.log 2>&1 || { tail -20 gpurun_out/plain_$tag.log; exit 1; }
DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s ${4:-16} -c $cnt -o gpurun_out/prof_$tag $B > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
