set -x
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r1_a.json 2> gpurun_out/bench_r1_a.err
$B > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1_a.csv $B > gpurun_out/ncu_l_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fecl_tc|uncl_|pack16" -s 32 -c 8 -o gpurun_out/prof_r1_a_step $B > gpurun_out/ncu_f_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ema_multi" -s 3 -c 1 -o gpurun_out/prof_r1_a_ema $B > gpurun_out/ncu_e_a.log 2>&1
tail -c 600 gpurun_out/bench_r1_a.json
