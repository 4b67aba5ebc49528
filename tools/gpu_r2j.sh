# round 2, call J (1 GPU): ncu evidence of the final build -- launch list, --set full of the step kernels, DRAM traffic
# per family from a RANGE replay (40 rotating launches per range, so write-backs are evicted and counted)
set -x
B="python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r2j.log 2>&1 || { tail -20 gpurun_out/plain_r2j.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2j.csv $B > gpurun_out/ncu_l_r2j.log 2>&1
DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_|pack16|uncl_|zero_fill" -s 30 -c 9 -o gpurun_out/prof_r2j_step $B > gpurun_out/ncu_f_r2j.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ema_" -s 3 -c 1 -o gpurun_out/prof_r2j_ema $B > gpurun_out/ncu_e_r2j.log 2>&1
P="python bench.py --profile-ranges"
DYCON_NO_PDL=1 $P > gpurun_out/ranges_plain_r2j.log 2>&1 && DYCON_NO_PDL=1 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ranges_r2j.csv $P > gpurun_out/ncu_r_r2j.log 2>&1
tail -3 gpurun_out/ncu_r_r2j.log; head -30 gpurun_out/ranges_r2j.csv | cut -c1-300
