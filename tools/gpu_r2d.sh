# round 2, call D (1 GPU): tests, then ncu --set full of the FeCL sweeps / backward, sorted+classes vs unsorted
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2d.log
B="python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_r2d.log 2>&1 || { tail -20 gpurun_out/plain_r2d.log; exit 1; }
DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_bwd|fecl_tc_sweep|fecl_rank|pack16" -s 40 -c 6 -o gpurun_out/prof_r2d_sorted $B > gpurun_out/ncu_r2d_sorted.log 2>&1
DYCON_FECL_SORT=0 DYCON_NO_PDL=1 ncu --set full --clock-control none --import-source on -k regex:"fecl_tc_bwd|fecl_tc_sweep|pack16" -s 35 -c 5 -o gpurun_out/prof_r2d_unsorted $B > gpurun_out/ncu_r2d_unsorted.log 2>&1
tail -3 gpurun_out/ncu_r2d_sorted.log gpurun_out/ncu_r2d_unsorted.log
ls -la gpurun_out/*.ncu-rep
