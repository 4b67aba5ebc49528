# round 2, call K (1 GPU): final-build record -- GPU suite, smoke, bench line, DRAM traffic per family (range replay)
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2k.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2k.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2k.log
P="python bench.py --profile-ranges"
DYCON_NO_PDL=1 $P > gpurun_out/ranges_plain_r2k.log 2>&1 && DYCON_NO_PDL=1 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ranges_r2k.csv $P > gpurun_out/ncu_r_r2k.log 2>&1
grep -v "^==" gpurun_out/ranges_r2k.csv | cut -c1-200 | tail -14
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2k_ref.json 2> gpurun_out/bench_r2k_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_r2k_ref.json
