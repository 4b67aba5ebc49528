# round 2, call A (1 GPU): full GPU test suite, parity report, bench
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r2a.log
python tools/parity_report.py --json gpurun_out/parity_r2a.json > gpurun_out/parity_r2a.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/parity_r2a.log
python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2a.json
