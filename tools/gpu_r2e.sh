# round 2, call E (1 GPU): fused SGD/EMA tests + device-clock timelines of P2 / backward (timeline build)
set -x
python -m pytest tests/test_gpu_sgd_ema.py tests/test_gpu_steplosses.py -q > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_r2e.log
python tools/timeline.py > gpurun_out/timeline_sorted.md 2> gpurun_out/timeline_sorted.err; echo "rc=$?"
python tools/timeline.py --unsorted > gpurun_out/timeline_unsorted.md 2> gpurun_out/timeline_unsorted.err; echo "rc=$?"
head -5 gpurun_out/timeline_sorted.md; tail -3 gpurun_out/timeline_sorted.err
