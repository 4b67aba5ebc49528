# round 2, call B (2 GPUs): multi-rank check of the fused exchange, bench at N=2 with the parity block
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mgc2_r2b.log 2>&1; echo "mgc rc=$?"; tail -5 gpurun_out/mgc2_r2b.log
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench2_r2b.json 2> gpurun_out/bench2_r2b.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench2_r2b.json'))
print('value', d['value']/1e9, 'ms/step', d['ms_per_step'], 'launches', d['gpu_launches'])
print('parity', d.get('parity'))
PY
tail -5 gpurun_out/bench2_r2b.err
