# round 2, call AL (2 GPUs): final build (L2 evict-first policy) -- multi-rank check (exchange, sharded modules, graph replay, global negatives) and
# config 2 sharded with the fp64 CPU oracle on the concatenated batch
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mgc2_r2al.log 2>&1; echo "mgc rc=$?"; tail -4 gpurun_out/mgc2_r2al.log
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 10 --no-parity-oracle > gpurun_out/bench2_r2al_brats.json 2> gpurun_out/bench2_r2al_brats.err; echo "brats rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench2_r2al_brats.json'))
    print('brats value', round(d['value']/1e9,2), 'Gvox/s us/step', round(d['ms_per_step']*1e3,1), 'launches', d['gpu_launches'], 'e2e', d['e2e']['ms_per_step'], 'parity', json.dumps(d.get('parity')))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench2_r2al_brats.err').read()[-1500:])
PY
