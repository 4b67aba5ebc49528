# round 2, call AP (N GPUs): config 2 sharded, final build (usage: gpurun --gpus N -- 'bash tools/gpu_r2ap.sh N')
N=${1:-4}
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547"
timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 10 --no-parity-oracle > gpurun_out/bench${N}_r2ap_brats.json 2> gpurun_out/bench${N}_r2ap_brats.err; echo "brats rc=$?"
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench${N}_r2ap_brats.json'))
    print('brats value', round(d['value']/1e9,2), 'Gvox/s us/step', round(d['ms_per_step']*1e3,1), 'launches', d['gpu_launches'], 'e2e', d['e2e']['ms_per_step'], 'parity', json.dumps(d.get('parity')))
except Exception as e:
    print('failed', e); print(open('gpurun_out/bench${N}_r2ap_brats.err').read()[-1500:])
PY
