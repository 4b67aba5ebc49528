# round 2, call Z (8 GPUs): final build -- weak scaling of config 2 at N = 8 and config 4 (ISLES22, global B = 8, one sample per GPU),
# parity against the unsharded kernels on rank 0's GPU
set -x
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $TR bench.py --gpus $N --steps 50 --warmup 10 --no-parity-oracle > gpurun_out/bench${N}_r2z_brats.json 2> gpurun_out/bench${N}_r2z_brats.err; echo "brats rc=$?"
timeout 400 $TR bench.py --gpus $N --steps 30 --warmup 5 --shape isles22 --batch $((8 / N)) --no-parity-oracle --no-e2e > gpurun_out/bench${N}_r2z_isles22.json 2> gpurun_out/bench${N}_r2z_isles22.err; echo "isles rc=$?"
python - <<PY
import json
for tag in ("brats","isles22"):
    try:
        d=json.load(open(f'gpurun_out/bench${N}_r2z_{tag}.json'))
        print(tag, 'value', round(d['value']/1e9,2), 'Gvox/s us/step', round(d['ms_per_step']*1e3,1), 'launches', d['gpu_launches'], 'parity', d.get('parity'))
    except Exception as e:
        print(tag, 'failed', e); print(open(f'gpurun_out/bench${N}_r2z_{tag}.err').read()[-1500:])
PY
