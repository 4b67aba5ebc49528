# round 2, call H (8 GPUs): weak scaling at N=8 (config 2 x 8), config 4 (ISLES22 global B=8 sharded), config 5
# (global negatives, B=32 over 8 GPUs) -- sharded-vs-unsharded parity on the GPU, CPU oracle left to the 2-GPU call
set -x
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 50 --warmup 10 --no-parity-oracle > gpurun_out/bench${N}_brats.json 2> gpurun_out/bench${N}_brats.err; echo "brats rc=$?"
timeout 600 $TR bench.py --gpus $N --steps 30 --warmup 5 --shape isles22 --batch $((8 / N)) --no-parity-oracle --no-e2e > gpurun_out/bench${N}_isles22.json 2> gpurun_out/bench${N}_isles22.err; echo "isles rc=$?"
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --global-negatives --no-parity-oracle --no-e2e > gpurun_out/bench${N}_gn.json 2> gpurun_out/bench${N}_gn.err; echo "gn rc=$?"
python - <<PY
import json
for tag in ("brats","isles22","gn"):
    try:
        d=json.load(open(f'gpurun_out/bench${N}_{tag}.json'))
        print(tag, 'value', round(d['value']/1e9,2), 'Gvox/s ms/step', round(d['ms_per_step']*1e3,1), 'launches', d['gpu_launches'], 'parity', d.get('parity'))
    except Exception as e:
        print(tag, 'failed', e); print(open(f'gpurun_out/bench${N}_{tag}.err').read()[-1500:])
PY
