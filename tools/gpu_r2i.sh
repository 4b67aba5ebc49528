# round 2, call I (2 GPUs): multi-rank check (exchange, sharded modules, graph replay, global negatives) and the three
# sharded configurations WITH the fp64 CPU oracle on the concatenated batch
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 300 $TR tools/multi_gpu_check.py > gpurun_out/mgc2_r2i.log 2>&1; echo "mgc rc=$?"; tail -3 gpurun_out/mgc2_r2i.log
timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 10 > gpurun_out/bench2_brats.json 2> gpurun_out/bench2_brats.err; echo "brats rc=$?"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --shape isles22 --batch 4 --no-e2e > gpurun_out/bench2_isles22.json 2> gpurun_out/bench2_isles22.err; echo "isles rc=$?"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 --global-negatives --no-e2e > gpurun_out/bench2_gn.json 2> gpurun_out/bench2_gn.err; echo "gn rc=$?"
python - <<PY
import json
for tag in ("brats","isles22","gn"):
    try:
        d=json.load(open(f'gpurun_out/bench2_{tag}.json'))
        print(tag, 'value', round(d['value']/1e9,2), 'Gvox/s ms/step', round(d['ms_per_step']*1e3,1), 'launches', d['gpu_launches'], 'parity', json.dumps(d.get('parity')))
    except Exception as e:
        print(tag, 'failed', e); print(open(f'gpurun_out/bench2_{tag}.err').read()[-1500:])
PY
