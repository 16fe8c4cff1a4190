set -x
mkdir -p gpurun_out
nproc; free -g | head -2
python tools/host_jitter.py 2 > gpurun_out/r2_jitter_8gpu.json; cat gpurun_out/r2_jitter_8gpu.json
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err
timeout 600 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 10 --warmup 3 --volume 2048 --no-e2e --no-verify > gpurun_out/r2_bench_n8_v2048.json 2> gpurun_out/r2_bench_n8_v2048.err
tail -3 gpurun_out/r2_bench_n8_v2048.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29713 bench.py --gpus 4 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
tail -3 gpurun_out/r2_bench_n4.err
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29714 bench.py --gpus 4 --steps 5 --warmup 3 --volume 2048 --no-e2e --no-verify > gpurun_out/r2_bench_n4_v2048.json 2> gpurun_out/r2_bench_n4_v2048.err
tail -3 gpurun_out/r2_bench_n4_v2048.err
python - <<'PY'
import json
for f in ('n8','n8_v2048','n4','n4_v2048'):
    try:
        p=json.load(open(f'gpurun_out/r2_bench_{f}.json'))
        print(f, round(p['value'],1), round(p['ms_per_step'],2), p['stage_ms'], p['config'].get('parity'), p['config']['known_answer'] and p['config']['known_answer']['match'])
        print('   e2e', p['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
python tools/host_jitter.py 2
