set -x
mkdir -p gpurun_out
B2M_TEST_ABORT=1 timeout 700 python -m pytest tests/test_slabs_nccl.py -m gpu -x -q -k "over_nccl and not other or failed_rank or two_devices" > gpurun_out/r2_nccl_tests2.txt 2>&1
tail -8 gpurun_out/r2_nccl_tests2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -3 gpurun_out/r2_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --volume 2048 --no-e2e > gpurun_out/r2_bench_n2_v2048.json 2> gpurun_out/r2_bench_n2_v2048.err
tail -3 gpurun_out/r2_bench_n2_v2048.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n2.json','gpurun_out/r2_bench_n2_v2048.json'):
    try:
        p=json.load(open(f))
        print(f, p['value'], p['ms_per_step'], p['stage_ms'], p['config'].get('parity'), p['config']['known_answer']['match'])
        print(p['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
B2M_COPY_THREADS=4 python tools/e2e_probe.py 1024 3
python tools/e2e_probe.py 1024 3
