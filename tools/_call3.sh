set -x
mkdir -p gpurun_out
nvidia-smi -L; nproc
timeout 1500 python -m pytest tests/test_slabs_nccl.py -m gpu -x -q > gpurun_out/r2_nccl_tests.txt 2>&1
tail -25 gpurun_out/r2_nccl_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
tail -5 gpurun_out/r2_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --volume 2048 --no-e2e > gpurun_out/r2_bench_n2_v2048.json 2> gpurun_out/r2_bench_n2_v2048.err
tail -5 gpurun_out/r2_bench_n2_v2048.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_n2.json','gpurun_out/r2_bench_n2_v2048.json'):
    try:
        p=json.load(open(f))
        print(f, p['value'], p['ms_per_step'], p['stage_ms'], p['config'].get('parity'), p['config']['known_answer'])
        print(p['roofline']['kernels_ms_per_step'])
        print(p['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
