set -x
mkdir -p gpurun_out
python tools/host_jitter.py 2 > gpurun_out/r2_jitter_2gpu.json; cat gpurun_out/r2_jitter_2gpu.json
timeout 900 python -m pytest tests/test_slabs_nccl.py -m gpu -q > gpurun_out/r2_final_nccl_tests.txt 2>&1
tail -6 gpurun_out/r2_final_nccl_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_final_bench_n2.json 2> gpurun_out/r2_final_bench_n2.err
tail -2 gpurun_out/r2_final_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --volume 2048 --no-e2e > gpurun_out/r2_final_bench_n2_v2048.json 2> gpurun_out/r2_final_bench_n2_v2048.err
tail -2 gpurun_out/r2_final_bench_n2_v2048.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_final_bench_n2.json','gpurun_out/r2_final_bench_n2_v2048.json'):
    try:
        p=json.load(open(f))
        print(f, round(p['value'],1), round(p['ms_per_step'],2), p['stage_ms'], p['config'].get('parity'), p['config']['known_answer']['match'])
        print('   e2e', p['e2e'])
    except Exception as e: print(f, 'ERR', e)
PY
CUDA_VISIBLE_DEVICES=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_smooth3|k_threshold' -s 2 -c 2 -o gpurun_out/r2_final_full_smooth python tools/profile_step.py 1024 2 > gpurun_out/r2_final_full_smooth.log 2>&1
tail -2 gpurun_out/r2_final_full_smooth.log
python tools/host_jitter.py 2
