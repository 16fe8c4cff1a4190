set -x
mkdir -p gpurun_out
nproc; nvidia-smi topo -m | head -8
run() { name=$1; shift; env "$@" B2M_SYNC_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 296$((RANDOM % 90 + 10)) bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --no-verify > gpurun_out/r2_diag_$name.json 2> gpurun_out/r2_diag_$name.err; grep "sync trace" gpurun_out/r2_diag_$name.err | tail -2; python -c "
import json;p=json.load(open('gpurun_out/r2_diag_$name.json'));print('$name',round(p['ms_per_step'],2),p['stage_ms'])"; }
run watch X=1
run poll B2M_SYNC_WATCH=0
run streamsync B2M_SYNC_WATCH=0 B2M_STREAM_POLL=0
run nccl_scalars B2M_SCALARS_NCCL=1
run noasync B2M_HALO_ASYNC=0
run notma B2M_SMOOTH_TMA=0
