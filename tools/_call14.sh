set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
run() { name=$1; shift; env "$@" timeout 400 $TR --master-port 297$((RANDOM % 90 + 10)) bench.py --gpus 8 --steps 3 --warmup 3 --no-verify --no-cpu > gpurun_out/r2_e2e8_$name.json 2> gpurun_out/r2_e2e8_$name.err; python -c "
import json;p=json.load(open('gpurun_out/r2_e2e8_$name.json'));e=p['e2e'];print('$name',round(p['ms_per_step'],2),round(e['ms_per_step'],1),e['breakdown']['h2d_ms'],e['breakdown']['device_ms'],e['breakdown']['d2h_ms'],e['copy_threads']);print(e['breakdown']['per_rank_h2d_device_d2h_ms'])"; }
run default X=1
run ring2m4 B2M_RING_CHUNK_KB=2048 B2M_RING_SLOTS=4
run ring1m8 B2M_RING_CHUNK_KB=1024 B2M_RING_SLOTS=8
run thr2 B2M_COPY_THREADS=2
run noprefault B2M_PREFAULT=0
