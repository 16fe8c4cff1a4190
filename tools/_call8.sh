set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_e.txt 2>&1
tail -5 gpurun_out/r2_gputests_e.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_bench_e.json'))
print(p['value'], p['ms_per_step'], p['stage_ms'])
print(p['roofline']['kernels_ms_per_step'])
PY
{
python tools/e2e_probe.py 1024 4
B2M_RING_CHUNK_KB=2048 python tools/e2e_probe.py 1024 4
B2M_RING_CHUNK_KB=1024 python tools/e2e_probe.py 1024 4
B2M_RING_CHUNK_KB=2048 B2M_RING_SLOTS=6 python tools/e2e_probe.py 1024 4
B2M_RING_CHUNK_KB=4096 B2M_RING_SLOTS=8 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=4 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=4 B2M_RING_CHUNK_KB=2048 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=8 B2M_RING_CHUNK_KB=2048 python tools/e2e_probe.py 1024 4
B2M_H2D_OVERLAP=0 python tools/e2e_probe.py 1024 4
} > gpurun_out/r2_e2e_probe.txt 2>&1
cat gpurun_out/r2_e2e_probe.txt
