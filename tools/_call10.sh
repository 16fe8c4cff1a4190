set -x
mkdir -p gpurun_out
python tools/host_jitter.py 2 > gpurun_out/r2_jitter_1gpu.json; cat gpurun_out/r2_jitter_1gpu.json
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_f.txt 2>&1
tail -5 gpurun_out/r2_gputests_f.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_bench_f.json'))
print(p['value'], p['ms_per_step'], p['stage_ms'])
print(p['roofline']['kernels_ms_per_step'])
PY
python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=4 python tools/e2e_probe.py 1024 4
python tools/host_jitter.py 2
