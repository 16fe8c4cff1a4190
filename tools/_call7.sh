set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_d.txt 2>&1
tail -8 gpurun_out/r2_gputests_d.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
B2M_SMOOTH_TMA=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_d_tma.json 2> gpurun_out/r2_bench_d_tma.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_d.json','gpurun_out/r2_bench_d_tma.json'):
    p=json.load(open(f))
    print(f, p['value'], p['ms_per_step'], p['stage_ms'])
    print(p['roofline']['kernels_ms_per_step'])
PY
