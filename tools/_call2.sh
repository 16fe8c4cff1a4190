set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_a.txt 2>&1
tail -15 gpurun_out/r2_gputests_a.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -3 gpurun_out/r2_bench_a.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_bench_a.json'))
print(p['value'], p['ms_per_step'], p['stage_ms'])
print(p['roofline']['kernels_ms_per_step'])
print(p['e2e'])
PY
