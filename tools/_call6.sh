set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_c.txt 2>&1
tail -12 gpurun_out/r2_gputests_c.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_bench_c.json'))
print(p['value'], p['ms_per_step'], p['stage_ms'])
print(p['roofline']['kernels_ms_per_step'])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cc_local|k_cc_border|k_mc_classify|k_tri|k_compact' -s 4 -c 7 -o gpurun_out/r2c_full python tools/profile_step.py 1024 2 > gpurun_out/r2c_full.log 2>&1
tail -2 gpurun_out/r2c_full.log
timeout 300 python tools/bench_atlas.py --workers 1,8,16 --steps 3 > gpurun_out/r2_atlas.json 2> gpurun_out/r2_atlas.err; tail -2 gpurun_out/r2_atlas.err; cat gpurun_out/r2_atlas.json | cut -c1-1500
