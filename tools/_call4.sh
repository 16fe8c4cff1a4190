set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_b.txt 2>&1
tail -15 gpurun_out/r2_gputests_b.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_b_tma.json 2> gpurun_out/r2_bench_b_tma.err
B2M_SMOOTH_TMA=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_bench_b_notma.json 2> gpurun_out/r2_bench_b_notma.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench_b_tma.json','gpurun_out/r2_bench_b_notma.json'):
    try:
        p=json.load(open(f))
        print(f, p['value'], p['ms_per_step'], p['stage_ms'])
        print(p['roofline']['kernels_ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_smooth3|k_cc_local|k_cc_border|k_mc_emit|k_mc_classify' -s 6 -c 7 -o gpurun_out/r2b_full python tools/profile_step.py 1024 2 > gpurun_out/r2b_full.log 2>&1
tail -3 gpurun_out/r2b_full.log
ls -la gpurun_out/*.ncu-rep
