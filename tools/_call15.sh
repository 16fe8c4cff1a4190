set -x
mkdir -p gpurun_out
python tools/host_jitter.py 2 > gpurun_out/r2_final_jitter.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_final_gputests.txt 2>&1
tail -5 gpurun_out/r2_final_gputests.txt
timeout 600 python bench.py > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -2 gpurun_out/r2_final_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_final_launches.csv python tools/profile_step.py 1024 2 > gpurun_out/r2_final_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_smooth3|k_threshold|k_cc_|k_dilate|k_mc_|k_scan3_apply|k_compact|k_tri_' -s 25 -c 25 -o gpurun_out/r2_final_full python tools/profile_step.py 1024 2 > gpurun_out/r2_final_full.log 2>&1
tail -2 gpurun_out/r2_final_full.log
timeout 300 python tools/bench_atlas.py --workers 1,8,16 --steps 3 > gpurun_out/r2_final_atlas.json 2> gpurun_out/r2_final_atlas.err
python - <<'PY'
import json
p=json.load(open('gpurun_out/r2_final_bench_n1.json'))
print(p['value'], p['ms_per_step'], p['stage_ms'], p['roofline']['frac'], p['roofline']['whole_step']['frac'])
print(p['e2e']); print(p['cpu_baseline'])
print(json.load(open('gpurun_out/r2_final_bench_ref.json'))['value'])
PY
