set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "registered or repeated_host or pinned_input" > gpurun_out/r2_gputests_g.txt 2>&1
tail -5 gpurun_out/r2_gputests_g.txt
{
python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=4 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=4 B2M_D2H_REGISTER=0 python tools/e2e_probe.py 1024 4
B2M_D2H_REGISTER=1 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=8 python tools/e2e_probe.py 1024 4
B2M_COPY_THREADS=8 B2M_D2H_REGISTER=0 python tools/e2e_probe.py 1024 4
} > gpurun_out/r2_e2e_probe2.txt 2>&1
grep total gpurun_out/r2_e2e_probe2.txt
