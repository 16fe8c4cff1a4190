set -x
mkdir -p gpurun_out
{
lscpu | head -45; nproc; cat /sys/kernel/mm/transparent_hugepage/enabled; cat /sys/kernel/mm/transparent_hugepage/defrag
numactl -H 2>/dev/null; ls /sys/devices/system/node/; cat /sys/devices/system/node/node*/cpulist
nvidia-smi topo -m; nvidia-smi -L
for d in /sys/bus/pci/devices/*; do [ "$(cat $d/vendor)" = 0x10de ] && echo $d $(cat $d/class) numa $(cat $d/numa_node) cpus $(cat $d/local_cpulist); done
free -g; ulimit -l; grep -i huge /proc/meminfo; taskset -p $$; cat /proc/self/status | grep -i allowed
uname -r; ls /dev/shm | head; df -h /dev/shm
} > gpurun_out/r2_sysinfo.txt 2>&1
nvcc -O2 -o /tmp/hcb tools/hostcopy_bench.cu -lpthread && /tmp/hcb > gpurun_out/r2_hostcopy.txt 2>&1
python tools/check_h2d_overlap.py 1024 > gpurun_out/r2_h2d_plain.txt 2>&1
B2M_H2D_OVERLAP=1 python tools/check_h2d_overlap.py 1024 > gpurun_out/r2_h2d_overlap.txt 2>&1
B2M_LIBPATH=$PWD/nii2mesh_b200/libb2m_alt.so timeout 900 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_emit_early_tests.txt 2>&1
tail -3 gpurun_out/r2_emit_early_tests.txt
cat gpurun_out/r2_h2d_plain.txt gpurun_out/r2_h2d_overlap.txt
