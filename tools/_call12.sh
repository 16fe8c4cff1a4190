set -x
mkdir -p gpurun_out
for v in default t488 t2168 t4164 t2816; do
  if [ $v = default ]; then unset B2M_LIBPATH; else export B2M_LIBPATH=$PWD/nii2mesh_b200/libb2m_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2_tile_$v.json 2> gpurun_out/r2_tile_$v.err
  python -c "
import json;p=json.load(open('gpurun_out/r2_tile_$v.json'));k=p['roofline']['kernels_ms_per_step'];print('$v',round(p['ms_per_step'],3),p['stage_ms']['cc'],{a:k[a] for a in ('cc_local','cc_border','cc_select','cc_flatten','cc_best')})"
done
for v in t488 t2168; do
  B2M_LIBPATH=$PWD/nii2mesh_b200/libb2m_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slabs.py -m gpu -x -q -k "front or cc or slabs or meshify_vs or gyroid or narrow" > gpurun_out/r2_tile_tests_$v.txt 2>&1; tail -3 gpurun_out/r2_tile_tests_$v.txt
done
