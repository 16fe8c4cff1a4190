#!/bin/bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference (neurolabusc/nii2mesh) from the sources where they lie
# (default /root/reference/src) into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Flags are exactly the reference Makefile's (src/Makefile:49,52): -O3, no -march, no -ffast-math,
# so gcc cannot contract to FMA and the bits match a stock reference build.
#   libref_lewiner.so : meshify() + stage functions, Lewiner MarchingCubes.c   (make lewiner)
#   libref_classic.so : same with -DUSE_CLASSIC_CUBES oldcubes.c                (default make)
#   mctest            : MarchingCubes.c self-test (10 analytic 60^3 surfaces)
#   nii2mesh_lewiner / nii2mesh_classic : the reference CLI
#   nii2mesh_b2m      : the reference CLI (main(), flag parsing, load_nii, atlas loop, nii2(), every mesh writer, quadric
#                       simplification - all from the reference's own sources, untouched) linked against libb2m.so for
#                       meshify(), setThreshold() and laplacian_smoothHC(): the drop-in of INTEGRATION.md section 2,
#                       exercised by tests/test_cli_dropin.py.  No reference hot-path object is linked: the macro renames
#                       below only move the reference's OWN definitions out of the way inside their translation units.
# meshify.c alone is compiled with -Dstatic= so that its file-local stage functions
# (quick_smooth, dilate, unify_vertices, remove_degenerate_triangles) are callable from tests.
set -euo pipefail
SRC="${1:-/root/reference/src}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$SRC/meshify.c" ]; then
  echo "build_ref: reference sources not found at $SRC (skipping)"; exit 0
fi
mkdir -p "$OUT/obj_l" "$OUT/obj_c"
CF="-O3 -fPIC -DNII2MESH -DHAVE_ZLIB -DHAVE_FORMATS -I$SRC"
for flavour in l c; do
  O="$OUT/obj_$flavour"
  EXTRA=""; MC="MarchingCubes.c"
  if [ $flavour = c ]; then EXTRA="-DUSE_CLASSIC_CUBES"; MC="oldcubes.c"; fi
  gcc $CF $EXTRA -Dstatic= -c "$SRC/meshify.c" -o "$O/meshify.o" 2>/dev/null
  for f in bwlabel.c radixsort.c base64.c isolevel.c quadric.c $MC; do
    gcc $CF $EXTRA -c "$SRC/$f" -o "$O/${f%.c}.o" 2>/dev/null
  done
done
gcc -shared -o "$OUT/libref_lewiner.so" "$OUT"/obj_l/*.o -lz -lm
gcc -shared -o "$OUT/libref_classic.so" "$OUT"/obj_c/*.o -lz -lm
gcc -O3 -DMC_SELF_TEST "$SRC/MarchingCubes.c" -o "$OUT/mctest" -lm 2>/dev/null
L="isolevel.c meshify.c quadric.c bwlabel.c radixsort.c nii2mesh.c base64.c"
(cd "$SRC" && gcc -O3 -DNII2MESH $L -DHAVE_FORMATS MarchingCubes.c -lm -lz -DHAVE_ZLIB -o "$OUT/nii2mesh_lewiner" 2>/dev/null)
(cd "$SRC" && gcc -O3 -DNII2MESH $L -DHAVE_FORMATS -DUSE_CLASSIC_CUBES oldcubes.c -lm -lz -DHAVE_ZLIB -o "$OUT/nii2mesh_classic" 2>/dev/null)
# ---- the drop-in build: reference CLI + writers over libb2m.so ----
B2M="$(cd "$HERE/../nii2mesh_b200" && pwd)"
if [ -f "$B2M/libb2m.so" ]; then
  O="$OUT/obj_d"; mkdir -p "$O"
  DF="-O3 -DNII2MESH -DHAVE_ZLIB -DHAVE_FORMATS -I$SRC -ffunction-sections -fdata-sections"
  # meshify.c holds the ten mesh writers AND the CPU hot path: its meshify() and the host utilities the library also
  # exports are renamed inside this translation unit only, so every call from nii2mesh.c binds to libb2m.so
  gcc $DF -Dmeshify=ref_cpu_meshify -Dapply_sform=ref_cpu_apply_sform -c "$SRC/meshify.c" -o "$O/meshify_io.o" 2>/dev/null
  # quadric.c: the edge-collapse simplification stays the reference's; its Laplacian smooth yields to the library's
  gcc $DF -Dlaplacian_smoothHC=ref_cpu_laplacian_smoothHC -c "$SRC/quadric.c" -o "$O/quadric.o" 2>/dev/null
  for f in nii2mesh.c base64.c bwlabel.c radixsort.c MarchingCubes.c; do gcc $DF -c "$SRC/$f" -o "$O/${f%.c}.o" 2>/dev/null; done
  # (bwlabel / radixsort / MarchingCubes only satisfy the references of the renamed, never-called ref_cpu_meshify;
  #  --gc-sections drops them from the binary.  isolevel.c is NOT linked: setThreshold() comes from libb2m.so)
  gcc -o "$OUT/nii2mesh_b2m" "$O"/*.o -Wl,--gc-sections -L"$B2M" -lb2m -Wl,-rpath,'$ORIGIN/../../nii2mesh_b200' -lm -lz
  rm -rf "$O"
fi
rm -rf "$OUT/obj_l" "$OUT/obj_c"
ls -la "$OUT"
