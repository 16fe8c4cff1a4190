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
rm -rf "$OUT/obj_l" "$OUT/obj_c"
ls -la "$OUT"
