#!/bin/bash
# Tuning builds of libplo_cuda.so with different compile-time knobs, for A/B timing in ONE gpurun call:
#   tools/build_variants.sh tag1 "-DPLO_MINB=3" tag2 "-DPLO_LEAF_BATCH=2" ...
# -> build/variants/libplo_cuda_<tag>.so (git-ignored, travels with gpurun); select with PLO_LIB=<path>.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
CS="$ROOT/planetary-lidar-odometry_b200/csrc"
OUT="$ROOT/build/variants"
mkdir -p "$OUT"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-fvisibility=hidden -I$ROOT/include -I$CS"
while [ $# -ge 2 ]; do
  tag="$1"; defs="$2"; shift 2
  d="$OUT/obj_$tag"; mkdir -p "$d"
  for f in plo_api index_build knn_project p2plane_solve frontend; do
    if [ "$f" = knn_project ] || [ "$f" = p2plane_solve ] || [ "$f" = index_build ] || [ ! -f "$CS/$f.o" ]; then nvcc $FLAGS $defs -c "$CS/$f.cu" -o "$d/$f.o" & else cp "$CS/$f.o" "$d/$f.o"; fi
  done
  wait
  nvcc -shared -cudart static -o "$OUT/libplo_cuda_$tag.so" "$d"/*.o 2>/dev/null
  rm -rf "$d"
  echo "built $OUT/libplo_cuda_$tag.so ($defs)"
done
