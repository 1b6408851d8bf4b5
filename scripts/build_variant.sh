#!/bin/bash
# build_variant.sh <name> [extra nvcc flags ...]  ->  real_robots_b200/variants/lib_<name>.so
# Tuning builds: the extra flags apply to $VARIANT_SRC (rr_render by default; rr_physics / rr_collide also work).
# Select one at run time with RR_LIB=real_robots_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../real_robots_b200/csrc"
name=$1; shift
src=${VARIANT_SRC:-rr_render}
mkdir -p ../variants
make -s rr_api.o rr_physics.o rr_collide.o rr_render.o
fmad="-fmad=false"; [ "$src" = rr_physics ] && fmad=""
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $fmad "$@" -c $src.cu -o /tmp/${src}_$name.o
objs=""; for o in rr_api rr_physics rr_render rr_collide; do if [ $o = $src ]; then objs="$objs /tmp/${src}_$name.o"; else objs="$objs $o.o"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so $objs -lcudart
echo built lib_$name.so
