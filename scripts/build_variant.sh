#!/bin/bash
# build_variant.sh <name> [extra nvcc flags for rr_render.cu ...]  ->  real_robots_b200/variants/lib_<name>.so
# Rasteriser tuning builds: select one at run time with RR_LIB=real_robots_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../real_robots_b200/csrc"
name=$1; shift
mkdir -p ../variants
make -s rr_api.o rr_physics.o rr_collide.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -fmad=false "$@" -c rr_render.cu -o /tmp/rr_render_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/lib_$name.so rr_api.o rr_physics.o /tmp/rr_render_$name.o rr_collide.o -lcudart
echo built lib_$name.so
