#!/bin/bash
# build/ab/<name>.so = libaip_b200 compiled with extra -D switches (A/B experiments; see tools/ab_kernels.py)
#   tools/build_variant.sh <name> [-DAIP_...]...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab/obj_$1
name=$1; shift
objs=()
for u in ml_audio_inpainting_b200/csrc/aip_*.cu; do
  o=build/ab/obj_$name/$(basename ${u%.cu}).o
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC "$@" -c -o $o $u &
  objs+=($o)
done
wait
nvcc -shared -o build/ab/$name.so "${objs[@]}"
echo built build/ab/$name.so "$@"
