#!/bin/bash
# build/ab/<name>.so = libaip_b200 compiled with extra -D switches (A/B experiments; see tools/ab_kernels.py)
#   tools/build_variant.sh <name> [-DAIP_...]...
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -shared -Xcompiler -fPIC "$@" \
     -o build/ab/$name.so ml_audio_inpainting_b200/csrc/aip_kernels.cu
echo built build/ab/$name.so "$@"
