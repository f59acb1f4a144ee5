#!/bin/bash
# Run ON the GPU box (through gpurun): plain bench first, then the ncu launch list of the same command and one
# `--set full` capture of each headline kernel.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
set -x
tag=${1:-r1}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stft512_fwd_kernel|istft512_kernel' -c 12 \
    -o gpurun_out/prof_${tag}_full -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --gl-clips 0 \
    > gpurun_out/ncu_full_${tag}.log 2>&1
tail -3 gpurun_out/ncu_full_${tag}.log
