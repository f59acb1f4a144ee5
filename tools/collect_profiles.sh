#!/bin/bash
# Run ON the GPU box (through gpurun): plain bench first, then the ncu launch list of the same command and one
# `--set full` capture of each headline kernel.  Outputs land in gpurun_out/ (copied into profiles/ afterwards).
set -x
tag=${1:-r2}
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || exit 1
P="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-clock-probe --model-clips 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${tag}.csv \
    $P > gpurun_out/ncu_launch_${tag}.log 2>&1
for k in stft512_fwd_kernel istft512_tma_kernel istft512_gl_tma_kernel; do
  ncu --set full --import-source on --clock-control none -k regex:$k -c 2 -o gpurun_out/prof_${tag}_$k -f \
      $P > gpurun_out/ncu_full_${tag}_$k.log 2>&1
  tail -2 gpurun_out/ncu_full_${tag}_$k.log
done
