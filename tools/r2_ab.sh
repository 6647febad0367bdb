#!/bin/bash
# A/B on one box: round-1 library vs current (PHASED).
mkdir -p gpurun_out
for rep in 1 2; do
  echo "== r01 lib"; RTRACE_B200_LIB=build/librtrace_b200_r01.so timeout 300 python tools/gpu_matrix.py 4 c2,c3_l9,c1,c4_l9
  echo "== current lib"; timeout 300 python tools/gpu_matrix.py 4 c2,c3_l9,c1,c4_l9
done 2>&1 | tee gpurun_out/r2_ab.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks.mem --format=csv | tee -a gpurun_out/r2_ab.txt
