#!/bin/bash
# parity suite + A/B of build/librtrace_b200_$BASE.so against the current library
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_ab.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_ab.log
tail -4 gpurun_out/r2_pytest_ab.log
for rep in 1 2; do
for n in ${BASE:-wdbase} cur; do
  lib=$PWD/build/librtrace_b200_$n.so; [ $n = cur ] && lib=$PWD/rust-tracer_b200/librtrace_b200.so
  echo "== $n"; RTRACE_B200_LIB=$lib timeout 300 python tools/gpu_matrix.py 4 ${CASES:-c2,c3_l9,c4_l9,c1,c2_l10}
done; done 2>&1 | tee gpurun_out/r2_ab.txt
