#!/bin/bash
# Kernel times of experiment builds (build/librtrace_b200_<name>.so) next to the default library.
# usage: abtest.sh "<names>" "<cases>" [variant]
for n in base $1; do
  lib=$PWD/build/librtrace_b200_$n.so; [ $n = base ] && lib=$PWD/rust-tracer_b200/librtrace_b200.so
  echo "== $n"
  RTRACE_B200_LIB=$lib timeout 300 python tools/gpu_matrix.py ${3:-4} $2 2>&1 | tail -n 12
done
