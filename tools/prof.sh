#!/bin/bash
# ncu --set full of the kernels matching $1 on case $2 (default c2), report -> gpurun_out/$3.ncu-rep
K=${1:-phase_shade_store}; CASE=${2:-c2}; NAME=${3:-prof}
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$K" -s ${SKIP:-4} -c ${COUNT:-2} -f -o gpurun_out/$NAME python tools/gpu_matrix.py ${VARIANT:-4} $CASE > gpurun_out/$NAME.log 2>&1
tail -3 gpurun_out/$NAME.log
