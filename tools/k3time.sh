#!/bin/bash
for n in base $1; do
lib=$PWD/build/librtrace_b200_$n.so; [ $n = base ] && lib=$PWD/rust-tracer_b200/librtrace_b200.so
echo "== $n"
RTRACE_B200_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:"${K:-phase_cull_shadow}" -s ${SKIP:-2} -c 1 --csv --log-file gpurun_out/k3_$n.csv python tools/gpu_matrix.py 4 ${CASE:-c2} > /dev/null 2>&1
grep -E "phase_" gpurun_out/k3_$n.csv | awk -F'","' '{print substr($5,1,30), $(NF-2), $NF}'
done
