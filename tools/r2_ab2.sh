#!/bin/bash
# per-kernel durations (ncu launch list) of one PHASED frame: round-1 library vs current
mkdir -p gpurun_out
for lib in r01 cur; do
 for c in c1 c3_l9; do
  if [ $lib = r01 ]; then export RTRACE_B200_LIB=build/librtrace_b200_r01.so; else unset RTRACE_B200_LIB; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,launch__registers_per_thread,smsp__cycles_active.avg --clock-control none -k regex:phase_ -s 8 -c 4 --csv --log-file gpurun_out/ab2_${lib}_$c.csv python tools/gpu_matrix.py 4 $c > /dev/null 2>&1
  echo "== $lib $c"
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ab2_${lib}_$c.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{})['k']=r[ki][:28]; d[r[ii]][r[mi]]=r[vi]
for i,v in d.items(): print(v)
PY
 done
done 2>&1 | tee gpurun_out/r2_ab2.txt
