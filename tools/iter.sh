#!/bin/bash
# One GPU iteration: parity tests, PHASED kernel times per tile shape, per-launch times (ncu) on C2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
: > gpurun_out/matrix.jsonl
for shape in ${SHAPES:-0 1}; do
  echo "{\"shape\": $shape}" >> gpurun_out/matrix.jsonl
  RTRACE_TILE_SHAPE=$shape timeout 300 python tools/gpu_matrix.py ${VARIANTS:-4} ${CASES:-c1,c2,c2_l9,c3_l9,c4_l9,c3_l10} >> gpurun_out/matrix.jsonl 2>&1
done
cat gpurun_out/matrix.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_iter.csv python tools/gpu_matrix.py 4 ${NCU_CASE:-c2} > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_iter.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{})['k']=r[ki][:40]; d[r[ii]][r[mi]]=r[vi]
for i,v in list(d.items())[-8:]:
    print(i, v)
PY
