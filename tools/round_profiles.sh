#!/bin/bash
# Round-end evidence: bench lines, ncu launch list of the bench command, ncu --set full of the PHASED kernels.
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for w in c1 c3 c4; do python bench.py --workload $w --steps 50 --warmup 5 > gpurun_out/bench_$w.json 2>/dev/null; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
python tools/gpu_matrix.py 1,3,4,0 c1,c2,c2_l9,c2_l10,c3_l9,c3_l10,c4_l9,c4_l10 > gpurun_out/matrix.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c2 ${R}_phased_c2
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c3_l9 ${R}_phased_c3
cat gpurun_out/bench_c2.json
