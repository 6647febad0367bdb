#!/bin/bash
# Round evidence in one GPU call (1 GPU): per-frame executed-FP32 / DRAM / pipe counters of every bench workload (what
# bench.py's roofline block reads from profiles/latest_summary.json), calibration of those counters on kernels of
# known flop count, ncu --set full captures of the four PHASED kernels on C2 and C3, the ncu launch list of the
# default bench command, the variant matrix.     bash tools/round_profiles.sh r02 ; python tools/collect_profiles.py r02
R=${1:-r02}
mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__cycles_active.avg,smsp__thread_inst_executed_per_inst_executed.ratio
for c in c2 c3 c4 c1; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:phase_ --csv --log-file gpurun_out/${R}_counters_$c.csv python tools/one_frame.py $c > gpurun_out/${R}_counters_$c.log 2>&1
done
# the orbit frames of c5 (every C5_STRIDE-th of the 120; collect_profiles.py interpolates between them): instruction and
# scalar-op counters only (one pass per launch)
F=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum
timeout 900 ncu --metrics $F --clock-control none -k regex:phase_ --csv --log-file gpurun_out/${R}_counters_c5.csv python tools/one_frame.py c5 120 ${C5_STRIDE:-1} > gpurun_out/${R}_counters_c5.log 2>&1
# calibration: the FP32 microbenchmark kernels (16 chains x iters x 256 threads x blocks; FFMA / FMUL+FADD / FFMA2 / FMUL2+FADD2)
timeout 300 ncu --metrics $F --clock-control none -k regex:fp32_peak --csv --log-file gpurun_out/${R}_counters_calib.csv python - > gpurun_out/${R}_counters_calib.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt
for mode in (0, 1, 2, 3):
    print(mode, rt.microbench_fp32(0, mode))
PY
# ncu --set full of one frame's four launches
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c2 ${R}_phased_c2
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c3_l9 ${R}_phased_c3
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c4_l9 ${R}_phased_c4
COUNT=4 SKIP=4 bash tools/prof.sh "phase_" c1 ${R}_phased_c1
# launch list of the default bench command (shares of the step must agree with the event times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_ncu_bench.log 2>&1
timeout 600 python tools/gpu_matrix.py 1,3,4,0 c1,c2,c2_l9,c2_l10,c3_l9,c3_l10,c4_l9,c4_l10 > gpurun_out/${R}_matrix.jsonl 2>&1
tail -3 gpurun_out/${R}_matrix.jsonl
# summarise here (ncu is on the box), keep only what fits the 64 MiB that travel back: the JSON / CSV summaries and
# the C3 report (for per-source-line analysis at home)
python tools/collect_profiles.py $R gpurun_out/profiles_$R > gpurun_out/${R}_collect.log 2>&1
tail -40 gpurun_out/${R}_collect.log
rm -f gpurun_out/${R}_phased_c1.ncu-rep gpurun_out/${R}_phased_c2.ncu-rep gpurun_out/${R}_phased_c4.ncu-rep gpurun_out/${R}_counters_c5.csv
du -sh gpurun_out
