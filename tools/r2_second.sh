#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_n2b.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_n2b.log
tail -6 gpurun_out/r2_pytest_n2b.log
for ch in 1 2 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ch bench.py --gpus 2 --steps 10 --warmup 3 --workload c4 --mode bands --band-chunks $ch > gpurun_out/r2_bands_n2_ch$ch.json 2> gpurun_out/r2_bands_n2_ch$ch.err; echo "bands ch$ch rc=$?"
tail -c 400 gpurun_out/r2_bands_n2_ch$ch.err
done
python - <<'PY'
import json
for ch in (1, 2, 4):
    try:
        d = json.loads(open("gpurun_out/r2_bands_n2_ch%d.json" % ch).read().strip().splitlines()[-1])
        print(ch, "value %.0f ms %.4f (unsynced %.0f) e2e %.0f (%.4f ms) verified %s / %s pcie %s" % (d["value"], d["ms_per_step"], d["config"]["unsynchronised_throughput_mrays_s"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("gathered_frame_verified"), d["e2e"].get("gathered_frame_verified"), d["e2e"]["pcie"]))
    except Exception as e:
        print(ch, "unreadable:", e)
PY
F=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum
timeout 300 ncu --metrics $F --clock-control none -k regex:fp32_peak --csv --log-file gpurun_out/r02_counters_calib.csv python - > gpurun_out/r02_counters_calib.log 2>&1 <<'PY'
import sys
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt
for mode in (0, 1, 2, 3):
    print(mode, rt.microbench_fp32(0, mode))
PY
tail -3 gpurun_out/r02_counters_calib.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_counters_calib.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
d={}
for r in rows[1:]:
    d.setdefault(r[ii],{})['k']=r[ki][:40]; d[r[ii]][r[mi].replace('smsp__sass_thread_inst_executed_op_','').replace('_pred_on.sum','')]=r[vi]
for i,v in d.items(): print(v)
PY
