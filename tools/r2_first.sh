#!/bin/bash
# Round-2 first GPU call (2 GPUs): the whole GPU suite incl. the multi-GPU tests, bench default at N=1 and N=2,
# the reference arm, and a copy-only D2H probe (pinned vs write-combined, 1 and 2 ranks).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.txt 2>&1; nproc >> gpurun_out/r2_gpus.txt; numactl -H >> gpurun_out/r2_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_n2.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_n2.log
tail -15 gpurun_out/r2_pytest_n2.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_n2.err
python - <<'PY' > gpurun_out/r2_d2h.txt 2>&1
import sys, os
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt
rt.set_device(0)
for nbytes in (1 << 20, 8 << 20, 24883200, 132710400):
    print(nbytes, "pinned %.1f GB/s" % rt.microbench_d2h(nbytes, 20), "write-combined %.1f GB/s" % rt.microbench_d2h(nbytes, 20, True))
PY
cat gpurun_out/r2_d2h.txt
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/r2_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f ms %.4f e2e %.0f (%.4f ms) pcie %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pcie")))
        for k, v in (d.get("also") or {}).items():
            print("   ", k, "value %.0f ms %.4f e2e %.0f (%.4f ms) verified %s / %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("gathered_frame_verified"), v["e2e"].get("gathered_frame_verified")), v["e2e"].get("pcie"))
        print("   clocks", d.get("clocks"))
    except Exception as e:
        print(n, "unreadable:", e)
PY
