#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_n2c.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_n2c.log
tail -6 gpurun_out/r2_pytest_n2c.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench n1 rc=$?"
tail -c 600 gpurun_out/r2_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"
tail -c 800 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/r2_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f ms %.4f e2e %.0f (%.4f ms) frames %s pcie %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["frames_per_rank"], {k: v for k, v in d["e2e"]["pcie"].items() if k != "note"}))
        for k, v in (d.get("also") or {}).items():
            print("   ", k, "value %.0f ms %.4f e2e %.0f (%.4f ms) verified %s / %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("gathered_frame_verified"), v["e2e"].get("gathered_frame_verified")), v["e2e"].get("frames_per_rank"))
        print("   clocks", d.get("clocks"), "roofline", {k: v for k, v in d["roofline"].items() if k in ("achieved", "frac", "traffic")})
    except Exception as e:
        print(n, "unreadable:", e)
PY
