#!/bin/bash
# 8-GPU check: multi-GPU tests, bench default at N=8 and N=4, CLI sweep/frame on 8 GPUs.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_cli.py -m gpu -q -x > gpurun_out/r2_pytest_n8.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_n8.log
tail -5 gpurun_out/r2_pytest_n8.log
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "bench n$n rc=$?"
tail -c 600 gpurun_out/r2_bench_n$n.err
done
python - <<'PY'
import json
for n in (4, 8):
    try:
        d = json.loads(open("gpurun_out/r2_bench_n%d.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f ms %.4f e2e %.0f (%.4f ms) pcie %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k: v for k, v in d["e2e"]["pcie"].items() if k != "note"}))
        for k, v in (d.get("also") or {}).items():
            print("   ", k, "value %.0f ms %.4f e2e %.0f (%.4f ms) verified %s / %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("gathered_frame_verified"), v["e2e"].get("gathered_frame_verified")), {k2: v2 for k2, v2 in v["e2e"]["pcie"].items() if k2 != "note"})
        print("   clocks", d.get("clocks"))
    except Exception as e:
        print(n, "unreadable:", e)
PY
cd /tmp && for g in 1 8; do /usr/bin/time -f "%e s wall" $GRAFT_REPO_ROOT/target/release/rtrace --width=3840 --height=2160 --samples-per-pixel=4 --level=9 --frames=24 --gpus=$g --stats sweep$g.tga; done 2>&1 | tail -6
sha256sum /tmp/sweep1.0007.tga /tmp/sweep8.0007.tga
for g in 1 8; do /usr/bin/time -f "%e s wall" $GRAFT_REPO_ROOT/target/release/rtrace --width=7680 --height=4320 --samples-per-pixel=4 --level=9 --gpus=$g --stats frame$g.tga; done 2>&1 | tail -6
sha256sum /tmp/frame1.tga /tmp/frame8.tga
