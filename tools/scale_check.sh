#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err; echo "bench n8 rc=$?"
tail -c 300 gpurun_out/r2_bench_n8b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n8b.json").read().strip().splitlines()[-1])
print(8, "value %.0f ms %.4f e2e %.0f (%.4f ms) frames %s pcie %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["frames_per_rank"], {k: v for k, v in d["e2e"]["pcie"].items() if k != "note"}))
for k, v in (d.get("also") or {}).items():
    print("   ", k, "value %.0f ms %.4f e2e %.0f (%.4f ms) verified %s / %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["e2e"]["ms_per_step"], v.get("gathered_frame_verified"), v["e2e"].get("gathered_frame_verified")), v["e2e"].get("frames_per_rank"), {k2: v2 for k2, v2 in v["e2e"]["pcie"].items() if k2 != "note"})
PY
cd /tmp && for g in 8 1; do ( time $GRAFT_REPO_ROOT/target/release/rtrace --width=3840 --height=2160 --samples-per-pixel=4 --level=9 --frames=48 --gpus=$g --stats sweep$g.tga ) 2>&1 | grep -E "rtrace-b200|real"; done
sha256sum /tmp/sweep1.0017.tga /tmp/sweep8.0017.tga
