#!/bin/bash
# Scaling curve on ONE 8-GPU box: the default bench line at N = 1, 2, 4, 8 (what the driver's SCALE run does), then
# `rtrace --frames --gpus 8` against 1 GPU (same files).   gpurun --gpus 8 -- bash tools/scale_check.sh
mkdir -p gpurun_out
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r02_scale_n$n.json 2> gpurun_out/r02_scale_n$n.err
  fi
  echo "bench n$n rc=$?"
done
python - <<'PY'
import json
base = {}
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open("gpurun_out/r02_scale_n%d.json" % n).read().strip().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable:", e); continue
    rows = [("c5", d)] + list((d.get("also") or {}).items())
    for k, v in rows:
        key = "c4" if k.startswith("c4") else k
        if n == 1: base[key] = (v["value"], v["e2e"]["value"])
        b = base.get(key, (None, None))
        print("N=%d %-9s value %9.0f (%.4f ms)%s  e2e %9.0f (%.4f ms)%s  pcie %.1f/%.1f GB/s  %s %s" % (
            n, k, v["value"], v["ms_per_step"], "  x%.2f" % (v["value"] / b[0]) if b[0] else "", v["e2e"]["value"], v["e2e"]["ms_per_step"],
            "  x%.2f" % (v["e2e"]["value"] / b[1]) if b[1] else "", v["e2e"]["pcie"]["achieved_gbs"], v["e2e"]["pcie"]["ceiling_gbs"],
            v["e2e"].get("frames_per_rank") or "", ("verified %s/%s" % (v.get("gathered_frame_verified"), v["e2e"].get("gathered_frame_verified"))) if "bands" in k else ""))
PY
cd /tmp && for g in 8 1; do ( time $GRAFT_REPO_ROOT/target/release/rtrace --width=3840 --height=2160 --samples-per-pixel=4 --level=9 --frames=48 --gpus=$g --stats sweep$g.tga ) 2>&1 | grep -E "rtrace-b200|real"; done
sha256sum /tmp/sweep1.0017.tga /tmp/sweep8.0017.tga
