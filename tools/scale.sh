#!/bin/bash
# Scaling evidence on one multi-GPU box: bench.py at N = 1, 2, .. $1 (frames = weak, bands = strong).
MAXN=${1:-2}
mkdir -p gpurun_out
for n in 1 2 4 8; do
  [ $n -gt $MAXN ] && break
  for mode in frames bands; do
    [ $n = 1 ] && [ $mode = bands ] && continue
    if [ $n = 1 ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py"; fi
    for w in c2 c4; do
      timeout 600 $cmd --gpus $n --steps 100 --warmup 5 --mode $mode --workload $w --no-cpu-baseline > gpurun_out/scale_${w}_${mode}_n$n.json 2> gpurun_out/scale_${w}_${mode}_n$n.err
      python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${w}_${mode}_n$n.json").read().strip().splitlines()[-1])
    print("$w $mode n=$n value %.0f e2e %.0f ms %.4f verified %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("gathered_frame_verified")))
except Exception as e:
    print("$w $mode n=$n FAILED", e)
PY
    done
  done
done
