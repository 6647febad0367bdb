#!/bin/bash
# Round-2 closing run on one GPU: the GPU test suite, memcheck, the round's profile capture, the default bench line and the reference arm.
# C5_STRIDE=3 captures every third orbit frame (tools/round_profiles.sh); SKIP_REF=1 leaves out the reference arm (host-only code:
# its committed line stays valid when only kernels changed).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_1gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_1gpu.log
SAN_TOOLS=memcheck bash tools/sanitize.sh > gpurun_out/r02_sanitize.log 2>&1; tail -4 gpurun_out/r02_sanitize.log
bash tools/round_profiles.sh r02 > gpurun_out/r02_round_profiles.log 2>&1; tail -5 gpurun_out/r02_round_profiles.log
cp profiles/latest_summary.json /tmp/old_summary.json; cp gpurun_out/profiles_r02/latest_summary.json profiles/latest_summary.json   # so that the bench line below carries the executed-work roofline
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_default_n1.json 2> gpurun_out/r02_bench_default_n1.err; echo "bench rc=$?"
[ -n "$SKIP_REF" ] && cp profiles/r02_bench_reference_arm.json gpurun_out/r02_bench_reference_arm.json
[ -z "$SKIP_REF" ] && { timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"; }
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default_n1.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r02_bench_reference_arm.json").read().strip().splitlines()[-1])
print("c5 value %.0f (%.4f ms) e2e %.0f roofline frac %s achieved %s ref-work %.2f; reference arm %.1f Mrays/s (%d steps)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"]["reference_work"]["frac_of_peak"], r["value"], r["steps"]))
for k, v in d["also"].items():
    print(k, "value %.0f (%.4f ms) e2e %.0f frac %s" % (v["value"], v["ms_per_step"], v["e2e"]["value"], v["roofline"]["frac"]))
print(d["clocks"], d["cpu_baseline"])
PY
du -sh gpurun_out
