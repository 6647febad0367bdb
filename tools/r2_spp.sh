#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_spp.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_spp.log
tail -15 gpurun_out/r2_pytest_spp.log
python - <<'PY' 2>&1 | tee gpurun_out/r2_spp_matrix.txt
import sys, json
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt
s = rt.Scene(level=9)
for spp in (4, 5, 6, 8):
    w, h = 3840, 2160
    p, sh = s.count_rays(w, h, spp)
    for v in (rt.VARIANT_LANE, rt.VARIANT_PHASED):
        rt.set_variant(v)
        best = min(rt.Renderer.render(rt.RenderOptions(w, h, spp), s, want_stats=True)[1].kernel_ms for _ in range(2))
        print(json.dumps({"case": "4K level 9 spp %d" % spp, "variant": v, "kernel_ms": round(best, 3), "grays_s": round((p + sh) / best / 1e6, 1)}), flush=True)
PY
for rep in 1 2; do
  echo "== r01 lib"; RTRACE_B200_LIB=build/librtrace_b200_r01.so timeout 300 python tools/gpu_matrix.py 4 c2,c3_l9,c1
  echo "== current lib"; timeout 300 python tools/gpu_matrix.py 4 c2,c3_l9,c1
done 2>&1 | tee gpurun_out/r2_ab.txt
