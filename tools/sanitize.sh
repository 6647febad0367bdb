#!/bin/bash
# compute-sanitizer over a small frame for every kernel variant (run on the GPU box):
#   gpurun -- bash tools/sanitize.sh
set -u
cd "$(dirname "$0")/.."
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt
for level, w, h, spp in ((8, 200, 120, 1), (8, 96, 64, 4), (10, 160, 90, 2), (5, 70, 33, 3)):
    s = rt.Scene(level=level)
    for v in (1, 2, 3, 4):
        rt.set_variant(v)
        rt.Renderer.render(rt.RenderOptions(w, h, spp), s)
        s.count_rays(w, h, spp)
    rt.set_variant(0)
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), s, 3)
print("sanitize-case done")
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py 2>&1 | tail -6
  echo "exit code: $?"
done
