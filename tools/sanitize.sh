#!/bin/bash
# compute-sanitizer over a small frame for every kernel variant (run on the GPU box):
#   gpurun -- bash tools/sanitize.sh
# SAN_ONLY=window limits the run to the render_region windows and previews; SAN_TOOLS="memcheck" picks the tools.
set -u
cd "$(dirname "$0")/.."
cat > /tmp/san_case.py <<'PY'
import sys
sys.path.insert(0, "rust-tracer_b200")
import os
import rtrace_b200 as rt
s = rt.Scene(level=6)
o = rt.RenderOptions(150, 70, 2)
for (l, b, r, t) in ((0, 0, 64, 64), (64, 0, 128, 64), (128, 64, 150, 70), (149, 69, 150, 70), (3, 5, 121, 66)):
    rt.Renderer.render_region(o, s, l, b, r, t)            # column window: per-lane kernel over the bucket only
for step in (1, 3, 8, 64, 200):
    rt.Renderer.render_preview(o, s, step)                  # undersampled preview: block stores clipped to the image
print("sanitize-case windows done")
if os.environ.get("SAN_ONLY") == "window":
    raise SystemExit(0)
for level, w, h, spp in ((8, 200, 120, 1), (8, 96, 64, 4), (10, 160, 90, 2), (5, 70, 33, 3), (7, 90, 50, 5), (6, 64, 40, 8)):
    s = rt.Scene(level=level)
    for v in (1, 2, 3, 4):
        rt.set_variant(v)
        rt.Renderer.render(rt.RenderOptions(w, h, spp), s)
        s.count_rays(w, h, spp)
    rt.set_variant(0)
    rt.Renderer.render_sweep(rt.RenderOptions(w, h, spp), s, 3)
    it = iter([(5, rt.orbit_camera(3, 40)), (6, None), (7, rt.orbit_camera(9, 40))])
    rt.Renderer.render_sweep_pull(rt.RenderOptions(w, h, spp), s, lambda: next(it, None), rgb=True)
print("sanitize-case done")
PY
for tool in ${SAN_TOOLS:-memcheck racecheck}; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san_case.py 2>&1 | tail -6
  echo "exit code: ${PIPESTATUS[0]}"
done
