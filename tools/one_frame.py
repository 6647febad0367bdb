"""Render exactly the frames of one benchmark workload once each (after one warm-up frame), for ncu captures:
    python tools/one_frame.py c2|c3|c4|c1|c5 [n_frames] [stride]
c5 renders orbit frames 0, stride, 2 stride, ... < n_frames (default: all 120) of the 3840x2160 / 4x4 / level-9 sweep, one
after the other.
The LAST 4 * frames kernel launches named phase_* are the frames (the warm-up frame comes first)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
sys.path.insert(0, ROOT)
import rtrace_b200 as rt
import bench
name = sys.argv[1]
w, h, spp, level = bench.WORKLOADS[name]
n = int(sys.argv[2]) if len(sys.argv) > 2 else (bench.ORBIT_FRAMES if name == "c5" else 1)
rt.set_device(0)
s = rt.Scene(level=level)
o = rt.RenderOptions(w, h, spp)
fb = rt.device_alloc(w * h * 4)
rt.Renderer.render_rows(o, s, out_ptr=fb)     # warm-up (allocates the scratch)
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
frames = list(range(0, n, stride))
for f in frames:
    cam = rt.make_camera(*bench.orbit_basis(f)) if name == "c5" else None
    _, st = rt.Renderer.render_rows(o, s, camera=cam, out_ptr=fb, want_stats=True)
print("rendered", len(frames), "frame(s) of", name, "variant", st.variant_used, "launches/frame", st.kernel_launches)
