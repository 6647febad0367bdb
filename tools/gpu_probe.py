"""Quick on-GPU probe: FP32 microbenchmarks and kernel times per variant / workload."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
import rtrace_b200 as rt

out = {"fp32_ffma_tflops": rt.microbench_fp32(0, 0), "fp32_fmul_fadd_tflops": rt.microbench_fp32(0, 1)}
cases = {"c1": (1024, 768, 4, 8), "c2": (3840, 2160, 1, 8), "c2_l10": (3840, 2160, 1, 10), "c3": (3840, 2160, 4, 9)}
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "2"])]
for name, (w, h, spp, level) in cases.items():
    s = rt.Scene(level=level)
    p, sh = s.count_rays(w, h, spp)
    for v in variants:
        rt.set_variant(v)
        best = 1e30
        for _ in range(4):
            _, st = rt.Renderer.render(rt.RenderOptions(w, h, spp), s, want_stats=True)
            best = min(best, st.kernel_ms)
        out["%s_v%d" % (name, v)] = {"kernel_ms": best, "mrays_s": (p + sh) / best / 1e3, "rays": p + sh}
        print(name, v, out["%s_v%d" % (name, v)], flush=True)
print(json.dumps(out))
