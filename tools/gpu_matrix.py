"""Kernel time of every variant on every benchmark configuration (one JSON line per cell)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
import rtrace_b200 as rt
cases = {"c1": (1024, 768, 4, 8), "c2": (3840, 2160, 1, 8), "c2_l9": (3840, 2160, 1, 9), "c2_l10": (3840, 2160, 1, 10),
         "s1": (640, 480, 1, 8), "s2": (256, 192, 2, 8), "s3": (1280, 720, 1, 8), "s4": (1920, 1080, 1, 8), "s5": (320, 240, 1, 8), "s6": (128, 96, 4, 8),
         "c3_l9": (3840, 2160, 4, 9), "c3_l10": (3840, 2160, 4, 10), "c4_l9": (7680, 4320, 4, 9), "c4_l10": (7680, 4320, 4, 10)}
only = sys.argv[2].split(",") if len(sys.argv) > 2 else list(cases)
variants = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["1", "3", "4"])]
for name in only:
    w, h, spp, level = cases[name]
    s = rt.Scene(level=level)
    p, sh = s.count_rays(w, h, spp)
    for v in variants:
        rt.set_variant(v)
        best = 1e30
        for _ in range(3):
            _, st = rt.Renderer.render_rows(rt.RenderOptions(w, h, spp), s, out_ptr=None, want_stats=True) if False else rt.Renderer.render(rt.RenderOptions(w, h, spp), s, want_stats=True)
            best = min(best, st.kernel_ms)
        print(json.dumps({"case": name, "variant": v, "kernel_ms": round(best, 4), "grays_s": round((p + sh) / best / 1e6, 2)}), flush=True)
