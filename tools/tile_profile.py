"""Phase breakdown of the TILE kernel (needs a -DRT_TILE_PROFILE build via RTRACE_B200_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
os.environ["RTRACE_PROFILE"] = "1"
import rtrace_b200 as rt
cases = {"c1": (1024, 768, 4, 8), "c2": (3840, 2160, 1, 8), "c2_l10": (3840, 2160, 1, 10), "c3": (3840, 2160, 4, 9)}
for name, (w, h, spp, level) in cases.items():
    s = rt.Scene(level=level)
    print(name, "(setup, pcull, ptest, shade, scull, stest, store | .. pcands scands tiles)", flush=True)
    s.count_rays(w, h, spp)
