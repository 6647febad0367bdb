"""Copy the round's GPU evidence from gpurun_out/ (scratch) into profiles/ (tracked):
bench lines, the ncu launch list of the bench command, ncu --set full summaries, the variant matrix."""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for w in ("c1", "c2", "c3", "c4"):
    src = os.path.join(G, "bench_%s.json" % w)
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, "%s_bench_%s_n1.json" % (R, w)))
if os.path.exists(os.path.join(G, "bench_ref.json")):
    shutil.copy(os.path.join(G, "bench_ref.json"), os.path.join(P, "%s_bench_reference_arm.json" % R))
shutil.copy(os.path.join(G, "%s_launches_c2.csv" % R), os.path.join(P, "%s_launches_c2.csv" % R))
caps = {}
for case in ("c2", "c3"):
    rep = os.path.join(G, "%s_phased_%s.ncu-rep" % (R, case))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    dec, recs, i = json.JSONDecoder(), [], 0
    while i < len(out):
        while i < len(out) and out[i].isspace():
            i += 1
        if i >= len(out):
            break
        obj, i = dec.raw_decode(out, i)
        obj["report"] = os.path.basename(obj["report"])
        recs.append(obj)
    caps[case] = recs[:4]   # one frame: the four launches
json.dump({"what": "ncu --set full --clock-control none of the four PHASED launches of one frame (cold caches: ncu flushes L2 "
                   "between replays); c2 = 3840x2160 1 spp level 8, c3 = 3840x2160 4x4 spp level 9",
           "captures": caps}, open(os.path.join(P, "%s_ncu_phased.json" % R), "w"), indent=1)
# dominant kernel of C2 and its DRAM bytes per launch (roofline.traffic in bench.py)
def num(s):
    v, u = s.split()[0], (s.split() + [""])[1]
    return float(v) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
dom = max(caps["c2"], key=lambda r: float(r["gpu__time_duration.sum"].split()[0]) * (1e3 if "ms" in r["gpu__time_duration.sum"] else 1))
def pct(r, k):
    return float(r[k].split()[0]) if k in r else None
frame_inst = sum(float(r["smsp__inst_executed.sum"].split()[0]) for r in caps["c2"])
json.dump({"c2": {"dominant_kernel": dom["kernel"], "dram_bytes_per_launch": num(dom["dram__bytes_read.sum"]) + num(dom["dram__bytes_write.sum"]),
                  "fma_pipe_cycles_active_pct": pct(dom, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                  "issue_active_pct": pct(dom, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                  "warp_instructions_per_frame": frame_inst,
                  "source": "profiles/%s_ncu_phased.json (cold-cache ncu replay)" % R}}, open(os.path.join(P, "latest_summary.json"), "w"), indent=1)
# variant matrix
names = {1: "LANE", 3: "TILE", 4: "PHASED", 0: "AUTO"}
res = {}
for line in open(os.path.join(G, "matrix.jsonl")):
    try:
        d = json.loads(line)
    except ValueError:
        continue
    if "case" in d:
        res.setdefault(d["case"], {})[names.get(d["variant"], str(d["variant"]))] = {"kernel_ms": d["kernel_ms"], "grays_s": d["grays_s"]}
old = json.load(open(os.path.join(P, "%s_variant_matrix.json" % R)))
old["results"] = res
json.dump(old, open(os.path.join(P, "%s_variant_matrix.json" % R), "w"), indent=1)
print(json.dumps(json.load(open(os.path.join(P, "latest_summary.json")))))
