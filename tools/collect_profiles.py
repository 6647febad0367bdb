"""Copy the round's GPU evidence from gpurun_out/ (scratch) into profiles/ (tracked) and build
profiles/latest_summary.json, which bench.py's roofline block reads:

    python tools/collect_profiles.py r02 [output directory, default profiles/]

Inputs (written by tools/round_profiles.sh R): R_counters_{c1,c2,c3,c4,c5,calib}.csv (ncu --metrics ... --csv launch
lists), R_phased_{c2,c3}.ncu-rep (ncu --set full), R_launches_bench.csv, R_matrix.jsonl.
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_hash, WORKLOADS)
from orbit_interp import interpolate_orbit, stride_of  # noqa: E402

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
# On the GPU box only gpurun_out/ travels back (<= 64 MiB): round_profiles.sh runs this there with an output
# directory under gpurun_out/ and drops the large .ncu-rep files afterwards; the result is copied into profiles/.
G = os.path.join(ROOT, "gpurun_out")
P = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
FADD, FMUL, FFMA = ("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k for k in ("fadd", "fmul", "ffma"))


def launches(path):
    """ncu --csv launch list -> [{kernel, metric: value, ...}] in launch order."""
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    out = {}
    for r in rows[rows.index(hdr) + 1:]:
        d = out.setdefault(r[ii], {"kernel": r[ki]})
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)  # bytes, ns
        d[r[mi]] = v
    return [out[k] for k in sorted(out, key=int)]


def flops(k, w):
    """FP32 flops of a launch: thread-level FADD + FMUL + 2 FFMA, each weighted by what the calibration found a
    packed instruction counts as."""
    return (k.get(FADD, 0.0) + k.get(FMUL, 0.0)) * w["addmul"] + 2.0 * k.get(FFMA, 0.0) * w["ffma"]


# ---- calibration: kernels of known flop count (rt_microbench_fp32: 148*8 blocks x 256 threads x 16 chains x iters) ----
calib = {"note": "ncu thread-level op counters on rt_microbench_fp32 kernels of known flop count"}
weights = {"addmul": 1.0, "ffma": 1.0}
cpath = os.path.join(G, "%s_counters_calib.csv" % R)
if os.path.exists(cpath):
    ks = launches(cpath)
    # per mode two launches: warm-up (2000 iterations) and timed (20000); 16 chains per thread, 2 flop per chain step
    per_mode = {}
    for k in ks:
        mode = int(k["kernel"].split("<")[1].split(">")[0].replace("(int)", ""))
        per_mode.setdefault(mode, []).append(k)
    for mode, lst in sorted(per_mode.items()):
        k = max(lst, key=lambda r: r.get("smsp__inst_executed.sum", 0))
        threads = 148 * 8 * 256
        true_flop = threads * 20000 * 16 * 2.0
        counted = k.get(FADD, 0) + k.get(FMUL, 0) + 2 * k.get(FFMA, 0)
        calib["mode%d" % mode] = {"kernel": k["kernel"], "true_flop": true_flop, "fadd": k.get(FADD, 0), "fmul": k.get(FMUL, 0),
                                  "ffma": k.get(FFMA, 0), "counted_flop": counted, "true_over_counted": true_flop / counted if counted else None}
    # modes 2 / 3 are the packed f32x2 chains: if the counters see one op per packed instruction, scale by the ratio
    for key, mode in (("ffma", 2), ("addmul", 3)):
        r = calib.get("mode%d" % mode, {}).get("true_over_counted")
        calib["packed_%s_ratio" % key] = r
    blind = all((calib.get("packed_%s_ratio" % k) or 0.0) > 100.0 for k in ("ffma", "addmul"))
    calib["conclusion"] = (
        "the thread-level op counters (smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on) are exact for scalar FFMA / "
        "FMUL / FADD chains and count NOTHING for packed FFMA2 / FMUL2 / FADD2 chains; executed flops are therefore summed per "
        "opcode from the ncu source page (tools/ncu_flops.py)" if blind else
        "packed f32x2 instructions are counted with the ratios above")

summary = {"kernel_source_hash": bench.kernel_source_hash(), "round": R, "calibration": calib, "workloads": {}}


def frame_record(ks, source):
    t = sum(k.get("gpu__time_duration.sum", 0.0) for k in ks)
    dom = max(ks, key=lambda k: k.get("gpu__time_duration.sum", 0.0))
    rec = {
        "fp32_flop_per_frame": sum(flops(k, weights) for k in ks),
        "warp_instructions_per_frame": sum(k.get("smsp__inst_executed.sum", 0.0) for k in ks),
        "dram_bytes_per_frame": sum(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0) for k in ks),
        "ncu_ns_per_frame": t,
        "dominant_kernel": dom["kernel"].split("(")[0],
        "dominant_kernel_share": dom.get("gpu__time_duration.sum", 0.0) / t if t else None,
        "source": source,
        "launches": [{"kernel": k["kernel"].split("(")[0].replace("void ", ""), "ns": k.get("gpu__time_duration.sum"),
                      "warp_inst": k.get("smsp__inst_executed.sum"), "fp32_flop": flops(k, weights),
                      "dram_bytes": k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0),
                      "fma_pipe_active_pct_of_elapsed": k.get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                      "issue_active_pct_of_elapsed": k.get("smsp__issue_active.avg.pct_of_peak_sustained_elapsed"),
                      "sm_cycles_active_over_elapsed": (k.get("smsp__cycles_active.avg", 0.0) / k["sm__cycles_elapsed.max"])
                      if k.get("sm__cycles_elapsed.max") else None,
                      "threads_per_instruction": k.get("smsp__thread_inst_executed_per_inst_executed.ratio")} for k in ks],
    }
    if t:
        for key, name in (("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "fma_pipe_active_pct"),
                          ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue_active_pct")):
            if all(key in k for k in ks):
                rec[name] = sum(k[key] * k.get("gpu__time_duration.sum", 0.0) for k in ks) / t   # time-weighted over the frame
    return rec


def opcode_flops(w):
    """Executed FP32 flops of one frame of workload w from the per-SASS-instruction counts of its ncu --set full
    report (tools/ncu_flops.py): the op counters above cannot see packed f32x2 instructions."""
    rep = os.path.join(G, "%s_phased_%s.ncu-rep" % (R, w))
    if not os.path.exists(rep):
        return None
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_flops.py"), rep], capture_output=True, text=True).stdout
    try:
        return json.loads(out)
    except ValueError:
        return None


for w in ("c1", "c2", "c3", "c4"):
    path = os.path.join(G, "%s_counters_%s.csv" % (R, w))
    if not os.path.exists(path):
        continue
    ks = [k for k in launches(path) if "phase_" in k["kernel"]]
    if len(ks) < 8:
        continue
    rec = frame_record(ks[-4:], "profiles/%s_counters_%s.csv (ncu --metrics, --clock-control none; the frame after the warm-up frame)" % (R, w))
    of = opcode_flops(w)
    if of and of.get("kernels"):
        rec["fp32_flop_per_frame_scalar_op_counters"] = rec["fp32_flop_per_frame"]   # blind to FFMA2 / FMUL2 / FADD2
        rec["fp32_flop_per_frame"] = of["fp32_flop_per_frame"]
        rec["fp32_pipe_lane_ops_per_frame"] = sum(k["fp32_pipe_lane_ops"] for k in of["kernels"])
        rec["flop_source"] = ("per-opcode thread-level execution counts of profiles/%s_ncu_phased.json's reports (tools/ncu_flops.py): "
                              "FFMA 2, FMUL/FADD 1, FFMA2 4, FMUL2/FADD2 2, MUFU 1" % R)
        rec["opcodes"] = {k["kernel"].split("<")[0].replace("void rt::", ""): k["fp_opcodes_thread_level"] for k in of["kernels"]}
        json.dump(of, open(os.path.join(P, "%s_flops_%s.json" % (R, w)), "w"), indent=1)
    summary["workloads"][w] = rec
    shutil.copy(path, os.path.join(P, "%s_counters_%s.csv" % (R, w)))
path = os.path.join(G, "%s_counters_c5.csv" % R)
if os.path.exists(path):
    ks = [k for k in launches(path) if "phase_" in k["kernel"]]
    n = bench.ORBIT_FRAMES
    n_cap = (len(ks) - 4) // 4                       # frames after the warm-up frame
    stride = stride_of(n, n_cap)
    if stride:
        ks = ks[-4 * n_cap:]
        cap = [sum(k.get("smsp__inst_executed.sum", 0.0) for k in ks[4 * i:4 * i + 4]) for i in range(n_cap)]
        winst = interpolate_orbit(cap, n, stride)    # stride 1: the capture itself
        c3 = summary["workloads"].get("c3", {})
        # flops of an orbit frame = its warp-instruction count x the flops per warp-instruction of frame 0 (= C3, whose
        # per-opcode counts are known): the kernels and their instruction mix are the same, only the amount of work moves
        per_inst = c3["fp32_flop_per_frame"] / c3["warp_instructions_per_frame"] if c3.get("warp_instructions_per_frame") else None
        summary["workloads"]["c5"] = {
            "fp32_flop_per_orbit_frame": [w_ * per_inst for w_ in winst] if per_inst else None,
            "warp_instructions_per_orbit_frame": winst,
            "fp32_flop_per_warp_instruction": per_inst,
            "captured_orbit_frames": list(range(0, n, stride)),
            "source": "profiles/%s_counters_c5.csv.gz (ncu instruction counts of %s of the 120 orbit frames%s) x flops per "
                      "warp-instruction of frame 0 (C3's per-opcode capture)"
                      % (R, "all" if stride == 1 else "every %s" % {2: "second", 3: "third", 4: "fourth"}.get(stride, "%dth" % stride),
                         "" if stride == 1 else ", linear in between")}
        subprocess.run("gzip -9 -c %s > %s" % (path, os.path.join(P, "%s_counters_c5.csv.gz" % R)), shell=True, check=False)
        for k in ("fma_pipe_active_pct", "issue_active_pct", "dominant_kernel", "dominant_kernel_share", "dram_bytes_per_frame"):
            if "c3" in summary["workloads"] and k in summary["workloads"]["c3"]:
                summary["workloads"]["c5"][k] = summary["workloads"]["c3"][k]   # frame 0 of the orbit is C3
json.dump(summary, open(os.path.join(P, "latest_summary.json"), "w"), indent=1)

# ---- ncu --set full summaries of the four launches ----
caps = {}
for case in ("c1", "c2", "c3", "c4"):
    rep = os.path.join(G, "%s_phased_%s.ncu-rep" % (R, case))
    if not os.path.exists(rep):
        continue
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
if caps:
    json.dump({"what": "ncu --set full --clock-control none of the four PHASED launches of one frame (cold caches: ncu flushes L2 "
                       "between replays); c1 = 1024x768 4x4 level 8, c2 = 3840x2160 1 spp level 8, c3 = 3840x2160 4x4 level 9, c4 = 7680x4320 4x4 level 9",
               "captures": caps}, open(os.path.join(P, "%s_ncu_phased.json" % R), "w"), indent=1)
for name in ("%s_launches_bench.csv" % R,):
    if os.path.exists(os.path.join(G, name)):
        shutil.copy(os.path.join(G, name), os.path.join(P, name))
# ---- variant matrix ----
names = {1: "LANE", 3: "TILE", 4: "PHASED", 0: "AUTO"}
res = {}
mpath = os.path.join(G, "%s_matrix.jsonl" % R)
if os.path.exists(mpath):
    for line in open(mpath):
        try:
            d = json.loads(line)
        except ValueError:
            continue
        if "case" in d:
            res.setdefault(d["case"], {})[names.get(d["variant"], str(d["variant"]))] = {"kernel_ms": d["kernel_ms"], "grays_s": d["grays_s"]}
    json.dump({"what": "kernel-only time (CUDA events in rt_stats, best of 3) per variant and configuration on one B200; "
                       "cN_lM = BASELINE config N at pyramid level M", "results": res},
              open(os.path.join(P, "%s_variant_matrix.json" % R), "w"), indent=1)
print(json.dumps({k: {kk: vv for kk, vv in v.items() if not isinstance(vv, list)} for k, v in summary["workloads"].items()}, indent=1)[:3000])
print(json.dumps(calib, indent=1)[:2500])
