"""Executed FP32 work of the kernels in an ncu report, from the per-SASS-instruction execution counts of the
source page (SourceCounters section of `--set full`):

    python tools/ncu_flops.py <report.ncu-rep> [kernel regex ...]      (default: the four PHASED kernels)

The thread-level op counters (smsp__sass_thread_inst_executed_op_{fadd,fmul,ffma}_pred_on) do not see the packed
f32x2 instructions of sm_100 at all (calibrated on rt_microbench_fp32: 0 counted for the FFMA2 / FMUL2 chains), so
flops are summed per opcode: thread-level executions x {FFMA 2, FMUL 1, FADD 1, FFMA2 4, FMUL2 2, FADD2 2, MUFU 1}.
Comparisons, selects, min/max and conversions are not counted.  Prints one JSON object.
"""
import collections
import csv
import io
import json
import subprocess
import sys

FLOP = {"FFMA": 2, "FMUL": 1, "FADD": 1, "FFMA2": 4, "FMUL2": 2, "FADD2": 2, "MUFU": 1}
# slots of the FP32 pipe an instruction occupies per thread: a packed instruction holds it for two
LANE_OPS = {"FFMA": 1, "FMUL": 1, "FADD": 1, "FFMA2": 2, "FMUL2": 2, "FADD2": 2}


def kernel_flops(rep, kre):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next((i for i, r in enumerate(rows) if "Source" in r and "Address" in r), None)
    if hdr_i is None:
        return None
    name = rows[hdr_i - 1][1] if hdr_i > 0 and len(rows[hdr_i - 1]) > 1 else kre
    hdr = rows[hdr_i]
    isrc, iw = hdr.index("Source"), hdr.index("Instructions Executed")
    it = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
    warp, thread = collections.Counter(), collections.Counter()
    tot_w = tot_t = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= max(isrc, iw):
            continue
        toks = r[isrc].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        try:
            w = int(r[iw] or 0)
            t = int(r[it] or 0) if it is not None else w * 32
        except ValueError:
            continue
        warp[op] += w
        thread[op] += t
        tot_w += w
        tot_t += t
    # The source page's counts cover every replay pass that collected them (a multiple of one launch): normalise
    # with the launch's own smsp__inst_executed.sum from the raw page.
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "-k", "regex:" + kre, "-c", "1"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    scale, one_launch = 1.0, None
    if len(rr) >= 3 and "smsp__inst_executed.sum" in rr[0]:
        one_launch = float(rr[2][rr[0].index("smsp__inst_executed.sum")].replace(",", ""))
        if tot_w:
            scale = one_launch / tot_w
    flops = sum(thread[o] * f for o, f in FLOP.items()) * scale
    lane_ops = sum(thread[o] * f for o, f in LANE_OPS.items()) * scale
    return {"kernel": name, "warp_instructions": tot_w * scale, "thread_instructions": tot_t * scale, "fp32_flop": flops,
            "fp32_pipe_lane_ops": lane_ops, "source_page_over_launch": (1.0 / scale) if scale else None,
            "fp_opcodes_thread_level": {o: thread[o] * scale for o in FLOP if thread[o]},
            "top_opcodes_warp_level": {o: n * scale for o, n in warp.most_common(12)}}


if __name__ == "__main__":
    rep = sys.argv[1]
    ks = sys.argv[2:] or ["phase_cull_primary", "phase_test_primary", "phase_cull_shadow", "phase_shade_store"]
    recs = [kernel_flops(rep, k) for k in ks]
    recs = [r for r in recs if r]
    print(json.dumps({"report": rep, "kernels": recs, "fp32_flop_per_frame": sum(r["fp32_flop"] for r in recs),
                      "warp_instructions_per_frame": sum(r["warp_instructions"] for r in recs)}, indent=1))
