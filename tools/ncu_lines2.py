"""Executed warp-instructions and stall samples of ONE kernel of an ncu report, per CUDA source line.

    python tools/ncu_lines2.py <report.ncu-rep> <lib.so> <ncu kernel regex> <mangled-name substring> [top] [focus file]

The SASS page of the report (`--page source --csv`) is joined by instruction offset with `nvdisasm -gi`
line info of the cubin; an instruction inlined from a header is attributed to the innermost frame
inside the focus file (default rt_phased.cu).
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kre, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
FOCUS = sys.argv[6] if len(sys.argv) > 6 else "rt_phased.cu"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
line_of = {}
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, block, in_block, cur_line = False, [], False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            inside = mangled in m.group(1)
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if not in_block:
                block, in_block = [], True
            block.append((os.path.basename(m.group(1)), int(m.group(2))))
            if m.group(4):
                block.append((os.path.basename(m.group(4)), int(m.group(5))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if in_block:
                own = [b for b in block if b[0] == FOCUS]
                cur_line = own[0] if own else (block[0] if block else None)
                in_block = False
            line_of[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "-c", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1])
hdr = rows[1]
ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
per_line, samp_line, ops = collections.Counter(), collections.Counter(), collections.Counter()
base = None
tot = tots = 0
for r in rows[2:]:
    try:
        addr = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = addr
    n, s = int(r[ii] or 0), int(r[isamp] or 0)
    key = line_of.get(addr - base, (None, ""))[0]
    per_line[key] += n
    samp_line[key] += s
    ops[r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]] += n
    tot += n
    tots += s
print("total warp-instructions %d, samples %d" % (tot, tots))
print("opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(18)))
src_cache = {}
for key, n in per_line.most_common(top):
    text = ""
    if key:
        for root, _, files in os.walk(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "rust-tracer_b200")):
            if key[0] in files:
                path = os.path.join(root, key[0])
                src_cache.setdefault(path, open(path).read().splitlines())
                text = src_cache[path][key[1] - 1].strip()[:100]
    print("%6.2f%% inst %6.2f%% samp  %s:%s  %s" % (100.0 * n / tot, 100.0 * samp_line[key] / max(tots, 1),
                                                  key[0] if key else "?", key[1] if key else "", text))
