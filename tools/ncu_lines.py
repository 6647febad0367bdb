"""Aggregate an ncu SASS source page (--page source --csv) per CUDA source line,
using nvdisasm line info of the cubin: where do the executed instructions go?

    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
FOCUS = sys.argv[5] if len(sys.argv) > 5 else "rt_tile.cu"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
line_of = {}
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur_fn, cur_line, inside = None, None, False
    block = []       # consecutive //## markers before an instruction: innermost frame first
    in_block = False
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            cur_fn = m.group(1)
            inside = kern in cur_fn
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
                # innermost frame that lies in the kernel's own file (not the math header)
                own = [b for b in block if b[0] == FOCUS]
                cur_line = own[0] if own else (block[0] if block else None)
                in_block = False
            line_of[int(m.group(1), 16)] = (cur_line, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
per_line = collections.Counter()
samp_line = collections.Counter()
base = None
tot = tots = 0
for r in rows[2:]:
    try:
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
    except ValueError:
        continue
    if base is None:
        base = addr
    off = addr - base
    n = int(r[ii] or 0)
    s = int(r[isamp] or 0)
    key = line_of.get(off, (None, ""))[0]
    per_line[key] += n
    samp_line[key] += s
    tot += n
    tots += s
print("total warp-instructions %d, samples %d" % (tot, tots))
src_cache = {}
for key, n in per_line.most_common(top):
    text = ""
    if key:
        path = None
        for root, _, files in os.walk(os.path.dirname(os.path.abspath(lib))):
            if key[0] in files:
                path = os.path.join(root, key[0])
        if path:
            src_cache.setdefault(path, open(path).read().splitlines())
            text = src_cache[path][key[1] - 1].strip()[:90]
    print("%6.2f%% inst %6.2f%% samp  %s  %s" % (100.0 * n / tot, 100.0 * samp_line[key] / max(tots, 1), key, text))
