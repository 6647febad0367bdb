"""Static SASS instruction count per source line of a kernel: python tools/sass_lines.py lib.so kernel-substr [top] [focus-file]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
FOCUS = sys.argv[4] if len(sys.argv) > 4 else "rt_tile.cu"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cnt = collections.Counter(); total = 0
for f in os.listdir(tmp):
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside = False; block = []; in_block = False; cur = None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            inside = kern in m.group(1); continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            if not in_block: block, in_block = [], True
            block.append((os.path.basename(m.group(1)), int(m.group(2))))
            if m.group(4): block.append((os.path.basename(m.group(4)), int(m.group(5))))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+", ln):
            if in_block:
                own = [b for b in block if b[0] == FOCUS]
                cur = own[0] if own else (block[0] if block else None)
                in_block = False
            cnt[cur] += 1; total += 1
print("total SASS instructions", total, "=", total * 16 / 1024, "KB")
src = {}
for key, n in cnt.most_common(top):
    text = ""
    if key:
        for root, _, files in os.walk(os.path.dirname(os.path.abspath(lib))):
            if key[0] in files:
                path = os.path.join(root, key[0]); src.setdefault(path, open(path).read().splitlines())
                text = src[path][key[1] - 1].strip()[:100]
    print("%5d  %s  %s" % (n, key, text))
