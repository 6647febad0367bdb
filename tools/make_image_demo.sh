#!/bin/bash
# BASELINE configs[0] as a user runs it: smoke(), then the reference's own `make image` command line (Makefile:7) on the
# GPU box, twice (the first run pays the CUDA context), with --stats and the sha256 of the written file next to the
# golden one (tests/golden/rtrace_output_1024x768.json: the reference's shipped image as a PPM).
mkdir -p gpurun_out
{
python -c 'import __graft_entry__ as g; g.smoke()'
for i in 1 2; do
  ( time ./target/release/rtrace --samples-per-pixel=4 --width=1024 --height=768 --stats out.tga ) 2>&1
done
sha256sum out.tga
python - <<'PY'
import hashlib, json
g = json.load(open("tests/golden/rtrace_output_1024x768.json"))
body = open("out.tga", "rb").read()
print("out.tga sha256", hashlib.sha256(body).hexdigest(), "== the reference image as a PPM:", hashlib.sha256(body).hexdigest() == g["ppm_sha256"])
PY
} > gpurun_out/r02_make_image_demo.log 2>&1
tail -12 gpurun_out/r02_make_image_demo.log
