#!/bin/bash
# FUSED vs PHASED: parity suite, kernel-time matrix, host launch overhead.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest_fused.log
tail -25 gpurun_out/r2_pytest_fused.log
timeout 600 python tools/gpu_matrix.py 4,5 ${CASES:-c1,c2,c2_l9,c2_l10,c3_l9,c3_l10,c4_l9,s1,s3,s4,s5} 2>&1 | tee gpurun_out/r2_matrix_fused.jsonl
python - <<'PY'
import sys, time
sys.path.insert(0, "rust-tracer_b200")
import rtrace_b200 as rt, torch
rt.set_device(0)
s = rt.Scene()
o = rt.RenderOptions(3840, 2160, 1)
fb = torch.zeros((2160, 3840, 4), dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for v in (4, 5):
    rt.set_variant(v)
    for _ in range(5): rt.Renderer.render_rows(o, s, out_ptr=fb.data_ptr(), stream=st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n): rt.Renderer.render_rows(o, s, out_ptr=fb.data_ptr(), stream=st)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("variant %d: host enqueue %.1f us per frame, back-to-back %.1f us per frame" % (v, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6))
PY
