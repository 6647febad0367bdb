#!/bin/bash
# One GPU call for a set of experiment builds (make variant NAME=x DEFS=...): the oracle parity tests on ONE
# candidate build, then kernel times of every build next to the in-tree library, twice over, interleaved.
#   gpurun -- bash tools/exp_ab.sh <parity build> "<builds>" "<cases>" [tag]
P=$1; NAMES=$2; CASES=${3:-c2,c3_l9,c4_l9}; TAG=${4:-exp}
mkdir -p gpurun_out
RTRACE_B200_LIB=$PWD/build/librtrace_b200_$P.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kats.py -m gpu -q -x \
  > gpurun_out/${TAG}_pytest_$P.log 2>&1; echo "pytest($P) rc=$?" | tee -a gpurun_out/${TAG}_pytest_$P.log
tail -3 gpurun_out/${TAG}_pytest_$P.log
for rep in 1 2; do
  for n in base $NAMES; do
    lib=$PWD/build/librtrace_b200_$n.so; [ $n = base ] && lib=$PWD/rust-tracer_b200/librtrace_b200.so
    RTRACE_B200_LIB=$lib timeout 200 python tools/gpu_matrix.py 4 $CASES 2>&1 | sed "s/^{/{\"build\": \"$n\", \"rep\": $rep, /"
  done
done | tee gpurun_out/${TAG}_ab.jsonl | python -c "
import sys, json, collections
t = collections.OrderedDict()
for ln in sys.stdin:
    try: d = json.loads(ln)
    except Exception: print(ln.rstrip()); continue
    t.setdefault(d['build'], {}).setdefault(d['case'], []).append(d['kernel_ms'])
base = t.get('base', {})
for b, cs in t.items():
    print(b.ljust(8), '  '.join('%s %s (%+.1f%%)' % (c, '/'.join('%.4f' % x for x in v), 100 * (min(v) / min(base[c]) - 1) if c in base else 0) for c, v in cs.items()))
"
