#!/bin/bash
# usage: bash tools/gpu_ab_libs.sh TAG NAME...   A/B of experiment libraries (audiossl_b200/libatst_b200_exp_NAME.so)
# against the default build on one box: GEMM kernel tests on each experiment, the isolated epilogue timings, then
# interleaved bench runs.
tag=$1; shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
cp audiossl_b200/libatst_b200.so /tmp/default.so
for name in "$@"; do
  cp audiossl_b200/libatst_b200_exp_$name.so audiossl_b200/libatst_b200.so
  timeout 400 python -m pytest tests/test_parity_gpu.py -q -x -k "gelu or gemm or residual or colsum" > gpurun_out/${tag}_test_$name.log 2>&1; echo "tests $name rc=$?"
  tail -2 gpurun_out/${tag}_test_$name.log | cut -c1-300
done
for name in default "$@"; do
  if [ $name = default ]; then cp /tmp/default.so audiossl_b200/libatst_b200.so; else cp audiossl_b200/libatst_b200_exp_$name.so audiossl_b200/libatst_b200.so; fi
  echo "== $name"; timeout 200 python tools/ab_gelu_half.py 15 2>&1 | tee gpurun_out/${tag}_ab_$name.log
done
for rep in a b; do
  for name in default "$@"; do
    if [ $name = default ]; then cp /tmp/default.so audiossl_b200/libatst_b200.so; else cp audiossl_b200/libatst_b200_exp_$name.so audiossl_b200/libatst_b200.so; fi
    timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 --breakdown gpurun_out/${tag}_breakdown_${name}_$rep.txt > gpurun_out/${tag}_bench_${name}_$rep.json 2> gpurun_out/${tag}_bench_${name}_$rep.err; echo "bench $name $rep rc=$?"
  done
done
cp /tmp/default.so audiossl_b200/libatst_b200.so
python - "$tag" default "$@" <<'PY'
import json, sys
tag, names = sys.argv[1], sys.argv[2:]
for rep in "ab":
    for n in names:
        try:
            d = json.loads(open("gpurun_out/%s_bench_%s_%s.json" % (tag, n, rep)).read().strip().splitlines()[-1])
            print("%-10s %s value %.1f ms %.2f e2e %.1f gemm_ms %.2f frac %.4f sm %s" % (n, rep, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
        except Exception as e:
            print(n, rep, "unreadable", e)
PY
