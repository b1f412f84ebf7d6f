#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests/test_parity_gpu.py -q -x -k "gelu or gemm or residual or colsum" > gpurun_out/c3_test_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -3 gpurun_out/c3_test_gemm.log | cut -c1-300
timeout 200 python tools/ab_gelu_half.py 15 > gpurun_out/c3_ab.log 2>&1; echo "ab rc=$?"
cat gpurun_out/c3_ab.log
cp audiossl_b200/libatst_b200.so /tmp/new.so
for rep in a b; do
  cp audiossl_b200/libatst_b200_old.so audiossl_b200/libatst_b200.so
  timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/c3_bench_old_$rep.json 2> gpurun_out/c3_bench_old_$rep.err; echo "bench old $rep rc=$?"
  cp /tmp/new.so audiossl_b200/libatst_b200.so
  timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 --breakdown gpurun_out/c3_breakdown_new_$rep.txt > gpurun_out/c3_bench_new_$rep.json 2> gpurun_out/c3_bench_new_$rep.err; echo "bench new $rep rc=$?"
done
python - <<'PY'
import json
for n in ("old_a", "new_a", "old_b", "new_b"):
    try:
        d = json.loads(open("gpurun_out/c3_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f ms %.2f e2e %.1f gemm_ms %.2f frac %.4f sm %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
