#!/bin/bash
# one gpurun call: kernel test of the fp16 gelu' epilogues, isolated A/B, in-step A/B, link-wise parity with the switch on
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.log 2>&1
timeout 400 python -m pytest tests/test_parity_gpu.py -q -x -k "gelu" > gpurun_out/c1_test_gelu.log 2>&1; echo "test_gelu rc=$?"
tail -3 gpurun_out/c1_test_gelu.log
timeout 200 python tools/ab_gelu_half.py 15 > gpurun_out/c1_ab_gelu_half.log 2>&1; echo "ab rc=$?"
cat gpurun_out/c1_ab_gelu_half.log
ATST_FUSE_GELU=7 timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/c1_bench_fg7.json 2> gpurun_out/c1_bench_fg7.err; echo "bench7 rc=$?"
ATST_FUSE_GELU=23 timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/c1_bench_fg23.json 2> gpurun_out/c1_bench_fg23.err; echo "bench23 rc=$?"
ATST_FUSE_GELU=7 timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/c1_bench_fg7b.json 2> gpurun_out/c1_bench_fg7b.err; echo "bench7b rc=$?"
ATST_FUSE_GELU=23 timeout 300 python bench.py --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/c1_bench_fg23b.json 2> gpurun_out/c1_bench_fg23b.err; echo "bench23b rc=$?"
python - <<'PY'
import json
for n in ("fg7", "fg23", "fg7b", "fg23b"):
    try:
        d = json.loads(open("gpurun_out/c1_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f ms %.2f e2e %.1f gemm_ms %.2f frac %.4f sm %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
ATST_FUSE_GELU=23 timeout 600 python -m pytest tests/test_parity_tf32_gpu.py -q -x -s > gpurun_out/c1_linkwise_fg23.log 2>&1; echo "linkwise23 rc=$?"
grep -E "links|passed|failed|Error|error" gpurun_out/c1_linkwise_fg23.log | cut -c1-250 | tail -20
timeout 100 python tools/probe_hbm.py > gpurun_out/c1_probe_hbm.log 2>&1; cat gpurun_out/c1_probe_hbm.log
