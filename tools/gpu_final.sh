#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -q -x -m gpu -s > gpurun_out/f_gpu_tests_verbose.log 2>&1; echo "gpu tests rc=$?"
tail -2 gpurun_out/f_gpu_tests_verbose.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"
for c in c1 c3 c4 c5; do
  timeout 300 python bench.py --config $c --no-cpu-baseline --no-augment --steps 8 --warmup 3 > gpurun_out/f_bench_$c.json 2> gpurun_out/f_bench_$c.err; echo "bench $c rc=$?"
done
timeout 200 python bench.py --config c1 --graph --no-cpu-baseline --no-augment --steps 20 --warmup 5 > gpurun_out/f_bench_c1_graph.json 2> gpurun_out/f_bench_c1_graph.err; echo "bench c1 graph rc=$?"
python - <<'PY'
import json
for n in ("n1", "c1", "c1_graph", "c3", "c4", "c5"):
    try:
        d = json.loads(open("gpurun_out/f_bench_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.1f ms %.2f e2e %.1f frac %.4f sm %s %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
