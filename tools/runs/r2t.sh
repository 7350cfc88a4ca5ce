#!/bin/bash
# round 2, visit T: bench.py at N = 1 after the OOK / generic chain / FIR tile work
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python bench.py > $O/r2t_bench_n1.json 2> $O/r2t_bench_n1.err; echo "bench exit $?"; tail -5 $O/r2t_bench_n1.err; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2t_bench_n1.json").read().strip().splitlines()[-1])
    print("value", j["value"], "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"], "e2e", j["e2e"]["value"], "e2e_u8", j["e2e_u8"]["value"], "cpu", j["cpu_baseline"]["value"])
    for k, v in j.get("extra", {}).items():
        print(k, round(v["value"]), "Ms/s", round(v["ms"], 3), "ms frac", round(v["roofline"]["frac"], 3), v["roofline"]["bound"])
    print(j["clocks"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2t_bench_ref.json 2> $O/r2t_bench_ref.err; echo "ref exit $?"; cut -c1-400 $O/r2t_bench_ref.json
