#!/bin/bash
# round 2, visit J: chain u8 instance, full GPU suite, bench N=1
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > $O/r2j_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 14 $O/r2j_pytest_gpu.log
timeout 600 python bench.py > $O/r2j_bench_n1.json 2> $O/r2j_bench_n1.err; echo "bench exit $?"; tail -5 $O/r2j_bench_n1.err; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/r2j_bench_n1.json").read().strip().splitlines()[-1])
    print("value", j["value"], "frac", j["roofline"]["frac"], "e2e", j["e2e"]["value"], "e2e_u8", j["e2e_u8"]["value"], "h2d", j["e2e"]["h2d_copy_gbs_per_gpu_all_ranks_copying"])
    for k, v in j.get("extra", {}).items():
        print(k, round(v["value"]), "Ms/s", round(v["ms"], 3), "ms frac", round(v["roofline"]["frac"], 3), v["roofline"]["bound"])
except Exception as e:
    print("parse failed", e)
PY
