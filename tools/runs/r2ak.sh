#!/bin/bash
# round 2, visit AK (8 GPUs, last tree of the round): bench.py under torchrun with the extra block and the scaling diagnostics; box topology
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
N=${1:-8}
(nvidia-smi topo -m; nproc; lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)"; cat /sys/devices/system/node/online; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor)" = "0x10de" ]; then echo "$d numa $(cat $d/numa_node) class $(cat $d/class)"; fi; done) > $O/r2ak_topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/r2ak_bench_n$N.json 2> $O/r2ak_bench_n$N.err; echo "bench n$N exit $?"; tail -5 $O/r2ak_bench_n$N.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2ak_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", j["value"], "ms", j["ms_per_step"], "frac", j["roofline"]["frac"], "e2e", j["e2e"]["value"], "e2e_u8", j["e2e_u8"]["value"], "h2d", j["e2e"]["h2d_copy_gbs_per_gpu_all_ranks_copying"])
    print(j.get("scaling_diagnostics"))
    for k, v in j.get("extra", {}).items():
        print(k, round(v["value"]), "Ms/s", round(v["ms"], 3), "ms frac", round(v["roofline"]["frac"], 3), v["roofline"]["bound"], v.get("gathered_to_rank0", ""))
    print(j["run"])
except Exception as e:
    print("parse failed", e)
PY
head -30 $O/r2ak_topo_n$N.txt
