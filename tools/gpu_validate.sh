#!/bin/bash
# One GPU-box visit that produces everything a round's profiles/ needs (run it through gpurun from the repo root):
#   parity suite, smoke, headline bench line, per-kernel table with the CPU legs, the ncu launch list of the bench
#   command (our kernels only) and one --set full capture of the chain kernel (DRAM traffic for roofline.traffic).
# Outputs land in gpurun_out/ under the prefix given as $1 (default "val"); copy what should be judged to profiles/.
set -u
P=${1:-val}
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q --durations=5 > $O/${P}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 $O/${P}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 900 python bench.py > $O/${P}_bench_n1.json 2> $O/${P}_bench_n1.err; echo "bench exit $?"; cut -c1-200 $O/${P}_bench_n1.json
timeout 900 python tools/bench_kernels.py --cpu > $O/${P}_kernels_table.jsonl 2> $O/${P}_kernels_table.err; echo "table exit $?"; cut -c1-120 $O/${P}_kernels_table.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chain_kernel|psd_reduce|unpack' -c 400 --csv \
    --log-file $O/${P}_launches_chain.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > $O/${P}_ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 --launch-skip 3 -o $O/${P}_chain_full -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/${P}_ncu_chain.log 2>&1; echo "ncu full exit $?"
