#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
: > $O/s8k_chain_grid.txt
for k in 1 2 4 16; do
  echo "LRC_CHAIN_GRID=$k" | tee -a $O/s8k_chain_grid.txt
  LRC_CHAIN_GRID=$k timeout 150 python bench.py --no-cpu --steps 20 2>>$O/s8k.err | tee -a $O/s8k_chain_grid.txt | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'])"
done
timeout 200 python tools/bench_kernels.py > $O/s8k_kernels_table.jsonl 2>>$O/s8k.err; cut -c1-130 $O/s8k_kernels_table.jsonl
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 1
