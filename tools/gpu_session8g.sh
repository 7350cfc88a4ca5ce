#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== stream reference points"; timeout 200 python tools/stream_refpoints.py 2>&1 | tee $O/s8g_stream_refpoints.txt | cut -c1-200
echo "== full table"; timeout 300 python tools/bench_kernels.py > $O/s8g_kernels_table.jsonl 2> $O/s8g_kernels_table.err; cut -c1-150 $O/s8g_kernels_table.jsonl
echo "== pytest"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 1
echo "== kpn"; ls kpn/
