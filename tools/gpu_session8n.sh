#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chain_kernel|psd_reduce|unpack' -c 400 --csv --log-file $O/s8n_launches_chain.csv python bench.py --steps 4 --warmup 3 --no-cpu > $O/s8n_ncu_bench.log 2>&1; echo "ncu list exit $?"
grep -c chain_kernel $O/s8n_launches_chain.csv
