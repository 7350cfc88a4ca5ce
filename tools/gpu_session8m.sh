#!/bin/bash
# final validation of the round: parity suite, smoke, headline bench (both arms), per-kernel table with the CPU legs,
# the ncu launch list of the bench command and one --set full capture of the chain kernel (traffic)
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > $O/s8m_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 2 $O/s8m_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 300 python bench.py > $O/s8m_bench_n1.json 2> $O/s8m_bench_n1.err; echo "bench exit $?"; cut -c1-200 $O/s8m_bench_n1.json
timeout 300 python tools/bench_kernels.py --cpu > $O/s8m_kernels_table.jsonl 2> $O/s8m_kernels_table.err; echo "table exit $?"; cut -c1-120 $O/s8m_kernels_table.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s8m_launches_chain.csv python bench.py --steps 4 --warmup 3 --no-cpu > $O/s8m_ncu_bench.log 2>&1; echo "ncu list exit $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -c 1 --launch-skip 3 -o $O/s8m_chain_full -f python bench.py --steps 2 --warmup 3 --no-cpu > $O/s8m_ncu_chain.log 2>&1; echo "ncu full exit $?"
