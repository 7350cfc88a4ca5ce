#!/bin/bash
# round 2, visit A: validate the staged pieces (nfft-16384 overlap-save, config-3 receiver ring) on hardware,
# box topology, per-kernel table, and one ncu --set full of the staged 16k kernel.
set -u
O=gpurun_out; mkdir -p $O; export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/r2a_smi.txt; nproc >> $O/r2a_smi.txt
numactl -H >> $O/r2a_smi.txt 2>&1; nvidia-smi topo -m >> $O/r2a_smi.txt 2>&1; lscpu >> $O/r2a_smi.txt 2>&1
LRC_FASTFIR_STAGED=1 LRC_TEST_STAGED=1 timeout 500 python -m pytest tests/test_gpu_ook_fastfir.py tests/test_gpu_fm_resample.py tests/test_gpu_kpn.py -m gpu -x -q > $O/r2a_staged_pytest.log 2>&1
echo "staged pytest exit $?"; tail -n 3 $O/r2a_staged_pytest.log
(cd kpn && LRC_TEST_STAGED=1 timeout 120 ./test_gpu_blocks > ../$O/r2a_kpn_blocks.log 2>&1; echo "kpn staged exit $?"; tail -n 5 ../$O/r2a_kpn_blocks.log)
LRC_FASTFIR_STAGED=1 timeout 200 python tools/fastfir16k_check.py > $O/r2a_ff16k.json 2> $O/r2a_ff16k.err; echo "16k exit $?"; cat $O/r2a_ff16k.json
timeout 300 python tools/bench_kernels.py > $O/r2a_kernels_table.jsonl 2> $O/r2a_kernels_table.err; echo "table exit $?"; cut -c1-160 $O/r2a_kernels_table.jsonl
LRC_FASTFIR_STAGED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fastfir16k -c 1 --launch-skip 4 -o $O/r2a_ff16k_full -f \
    python tools/fastfir16k_check.py > $O/r2a_ncu_ff16k.log 2>&1; echo "ncu 16k exit $?"
