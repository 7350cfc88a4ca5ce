#!/bin/bash
# One GPU-box visit: parity suite, headline bench, per-kernel table (A/B of the packed kernels), ncu of the new
# kernels, compute-sanitizer on the tests that drive them.  Every step has its own timeout and log under
# gpurun_out/; steps are ordered by importance (the box budget may cut the tail).
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/s8_gpu.txt 2>&1

echo "== pytest -m gpu"; date +%s > $O/s8_t0
timeout 600 python -m pytest tests -m gpu -x -q --durations=12 > $O/s8_pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $O/s8_pytest_gpu.log
tail -5 $O/s8_pytest_gpu.log

echo "== bench.py"
timeout 300 python bench.py > $O/s8_bench_n1.json 2> $O/s8_bench_n1.err
echo "bench exit $?"; cat $O/s8_bench_n1.json | cut -c1-400

echo "== kernel table"
timeout 240 python tools/bench_kernels.py > $O/s8_kernels_table.jsonl 2> $O/s8_kernels_table.err
echo "table exit $?"; cut -c1-160 $O/s8_kernels_table.jsonl
echo "== A/B scalar resampler"
LRC_RS_VARIANT=0 timeout 120 python tools/bench_kernels.py --only fm,fmchain > $O/s8_kernels_rs_scalar.jsonl 2>> $O/s8_kernels_table.err
cut -c1-160 $O/s8_kernels_rs_scalar.jsonl

echo "== ncu --set full: packed u8 FIR + packed resampler"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'fir_tile_kernel|resample_dec2_kernel' -c 6 \
    -o $O/s8_fir_u8_rs_dec2 -f python tools/bench_kernels.py --quick --only fir,fm > $O/s8_ncu.log 2>&1
echo "ncu exit $?"

echo "== compute-sanitizer memcheck on the tests of the new kernels"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fm_resample.py -x -q \
    > $O/s8_sanitizer_fm_resample.log 2>&1
echo "sanitizer(fm_resample) exit $?" | tee -a $O/s8_sanitizer_fm_resample.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_core.py -x -q -k "fir" \
    > $O/s8_sanitizer_fir.log 2>&1
echo "sanitizer(fir) exit $?" | tee -a $O/s8_sanitizer_fir.log
tail -3 $O/s8_sanitizer_fm_resample.log $O/s8_sanitizer_fir.log
