#!/bin/bash
# A/B of the packed resampler configurations (threads x stages) + parity of each
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
for cfg in 1281 1282 641 642; do
  echo "== LRC_RS_CFG=$cfg"
  LRC_RS_CFG=$cfg timeout 300 python -m pytest tests/test_gpu_fm_resample.py -x -q 2>&1 | tail -n 1
  LRC_RS_CFG=$cfg timeout 200 python tools/bench_kernels.py --only fm,fmchain 2>> $O/s8f.err | grep -v "FM discr" | tee -a $O/s8f_rs_cfg_$cfg.jsonl | cut -c1-150
done
echo "== scalar"
LRC_RS_VARIANT=0 timeout 200 python tools/bench_kernels.py --only fm,fmchain 2>> $O/s8f.err | grep -v "FM discr" | tee -a $O/s8f_rs_scalar.jsonl | cut -c1-150
