#!/bin/bash
# round 2, visit BB: the two slicer forms at 4096 streams with quiet blocks skipped
set -u
export PYTHONUNBUFFERED=1
for kc in 1 0 1 0; do echo "streams 4096 KC=$kc"; LRC_OOK_KC=$kc timeout 200 python tools/bench_kernels.py --only ook --ook-streams 4096 2>/dev/null | tail -1 | cut -c1-120; done
for n in 3072 1024; do for kc in 1 0; do echo "streams $n KC=$kc"; LRC_OOK_KC=$kc timeout 200 python tools/bench_kernels.py --only ook --ook-streams $n 2>/dev/null | tail -1 | cut -c1-120; done; done
