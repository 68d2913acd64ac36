#!/bin/bash
# Developer helper (under gpurun): GPU parity tests, a stage-timed probe at N = 2^20, and one ncu --set full capture of the
# force pass.   usage: bash scripts/gpu_round.sh <tag> [tests|notests] [ncu|noncu]
tag=$1; tests=${2:-tests}; prof=${3:-ncu}
cd "$(dirname "$0")/.."
if [ "$tests" = tests ]; then python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log; fi
python scripts/perf_probe.py 1048576 3 300 200 > gpurun_out/${tag}_probe.log 2>&1; grep -E "unprofiled|force|rebuild  |kick|rror" gpurun_out/${tag}_probe.log
python scripts/perf_probe.py 65536 3 300 200 2>&1 | grep -E "unprofiled|force|rebuild  |kick|rror" | tee gpurun_out/${tag}_probe65k.log
if [ "$prof" = ncu ]; then
  ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 70 -c 2 -o gpurun_out/${tag}_prof python scripts/perf_probe.py 1048576 3 40 60 > gpurun_out/${tag}_ncu.log 2>&1; tail -2 gpurun_out/${tag}_ncu.log
fi
