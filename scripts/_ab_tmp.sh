python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
echo "== arranged (default)"; python scripts/perf_probe.py 1048576 3 300 2>&1 | grep -E "unprofiled|force|rebuild|kick|rror"
echo "== natural order"; LJGPU_NO_ARRANGE=1 python scripts/perf_probe.py 1048576 3 300 2>&1 | grep -E "unprofiled|force|rebuild|kick|rror"
echo "== 65536 arranged"; python scripts/perf_probe.py 65536 3 300 2>&1 | grep -E "unprofiled|force|rebuild|kick|rror"
