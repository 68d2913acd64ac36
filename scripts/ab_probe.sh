#!/bin/bash
# Developer helper: A/B a list of library variants (build/variants/libljgpu_<name>.so; "default" = the in-tree library)
# with scripts/perf_probe.py at N = 2^20.   usage (under gpurun): bash scripts/ab_probe.sh default w8s1 w8s2 ...
cd "$(dirname "$0")/.."
V=$PWD/lennard-jones-live-simulator_b200/build/variants
for name in "$@"; do
    echo "== $name"
    if [ "$name" = default ]; then lib=""; else lib="LJGPU_LIBRARY=$V/libljgpu_$name.so"; fi
    env $lib timeout 120 python scripts/perf_probe.py ${PROBE_N:-1048576} 3 ${PROBE_STEPS:-300} ${PROBE_WARM:-200} ${PROBE_DT:-0.001} 2>&1 | grep -E "unprofiled|force|rebuild  |kick|Error|error"
done
