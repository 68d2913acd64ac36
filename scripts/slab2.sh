#!/bin/bash
# Developer helper (run under `gpurun --gpus N`): slab parity worker + bench on N GPUs.   usage: bash scripts/slab2.sh N [bench args...]
N=${1:-2}; shift
cd "$(dirname "$0")/.."
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== slab worker, $N GPUs"
run 29611 tests/slab_worker.py 65536 10 150 2>&1 | grep -E "SLAB OK|Error|error|assert" | head -20
echo "== bench, $N GPUs"
run 29612 bench.py --gpus $N --steps 400 --warmup 50 "$@" 2> gpurun_out/slab_bench_n$N.err | tail -1 > gpurun_out/slab_bench_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/slab_bench_n$N.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "parity")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print({k: (v["us_per_call"], v["calls"]) for k, v in d["whole_step"]["stage_window"]["stages"].items()})
print({k: v.get("value") for k, v in d["configs"].items()})
PY
tail -5 gpurun_out/slab_bench_n$N.err
