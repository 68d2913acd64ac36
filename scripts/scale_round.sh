#!/bin/bash
# Developer helper (under `gpurun --gpus N`): slab parity worker, then bench.py as the driver launches it (20 steps) and over a long
# window (1000 steps).   usage: bash scripts/scale_round.sh N tag
N=$1; tag=$2
cd "$(dirname "$0")/.."
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== slab worker, $N GPUs"
run 29611 tests/slab_worker.py 262144 5 120 200 2 2>&1 | grep -E "SLAB OK|Error|error|assert" | head -5
for k in 20:5 1000:100; do
  steps=${k%%:*}; warm=${k##*:}
  run 29612 bench.py --gpus $N --steps $steps --warmup $warm 2> gpurun_out/${tag}_n${N}_s${steps}.err | tail -1 > gpurun_out/${tag}_n${N}_s${steps}.json
  python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_n${N}_s${steps}.json").read())
print("steps $steps:", {k: d[k] for k in ("value", "ms_per_step")}, "parity", d["parity"].get("max_rel_err"), d["parity"]["ok"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("   ", {k: (round(v["us_per_call"], 1) if v["us_per_call"] else None, v["calls"]) for k, v in d["whole_step"]["stage_window"]["stages"].items()})
print("   ", {k: (v.get("value"), v.get("ms_per_step")) for k, v in d.get("configs", {}).items()}, "roofline frac", d["roofline"]["frac"])
PY
done
