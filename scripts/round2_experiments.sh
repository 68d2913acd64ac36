#!/bin/bash
# Prepared A/B runs for the experiments that were written after round 1's GPU budget was spent.
# Build the variants HERE (no GPU needed), then run the printed command under gpurun in ONE call (about 1 GPU-minute):
#
#   scripts/round2_experiments.sh build
#   gpurun --timeout 600 -- 'bash scripts/round2_experiments.sh run'
#
# Parity of the pair path first: LJGPU_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_parity.py -m gpu -k experimental_pair
set -e
cd "$(dirname "$0")/.."
V=lennard-jones-live-simulator_b200/build/variants
case "$1" in
build)
    scripts/build_variant.sh prune2 -DLJ_TILE_PRUNE_MINBLOCKS=2
    scripts/build_variant.sh pair_g1b3 -DLJ_PAIR_GROUP=1 -DLJ_PAIR_MINBLOCKS=3
    scripts/build_variant.sh pair_g2b3 -DLJ_PAIR_MINBLOCKS=3
    scripts/build_variant.sh pair_bx5 -DLJ_PAIR_BX=5
    scripts/build_variant.sh pair_bx4g1b3 -DLJ_PAIR_BX=4 -DLJ_PAIR_GROUP=1 -DLJ_PAIR_MINBLOCKS=3
    ;;
run)
    probe() { echo "== $1"; shift; env "$@" timeout 90 python scripts/perf_probe.py 1048576 3 300 2>&1 | grep -E "unprofiled|force|ghost|rebuild  |Error|error"; }
    LJGPU_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -k experimental_pair -x -q 2>&1 | tail -3
    probe "default" X=1
    probe "prune pass with its own register budget" LJGPU_LIBRARY=$PWD/$V/libljgpu_prune2.so
    probe "per-atom path, lists in bank-class order" LJGPU_LIST_ORDER=1
    probe "pair path, lists in bank-class order" LJGPU_PAIRLIST=1 LJGPU_LIST_ORDER=1
    probe "pair path 1 entry x 2 atoms 3 CTAs/SM, lists in bank-class order" LJGPU_PAIRLIST=1 LJGPU_LIST_ORDER=1 LJGPU_LIBRARY=$PWD/$V/libljgpu_pair_g1b3.so
    probe "pair path (2 entries x 2 atoms, 2 CTAs/SM, 120 regs)" LJGPU_PAIRLIST=1
    probe "pair path without the chain ordering" LJGPU_PAIRLIST=1 LJGPU_PAIR_ORDER=0
    probe "pair path without the lane balance" LJGPU_PAIRLIST=1 LJGPU_PAIR_BALANCE=0
    probe "pair path, 1 entry x 2 atoms, 3 CTAs/SM" LJGPU_PAIRLIST=1 LJGPU_LIBRARY=$PWD/$V/libljgpu_pair_g1b3.so
    probe "pair path, 2 entries x 2 atoms, 3 CTAs/SM (spills)" LJGPU_PAIRLIST=1 LJGPU_LIBRARY=$PWD/$V/libljgpu_pair_g2b3.so
    probe "pair path, bricks 5x2x2" LJGPU_PAIRLIST=1 LJGPU_LIBRARY=$PWD/$V/libljgpu_pair_bx5.so
    probe "pair path, bricks 4x2x2, 1 entry x 2 atoms, 3 CTAs/SM" LJGPU_PAIRLIST=1 LJGPU_LIBRARY=$PWD/$V/libljgpu_pair_bx4g1b3.so
    ;;
run2)   # two GPUs: gpurun --gpus 2 --timeout 600 -- 'bash scripts/round2_experiments.sh run2'
    for fused in "" 1; do
        echo "== 2 GPUs, LJGPU_SLAB_FUSED_PUSH=$fused"
        env ${fused:+LJGPU_SLAB_FUSED_PUSH=1} timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
            --master-port 29577 bench.py --gpus 2 --steps 2000 --warmup 200 --no-cpu 2>/dev/null | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_step')}, d['whole_step']['stage_us_per_call']) for d in (json.loads(l) for l in sys.stdin if l.startswith('{'))]"
    done
    ;;
*) echo "usage: $0 build|run|run2"; exit 2;;
esac
