cd /root/repo
V=$PWD/lennard-jones-live-simulator_b200/build/variants
run() { echo "== $1"; shift; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 scripts/slab_repro.py 262144 1200 2>&1 | grep -E "steps|REPRO|LJGpuError" | tail -4; }
run default A=1
run lookahead16 LJGPU_BRICK_LOOKAHEAD=16
run nograph LJGPU_NO_GRAPH=1
run old LJGPU_LIBRARY=$V/libljgpu_old.so
