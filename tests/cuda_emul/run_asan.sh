#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Memory checking without compute-sanitizer: the emulated product library built with
# AddressSanitizer (global memory = heap blocks with red zones, dynamic shared memory likewise), then the parity
# subset, the edge cases and the slab worker run against it.  Takes ~10 minutes.
#   tests/cuda_emul/run_asan.sh
set -e
cd "$(dirname "$0")/../.."
python tests/cuda_emul/build_emul_lib.py >/dev/null                     # plain build: provides the stand-in NCCL
LIB=$(python tests/cuda_emul/build_emul_lib.py --asan | tail -1)
export LJGPU_LIBRARY=$LIB LJGPU_NO_GRAPH=1
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)"
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1
export LD_LIBRARY_PATH=$PWD/tests/cuda_emul/_build:$LD_LIBRARY_PATH
T=tests/test_gpu_parity.py
python -m pytest $T -m gpu -x -q -p no:cacheprovider \
    --deselect "$T::test_ten_step_window[n4096-allpairs]" --deselect "$T::test_window_from_equilibrated_liquid[1]" \
    --deselect "$T::test_cell_ids_and_neighbor_counts_bit_exact[1]" --deselect $T::test_nve_energy_conservation_and_thermostat \
    --deselect $T::test_full_size_n1048576_against_reference_golden_vectors --deselect $T::test_full_size_liquid_against_reference_golden_vectors \
    --deselect $T::test_long_run_statistics_match_the_reference
python tests/emul_edge_worker.py | tail -1
for w in 2 4; do python tests/slab_emul_worker.py $w 32768 3 70 200 2 | tail -1; done
LJGPU_SLAB_NCCL=1 python tests/slab_emul_worker.py 2 32768 3 70 200 2 | tail -1
echo "ASAN RUN CLEAN"
