"""The product's host code and kernels together, without a GPU: lj_sim.cu + lj_sort.cu + the device headers compiled
by g++ against the CUDA-on-CPU shim (tests/cuda_emul/build_emul_lib.py) into a library with the C ABI of
include/ljgpu.h, and the `-m gpu` parity tests run against THAT library in a subprocess (LJGPU_LIBRARY).
Test infrastructure only: it checks logic (binning, lists, batching, re-binning, publication, probes) where no
device is available; the shipped library has no CPU path and the parity claims rest on the runs on the B200."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "cuda_emul"))

# the O(N^2) all-pairs cases and the long NVE run take minutes under emulation; the N = 2^20 case needs the real thing
SKIP = [
    "test_ten_step_window[n4096-allpairs]",
    "test_window_from_equilibrated_liquid[1]",
    "test_cell_ids_and_neighbor_counts_bit_exact[1]",
    "test_nve_energy_conservation_and_thermostat",
    "test_full_size_n1048576_against_reference_golden_vectors",
    "test_liquid_window_strict_1e12[n65536]",
    "test_liquid_window_strict_1e12[n1048576]",
    "test_product_list_counts_match_reference",
    "test_full_size_liquid_against_reference_golden_vectors",
    "test_long_run_statistics_match_the_reference",
]


def test_gpu_parity_suite_against_the_emulated_library():
    import build_emul_lib
    lib = build_emul_lib.build()
    env = dict(os.environ, LJGPU_LIBRARY=lib, LJGPU_NO_GRAPH="1")
    cmd = [sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"]
    for name in SKIP:
        cmd += ["--deselect", f"tests/test_gpu_parity.py::{name}"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-2000:]
    assert " passed" in out.stdout and "failed" not in out.stdout


import pytest  # noqa: E402


@pytest.mark.parametrize("world,nccl_loop", [(2, False), (4, False), (2, True)], ids=["2-peer-memory", "4-peer-memory", "2-nccl-loop"])
def test_slab_path_on_the_emulated_library(world, nccl_loop):
    """The multi-device path with the "devices" as threads of one process (tests/slab_emul_worker.py): halo push and
    mailbox all-reduce over "peer memory", migration through the stand-in NCCL, pruning, re-binnings - same checks as
    the multi-GPU worker (tests/slab_worker.py).  With four devices the lower and upper neighbour are different ranks."""
    import build_emul_lib
    lib = build_emul_lib.build()
    env = dict(os.environ, LJGPU_LIBRARY=lib, LJGPU_NO_GRAPH="1",
               LD_LIBRARY_PATH=os.path.join(HERE, "cuda_emul", "_build") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("LJGPU_PRUNE_SKIN", None)
    if nccl_loop:
        env["LJGPU_SLAB_NCCL"] = "1"
    else:
        env.pop("LJGPU_SLAB_NCCL", None)
    cmd = [sys.executable, os.path.join(HERE, "slab_emul_worker.py"), str(world), "32768", "5", "80", "200", "2"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0 and "SLAB EMUL OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_slab_failure_on_one_device_is_reported_by_all():
    """A capacity check that fails on one device (injected: LJGPU_TEST_FAIL_RANK) must end the collective call on every
    device with the same error, not leave the others waiting in a collective."""
    import build_emul_lib
    lib = build_emul_lib.build()
    env = dict(os.environ, LJGPU_LIBRARY=lib, LJGPU_NO_GRAPH="1", LJGPU_TEST_FAIL_RANK="1",
               LD_LIBRARY_PATH=os.path.join(HERE, "cuda_emul", "_build") + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    cmd = [sys.executable, os.path.join(HERE, "slab_emul_worker.py"), "3", "32768", "5", "80", "150", "2"]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SLAB EMUL FAILED TOGETHER" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_edge_cases_on_the_emulated_library():
    """Two atoms, three cells per axis, mostly empty cells, non-unit sigma / epsilon / mass, tiny and wide shells,
    negative tau, a large time step: all three force kernels against the oracle (tests/emul_edge_worker.py)."""
    import build_emul_lib
    lib = build_emul_lib.build()
    env = dict(os.environ, LJGPU_LIBRARY=lib, LJGPU_NO_GRAPH="1")
    for k in ("LJGPU_PRUNE_SKIN", "LJGPU_PAIRLIST", "LJGPU_BRICK"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, os.path.join(HERE, "emul_edge_worker.py")], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "EDGE CASES OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def test_cpp_adapter_on_the_emulated_library(tmp_path):
    """The C++ drop-in class (host/lennardjonessimulation.cpp, three rotating mirror buffers over ljgpu_step_publish,
    mirror modes, parameter reload, movie file, driver thread) linked against the emulated library: the same
    executable as tests/test_adapter_cpp.py runs on the GPU, with a shorter driver-thread run."""
    import build_emul_lib
    lib = build_emul_lib.build()
    pkg = os.path.join(ROOT, "lennard-jones-live-simulator_b200")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblj_oracle.so"], stdout=subprocess.DEVNULL)
    exe = str(tmp_path / "test_adapter_emul")
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", os.path.join(HERE, "cpp", "test_adapter.cpp"),
        os.path.join(pkg, "host", "lennardjonessimulation.cpp"), os.path.join(pkg, "host", "threadintegrate.cpp"),
        "-I", os.path.join(pkg, "host"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
        lib, "-L", os.path.join(ROOT, "oracle"), "-llj_oracle",
        f"-Wl,-rpath,{os.path.dirname(lib)}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}", "-lpthread", "-o", exe])
    env = dict(os.environ, LJGPU_NO_GRAPH="1", LJGPU_ADAPTER_TEST_ITERS="40")
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "OK gpu" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
