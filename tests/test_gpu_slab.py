"""Multi-device parity: the z-slab path (NCCL halo + migration + all-reduce) against the CPU oracle.
Needs at least two GPUs; skipped otherwise."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_slab_path_matches_oracle(world):
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29510 + world),
           os.path.join(ROOT, "tests", "slab_worker.py"), "65536", "10", "150"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SLAB OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
