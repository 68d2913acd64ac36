"""The C++ host adapter (same public interface as the reference's LennardJonesSimulation) is compiled
against libljgpu.so and run as the reference's callers would use it."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lennard-jones-live-simulator_b200")


@pytest.fixture(scope="module")
def adapter_exe(tmp_path_factory):
    subprocess.check_call(["make", "-C", PKG, "libljadapter.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblj_oracle.so"], stdout=subprocess.DEVNULL)
    exe = str(tmp_path_factory.mktemp("adapter") / "test_adapter")
    subprocess.check_call([
        "g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_adapter.cpp"),
        "-I", os.path.join(PKG, "host"), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
        "-L", PKG, "-lljadapter", "-lljgpu", "-L", os.path.join(ROOT, "oracle"), "-llj_oracle",
        f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{os.path.join(ROOT, 'oracle')}", "-lpthread", "-o", exe])
    return exe


@pytest.mark.skipif(bool(glob.glob("/dev/nvidia[0-9]*")), reason="a GPU is present")
def test_adapter_compiles_and_fails_loudly_without_a_device(adapter_exe):
    out = subprocess.run([adapter_exe, "nogpu"], capture_output=True, text=True)
    assert out.returncode == 0 and "OK nogpu" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_adapter_against_oracle_on_gpu(adapter_exe):
    out = subprocess.run([adapter_exe, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "OK gpu" in out.stdout, out.stdout + out.stderr
