"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/ljgpu.h
declares, its host-side pieces agree with the oracle, and without a device it fails loudly."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "ljgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ljgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lj, ljlib):
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(ljlib, n), f"libljgpu.so does not export {n}"
    assert sorted(lj.EXPORTED_SYMBOLS) == names


def test_library_has_no_unexpected_dependencies(lj):
    """The product must not link the oracle or torch; CUDA runtime is linked statically."""
    out = subprocess.run(["ldd", lj.library_path()], capture_output=True, text=True).stdout
    assert "lj_oracle" not in out and "ljref" not in out and "torch" not in out


def test_library_contains_sm100a_code(lj):
    out = subprocess.run(["cuobjdump", "-lelf", lj.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_params_struct_layout_matches_header(lj, tmp_path):
    """The ctypes mirror of struct ljgpu_params / ljgpu_timers against the C compiler's view of the header."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ljgpu.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(ljgpu_params), offsetof(ljgpu_params, tau),'
                   ' offsetof(ljgpu_params, device), sizeof(ljgpu_timers), offsetof(ljgpu_timers, kernel_launches));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, off_tau, off_dev, tsize, off_launch = map(int, subprocess.check_output([str(exe)]).split())
    from importlib import import_module
    b = import_module(lj.__name__ + ".binding")
    assert C.sizeof(lj.LJParams) == size
    assert lj.LJParams.tau.offset == off_tau and lj.LJParams.device.offset == off_dev
    assert C.sizeof(b.LJTimers) == tsize and b.LJTimers.kernel_launches.offset == off_launch


def test_host_initialize_is_the_reference_start_state(lj):
    for cfg in (ol.DEFAULT, ol.config(4096), ol.config(3000, rho=0.6)):
        o = ol.Oracle(cfg)
        lj.host_rng_reset()
        st = lj.host_initialize(lj.LennardJonesParameters(**cfg))
        for a, b in zip(st, o.state()):
            assert np.array_equal(a, b)


def test_host_generator_continues_across_simulations_like_the_reference(lj):
    """get_gauss() keeps one process-wide generator (lennardjonessimulation.cpp:249-250)."""
    p = lj.LennardJonesParameters(**ol.DEFAULT)
    lj.host_rng_reset()
    a = lj.host_initialize(p)
    b = lj.host_initialize(p)
    assert not np.array_equal(a[3], b[3])
    o1 = ol.Oracle(ol.DEFAULT)
    o2 = ol.Oracle(ol.DEFAULT, reset_rng=False)
    assert np.array_equal(b[3], o2.velocities()[0]) and np.array_equal(a[3], o1.velocities()[0])


def test_maxwell_boltzmann_matches_graphwidget_formula(lj):
    got = lj.maxwell_boltzmann(1.0, 1.0, 1000)
    want = ol.oracle_maxwell_boltzmann(1.0, 1.0, 1000)
    assert np.array_equal(got, want)
    assert abs(got.sum() - 1000) < 1.0          # left-edge rule of a normalised density


def test_parameter_map_behaves_like_the_reference(lj):
    p = lj.LennardJonesParameters.default()
    assert p.get_param("cell_length") == 14.938 and p.get_param("nr_particles", int) == 1000
    with pytest.raises(RuntimeError, match="Could not find parameter: nope"):
        p.get_param("nope")
    p.set_param("nr_particles", "2000")         # the New-simulation dialog stores strings (dialognewsimulation.cpp:73-78)
    assert p.get_param("nr_particles", int) == 2000


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_cpu_fallback(lj, ljlib):
    assert lj.device_count() == 0
    with pytest.raises(lj.LJGpuError, match="no CUDA device"):
        lj.LennardJonesSimulation(lj.LennardJonesParameters.default())
    assert ljlib.ljgpu_step(None, 1, 0.001) != 0
    assert b"null" in ljlib.ljgpu_last_error()


def test_argument_validation_happens_before_device_use(lj, ljlib):
    st = lj.LennardJonesParameters(**dict(ol.DEFAULT, cell_length=8.0)).to_struct()
    arrs = [np.zeros(1000) for _ in range(6)]
    h = C.c_void_p()
    rc = ljlib.ljgpu_create(C.byref(st), *[a.ctypes.data_as(C.c_void_p) for a in arrs], C.byref(h))
    assert rc == -1 and b"27-cell stencil" in ljlib.ljgpu_last_error()
    assert ljlib.ljgpu_create(None, *[None] * 6, C.byref(h)) == -1


def test_adapter_compiles_against_the_references_qt_parameter_map(tmp_path):
    """LJGPU_WITH_QT: the drop-in class built against the reference's OWN lennardjonesparameters.h (QMap<QString, QVariant>;
    Qt types from oracle/shim here) - what MainWindow / DialogNewSimulation fill.  Runs where /root/reference exists."""
    ref = "/root/reference/src/simulator"
    if not os.path.exists(os.path.join(ref, "lennardjonesparameters.h")):
        pytest.skip("reference sources not present (GPU box)")
    pkg = os.path.join(ROOT, "lennard-jones-live-simulator_b200")
    exe = str(tmp_path / "test_qt_switch")
    cmd = ["g++", "-std=c++17", "-O1", "-DLJGPU_WITH_QT", f'-DLJGPU_QT_PARAMETERS_HEADER="{ref}/lennardjonesparameters.h"',
           "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(pkg, "host"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_qt_switch.cpp"), os.path.join(pkg, "host", "lennardjonessimulation.cpp"),
           os.path.join(ref, "lennardjonesparameters.cpp"), "-L", pkg, "-lljgpu", f"-Wl,-rpath,{pkg}", "-lpthread", "-o", exe]
    subprocess.run(cmd, check=True, timeout=300)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "QT SWITCH OK" in out.stdout, out.stdout + out.stderr
