"""ljgpu - B200 (sm_100a) Lennard-Jones integrator behind the reference's simulator interface.

The product is the C-ABI shared library ``libljgpu.so`` (``include/ljgpu.h``) built from ``csrc/`` and
``host/``; ``binding.py`` is a thin ctypes view of it used by ``tests/`` and ``bench.py``.  The
directory name contains hyphens, so import it with
``importlib.import_module("lennard-jones-live-simulator_b200")``.
"""
from .binding import (  # noqa: F401
    LJParams, LennardJonesParameters, LennardJonesSimulation, LJGpuError, load_library, library_path,
    KERNEL_AUTO, KERNEL_ALLPAIRS, KERNEL_CELL, KERNEL_VERLET, host_initialize, host_rng_reset,
    maxwell_boltzmann, device_count, EXPORTED_SYMBOLS, slab_layers, slab_unique_id, HostBuffer, measure_fp64_peak, SharedFrames,
)
