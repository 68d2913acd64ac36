"""ctypes view of libljgpu.so with the method names of the reference's LennardJonesSimulation.

Reference interface mirrored: /root/reference/src/simulator/lennardjonessimulation.h:113-155 and
lennardjonesparameters.h:32-69.  There is no CPU fallback: if the CUDA library is missing or no
device is usable the calls raise.
"""
import ctypes as C
import os

import numpy as np

KERNEL_AUTO, KERNEL_ALLPAIRS, KERNEL_CELL, KERNEL_VERLET = 0, 1, 2, 3

# every symbol include/ljgpu.h declares
EXPORTED_SYMBOLS = [
    "ljgpu_last_error", "ljgpu_device_count", "ljgpu_host_initialize", "ljgpu_host_rng_reset",
    "ljgpu_create", "ljgpu_destroy", "ljgpu_set_state", "ljgpu_update_params", "ljgpu_step",
    "ljgpu_step_n", "ljgpu_get_positions", "ljgpu_get_velocities", "ljgpu_get_forces",
    "ljgpu_get_energies", "ljgpu_get_energy_history", "ljgpu_get_speed_histogram",
    "ljgpu_maxwell_boltzmann", "ljgpu_get_cell_ids", "ljgpu_get_neighbor_counts", "ljgpu_get_list_counts",
    "ljgpu_write_movie_frame", "ljgpu_set_profiling", "ljgpu_get_timers", "ljgpu_reset_timers",
    "ljgpu_get_layout", "ljgpu_step_n_timed", "ljgpu_get_list_stats", "ljgpu_alloc_host", "ljgpu_free_host", "ljgpu_measure_fp64_peak", "ljgpu_register_host", "ljgpu_unregister_host", "ljgpu_set_sample_stride", "ljgpu_get_graph_stats", "ljgpu_get_speed_statistics",
    "ljgpu_slab_layers", "ljgpu_slab_unique_id", "ljgpu_create_slab", "ljgpu_slab_info",
    "ljgpu_step_publish", "ljgpu_publish_wait",
]


class LJGpuError(RuntimeError):
    pass


class HostBuffer:
    """Page-locked host memory from ljgpu_alloc_host viewed as a float64 numpy array (freed on close / gc)."""

    def __init__(self, count):
        self._lib = load_library()
        p = C.c_void_p()
        rc = self._lib.ljgpu_alloc_host(C.c_size_t(count * 8), C.byref(p))
        if rc != 0:
            raise LJGpuError(self._lib.ljgpu_last_error().decode())
        self.ptr = p.value
        self.array = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(count,))

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.ljgpu_free_host(C.c_void_p(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SharedFrames:
    """Two position frames (x | y | z, original atom order, 3 N doubles each) in host memory that every rank's process
    can write: page-locked memory on one device, POSIX shared memory registered with CUDA (ljgpu_register_host) by
    every rank on slabs - there ljgpu_step_publish lets each device deliver only its slice of the frame, so the
    consumer's view of the whole system is assembled by `world` PCIe links in parallel.  `dist` = torch.distributed
    (or None): used to agree on the segment names."""

    def __init__(self, n, world=1, rank=0, dist=None):
        self._lib = load_library()
        self.n, self.world, self.rank = int(n), int(world), int(rank)
        self._shm, self._host, self.frames = [], [], []
        if self.world == 1:
            for _ in range(2):
                hb = HostBuffer(3 * self.n)
                self._host.append(hb)
                self.frames.append(hb.array)
            return
        from multiprocessing import shared_memory
        names = [None, None]
        if self.rank == 0:
            for b in range(2):
                seg = shared_memory.SharedMemory(create=True, size=3 * self.n * 8)
                self._shm.append(seg)
                names[b] = seg.name
        dist.broadcast_object_list(names, src=0)
        if self.rank != 0:
            for b in range(2):
                self._shm.append(shared_memory.SharedMemory(name=names[b]))
        for seg in self._shm:
            arr = np.ndarray((3 * self.n,), dtype=np.float64, buffer=seg.buf)
            if self.rank == 0:
                arr[:] = 0.0                       # touch the pages before they are pinned
            self.frames.append(arr)
        dist.barrier()
        for arr in self.frames:
            _check(self._lib, self._lib.ljgpu_register_host(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)))

    def pointers(self):
        """[[x, y, z] of frame 0, [x, y, z] of frame 1] as raw addresses for integrate_publish."""
        return [[f.ctypes.data + k * self.n * 8 for k in range(3)] for f in self.frames]

    def xyz(self, b):
        f = self.frames[b]
        return f[:self.n], f[self.n:2 * self.n], f[2 * self.n:]

    def frame_is_finite(self, b):
        return bool(np.isfinite(self.frames[b]).all())

    def close(self):
        for arr in self.frames if self.world > 1 else []:
            self._lib.ljgpu_unregister_host(C.c_void_p(arr.ctypes.data))
        self.frames = []
        for hb in self._host:
            hb.close()
        self._host = []
        for seg in self._shm:
            try:
                seg.close()
                if self.rank == 0:
                    seg.unlink()
            except Exception:
                pass
        self._shm = []


class GraphStats(C.Structure):
    """struct ljgpu_graph_stats (include/ljgpu.h)."""
    _fields_ = [("mean", C.c_double * 3), ("stddev", C.c_double * 3), ("min", C.c_double * 3), ("max", C.c_double * 3),
                ("n_samples", C.c_uint32), ("reserved", C.c_uint32)]


class LJParams(C.Structure):
    """struct ljgpu_params (include/ljgpu.h)."""
    _fields_ = [
        ("cell_length", C.c_double), ("nr_particles", C.c_int32), ("force_kernel", C.c_int32),
        ("kT", C.c_double), ("mass", C.c_double), ("sigma", C.c_double), ("rcut", C.c_double),
        ("epsilon", C.c_double), ("tau", C.c_double), ("shell", C.c_double), ("stepsize", C.c_double),
        ("device", C.c_int32), ("reserved", C.c_int32),
    ]


class LJTimers(C.Structure):
    _fields_ = [(k, C.c_double) for k in
                ("ms_predict", "ms_kick_drift", "ms_force_prune", "ms_force", "ms_kick2", "ms_rebuild", "ms_halo")] + \
               [(k, C.c_uint64) for k in
                ("n_predict", "n_kick_drift", "n_force_prune", "n_force", "n_kick2", "n_rebuild", "n_halo",
                 "steps", "rebuilds", "kernel_launches")]


_LIB = None


def library_path():
    """In-tree libljgpu.so; LJGPU_LIBRARY overrides it (developer experiments with build variants)."""
    return os.environ.get("LJGPU_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libljgpu.so")


def load_library():
    """Load libljgpu.so (built in-tree by __graft_entry__.build() / make). Raises if absent."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise LJGpuError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    vp = C.c_void_p
    lib.ljgpu_last_error.restype = C.c_char_p
    lib.ljgpu_device_count.restype = C.c_int
    lib.ljgpu_host_initialize.argtypes = [C.POINTER(LJParams)] + [vp] * 6
    lib.ljgpu_create.argtypes = [C.POINTER(LJParams)] + [vp] * 6 + [C.POINTER(vp)]
    lib.ljgpu_destroy.argtypes = [vp]
    lib.ljgpu_destroy.restype = None
    lib.ljgpu_set_state.argtypes = [vp] + [vp] * 6
    lib.ljgpu_update_params.argtypes = [vp, C.POINTER(LJParams)]
    lib.ljgpu_step.argtypes = [vp, C.c_uint32, C.c_double]
    lib.ljgpu_step_n.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_double]
    for name in ("ljgpu_get_positions", "ljgpu_get_velocities", "ljgpu_get_forces"):
        getattr(lib, name).argtypes = [vp, vp, vp, vp]
    lib.ljgpu_get_energies.argtypes = [vp, dp, dp, dp]
    lib.ljgpu_get_energy_history.argtypes = [vp, C.c_uint32, vp, vp, vp, C.POINTER(C.c_uint32)]
    lib.ljgpu_get_speed_histogram.argtypes = [vp, C.c_uint32, C.c_double, vp]
    lib.ljgpu_maxwell_boltzmann.argtypes = [C.c_double, C.c_double, C.c_int32, C.c_uint32, C.c_double, vp]
    lib.ljgpu_get_cell_ids.argtypes = [vp, vp, vp]
    lib.ljgpu_get_neighbor_counts.argtypes = [vp, C.c_double, C.c_int, vp]
    lib.ljgpu_get_list_counts.argtypes = [vp, C.c_int, vp]
    lib.ljgpu_write_movie_frame.argtypes = [vp, C.c_char_p, C.c_int]
    lib.ljgpu_set_profiling.argtypes = [vp, C.c_int]
    lib.ljgpu_get_timers.argtypes = [vp, C.POINTER(LJTimers)]
    lib.ljgpu_reset_timers.argtypes = [vp]
    lib.ljgpu_get_layout.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                     C.POINTER(C.c_uint64)]
    lib.ljgpu_step_n_timed.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_double, dp]
    lib.ljgpu_get_list_stats.argtypes = [vp, dp, C.POINTER(C.c_uint32)]
    lib.ljgpu_measure_fp64_peak.argtypes = [C.c_int, dp]
    lib.ljgpu_register_host.argtypes = [vp, C.c_size_t]
    lib.ljgpu_set_sample_stride.argtypes = [vp, C.c_uint32]
    lib.ljgpu_get_graph_stats.argtypes = [vp, C.c_int, C.POINTER(GraphStats)]
    lib.ljgpu_get_speed_statistics.argtypes = [vp, C.c_uint32, C.c_double, vp, vp, dp, C.POINTER(C.c_uint32)]
    lib.ljgpu_unregister_host.argtypes = [vp]
    lib.ljgpu_step_publish.argtypes = [vp, C.c_uint32, C.c_double, vp, vp, vp]
    lib.ljgpu_publish_wait.argtypes = [vp]
    lib.ljgpu_alloc_host.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.ljgpu_free_host.argtypes = [vp]
    lib.ljgpu_slab_layers.argtypes = [C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    lib.ljgpu_slab_unique_id.argtypes = [vp]
    lib.ljgpu_create_slab.argtypes = [C.POINTER(LJParams), C.c_int, C.c_int, vp] + [vp] * 6 + [C.POINTER(vp)]
    lib.ljgpu_slab_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    _LIB = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise LJGpuError(f"ljgpu error {rc}: {lib.ljgpu_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def device_count():
    return load_library().ljgpu_device_count()


class LennardJonesParameters:
    """String-keyed parameter set with the reference's ten keys and its error behaviour:
    get_param of a missing key raises RuntimeError("Could not find parameter: <name>")
    (lennardjonesparameters.h:52-58)."""
    KEYS = ("cell_length", "nr_particles", "kT", "mass", "sigma", "rcut", "epsilon", "tau", "shell", "stepsize")

    def __init__(self, **kw):
        self._p = {}
        for k, v in kw.items():
            self.set_param(k, v)

    def set_param(self, name, value):
        self._p[name] = value

    def get_param(self, name, typ=float):
        if name not in self._p:
            raise RuntimeError("Could not find parameter: " + name)
        v = self._p[name]
        if typ is int:
            return int(round(float(v)))      # QVariant::value<int>() of a double/string rounds
        return float(v)

    @classmethod
    def default(cls):
        """The shipped default configuration, src/gui/mainwindow.cpp:67-77."""
        return cls(cell_length=14.938, nr_particles=1000, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5,
                   epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)

    def to_struct(self, force_kernel=KERNEL_AUTO, device=-1):
        g = self.get_param
        return LJParams(g("cell_length"), g("nr_particles", int), force_kernel, g("kT"), g("mass"), g("sigma"),
                        g("rcut"), g("epsilon"), g("tau"), g("shell"), g("stepsize"), device, 0)


def slab_layers(cells_per_axis, world, rank):
    """[first, end) cell layers along z that `rank` of `world` owns (ljgpu_slab_layers)."""
    lo, hi = C.c_uint32(), C.c_uint32()
    rc = load_library().ljgpu_slab_layers(cells_per_axis, world, rank, C.byref(lo), C.byref(hi))
    if rc != 0:
        raise LJGpuError(f"ljgpu_slab_layers failed ({rc})")
    return lo.value, hi.value


def slab_unique_id():
    """128-byte NCCL id, to be created on rank 0 and handed to every rank."""
    buf = C.create_string_buffer(128)
    lib = load_library()
    _check(lib, lib.ljgpu_slab_unique_id(buf))
    return buf.raw


def measure_fp64_peak(device=-1):
    """Measured DFMA throughput of a device in TFLOP/s (the force kernels' roofline denominator)."""
    lib = load_library()
    out = C.c_double()
    _check(lib, lib.ljgpu_measure_fp64_peak(int(device), C.byref(out)))
    return out.value


def host_rng_reset():
    load_library().ljgpu_host_rng_reset()


def host_initialize(params, force_kernel=KERNEL_AUTO):
    """initialize() of the reference (lennardjonessimulation.cpp:185-246) -> six float64 arrays."""
    lib = load_library()
    st = params.to_struct(force_kernel) if isinstance(params, LennardJonesParameters) else params
    n = st.nr_particles
    arrs = [np.empty(n, dtype=np.float64) for _ in range(6)]
    rc = lib.ljgpu_host_initialize(C.byref(st), *[_ptr(a) for a in arrs])
    if rc != 0:
        raise LJGpuError(f"ljgpu_host_initialize failed ({rc})")
    return arrs


def maxwell_boltzmann(kT, mass, n, nbins=100, vmax=10.0):
    out = np.empty(nbins, dtype=np.float64)
    rc = load_library().ljgpu_maxwell_boltzmann(kT, mass, n, nbins, vmax, _ptr(out))
    if rc != 0:
        raise LJGpuError(f"ljgpu_maxwell_boltzmann failed ({rc})")
    return out


class LennardJonesSimulation:
    """Same public surface as the reference class (lennardjonessimulation.h:113-155), GPU-backed."""

    def __init__(self, params, state=None, force_kernel=KERNEL_AUTO, device=-1, slab=None):
        """slab = (rank, world, nccl_unique_id_bytes) selects the multi-device z-slab path (collective)."""
        self._lib = load_library()
        self._h = C.c_void_p()
        self.params = params
        self._struct = params.to_struct(force_kernel, device)
        self.n = self._struct.nr_particles
        if state is None:
            state = host_initialize(self._struct)
        state = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
        if len(state) != 6 or any(a.shape != (self.n,) for a in state):
            raise ValueError("state must be six arrays of nr_particles float64")
        if slab is None:
            _check(self._lib, self._lib.ljgpu_create(C.byref(self._struct), *[_ptr(a) for a in state], C.byref(self._h)))
        else:
            rank, world, uid = slab
            idbuf = C.create_string_buffer(bytes(uid), 128)
            _check(self._lib, self._lib.ljgpu_create_slab(C.byref(self._struct), rank, world, idbuf,
                                                          *[_ptr(a) for a in state], C.byref(self._h)))

    def close(self):
        if self._h:
            self._lib.ljgpu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- stepping ----------------------------------------------------------------------------
    def integrate(self, step, dt):
        _check(self._lib, self._lib.ljgpu_step(self._h, step, dt))

    def integrate_n(self, first, n, dt):
        _check(self._lib, self._lib.ljgpu_step_n(self._h, first, n, dt))

    def integrate_publish(self, step, dt, px, py, pz):
        """integrate() + publish_state(): the step is complete on return, its positions are in flight into the
        caller's (pinned) buffers given as raw addresses; the previous call's delivery is complete."""
        _check(self._lib, self._lib.ljgpu_step_publish(self._h, step, dt, px, py, pz))

    def publish_wait(self):
        _check(self._lib, self._lib.ljgpu_publish_wait(self._h))

    def integrate_n_timed(self, first, n, dt):
        """integrate_n + device milliseconds (CUDA events on the integrator's stream)."""
        ms = C.c_double()
        _check(self._lib, self._lib.ljgpu_step_n_timed(self._h, first, n, dt, C.byref(ms)))
        return ms.value

    def set_state(self, state):
        state = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
        _check(self._lib, self._lib.ljgpu_set_state(self._h, *[_ptr(a) for a in state]))

    def set_param(self, name, value):
        """LennardJonesParameters::set_param on a live simulation (kT, tau, mass, sigma, epsilon)."""
        self.params.set_param(name, value)
        st = self.params.to_struct(self._struct.force_kernel, self._struct.device)
        _check(self._lib, self._lib.ljgpu_update_params(self._h, C.byref(st)))
        self._struct = st

    # -- queries -----------------------------------------------------------------------------
    def get_params(self):
        return self.params

    def get_dims(self):
        length = self._struct.cell_length
        return (length, length, length)

    def get_particle_count(self):
        return self.n

    def _get3(self, fn):
        a = [np.empty(self.n, dtype=np.float64) for _ in range(3)]
        _check(self._lib, fn(self._h, *[_ptr(x) for x in a]))
        return a

    def get_positions_soa(self):
        return self._get3(self._lib.ljgpu_get_positions)

    def get_positions_into(self, px, py, pz):
        """Positions into caller buffers given as raw addresses (e.g. pinned memory)."""
        _check(self._lib, self._lib.ljgpu_get_positions(self._h, px, py, pz))

    def get_list_stats(self):
        m, mx = C.c_double(), C.c_uint32()
        _check(self._lib, self._lib.ljgpu_get_list_stats(self._h, C.byref(m), C.byref(mx)))
        return m.value, mx.value

    def get_velocities_soa(self):
        return self._get3(self._lib.ljgpu_get_velocities)

    def get_forces_soa(self):
        return self._get3(self._lib.ljgpu_get_forces)

    def get_positions_copy(self):
        return np.stack(self.get_positions_soa(), axis=1)

    def get_velocities_copy(self):
        return np.stack(self.get_velocities_soa(), axis=1)

    def get_energies(self):
        e = [C.c_double() for _ in range(3)]
        _check(self._lib, self._lib.ljgpu_get_energies(self._h, *[C.byref(x) for x in e]))
        return tuple(x.value for x in e)

    def get_ekin(self):
        return self.get_energies()[0]

    def get_epot(self):
        return self.get_energies()[1]

    def get_etot(self):
        return self.get_energies()[2]

    def get_energy_history(self, max_entries=4096):
        steps = np.empty(max_entries, dtype=np.uint32)
        ek = np.empty(max_entries, dtype=np.float64)
        ep = np.empty(max_entries, dtype=np.float64)
        n = C.c_uint32()
        _check(self._lib, self._lib.ljgpu_get_energy_history(self._h, max_entries, _ptr(steps), _ptr(ek), _ptr(ep), C.byref(n)))
        return steps[:n.value], ek[:n.value], ep[:n.value]

    def get_speed_histogram(self, nbins=100, vmax=10.0):
        out = np.zeros(nbins, dtype=np.uint64)
        _check(self._lib, self._lib.ljgpu_get_speed_histogram(self._h, nbins, vmax, _ptr(out)))
        return out

    def get_cell_ids(self):
        out = np.empty(self.n, dtype=np.uint32)
        dims = (C.c_uint32 * 3)()
        _check(self._lib, self._lib.ljgpu_get_cell_ids(self._h, _ptr(out), dims))
        return out, tuple(dims)

    def get_neighbor_counts(self, cutoff, inclusive=True):
        out = np.empty(self.n, dtype=np.uint32)
        _check(self._lib, self._lib.ljgpu_get_neighbor_counts(self._h, cutoff, 1 if inclusive else 0, _ptr(out)))
        return out

    def set_sample_stride(self, stride):
        """Every `stride`-th step's energies enter the series the GraphWidget plots (default 1000); starts a new series."""
        _check(self._lib, self._lib.ljgpu_set_sample_stride(self._h, int(stride)))

    def get_graph_stats(self, reference_truncation=False):
        """GraphWidget's running statistics of the sampled energy series, computed on the device:
        {"n": samples, "ekin" / "epot" / "etot": {"mean", "stddev", "min", "max"}}."""
        st = GraphStats()
        _check(self._lib, self._lib.ljgpu_get_graph_stats(self._h, 1 if reference_truncation else 0, C.byref(st)))
        out = {"n": int(st.n_samples)}
        for k, name in enumerate(("ekin", "epot", "etot")):
            out[name] = {"mean": st.mean[k], "stddev": st.stddev[k], "min": st.min[k], "max": st.max[k]}
        return out

    def get_speed_statistics(self, nbins=100, vmax=10.0):
        """Speed histogram, Maxwell-Boltzmann overlay and their chi-square, all from the device."""
        counts = np.zeros(nbins, dtype=np.uint64)
        expected = np.zeros(nbins, dtype=np.float64)
        chi2, used = C.c_double(), C.c_uint32()
        _check(self._lib, self._lib.ljgpu_get_speed_statistics(self._h, nbins, vmax, _ptr(counts), _ptr(expected), C.byref(chi2), C.byref(used)))
        return counts, expected, chi2.value, int(used.value)

    def get_list_counts(self, inner=False):
        """Per-atom lengths of the list the force kernel walks: the list of the last re-binning (the reference's
        neighbor_list) or, inner=True, the pruned inner list."""
        out = np.empty(self.n, dtype=np.uint32)
        _check(self._lib, self._lib.ljgpu_get_list_counts(self._h, 1 if inner else 0, _ptr(out)))
        return out

    def write_to_movie_file(self, path, create=False):
        _check(self._lib, self._lib.ljgpu_write_movie_frame(self._h, os.fsencode(path), 1 if create else 0))

    # -- introspection -------------------------------------------------------------------------
    def set_profiling(self, on):
        """False/0 off, True/1 every step, k > 1 every k-th step (others replay the CUDA graph)."""
        _check(self._lib, self._lib.ljgpu_set_profiling(self._h, int(on)))

    def get_timers(self):
        t = LJTimers()
        _check(self._lib, self._lib.ljgpu_get_timers(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in LJTimers._fields_}

    def reset_timers(self):
        _check(self._lib, self._lib.ljgpu_reset_timers(self._h))

    def slab_info(self):
        r, w, lo, nl, own, halo = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
        _check(self._lib, self._lib.ljgpu_slab_info(self._h, C.byref(r), C.byref(w), C.byref(lo), C.byref(nl),
                                                    C.byref(own), C.byref(halo)))
        return {"rank": r.value, "world": w.value, "first_layer": lo.value, "n_layers": nl.value,
                "atoms_owned": own.value, "atoms_halo": halo.value}

    def get_layout(self):
        k, nc, cap, g = C.c_int32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        _check(self._lib, self._lib.ljgpu_get_layout(self._h, C.byref(k), C.byref(nc), C.byref(cap), C.byref(g)))
        return {"force_kernel": k.value, "cells_per_axis": nc.value, "list_capacity": cap.value, "inner_list_steps": g.value}
