"""ctypes wrappers for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

* ``Oracle``   - oracle/liblj_oracle.so, the plain-C restatement (oracle/lj_oracle.c)
* ``RefSim``   - oracle/_ref/libljref.so, the reference's own sources compiled unmodified
                 (present when built in the authoring container; it travels to the GPU box as a
                 prebuilt file, /root/reference itself does not)
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liblj_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libljref.so")

DEFAULT = dict(cell_length=14.938, nr_particles=1000, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5,
               epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)


def box_for(n, rho):
    """L = cbrt(N / rho*), SURVEY.md 8(d)."""
    return float(np.cbrt(n / rho))


def config(n, rho=0.8, kT=1.0, **kw):
    d = dict(DEFAULT)
    d.update(cell_length=box_for(n, rho), nr_particles=n, kT=kT)
    d.update(kw)
    return d


class OParams(C.Structure):
    _fields_ = [("cell_length", C.c_double), ("nr_particles", C.c_int), ("kT", C.c_double), ("mass", C.c_double),
                ("sigma", C.c_double), ("rcut", C.c_double), ("epsilon", C.c_double), ("tau", C.c_double),
                ("shell", C.c_double), ("stepsize", C.c_double)]


def _build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ROOT, "oracle", "lj_oracle.c")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblj_oracle.so"],
                              stdout=subprocess.DEVNULL)


_O = None
_R = None


def oracle_lib():
    global _O
    if _O is None:
        _build_oracle()
        lib = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        lib.ljo_create.restype = vp
        lib.ljo_create.argtypes = [C.POINTER(OParams)]
        lib.ljo_create_from_state.restype = vp
        lib.ljo_create_from_state.argtypes = [C.POINTER(OParams)] + [vp] * 6
        lib.ljo_destroy.argtypes = [vp]
        lib.ljo_set_tau.argtypes = [vp, C.c_double]
        lib.ljo_integrate.argtypes = [vp, C.c_uint, C.c_double]
        lib.ljo_run.argtypes = [vp, C.c_uint, C.c_uint, C.c_double]
        lib.ljo_run.restype = C.c_double
        lib.ljo_n.argtypes = [vp]
        for nm in ("positions", "velocities", "forces", "dij"):
            getattr(lib, "ljo_get_" + nm).argtypes = [vp, vp, vp, vp]
        lib.ljo_get_energies.argtypes = [vp] + [C.POINTER(C.c_double)] * 3
        lib.ljo_list_size.argtypes = [vp]
        lib.ljo_list_size.restype = C.c_uint64
        lib.ljo_rebuild_count.argtypes = [vp]
        lib.ljo_rebuild_count.restype = C.c_uint64
        lib.ljo_neighbor_counts.argtypes = [vp, vp]
        lib.ljo_incutoff_counts.argtypes = [vp, vp]
        lib.ljo_cell_ids.argtypes = [vp, vp, vp]
        lib.ljo_force_abs_sums.argtypes = [vp, vp]
        lib.ljo_epot_long_double.argtypes = [vp]
        lib.ljo_epot_long_double.restype = C.c_double
        lib.ljo_cell_ids_of.argtypes = [C.POINTER(OParams), C.c_int, vp, vp, vp, vp, vp]
        lib.ljo_pair_counts_bruteforce.argtypes = [C.c_double, C.c_int, vp, vp, vp, C.c_double, vp]
        lib.ljo_speed_histogram.argtypes = [C.c_int, vp, vp, vp, C.c_uint, C.c_double, vp]
        lib.ljo_maxwell_boltzmann.argtypes = [C.c_double, C.c_double, C.c_int, C.c_uint, C.c_double, vp]
        lib.ljo_write_movie.argtypes = [vp, C.c_char_p, C.c_int]
        _O = lib
    return _O


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _R
    if _R is None:
        lib = C.CDLL(REF_SO)
        vp = C.c_void_p
        lib.ljref_create.restype = vp
        lib.ljref_create.argtypes = [C.c_double, C.c_int] + [C.c_double] * 8
        lib.ljref_destroy.argtypes = [vp]
        lib.ljref_set_param.argtypes = [vp, C.c_char_p, C.c_double]
        lib.ljref_integrate.argtypes = [vp, C.c_uint, C.c_double]
        lib.ljref_run.argtypes = [vp, C.c_uint, C.c_uint, C.c_double]
        lib.ljref_run.restype = C.c_double
        lib.ljref_n.argtypes = [vp]
        for nm in ("positions", "velocities", "forces", "dij"):
            getattr(lib, "ljref_get_" + nm).argtypes = [vp, vp, vp, vp]
        lib.ljref_get_energies.argtypes = [vp] + [C.POINTER(C.c_double)] * 3
        lib.ljref_set_state.argtypes = [vp] + [vp] * 6
        lib.ljref_list_size.argtypes = [vp]
        lib.ljref_list_size.restype = C.c_ulonglong
        lib.ljref_neighbor_counts.argtypes = [vp, vp]
        lib.ljref_incutoff_counts.argtypes = [vp, vp]
        lib.ljref_cell_ids.argtypes = [vp, vp, vp]
        lib.ljref_write_movie.argtypes = [vp, C.c_char_p, C.c_int]
        _R = lib
    return _R


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Base:
    prefix = None

    def _get3(self, name):
        a = [np.empty(self.n, dtype=np.float64) for _ in range(3)]
        getattr(self.lib, f"{self.prefix}_get_{name}")(self.h, *[_p(x) for x in a])
        return a

    def positions(self):
        return self._get3("positions")

    def velocities(self):
        return self._get3("velocities")

    def forces(self):
        return self._get3("forces")

    def dij(self):
        return self._get3("dij")

    def state(self):
        return self.positions() + self.velocities()

    def energies(self):
        e = [C.c_double() for _ in range(3)]
        getattr(self.lib, f"{self.prefix}_get_energies")(self.h, *[C.byref(x) for x in e])
        return tuple(x.value for x in e)

    def integrate(self, step, dt):
        getattr(self.lib, f"{self.prefix}_integrate")(self.h, step, dt)

    def run(self, first, nsteps, dt):
        return getattr(self.lib, f"{self.prefix}_run")(self.h, first, nsteps, dt)

    def list_size(self):
        return int(getattr(self.lib, f"{self.prefix}_list_size")(self.h))

    def neighbor_counts(self):
        out = np.empty(self.n, dtype=np.uint32)
        getattr(self.lib, f"{self.prefix}_neighbor_counts")(self.h, _p(out))
        return out

    def incutoff_counts(self):
        out = np.empty(self.n, dtype=np.uint32)
        getattr(self.lib, f"{self.prefix}_incutoff_counts")(self.h, _p(out))
        return out

    def cell_ids(self):
        out = np.empty(self.n, dtype=np.uint32)
        dims = np.zeros(3, dtype=np.uint32)
        getattr(self.lib, f"{self.prefix}_cell_ids")(self.h, _p(out), _p(dims))
        return out, tuple(int(d) for d in dims)

    def write_movie(self, path, create):
        getattr(self.lib, f"{self.prefix}_write_movie")(self.h, os.fsencode(path), 1 if create else 0)


class Oracle(_Base):
    """The plain-C restatement. ``state`` = six arrays to start from, else initialize()."""
    prefix = "ljo"

    def __init__(self, cfg, state=None, reset_rng=True):
        self.lib = oracle_lib()
        self.cfg = dict(cfg)
        self.p = OParams(**{k: cfg[k] for k, _ in OParams._fields_})
        self.n = cfg["nr_particles"]
        if state is None:
            if reset_rng:
                self.lib.ljo_rng_reset()
            self.h = C.c_void_p(self.lib.ljo_create(C.byref(self.p)))
        else:
            state = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
            self.h = C.c_void_p(self.lib.ljo_create_from_state(C.byref(self.p), *[_p(a) for a in state]))

    def set_tau(self, tau):
        self.lib.ljo_set_tau(self.h, tau)

    def force_abs_sums(self):
        out = np.empty(self.n, dtype=np.float64)
        self.lib.ljo_force_abs_sums(self.h, _p(out))
        return out

    def epot_long_double(self):
        return float(self.lib.ljo_epot_long_double(self.h))

    def rebuild_count(self):
        return int(self.lib.ljo_rebuild_count(self.h))

    def close(self):
        if self.h:
            self.lib.ljo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefSim(_Base):
    """The reference's own integrator. Its velocity generator is a process-wide static that cannot
    be reseeded: only the first RefSim of a process sees the default-seeded stream."""
    prefix = "ljref"

    def __init__(self, cfg, state=None):
        self.lib = ref_lib()
        self.cfg = dict(cfg)
        self.n = cfg["nr_particles"]
        self.h = C.c_void_p(self.lib.ljref_create(cfg["cell_length"], cfg["nr_particles"], cfg["kT"], cfg["mass"],
                                                  cfg["sigma"], cfg["rcut"], cfg["epsilon"], cfg["tau"],
                                                  cfg["shell"], cfg["stepsize"]))
        if state is not None:
            state = [np.ascontiguousarray(a, dtype=np.float64) for a in state]
            self.lib.ljref_set_state(self.h, *[_p(a) for a in state])

    def set_tau(self, tau):
        self.lib.ljref_set_param(self.h, b"tau", tau)

    def close(self):
        if self.h:
            self.lib.ljref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def oracle_cell_ids_of(cfg, x, y, z):
    lib = oracle_lib()
    p = OParams(**{k: cfg[k] for k, _ in OParams._fields_})
    n = len(x)
    out = np.empty(n, dtype=np.uint32)
    dims = np.zeros(3, dtype=np.uint32)
    x, y, z = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z)]
    lib.ljo_cell_ids_of(C.byref(p), n, _p(x), _p(y), _p(z), _p(out), _p(dims))
    return out, tuple(int(d) for d in dims)


def oracle_pair_counts(L, x, y, z, cutoff):
    lib = oracle_lib()
    n = len(x)
    out = np.empty(n, dtype=np.uint32)
    x, y, z = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z)]
    lib.ljo_pair_counts_bruteforce(L, n, _p(x), _p(y), _p(z), cutoff, _p(out))
    return out


def oracle_speed_histogram(vx, vy, vz, nbins=100, vmax=10.0):
    lib = oracle_lib()
    vx, vy, vz = [np.ascontiguousarray(a, dtype=np.float64) for a in (vx, vy, vz)]
    out = np.zeros(nbins, dtype=np.uint64)
    lib.ljo_speed_histogram(len(vx), _p(vx), _p(vy), _p(vz), nbins, vmax, _p(out))
    return out


def oracle_maxwell_boltzmann(kT, m, n, nbins=100, vmax=10.0):
    lib = oracle_lib()
    out = np.empty(nbins, dtype=np.float64)
    lib.ljo_maxwell_boltzmann(kT, m, n, nbins, vmax, _p(out))
    return out


def force_rel_error(f_test, f_ref, abs_sums=None):
    """Per-atom |df_i| / max(|f_i|, f_rms, 1e-2 * sum_j |f_ij|).

    A net force is a sum of ~50 pair terms that partly cancel; on the perfect start lattice it
    cancels completely and what is left is summation-order noise of a few ulps of the summed
    magnitudes (hard part 2), so the denominator is floored by the rms force
    and by a hundredth of the atom's summed pair-force magnitudes. In a liquid |f_i| is a fifth to a
    half of that sum and the floor is idle."""
    ft = np.stack(f_test, axis=1)
    fr = np.stack(f_ref, axis=1)
    mag = np.linalg.norm(fr, axis=1)
    rms = np.sqrt(np.mean(mag ** 2))
    scale = np.maximum(mag, max(rms, 1e-300))
    if abs_sums is not None:
        scale = np.maximum(scale, 1e-2 * abs_sums)
    return np.linalg.norm(ft - fr, axis=1) / scale
