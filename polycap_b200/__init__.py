"""polycap_b200 -- Python host side of libpolycap_b200.so.

Mirrors the reference's Cython module (/root/reference/python/polycap.pyx): the classes
``Profile, Rng, Description, Photon, Leak, Source, TransmissionEfficiencies`` keep its names,
argument order, return conventions and exception mapping (polycap.pyx:94-109), but bind the
C ABI of include/polycap.h through ctypes instead of Cython.  Everything that traces a photon
goes to the CUDA kernels inside the shared library; there is no Python or CPU implementation of
the transport here, and importing this module on a machine whose library is missing fails.

Additive (include/polycap_b200.h): ``launch_batch`` (batched replay), ``Source.run_ids``
(id-addressed Philox runs), ``Session`` (device-resident tallies for multi-GPU reductions).
"""
import ctypes as C
import os
import re
from collections import namedtuple

import numpy as np

__version__ = "1.2"

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POLYCAP_B200_LIB") or os.path.join(_HERE, "libpolycap_b200.so")  # env override: kernel-variant experiments

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "polycap_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C polycap_b200/csrc`; there is no CPU fallback" % LIB_PATH)

_lib = C.CDLL(LIB_PATH)

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)

N_COUNTERS = 16
COUNTER_NAMES = ("ids", "exit", "not_entered", "not_transmitted", "uncounted", "refl_exit", "refl_all", "seg_steps", "solves", "launched",
                 "q_batches", "q_lanes", "q_empty_polls", "q_lost_races", "q_hit_batches", "wall_steps")


class _Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class _Error(C.Structure):
    _fields_ = [("code", C.c_int), ("message", C.c_char_p)]


class _Leak(C.Structure):  # include/polycap-photon.h:52-59
    _fields_ = [("coords", _Vec3), ("direction", _Vec3), ("elecv", _Vec3), ("n_energies", C.c_size_t),
                ("weight", c_double_p), ("n_refl", C.c_int64)]


_ErrPP = C.POINTER(C.POINTER(_Error))
_LeakPP = C.POINTER(C.POINTER(_Leak))
_Vec3P = C.POINTER(_Vec3)

# polycap.pyx:94-102
_ERROR_MAP = {0: MemoryError, 1: ValueError, 2: IOError, 3: IOError, 4: TypeError, 5: NotImplementedError, 6: RuntimeError}


def _sig(name, restype, argtypes):
    fn = getattr(_lib, name)
    fn.restype = restype
    fn.argtypes = argtypes
    return fn


_v = C.c_void_p
_profile_new = _sig("polycap_profile_new", _v, [C.c_int] + [C.c_double] * 7 + [_ErrPP])
_profile_new_from_arrays = _sig("polycap_profile_new_from_arrays", _v, [C.c_int, c_double_p, c_double_p, c_double_p, _ErrPP])
_profile_new_from_file = _sig("polycap_profile_new_from_file", _v, [C.c_char_p] * 3 + [_ErrPP])
_profile_get = {k: _sig("polycap_profile_get_" + k, C.c_bool, [_v, C.POINTER(C.c_size_t), C.POINTER(c_double_p), _ErrPP]) for k in ("ext", "cap", "z")}
_profile_free = _sig("polycap_profile_free", None, [_v])
_rng_new = _sig("polycap_rng_new", _v, [])
_rng_new_with_seed = _sig("polycap_rng_new_with_seed", _v, [C.c_ulong])
_rng_free = _sig("polycap_rng_free", None, [_v])
_description_new = _sig("polycap_description_new", _v, [_v, C.c_double, C.c_int64, C.c_uint, C.POINTER(C.c_int), c_double_p, C.c_double, _ErrPP])
_description_free = _sig("polycap_description_free", None, [_v])
_photon_new = _sig("polycap_photon_new", _v, [_v, _Vec3, _Vec3, _Vec3, _ErrPP])
_photon_free = _sig("polycap_photon_free", None, [_v])
_photon_launch = _sig("polycap_photon_launch", C.c_int, [_v, C.c_size_t, c_double_p, C.POINTER(c_double_p), C.c_bool, _ErrPP])
_photon_vec = {k: _sig("polycap_photon_get_" + k, _Vec3, [_v]) for k in
               ("start_coords", "start_direction", "start_electric_vector", "exit_coords", "exit_direction", "exit_electric_vector")}
_photon_get_dtravel = _sig("polycap_photon_get_dtravel", C.c_double, [_v])
_photon_get_irefl = _sig("polycap_photon_get_irefl", C.c_int64, [_v])
_photon_get_extleak = _sig("polycap_photon_get_extleak_data", C.c_bool, [_v, C.POINTER(_LeakPP), c_int64_p, _ErrPP])
_photon_get_intleak = _sig("polycap_photon_get_intleak_data", C.c_bool, [_v, C.POINTER(_LeakPP), c_int64_p, _ErrPP])
_leak_free = _sig("polycap_leak_free", None, [C.POINTER(_Leak)])
_source_new = _sig("polycap_source_new", _v, [_v] + [C.c_double] * 8 + [C.c_size_t, c_double_p, _ErrPP])
_source_new_from_file = _sig("polycap_source_new_from_file", _v, [C.c_char_p, _ErrPP])
_source_free = _sig("polycap_source_free", None, [_v])
_source_get_photon = _sig("polycap_source_get_photon", _v, [_v, _v, _ErrPP])
_source_get_te = _sig("polycap_source_get_transmission_efficiencies", _v, [_v, C.c_int, C.c_int, C.c_bool, _v, _ErrPP])
_te_free = _sig("polycap_transmission_efficiencies_free", None, [_v])
_te_write_hdf5 = _sig("polycap_transmission_efficiencies_write_hdf5", C.c_bool, [_v, C.c_char_p, _ErrPP])
_te_get_data = _sig("polycap_transmission_efficiencies_get_data", C.c_bool, [_v, C.POINTER(C.c_size_t), C.POINTER(c_double_p), C.POINTER(c_double_p), _ErrPP])
_te_get_extleak = _sig("polycap_transmission_efficiencies_get_extleak_data", C.c_bool, [_v, C.POINTER(_LeakPP), c_int64_p, _ErrPP])
_te_get_intleak = _sig("polycap_transmission_efficiencies_get_intleak_data", C.c_bool, [_v, C.POINTER(_LeakPP), c_int64_p, _ErrPP])
_te_get_start = _sig("polycap_transmission_efficiencies_get_start_data", C.c_bool, [_v, c_int64_p, c_int64_p] + [C.POINTER(_Vec3P)] * 4 + [_ErrPP])
_te_get_exit = _sig("polycap_transmission_efficiencies_get_exit_data", C.c_bool,
                    [_v, c_int64_p] + [C.POINTER(_Vec3P)] * 3 + [C.POINTER(c_int64_p), C.POINTER(c_double_p), C.POINTER(C.c_size_t),
                                                               C.POINTER(C.POINTER(c_double_p)), _ErrPP])
_free = _sig("polycap_free", None, [_v])
_error_free = _sig("polycap_error_free", None, [C.POINTER(_Error)])
# additive
_device_count = _sig("polycap_b200_device_count", C.c_int, [])
_last_error = _sig("polycap_b200_last_error", C.c_char_p, [])
_provider = _sig("polycap_b200_optconst_provider", C.c_char_p, [])
_optical_constants = _sig("polycap_b200_optical_constants", C.c_bool, [_v, C.c_size_t, c_double_p, c_double_p, c_double_p, _ErrPP])
_launch_batch = _sig("polycap_b200_photon_launch_batch", C.c_bool,
                     [_v, C.c_size_t, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_bool,
                      C.POINTER(C.c_int), c_double_p, c_double_p, c_double_p, c_int64_p, c_double_p, c_double_p, c_int64_p, _ErrPP])
_launch_batch_leaks = _sig("polycap_b200_photon_launch_batch_leaks", C.c_bool,
                           [_v, C.c_size_t, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p,
                            C.POINTER(C.c_int), c_double_p, c_double_p, c_double_p, c_int64_p, c_double_p, c_double_p, c_int64_p,
                            C.POINTER(c_double_p), c_int64_p, _ErrPP])
_run_ids = _sig("polycap_b200_source_run_ids", C.c_bool,
                [_v, C.c_uint64, C.c_int64, C.c_int64, C.c_bool, c_double_p, c_int64_p, C.POINTER(C.c_int8), c_double_p, _ErrPP])
_sample_ids = _sig("polycap_b200_source_sample_ids", C.c_bool, [_v, C.c_uint64, C.c_int64, C.c_int64, c_double_p, _ErrPP])
_session_new = _sig("polycap_b200_session_new", _v, [_v, C.c_int, C.c_uint64, _ErrPP])
_session_free = _sig("polycap_b200_session_free", None, [_v])
_session_run = _sig("polycap_b200_session_run", C.c_bool, [_v, C.c_int64, C.c_int64, C.c_int64, _v, _v, _v, _ErrPP])
_session_tune = _sig("polycap_b200_session_tune", None, [_v, C.c_int, C.c_int, C.c_int, C.c_int])
_session_timing = _sig("polycap_b200_session_timing", None, [_v, C.c_int])
_session_trace_ms = _sig("polycap_b200_session_trace_ms", C.c_double, [_v, c_int64_p])
_session_kernels = _sig("polycap_b200_session_kernels", C.c_int64, [_v, c_int64_p])
_stream_new = _sig("polycap_b200_stream_new", _v, [_v, C.c_uint64, _ErrPP])
_stream_free = _sig("polycap_b200_stream_free", None, [_v])
_stream_next = _sig("polycap_b200_stream_next", C.c_int64, [_v, C.c_int64, C.c_int64, c_double_p, c_double_p, _ErrPP])
_stream_totals = _sig("polycap_b200_stream_totals", C.c_bool, [_v, c_int64_p, c_double_p, c_int64_p, _ErrPP])
_last_run_counters = _sig("polycap_b200_last_run_counters", C.c_bool, [c_int64_p])
_dfma_peak = _sig("polycap_b200_dfma_peak_tflops", C.c_double, [C.c_int, C.c_int])
_selftest_division = _sig("polycap_b200_selftest_division", C.c_int64, [C.c_int, C.c_int64, C.c_uint64, C.c_int])


def _raise(err):
    """polycap_set_exception of polycap.pyx:104-109"""
    if not err:
        return
    code, msg = err.contents.code, (err.contents.message or b"").decode("utf-8", "replace")
    _error_free(err)
    raise _ERROR_MAP.get(code, RuntimeError)(msg)


def _err():
    return C.POINTER(_Error)()


def _dp(a):
    return a.ctypes.data_as(c_double_p)


VectorTuple = namedtuple("VectorTuple", "x y z")


def _np2vec(a, name):
    a = np.asarray(a, dtype=np.double)
    if a.ndim != 1 or a.size != 3:
        raise ValueError("%s must have exactly three elements" % name)
    return _Vec3(a[0], a[1], a[2])


def _vec2tuple(v):
    return VectorTuple(v.x, v.y, v.z)


def device_count():
    return _device_count()


def optconst_provider():
    return _provider().decode()


def dfma_peak_tflops(device=-1, iters=200000):
    return _dfma_peak(device, iters)


def selftest_division(n=1 << 24, seed=1, span=40, device=-1):
    """bit mismatches between the kernels' shared-divisor division and IEEE a / b over n operand pairs (0 expected)"""
    return _selftest_division(device, n, seed, span)


_SYMBOLS = ("H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr Rb Sr Y Zr Nb Mo Tc Ru Rh Pd "
            "Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac "
            "Th Pa U Np Pu Am Cm Bk Cf Es Fm Md No Lr Rf Db Sg Bh Hs Mt").split()
_ATOMIC_WEIGHT = {5: 10.811, 8: 15.9994, 11: 22.98977, 13: 26.981538, 14: 28.0855}


def ensure_int(x):
    """polycap.pyx:281-289 (SymbolToAtomicNumber without xraylib)"""
    if isinstance(x, str):
        if x not in _SYMBOLS:
            raise ValueError("Invalid chemical symbol %s" % x)
        return _SYMBOLS.index(x) + 1
    return int(x)


def _parse_formula(s):
    """Mass fractions of a flat chemical formula such as 'SiO2' (polycap.pyx:330-342 uses xraylib's CompoundParser)."""
    toks = re.findall(r"([A-Z][a-z]?)(\d*\.?\d*)", s)
    if not toks or "".join(a + b for a, b in toks) != s:
        raise ValueError("composition is not a valid chemical formula")
    masses = {}
    for sym, cnt in toks:
        z = ensure_int(sym)
        if z not in _ATOMIC_WEIGHT:
            raise NotImplementedError("formula parsing without xraylib covers B, O, Na, Al, Si only; pass a {Z: weight} dict")
        masses[z] = masses.get(z, 0.0) + _ATOMIC_WEIGHT[z] * (float(cnt) if cnt else 1.0)
    tot = sum(masses.values())
    return {z: m / tot for z, m in masses.items()}


class Profile:
    """polycap.pyx:111-259"""
    CONICAL, PARABOLOIDAL, ELLIPSOIDAL = 0, 1, 2

    def __init__(self, type, length, rad_ext_upstream, rad_ext_downstream, rad_int_upstream, rad_int_downstream,
                 focal_dist_upstream, focal_dist_downstream, ignore=False):
        self._profile = None
        if ignore:
            return
        e = _err()
        self._profile = _profile_new(int(type), length, rad_ext_upstream, rad_ext_downstream, rad_int_upstream, rad_int_downstream,
                                     focal_dist_upstream, focal_dist_downstream, C.byref(e))
        _raise(e)

    def __del__(self):
        if getattr(self, "_profile", None):
            _profile_free(self._profile)
            self._profile = None

    @classmethod
    def new_from_arrays(cls, ext, cap, z):
        ext, cap, z = (np.ascontiguousarray(np.atleast_1d(a), dtype=np.double) for a in (ext, cap, z))
        for a, n in ((ext, "ext"), (cap, "cap"), (z, "z")):
            if a.ndim != 1:
                raise ValueError("%s must be a 1D array" % n)
        if not (ext.size == cap.size == z.size):
            raise ValueError("ext, cap and z must have identical lengths")
        rv = cls(0, 0, 0, 0, 0, 0, 0, 0, ignore=True)
        e = _err()
        rv._profile = _profile_new_from_arrays(ext.size - 1, _dp(ext), _dp(cap), _dp(z), C.byref(e))
        _raise(e)
        return rv

    @classmethod
    def new_from_file(cls, single_cap_profile_file, central_axis_file, external_shape_file):
        rv = cls(0, 0, 0, 0, 0, 0, 0, 0, ignore=True)
        e = _err()
        rv._profile = _profile_new_from_file(single_cap_profile_file.encode(), central_axis_file.encode(), external_shape_file.encode(), C.byref(e))
        _raise(e)
        return rv

    def _get(self, which):
        n, p, e = C.c_size_t(0), c_double_p(), _err()
        _profile_get[which](self._profile, C.byref(n), C.byref(p), C.byref(e))
        _raise(e)
        out = np.ctypeslib.as_array(p, (n.value + 1,)).copy()
        _free(p)
        return out

    @property
    def get_ext(self):
        return self._get("ext")

    @property
    def get_cap(self):
        return self._get("cap")

    @property
    def get_z(self):
        return self._get("z")


class Rng:
    """polycap.pyx:262-278"""

    def __init__(self, seed=None):
        self._rng = _rng_new() if seed is None else _rng_new_with_seed(int(seed))

    def __del__(self):
        if getattr(self, "_rng", None):
            _rng_free(self._rng)
            self._rng = None


class Description:
    """polycap.pyx:292-362"""

    def __init__(self, profile, sig_rough, n_cap, composition, density):
        self._description = None
        if profile is None or not isinstance(profile, Profile):
            raise TypeError("profile must be a Profile")
        if composition is None:
            raise TypeError("composition must be a dictionary or a string")
        if isinstance(composition, dict):
            if len(composition) == 0:
                raise ValueError("composition cannot be empty")
            iz = [ensure_int(k) for k in composition.keys()]
            wi = [float(v) for v in composition.values()]
        elif isinstance(composition, str):
            comp = _parse_formula(composition)
            iz, wi = list(comp.keys()), list(comp.values())
        else:
            raise TypeError("composition must be a dictionary or a string")
        izs = (C.c_int * len(iz))(*iz)
        wis = (C.c_double * len(wi))(*wi)
        e = _err()
        self._description = _description_new(profile._profile, sig_rough, int(n_cap), len(iz), izs, wis, density, C.byref(e))
        _raise(e)

    def __del__(self):
        if getattr(self, "_description", None):
            _description_free(self._description)
            self._description = None

    def optical_constants(self, energies):
        """amu [1/cm] and scatf per energy, as the library derives them (replaces polycap_photon_scatf, photon.c:22-94)"""
        en = np.ascontiguousarray(np.atleast_1d(energies), dtype=np.double)
        amu, scatf = np.empty_like(en), np.empty_like(en)
        e = _err()
        _optical_constants(self._description, en.size, _dp(en), _dp(amu), _dp(scatf), C.byref(e))
        _raise(e)
        return amu, scatf


class Leak:
    """polycap.pyx:750-822; owns a copy of the C record"""

    def __init__(self, rec):
        self.coords = _vec2tuple(rec.coords)
        self.direction = _vec2tuple(rec.direction)
        self.elecv = _vec2tuple(rec.elecv)
        self.n_refl = rec.n_refl
        w = np.ctypeslib.as_array(rec.weight, (rec.n_energies,)).copy()
        w.flags.writeable = False
        self.weight = w


def _take_leaks(getter, handle):
    arr, n, e = _LeakPP(), C.c_int64(0), _err()
    ok = getter(handle, C.byref(arr), C.byref(n), C.byref(e))
    if not ok and n.value == 0:
        if e:
            _error_free(e)
        return []
    _raise(e)
    out = [Leak(arr[i].contents) for i in range(n.value)]
    for i in range(n.value):
        _leak_free(arr[i])
    _free(arr)
    return out


class Photon:
    """polycap.pyx:826-1038"""

    def __init__(self, description, start_coords, start_direction, start_electric_vector, ignore=False):
        self._photon = None
        self._description_py = description
        if ignore:
            return
        if description is None:
            raise ValueError("description cannot be None")
        c = _np2vec(start_coords, "start_coords")
        d = _np2vec(start_direction, "start_direction")
        v = _np2vec(start_electric_vector, "start_electric_vector")
        e = _err()
        self._photon = _photon_new(description._description, c, d, v, C.byref(e))
        _raise(e)

    def __del__(self):
        if getattr(self, "_photon", None):
            _photon_free(self._photon)
            self._photon = None

    def launch(self, energies, leak_calc=False):
        if energies is None:
            raise TypeError("energies cannot be None")
        en = np.ascontiguousarray(np.atleast_1d(np.asarray(energies, dtype=np.double)))
        if en.ndim != 1:
            raise ValueError("energies must be a 1D array")
        w, e = c_double_p(), _err()
        rv = _photon_launch(self._photon, en.size, _dp(en), C.byref(w), bool(leak_calc), C.byref(e))
        try:
            _raise(e)
            self.last_return = rv
            if rv == 2:
                return None
            return np.ctypeslib.as_array(w, (en.size,)).copy() if w else None
        finally:
            if w:
                _free(w)

    @property
    def extleak_data(self):
        return _take_leaks(_photon_get_extleak, self._photon) or None

    @property
    def intleak_data(self):
        return _take_leaks(_photon_get_intleak, self._photon) or None

    def _vec(self, k):
        return _vec2tuple(_photon_vec[k](self._photon))

    exit_coords = get_exit_coords = property(lambda self: self._vec("exit_coords"))
    start_coords = property(lambda self: self._vec("start_coords"))
    exit_direction = get_exit_direction = property(lambda self: self._vec("exit_direction"))
    start_direction = property(lambda self: self._vec("start_direction"))
    exit_electric_vector = get_exit_electric_vector = property(lambda self: self._vec("exit_electric_vector"))
    start_electric_vector = property(lambda self: self._vec("start_electric_vector"))
    i_refl = property(lambda self: _photon_get_irefl(self._photon))
    d_travel = property(lambda self: _photon_get_dtravel(self._photon))


class TransmissionEfficiencies:
    """polycap.pyx:367-747; produced by Source.get_transmission_efficiencies"""

    def __init__(self):
        self._trans_eff = None
        self._exit = None
        self._start = None

    @staticmethod
    def _create(handle, source_py):
        if not handle:
            return None
        rv = TransmissionEfficiencies()
        rv._trans_eff = handle
        rv._source_py = source_py  # the C object borrows the source (source.c:682)
        n, en, ef, e = C.c_size_t(0), c_double_p(), c_double_p(), _err()
        _te_get_data(handle, C.byref(n), C.byref(en), C.byref(ef), C.byref(e))
        _raise(e)
        rv._energies_np = np.ctypeslib.as_array(en, (n.value,)).copy()
        rv._efficiencies_np = np.ctypeslib.as_array(ef, (n.value,)).copy()
        rv._energies_np.flags.writeable = False
        rv._efficiencies_np.flags.writeable = False
        _free(en)
        _free(ef)
        return rv

    def __del__(self):
        if getattr(self, "_trans_eff", None):
            _te_free(self._trans_eff)
            self._trans_eff = None

    def write_hdf5(self, filename):
        if filename is None:
            raise TypeError("filename cannot be None")
        e = _err()
        _te_write_hdf5(self._trans_eff, filename.encode(), C.byref(e))
        _raise(e)

    @property
    def data(self):
        return (self._energies_np, self._efficiencies_np)

    @property
    def extleak_data(self):
        return iter(_take_leaks(_te_get_extleak, self._trans_eff))

    @property
    def intleak_data(self):
        return iter(_take_leaks(_te_get_intleak, self._trans_eff))

    def _load_exit(self):
        if self._exit is not None:
            return self._exit
        n, nE, e = C.c_int64(0), C.c_size_t(0), _err()
        co, di, ev = _Vec3P(), _Vec3P(), _Vec3P()
        nr, dt, w = c_int64_p(), c_double_p(), C.POINTER(c_double_p)()
        _te_get_exit(self._trans_eff, C.byref(n), C.byref(co), C.byref(di), C.byref(ev), C.byref(nr), C.byref(dt), C.byref(nE), C.byref(w), C.byref(e))
        _raise(e)
        N = n.value

        def vecs(p):
            a = np.ctypeslib.as_array(C.cast(p, c_double_p), (N, 3)).copy() if N else np.zeros((0, 3))
            _free(p)
            return a
        out = dict(coords=vecs(co), direction=vecs(di), elecv=vecs(ev))
        out["n_refl"] = np.ctypeslib.as_array(nr, (N,)).copy() if N else np.zeros(0, dtype=np.int64)
        out["d_travel"] = np.ctypeslib.as_array(dt, (N,)).copy() if N else np.zeros(0)
        wts = np.empty((N, nE.value))
        for i in range(N):
            wts[i] = np.ctypeslib.as_array(w[i], (nE.value,))
            _free(w[i])
        wts.flags.writeable = False
        out["weights"] = wts
        _free(nr)
        _free(dt)
        _free(w)
        self._exit = out
        return out

    def _load_start(self):
        if self._start is not None:
            return self._start
        ns, ne, e = C.c_int64(0), C.c_int64(0), _err()
        co, di, ev, sc = _Vec3P(), _Vec3P(), _Vec3P(), _Vec3P()
        _te_get_start(self._trans_eff, C.byref(ns), C.byref(ne), C.byref(co), C.byref(di), C.byref(ev), C.byref(sc), C.byref(e))
        _raise(e)
        N = ne.value

        def vecs(p):
            a = np.ctypeslib.as_array(C.cast(p, c_double_p), (N, 3)).copy() if N else np.zeros((0, 3))
            _free(p)
            return a
        self._start = dict(coords=vecs(co), direction=vecs(di), elecv=vecs(ev), src=vecs(sc), n_start=ns.value, n_exit=N)
        return self._start

    exit_coords = property(lambda self: (VectorTuple(*r) for r in self._load_exit()["coords"]))
    exit_direction = property(lambda self: (VectorTuple(*r) for r in self._load_exit()["direction"]))
    exit_elecv = property(lambda self: (VectorTuple(*r) for r in self._load_exit()["elecv"]))
    n_refl = property(lambda self: (int(v) for v in self._load_exit()["n_refl"]))
    d_travel = property(lambda self: (float(v) for v in self._load_exit()["d_travel"]))
    exit_weights = property(lambda self: (row for row in self._load_exit()["weights"]))
    start_coords = property(lambda self: (VectorTuple(*r) for r in self._load_start()["coords"]))
    start_direction = property(lambda self: (VectorTuple(*r) for r in self._load_start()["direction"]))
    start_elecv = property(lambda self: (VectorTuple(*r) for r in self._load_start()["elecv"]))
    src_start_coords = property(lambda self: (VectorTuple(*r) for r in self._load_start()["src"]))

    # N4 of SURVEY 8(f): whole-array access without one Python object per photon
    exit_arrays = property(lambda self: self._load_exit())
    start_arrays = property(lambda self: self._load_start())

    @property
    def records(self):
        """Zero-copy numpy views of the result's own structure-of-arrays storage (struct _polycap_images,
        src/polycap-private.h:156-181): no re-packing into vector arrays, no per-photon objects.  Read-only; the views keep this
        object alive.  Keys are the struct's field names; two-/three-component fields are tuples of arrays."""
        class _Images(C.Structure):
            _fields_ = [("i_start", C.c_int64), ("i_exit", C.c_int64), ("src_start_coords", c_double_p * 2), ("pc_start_coords", c_double_p * 2),
                        ("pc_start_dir", c_double_p * 2), ("pc_start_elecv", c_double_p * 2), ("pc_exit_coords", c_double_p * 3),
                        ("pc_exit_dir", c_double_p * 2), ("pc_exit_elecv", c_double_p * 2), ("pc_exit_nrefl", c_int64_p),
                        ("pc_exit_dtravel", c_double_p), ("exit_coord_weights", c_double_p)]

        class _Te(C.Structure):
            _fields_ = [("n_energies", C.c_size_t), ("energies", c_double_p), ("efficiencies", c_double_p), ("images", C.POINTER(_Images))]
        te = C.cast(self._trans_eff, C.POINTER(_Te)).contents
        im = te.images.contents
        n, nE = int(im.i_exit), int(te.n_energies)

        def view(ptr, shape):
            a = np.ctypeslib.as_array(ptr, shape) if n else np.zeros(shape, dtype=np.int64 if ptr._type_ is C.c_int64 else np.float64)
            a.flags.writeable = False
            return a
        out = {"i_start": int(im.i_start), "i_exit": n, "_owner": self}
        for name, _ in _Images._fields_[2:9]:
            out[name] = tuple(view(q, (n,)) for q in getattr(im, name))
        out["pc_exit_nrefl"] = view(im.pc_exit_nrefl, (n,))
        out["pc_exit_dtravel"] = view(im.pc_exit_dtravel, (n,))
        out["exit_coord_weights"] = view(im.exit_coord_weights, (n, nE))
        return out


class Source:
    """polycap.pyx:1042-1152"""

    def __init__(self, description, d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol, energies, _handle=None):
        self._source = None
        if _handle is not None:
            self._source = _handle
            return
        if description is None or energies is None:
            raise TypeError("description and energies cannot be None")
        en = np.ascontiguousarray(np.atleast_1d(np.asarray(energies, dtype=np.double)))
        if en.ndim != 1:
            raise ValueError("energies must be a 1D array")
        self.energies = en.copy()
        e = _err()
        self._source = _source_new(description._description, d_source, src_x, src_y, src_sigx, src_sigy, src_shiftx, src_shifty, hor_pol,
                                   en.size, _dp(en), C.byref(e))
        _raise(e)

    @classmethod
    def new_from_file(cls, filename):
        """polycap_source_new_from_file (source.c:228-445); profile file names resolve against the CWD"""
        e = _err()
        h = _source_new_from_file(filename.encode(), C.byref(e))
        _raise(e)
        return cls(None, 0, 0, 0, 0, 0, 0, 0, 0, None, _handle=h)

    def __del__(self):
        if getattr(self, "_source", None):
            _source_free(self._source)
            self._source = None

    def get_photon(self, rng):
        if rng is None:
            raise TypeError("rng cannot be None")
        e = _err()
        h = _source_get_photon(self._source, rng._rng, C.byref(e))
        _raise(e)
        rv = Photon(None, None, None, None, ignore=True)
        rv._photon = h
        rv._source_py = self  # the photon borrows the source's description
        return rv

    def get_transmission_efficiencies(self, max_threads, n_photons, leak_calc=False):
        e = _err()
        h = _source_get_te(self._source, int(max_threads), int(n_photons), bool(leak_calc), None, C.byref(e))
        _raise(e)
        return TransmissionEfficiencies._create(h, self)

    # ---- additive -------------------------------------------------------------------------
    def n_energies(self):
        class _S(C.Structure):  # head of struct _polycap_source (private.h:108-122)
            _fields_ = [("description", _v), ("rng", _v)] + [(k, C.c_double) for k in "abcdefgh"] + [("n_energies", C.c_size_t)]
        return C.cast(self._source, C.POINTER(_S)).contents.n_energies

    def sample_ids(self, seed, id0, n_ids):
        """start states the device draws for ids [id0, id0+n_ids): [n][11] = coords, direction, elecv, source x, y"""
        out = np.zeros((n_ids, 11))
        e = _err()
        _sample_ids(self._source, seed, id0, n_ids, _dp(out), C.byref(e))
        _raise(e)
        return out

    def run_ids(self, seed, id0, n_ids, records=False, status=True):
        """Photon ids [id0, id0+n_ids) with Philox seed `seed`: (sum_weights, counters dict, status, rec)"""
        nE = self.n_energies()
        sw = np.zeros(nE)
        cnt = np.zeros(N_COUNTERS, dtype=np.int64)
        st = np.zeros(n_ids, dtype=np.int8) if (status or records) else None
        rec = np.zeros((n_ids, 17 + nE)) if records else None
        e = _err()
        _run_ids(self._source, seed, id0, n_ids, False, _dp(sw), cnt.ctypes.data_as(c_int64_p),
                 st.ctypes.data_as(C.POINTER(C.c_int8)) if st is not None else None, _dp(rec) if rec is not None else None, C.byref(e))
        _raise(e)
        return sw, dict(zip(COUNTER_NAMES, cnt.tolist())), st, rec


def launch_batch(description, energies, start_coords, start_direction, start_electric_vector, amu=None, scatf=None, leak_calc=False):
    """Batched polycap_photon_launch: returns dict(ret, exit_coords, exit_direction, exit_electric_vector, i_refl, d_travel, weights, counters);
    with leak_calc=True additionally leaks = per-photon dict(ext=[...], int=[...]) of dict(coords, direction, elecv, n_refl, weight)."""
    en = np.ascontiguousarray(np.atleast_1d(energies), dtype=np.double)
    c = np.ascontiguousarray(start_coords, dtype=np.double).reshape(-1, 3)
    d = np.ascontiguousarray(start_direction, dtype=np.double).reshape(-1, 3)
    v = np.ascontiguousarray(start_electric_vector, dtype=np.double).reshape(-1, 3)
    n = c.shape[0]
    if d.shape[0] != n or v.shape[0] != n:
        raise ValueError("start arrays must have identical lengths")
    a = np.ascontiguousarray(amu, dtype=np.double) if amu is not None else None
    s = np.ascontiguousarray(scatf, dtype=np.double) if scatf is not None else None
    ret = np.zeros(n, dtype=np.int32)
    ec, ed, ev = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    ir = np.zeros(n, dtype=np.int64)
    dt = np.zeros(n)
    w = np.zeros((n, en.size))
    cnt = np.zeros(N_COUNTERS, dtype=np.int64)
    e = _err()
    common = [description._description, en.size, _dp(en), _dp(a) if a is not None else None, _dp(s) if s is not None else None, n, _dp(c), _dp(d), _dp(v)]
    outs = [ret.ctypes.data_as(C.POINTER(C.c_int)), _dp(ec), _dp(ed), _dp(ev), ir.ctypes.data_as(c_int64_p), _dp(dt), _dp(w)]
    out = dict(ret=ret, exit_coords=ec, exit_direction=ed, exit_electric_vector=ev, i_refl=ir, d_travel=dt, weights=w)
    if not leak_calc:
        _launch_batch(*common, False, *outs, cnt.ctypes.data_as(c_int64_p), C.byref(e))
        _raise(e)
        out["counters"] = dict(zip(COUNTER_NAMES, cnt.tolist()))
        return out
    lcount = np.zeros((n, 2), dtype=np.int64)
    recs, nrec = c_double_p(), C.c_int64(0)
    _launch_batch_leaks(*common, *outs, lcount.ctypes.data_as(c_int64_p), C.byref(recs), C.byref(nrec), C.byref(e))
    _raise(e)
    stride = 11 + en.size
    R = np.ctypeslib.as_array(recs, (nrec.value, stride)).copy() if nrec.value else np.zeros((0, stride))
    if recs:
        _free(recs)
    leaks, k = [], 0
    for i in range(n):
        m = int(lcount[i].sum())
        per = {"ext": [], "int": []}
        for r in R[k:k + m]:
            per["ext" if r[0] == 0 else "int"].append(dict(coords=r[1:4].copy(), direction=r[4:7].copy(), elecv=r[7:10].copy(), n_refl=int(r[10]),
                                                            weight=r[11:].copy()))
        leaks.append(per)
        k += m
    out["leaks"] = leaks
    out["leak_count"] = lcount
    return out


def last_run_counters():
    """work counters of this process's last Source.get_transmission_efficiencies call (dict by COUNTER_NAMES), None when there was none"""
    cnt = np.zeros(N_COUNTERS, dtype=np.int64)
    if not _last_run_counters(cnt.ctypes.data_as(c_int64_p)):
        return None
    return dict(zip(COUNTER_NAMES, cnt.tolist()))


class Stream:
    """Exit records of a run chunk by chunk (include/polycap_b200.h: polycap_b200_stream_*): the id-ordered photon sequence of `seed`
    continues from call to call, the tallies accumulate on the device(s)."""

    def __init__(self, source, seed=20000):
        self._source_py = source
        self._nE = source.n_energies()
        e = _err()
        self._stream = _stream_new(source._source, int(seed), C.byref(e))
        _raise(e)

    def __del__(self):
        if getattr(self, "_stream", None):
            _stream_free(self._stream)
            self._stream = None

    def next(self, max_records, max_ids=0, records=True):
        """-> (fields [17][n], weights [n][nE]) of the next n <= max_records exit photons (records=False: just n, tally only)"""
        f = np.zeros((17, max_records)) if records else None
        w = np.zeros((max_records, self._nE)) if records else None
        e = _err()
        n = _stream_next(self._stream, int(max_records), int(max_ids), _dp(f) if records else None, _dp(w) if records else None, C.byref(e))
        _raise(e)
        if not records:
            return n
        return f[:, :n], w[:n]

    def totals(self):
        """-> (photon ids consumed, sum of exit weights per energy, counters dict)"""
        ids = C.c_int64(0)
        sw = np.zeros(self._nE)
        cnt = np.zeros(N_COUNTERS, dtype=np.int64)
        e = _err()
        _stream_totals(self._stream, C.byref(ids), _dp(sw), cnt.ctypes.data_as(c_int64_p), C.byref(e))
        _raise(e)
        return ids.value, sw, dict(zip(COUNTER_NAMES, cnt.tolist()))


class Session:
    """Device-resident tallies (include/polycap_b200.h): d_sum_weights / d_counters are device pointers (ints),
    e.g. torch tensors' data_ptr(), so a caller can all-reduce them with NCCL without a host round trip."""

    def __init__(self, source, device=-1, seed=20000):
        self._source_py = source
        e = _err()
        self._session = _session_new(source._source, int(device), int(seed), C.byref(e))
        _raise(e)

    def __del__(self):
        if getattr(self, "_session", None):
            _session_free(self._session)
            self._session = None

    def run(self, id0, n_ids, d_sum_weights, d_counters, wave_size=1 << 20, stream=0):
        e = _err()
        _session_run(self._session, int(id0), int(n_ids), int(wave_size), C.c_void_p(stream), C.c_void_p(d_sum_weights), C.c_void_p(d_counters), C.byref(e))
        _raise(e)

    def timing(self, on=True):
        _session_timing(self._session, 1 if on else 0)

    def trace_ms(self):
        """(summed device ms, launches) of the trace kernels timed since the last call; synchronises on them"""
        n = C.c_int64(0)
        ms = _session_trace_ms(self._session, C.byref(n))
        return ms, n.value

    def kernels(self):
        """(all kernels, trace kernels) launched by this session so far"""
        n = C.c_int64(0)
        tot = _session_kernels(self._session, C.byref(n))
        return tot, n.value

    def tune(self, thr_cand=-1, thr_refl=-1, thr_new=-1, ctas_per_sm=-1):
        _session_tune(self._session, thr_cand, thr_refl, thr_new, ctas_per_sm)
