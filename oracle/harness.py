"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Ref``    -- oracle/_ref/libpolycap_ref.so: the reference's own sources (compiled in place
               from /root/reference by oracle/Makefile, never copied).  Struct layouts follow
               /root/reference/src/polycap-private.h:88-145.
* ``Oracle`` -- oracle/liboracle.so: the C restatement in oracle/polycap_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libpolycap_ref.so")
ORC_SO = os.path.join(HERE, "liboracle.so")

c_double_p = C.POINTER(C.c_double)


class Vec3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]

    def t(self):
        return (self.x, self.y, self.z)


def vec3(x, y, z):
    return Vec3(float(x), float(y), float(z))


# --------------------------------------------------------------------------- reference build
class RefProfile(C.Structure):
    _fields_ = [("nmax", C.c_int), ("z", c_double_p), ("cap", c_double_p), ("ext", c_double_p)]


class RefDescription(C.Structure):
    _fields_ = [("sig_rough", C.c_double), ("n_cap", C.c_int64), ("open_area", C.c_double),
                ("nelem", C.c_uint), ("iz", C.POINTER(C.c_int)), ("wi", c_double_p),
                ("density", C.c_double), ("profile", C.POINTER(RefProfile))]


class RefSource(C.Structure):
    _fields_ = [("description", C.POINTER(RefDescription)), ("rng", C.c_void_p),
                ("d_source", C.c_double), ("src_x", C.c_double), ("src_y", C.c_double),
                ("src_sigx", C.c_double), ("src_sigy", C.c_double), ("src_shiftx", C.c_double),
                ("src_shifty", C.c_double), ("hor_pol", C.c_double), ("n_energies", C.c_size_t),
                ("energies", c_double_p)]


class RefLeak(C.Structure):
    _fields_ = [("coords", Vec3), ("direction", Vec3), ("elecv", Vec3), ("n_energies", C.c_size_t),
                ("weight", c_double_p), ("n_refl", C.c_int64)]


class RefPhoton(C.Structure):
    _fields_ = [("description", C.POINTER(RefDescription)),
                ("extleak", C.POINTER(C.POINTER(RefLeak))), ("intleak", C.POINTER(C.POINTER(RefLeak))),
                ("n_extleak", C.c_int64), ("n_intleak", C.c_int64),
                ("start_coords", Vec3), ("start_direction", Vec3), ("start_electric_vector", Vec3),
                ("exit_coords", Vec3), ("exit_direction", Vec3), ("exit_electric_vector", Vec3),
                ("src_start_coords", Vec3), ("n_energies", C.c_size_t), ("energies", c_double_p),
                ("weight", c_double_p), ("amu", c_double_p), ("scatf", c_double_p),
                ("i_refl", C.c_int64), ("d_travel", C.c_double)]


class RefImages(C.Structure):
    _fields_ = [("i_start", C.c_int64), ("i_exit", C.c_int64),
                ("src_start_coords", c_double_p * 2), ("pc_start_coords", c_double_p * 2),
                ("pc_start_dir", c_double_p * 2), ("pc_start_elecv", c_double_p * 2),
                ("pc_exit_coords", c_double_p * 3), ("pc_exit_dir", c_double_p * 2),
                ("pc_exit_elecv", c_double_p * 2), ("pc_exit_nrefl", C.POINTER(C.c_int64)),
                ("pc_exit_dtravel", c_double_p), ("exit_coord_weights", c_double_p),
                ("i_extleak", C.c_int64), ("extleak_coords", c_double_p * 3), ("extleak_dir", c_double_p * 2),
                ("extleak_coord_weights", c_double_p), ("extleak_n_refl", C.POINTER(C.c_int64)),
                ("i_intleak", C.c_int64), ("intleak_coords", c_double_p * 3), ("intleak_dir", c_double_p * 2),
                ("intleak_elecv", c_double_p * 2), ("intleak_coord_weights", c_double_p),
                ("intleak_n_refl", C.POINTER(C.c_int64))]


class RefEfficiencies(C.Structure):
    _fields_ = [("n_energies", C.c_size_t), ("energies", c_double_p), ("efficiencies", c_double_p),
                ("images", C.POINTER(RefImages)), ("source", C.POINTER(RefSource))]


def ref_available():
    return os.path.exists(REF_SO)


class Ref:
    """Thin access to the reference's public and TEST_BUILD-private entry points."""

    def __init__(self):
        os.environ.setdefault("POLYCAP_B200_QUIET", "1")
        L = self.L = C.CDLL(REF_SO)
        P = C.POINTER
        L.polycap_profile_new.restype = P(RefProfile)
        L.polycap_profile_new.argtypes = [C.c_int] + [C.c_double] * 7 + [C.c_void_p]
        L.polycap_profile_new_from_file.restype = P(RefProfile)
        L.polycap_profile_new_from_file.argtypes = [C.c_char_p] * 3 + [C.c_void_p]
        L.polycap_profile_new_from_arrays.restype = P(RefProfile)
        L.polycap_profile_new_from_arrays.argtypes = [C.c_int, c_double_p, c_double_p, c_double_p, C.c_void_p]
        L.polycap_profile_free.argtypes = [P(RefProfile)]
        L.polycap_description_new.restype = P(RefDescription)
        L.polycap_description_new.argtypes = [P(RefProfile), C.c_double, C.c_int64, C.c_uint, P(C.c_int), c_double_p,
                                              C.c_double, C.c_void_p]
        L.polycap_description_free.argtypes = [P(RefDescription)]
        L.polycap_source_new.restype = P(RefSource)
        L.polycap_source_new.argtypes = [P(RefDescription)] + [C.c_double] * 8 + [C.c_size_t, c_double_p, C.c_void_p]
        L.polycap_source_new_from_file.restype = P(RefSource)
        L.polycap_source_new_from_file.argtypes = [C.c_char_p, C.c_void_p]
        L.polycap_source_free.argtypes = [P(RefSource)]
        L.polycap_source_get_photon.restype = P(RefPhoton)
        L.polycap_source_get_photon.argtypes = [P(RefSource), C.c_void_p, C.c_void_p]
        L.polycap_source_get_transmission_efficiencies.restype = P(RefEfficiencies)
        L.polycap_source_get_transmission_efficiencies.argtypes = [P(RefSource), C.c_int, C.c_int, C.c_bool, C.c_void_p, C.c_void_p]
        L.polycap_transmission_efficiencies_free.argtypes = [P(RefEfficiencies)]
        L.polycap_rng_new_with_seed.restype = C.c_void_p
        L.polycap_rng_new_with_seed.argtypes = [C.c_ulong]
        L.polycap_rng_free.argtypes = [C.c_void_p]
        L.polycap_photon_new.restype = P(RefPhoton)
        L.polycap_photon_new.argtypes = [P(RefDescription), Vec3, Vec3, Vec3, C.c_void_p]
        L.polycap_photon_free.argtypes = [P(RefPhoton)]
        L.polycap_photon_launch.restype = C.c_int
        L.polycap_photon_launch.argtypes = [P(RefPhoton), C.c_size_t, c_double_p, P(c_double_p), C.c_bool, C.c_void_p]
        L.polycap_photon_scatf.argtypes = [P(RefPhoton), C.c_void_p]
        L.polycap_free.argtypes = [C.c_void_p]
        L.polycap_capil_segment.restype = C.c_int
        L.polycap_capil_segment.argtypes = [Vec3, Vec3, C.c_double, C.c_double, Vec3, Vec3, Vec3, P(Vec3), P(Vec3), C.c_void_p]
        L.polycap_refl_polar.restype = C.c_double
        L.polycap_refl_polar.argtypes = [C.c_double] * 4 + [Vec3, P(RefPhoton), P(Vec3), C.c_void_p]
        L.polycap_capil_reflect.restype = C.c_int
        L.polycap_capil_reflect.argtypes = [P(RefPhoton), Vec3, C.c_bool, C.c_void_p]
        L.polycap_capil_trace.restype = C.c_int
        L.polycap_capil_trace.argtypes = [P(C.c_int), P(RefPhoton), P(RefDescription), c_double_p, c_double_p, C.c_bool, C.c_void_p]
        L.polycap_capil_trace_wall.restype = C.c_int
        L.polycap_capil_trace_wall.argtypes = [P(RefPhoton), c_double_p, P(C.c_int), P(C.c_int), C.c_void_p]
        L.polycap_photon_within_pc_boundary.restype = C.c_int
        L.polycap_photon_within_pc_boundary.argtypes = [C.c_double, Vec3, C.c_void_p]

    # -- convenience ---------------------------------------------------------
    def description(self, profile, sig_rough, n_cap, iz, wi, density):
        izs = (C.c_int * len(iz))(*iz)
        wis = (C.c_double * len(wi))(*wi)
        return self.L.polycap_description_new(profile, sig_rough, n_cap, len(iz), izs, wis, density, None)

    def source(self, desc, d_source, src_x, src_y, sigx, sigy, shiftx, shifty, hor_pol, energies):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        return self.L.polycap_source_new(desc, d_source, src_x, src_y, sigx, sigy, shiftx, shifty, hor_pol,
                                         e.size, e.ctypes.data_as(c_double_p), None)

    @staticmethod
    def profile_arrays(profile):
        p = profile.contents
        n = p.nmax + 1
        return (p.nmax, np.ctypeslib.as_array(p.z, (n,)).copy(), np.ctypeslib.as_array(p.cap, (n,)).copy(),
                np.ctypeslib.as_array(p.ext, (n,)).copy())

    def optical_constants(self, desc, energies):
        """amu[], scatf[] exactly as polycap_photon_scatf computes them in this build."""
        e = np.ascontiguousarray(energies, dtype=np.float64)
        ph = self.L.polycap_photon_new(desc, vec3(0, 0, 0), vec3(0, 0, 1), vec3(1, 0, 0), None)
        ph.contents.n_energies = e.size
        # energies must be malloc'd because polycap_photon_free frees them
        libc = C.CDLL(None)
        libc.malloc.restype = C.c_void_p
        mem = libc.malloc(C.c_size_t(8 * e.size))
        C.memmove(mem, e.ctypes.data, 8 * e.size)
        ph.contents.energies = C.cast(mem, c_double_p)
        self.L.polycap_photon_scatf(ph, None)
        amu = np.ctypeslib.as_array(ph.contents.amu, (e.size,)).copy()
        scatf = np.ctypeslib.as_array(ph.contents.scatf, (e.size,)).copy()
        self.L.polycap_photon_free(ph)
        return amu, scatf

    def launch(self, photon, energies, leak=False):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        wp = c_double_p()
        ret = self.L.polycap_photon_launch(photon, e.size, e.ctypes.data_as(c_double_p), C.byref(wp), leak, None)
        w = None
        if wp:
            w = np.ctypeslib.as_array(wp, (e.size,)).copy()
            self.L.polycap_free(wp)
        return ret, w

    @staticmethod
    def leaks(photon):
        p = photon.contents
        out = {"ext": [], "int": []}
        for key, arr, n in (("ext", p.extleak, p.n_extleak), ("int", p.intleak, p.n_intleak)):
            for i in range(n):
                lk = arr[i].contents
                out[key].append(dict(coords=lk.coords.t(), direction=lk.direction.t(), elecv=lk.elecv.t(), n_refl=lk.n_refl,
                                     weight=np.ctypeslib.as_array(lk.weight, (lk.n_energies,)).copy()))
        return out


# --------------------------------------------------------------------------- restatement
class OrcOptic(C.Structure):
    _fields_ = [("nmax", C.c_int), ("z", c_double_p), ("cap", c_double_p), ("ext", c_double_p),
                ("sig_rough", C.c_double), ("n_cap", C.c_int64), ("density", C.c_double),
                ("n_energies", C.c_size_t), ("energies", c_double_p), ("amu", c_double_p), ("scatf", c_double_p)]


class OrcLeak(C.Structure):
    _fields_ = [("coords", Vec3), ("direction", Vec3), ("elecv", Vec3), ("n_refl", C.c_int64), ("weight", c_double_p)]


class OrcPhoton(C.Structure):
    _fields_ = [("start_coords", Vec3), ("start_direction", Vec3), ("start_electric_vector", Vec3),
                ("exit_coords", Vec3), ("exit_direction", Vec3), ("exit_electric_vector", Vec3),
                ("src_start_coords", Vec3), ("i_refl", C.c_int64), ("d_travel", C.c_double),
                ("weight", c_double_p), ("extleak", C.POINTER(OrcLeak)), ("intleak", C.POINTER(OrcLeak)),
                ("n_extleak", C.c_int64), ("n_intleak", C.c_int64)]


class OrcSource(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("d_source", "src_x", "src_y", "src_sigx", "src_sigy", "src_shiftx", "src_shifty", "hor_pol")]


class OrcCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_ids", "n_exit", "n_not_entered", "n_not_transmitted", "n_uncounted",
                                         "sum_irefl_exit", "sum_irefl_all", "n_segments")]


class OrcMT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int)]


class OrcPhilox(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("id", C.c_uint64), ("draw", C.c_uint32)]


def oracle_available():
    return os.path.exists(ORC_SO)


class Optic:
    """Owns the numpy arrays behind an OrcOptic."""

    def __init__(self, nmax, z, cap, ext, sig_rough, n_cap, density, energies, amu, scatf):
        self.arr = [np.ascontiguousarray(a, dtype=np.float64) for a in (z, cap, ext, energies, amu, scatf)]
        z, cap, ext, energies, amu, scatf = self.arr
        self.nE = energies.size
        self.c = OrcOptic(int(nmax), z.ctypes.data_as(c_double_p), cap.ctypes.data_as(c_double_p), ext.ctypes.data_as(c_double_p),
                          float(sig_rough), int(n_cap), float(density), energies.size, energies.ctypes.data_as(c_double_p),
                          amu.ctypes.data_as(c_double_p), scatf.ctypes.data_as(c_double_p))


class Oracle:
    def __init__(self):
        L = self.L = C.CDLL(ORC_SO)
        P = C.POINTER
        L.orc_within_pc_boundary.restype = C.c_int
        L.orc_within_pc_boundary.argtypes = [C.c_double, Vec3]
        L.orc_capil_segment.restype = C.c_int
        L.orc_capil_segment.argtypes = [Vec3, Vec3, C.c_double, C.c_double, Vec3, Vec3, Vec3, P(Vec3), P(Vec3)]
        L.orc_refl_polar.restype = C.c_double
        L.orc_refl_polar.argtypes = [C.c_double] * 4 + [Vec3, P(OrcPhoton), P(Vec3)]
        L.orc_capil_reflect.restype = C.c_int
        L.orc_capil_reflect.argtypes = [P(OrcOptic), P(OrcPhoton), Vec3, C.c_int]
        L.orc_capil_trace_wall.restype = C.c_int
        L.orc_capil_trace_wall.argtypes = [P(OrcOptic), P(OrcPhoton), c_double_p, P(C.c_int), P(C.c_int)]
        L.orc_capil_trace.restype = C.c_int
        L.orc_capil_trace.argtypes = [P(OrcOptic), P(C.c_int), P(OrcPhoton), c_double_p, c_double_p, C.c_int]
        L.orc_photon_launch.restype = C.c_int
        L.orc_photon_launch.argtypes = [P(OrcOptic), P(OrcPhoton), C.c_int]
        L.orc_source_get_photon.argtypes = [P(OrcSource), P(OrcOptic), C.c_void_p, C.c_void_p, P(OrcPhoton)]
        L.orc_photon_clear_leaks.argtypes = [P(OrcPhoton)]
        L.orc_mt_seed.argtypes = [P(OrcMT), C.c_ulong]
        L.orc_mt_uniform.restype = C.c_double
        L.orc_mt_uniform.argtypes = [C.c_void_p]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_void_p]
        L.orc_philox4x32_10.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
        L.orc_run_starts.argtypes = [P(OrcOptic), C.c_int64, c_double_p, C.c_int, C.c_int, c_double_p, P(OrcCounters), P(C.c_int8), c_double_p]
        L.orc_run_ids.argtypes = [P(OrcSource), P(OrcOptic), C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                  c_double_p, P(OrcCounters), P(C.c_int8), c_double_p]

    def run_starts(self, optic, starts, leak=False, threads=1, records=False):
        starts = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 11)
        n, nE = starts.shape[0], optic.nE
        sw = np.zeros(nE)
        cnt = OrcCounters()
        status = np.zeros(n, dtype=np.int8)
        rec = np.zeros((n, 17 + nE)) if records else None
        self.L.orc_run_starts(C.byref(optic.c), n, starts.ctypes.data_as(c_double_p), int(leak), threads, sw.ctypes.data_as(c_double_p),
                              C.byref(cnt), status.ctypes.data_as(C.POINTER(C.c_int8)), rec.ctypes.data_as(c_double_p) if records else None)
        return sw, {n_: getattr(cnt, n_) for n_, _ in OrcCounters._fields_}, status, rec

    def photon(self, optic, start, direction, elecv):
        ph = OrcPhoton()
        ph.start_coords = ph.exit_coords = vec3(*start)
        ph.start_direction = ph.exit_direction = vec3(*direction)
        ph.start_electric_vector = ph.exit_electric_vector = vec3(*elecv)
        w = np.ones(optic.nE, dtype=np.float64)
        ph.weight = w.ctypes.data_as(c_double_p)
        ph._w = w  # keep alive
        return ph

    def launch(self, optic, ph, leak=False):
        ret = self.L.orc_photon_launch(C.byref(optic.c), C.byref(ph), int(leak))
        return ret, ph._w.copy()

    @staticmethod
    def leaks(ph, nE):
        out = {"ext": [], "int": []}
        for key, arr, n in (("ext", ph.extleak, ph.n_extleak), ("int", ph.intleak, ph.n_intleak)):
            for i in range(n):
                lk = arr[i]
                out[key].append(dict(coords=lk.coords.t(), direction=lk.direction.t(), elecv=lk.elecv.t(), n_refl=lk.n_refl,
                                     weight=np.ctypeslib.as_array(lk.weight, (nE,)).copy()))
        return out

    def run_ids(self, source, optic, seed, id0, n_ids, leak=False, threads=1, records=False):
        nE = optic.nE
        sw = np.zeros(nE)
        cnt = OrcCounters()
        status = np.zeros(n_ids, dtype=np.int8)
        rec = np.zeros((n_ids, 17 + nE)) if records else None
        self.L.orc_run_ids(C.byref(source), C.byref(optic.c), seed, id0, n_ids, int(leak), threads,
                           sw.ctypes.data_as(c_double_p), C.byref(cnt), status.ctypes.data_as(C.POINTER(C.c_int8)),
                           rec.ctypes.data_as(c_double_p) if records else None)
        counters = {n: getattr(cnt, n) for n, _ in OrcCounters._fields_}
        return sw, counters, status, rec
