"""polycap_b200/csrc/pcb_math.h -- the arithmetic the CUDA kernels compile -- built for the HOST and compared bit for bit
with glibc (the reference's libm) and with the oracle.  The transport is chaotic (one ulp in a normal or in sin/cos
of the grazing angle grows geometrically over tens of reflections), so the 1e-9 replay parity rests on these
identities.  The host build is test infrastructure; the library itself has no CPU path.

glibc dispatches acos/sin/cos/atan to FMA variants on FMA-capable x86-64 CPUs; pcb_math.h restates those variants,
so the identities below are asserted only where the CPU has FMA (the golden corpus was generated on such a CPU)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _cpu_has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
    except OSError:
        return False


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hm") / "libhostmath.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                           "-I" + os.path.join(ROOT, "polycap_b200", "csrc"), os.path.join(ROOT, "tests", "host_math.cpp"), "-o", out])
    L = C.CDLL(out)
    for f in (L.hm_trig_mismatches, L.hm_small_trig_mismatches):
        f.argtypes = [C.c_long, C.c_double, C.c_double, C.c_long * 3]
    L.hm_complex_mismatches.argtypes = [C.c_long, C.c_long * 3]
    L.hm_refl.restype = C.c_double
    L.hm_refl.argtypes = [C.POINTER(C.c_double)] * 5
    L.hm_segment.restype = C.c_int
    L.hm_segment.argtypes = [C.POINTER(C.c_double)] * 2
    return L


needs_fma = pytest.mark.skipif(not _cpu_has_fma(), reason="glibc uses its non-FMA libm variants on this CPU")


@needs_fma
def test_grazing_angle_trig_is_glibc_exact(hm):
    """theta = acos(n.d), sin(theta), cos(theta) of capil.c:487,497-498 for grazing cosines 1e-7 .. 0.124"""
    out = (C.c_long * 3)()
    for lo, hi in ((1e-7, 1e-4), (1e-4, 3e-3), (3e-3, 0.03), (0.03, 0.124)):
        hm.hm_trig_mismatches(300000, lo, hi, out)
        assert list(out) == [0, 0, 0], (lo, hi, list(out))


@needs_fma
def test_taper_angle_trig_is_glibc_exact(hm):
    """gamma = atan(dr/len), sin(gamma), cos(gamma) of capil.c:236-240 (wall taper, |gamma| << 1)"""
    out = (C.c_long * 3)()
    for lo, hi in ((1e-12, 1e-8), (1e-8, 1e-6), (1e-6, 1e-4), (1e-4, 1e-2), (1e-2, 0.0624)):
        hm.hm_small_trig_mismatches(300000, lo, hi, out)
        assert list(out) == [0, 0, 0], (lo, hi, list(out))


@needs_fma
def test_obtuse_incidence_predicate_matches_acos(hm):
    """capil.c:1270-1273 rejects a root when acos(n.d) > M_PI/2; the kernels evaluate that as one comparison"""
    hm.hm_obtuse_mismatches.restype = C.c_long
    hm.hm_obtuse_mismatches.argtypes = [C.c_long]
    assert hm.hm_obtuse_mismatches(300000) == 0


def test_complex_helpers_match_c99(hm):
    """csqrt / cabs / 1/z as libgcc + glibc evaluate them for the Fresnel term (capil.c:497-511)"""
    out = (C.c_long * 3)()
    hm.hm_complex_mismatches(500000, out)
    assert list(out) == [0, 0, 0], list(out)


@needs_fma
def test_reflection_chain_matches_oracle_bitwise(hm, harness, orc):
    """refl_geometry + fresnel_rtot (energy-independent part hoisted, per-energy constants precomputed) ==
    polycap_refl_polar as restated by the oracle, bit for bit, on random grazing reflections"""
    rng = np.random.default_rng(5)
    HC, N_AVOG, R0 = 1.23984193E-7, 6.022098e+23, 2.8179403227e-13
    nbad = 0
    for _ in range(4000):
        e = rng.uniform(1, 40)
        scatf, amu, rho = rng.uniform(0.45, 0.55), rng.uniform(1, 4000) / e, 2.23
        graze = 10 ** rng.uniform(-4.5, -1.2)
        az = rng.uniform(0, 2 * np.pi)
        sn = np.array([np.cos(az), np.sin(az), rng.uniform(-1e-4, 1e-4)])
        sn /= np.linalg.norm(sn)
        tang = np.array([0, 0, 1.0]) - sn[2] * sn
        tang /= np.linalg.norm(tang)
        d = np.cos(graze) * tang + np.sin(graze) * sn
        d /= np.linalg.norm(d)
        ev = np.cross(d, rng.normal(size=3))
        ev /= np.linalg.norm(ev)
        # per-energy constants exactly as pcb_econst.c derives them (C99 complex, reference order): via the oracle's numbers
        alfa = (HC / e) * (HC / e) * ((N_AVOG * R0 * rho) / (2 * np.pi)) * scatf
        beta = (HC) / (4. * np.pi) * (amu / e)
        n = complex(1.0 - alfa, beta)
        ph = harness.OrcPhoton()
        ph.exit_direction = harness.vec3(*d)
        ph.exit_electric_vector = harness.vec3(*ev)
        out_o = harness.Vec3()
        r_o = orc.L.orc_refl_polar(e, rho, scatf, amu, harness.vec3(*sn), C.byref(ph), C.byref(out_o))
        ninv = _c99_inv(n)
        i2 = _c99_mul(ninv, ninv)
        ec = (C.c_double * 4)(n.real, n.imag, i2.real, i2.imag)
        evc = (C.c_double * 3)(*ev)
        new_ev = (C.c_double * 3)()
        r_h = hm.hm_refl((C.c_double * 3)(*sn), (C.c_double * 3)(*d), evc, ec, new_ev)
        if r_h != r_o or tuple(new_ev) != out_o.t():
            nbad += 1
    assert nbad == 0, nbad


def _c99_mul(a, b):  # (a+ib)(c+id) without FMA, as gcc -O2 emits it for finite operands
    return complex(a.real * b.real - a.imag * b.imag, a.real * b.imag + a.imag * b.real)


def _c99_inv(z):  # libgcc __divdc3(1, 0, c, d), the branch for normally scaled operands
    c, d = z.real, z.imag
    if abs(c) < abs(d):
        ratio = c / d
        denom = (c * ratio) + d
        return complex(((1.0 * ratio) + 0.0) / denom, ((0.0 * ratio) - 1.0) / denom)
    ratio = d / c
    denom = (d * ratio) + c
    return complex(((0.0 * ratio) + 1.0) / denom, (0.0 - (1.0 * ratio)) / denom)


@needs_fma
def test_segment_matches_oracle_bitwise(hm, harness, orc):
    """ray/cone solve + surface normal (capil.c:105-254) on realistic grazing hits in an off-axis capillary"""
    g = load_golden("ellip")
    z, cap, ext = g["z"], g["cap"], g["ext"]
    rng = np.random.default_rng(1)
    hexz = ext / (2 * 0.86602540378443864676 * (258 + 1))
    hits = 0
    for _ in range(6000):
        i = int(rng.integers(0, 998))
        A, B = (2 * 37 + 11) * 0.86602540378443864676, 11 * 1.5
        c0 = (A * hexz[i], B * hexz[i], z[i])
        c1 = (A * hexz[i + 1], B * hexz[i + 1], z[i + 1])
        phi = rng.uniform(0, 2 * np.pi)
        rr = cap[i] * rng.uniform(0.5, 0.999)
        p0 = (c0[0] + rr * np.cos(phi), c0[1] + rr * np.sin(phi), z[i])
        th = rng.uniform(1e-4, 3e-3)
        psi = phi + rng.uniform(-1.0, 1.0)
        dz = z[i + 1] - z[i]
        d = (np.sin(th) * np.cos(psi) + (c1[0] - c0[0]) / dz, np.sin(th) * np.sin(psi) + (c1[1] - c0[1]) / dz, np.cos(th))
        p1 = (p0[0] + d[0] * dz / d[2], p0[1] + d[1] * dz / d[2], z[i + 1])
        last = (p0[0], p0[1], z[i] - 0.01)
        inp = (C.c_double * 20)(*c0, *c1, cap[i], cap[i + 1], *p0, *p1, *d, *last)
        out = (C.c_double * 6)()
        rh = hm.hm_segment(inp, out)
        ho, no = harness.Vec3(*last), harness.Vec3()
        ro = orc.L.orc_capil_segment(harness.vec3(*c0), harness.vec3(*c1), cap[i], cap[i + 1], harness.vec3(*p0), harness.vec3(*p1),
                                     harness.vec3(*d), C.byref(ho), C.byref(no))
        assert rh == ro
        if ro == 1:
            hits += 1
            assert tuple(out) == (*ho.t(), *no.t())
    assert hits > 300


def test_high_word_screens_never_contradict_the_float_comparisons(hm):
    """the high-word screens of the energy loop (pcb_refl_suspect_hi, pcb_weight_screen_hi) either agree with r < 0 || r > 1 and
    w >= 1e-4 or say "undecided", for every bit pattern class"""
    hm.hm_predicate_mismatches.restype = C.c_long
    hm.hm_predicate_mismatches.argtypes = [C.c_long]
    assert hm.hm_predicate_mismatches(2_000_000) == 0
