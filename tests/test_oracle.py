"""The oracle (oracle/polycap_oracle.c) pinned against (i) the golden vectors of the reference's own
tests, (ii) the replay corpus generated from the reference itself
(tests/golden/make_golden.py), (iii) the compiled reference (oracle/_ref) where it is present."""
import ctypes as C

import numpy as np
import pytest

from conftest import CORPUS, load_golden, optic_of, source_of, rel_err

HC, N_AVOG, R0 = 1.23984193E-7, 6.022098e+23, 2.8179403227e-13


# ------------------------------------------------------------------ golden vectors of the reference's tests
def test_segment_golden(harness, orc):  # tests/capil.c:29-65
    v = harness.vec3
    hit, sn = harness.Vec3(0, 0, 0), harness.Vec3()
    r = orc.L.orc_capil_segment(v(0, 0, 0), v(0, 0, 0.1), 0.005, 0.005, v(0, 0, 0), v(0.005, -0.005, 0.1), v(0.005, -0.005, 0.1),
                                C.byref(hit), C.byref(sn))
    assert r == 1
    assert np.allclose(hit.t(), (0.003536, -0.003536, 0.070711), atol=1e-5)
    assert np.allclose(sn.t(), (0.707107, -0.707107, 0.0), atol=1e-5)


@pytest.mark.parametrize("theta,expect", [(2e-3, 0.984522), (3.1e-3, 0.496310), (np.pi / 2, 0.0)])
@pytest.mark.parametrize("ev", [(1, 0, 0), (0, 1, 0), (np.sqrt(0.5), np.sqrt(0.5), 0)])
def test_refl_polar_golden(harness, orc, theta, expect, ev):  # tests/capil.c:107-231, literal scatf/amu
    ph = harness.OrcPhoton()
    ph.exit_direction = harness.vec3(0, np.sin(theta), np.cos(theta))  # plane of reflection is the yz plane
    ph.exit_electric_vector = harness.vec3(*ev)
    sn = harness.vec3(0, 1, 0)
    out = harness.Vec3()
    r = orc.L.orc_refl_polar(10.0, 2.23, 0.503696, 42.544677, sn, C.byref(ph), C.byref(out))
    assert abs(r - expect) < 1e-5


def test_within_boundary_golden(harness, orc):  # tests/photon.c:166-190
    v = harness.vec3
    assert orc.L.orc_within_pc_boundary(0.05, v(0.025, 0.025, 0)) == 1
    assert orc.L.orc_within_pc_boundary(0.05, v(0.075, 0.025, 0)) == 0
    assert orc.L.orc_within_pc_boundary(0.05, v(0.05, 0.05, 0)) == 0
    assert orc.L.orc_within_pc_boundary(-1.0, v(0, 0, 0)) == -1


def _ellip_optic(harness, energies=(10.0,), amu=(42.544635,), scatf=(0.503696,)):
    g = load_golden("ellip")
    return harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], 0.0, 200000, 2.23, energies, amu, scatf)


def test_trace_golden(harness, orc):  # tests/capil.c:405-493
    o = _ellip_optic(harness)
    n = o.c.nmax + 1
    zeros = np.zeros(n)
    ph = orc.photon(o, (0, 0, 0), (3e-5, 3e-5, 0.999), (0.5, 0.5, 0))
    ix = C.c_int(0)
    r = orc.L.orc_capil_trace(C.byref(o.c), C.byref(ix), C.byref(ph), zeros.ctypes.data_as(harness.c_double_p), zeros.ctypes.data_as(harness.c_double_p), 0)
    assert r == 1 and ix.value == 552 and ph.i_refl == 1
    assert abs(ph._w[0] - 0.999585) < 1e-5
    assert np.allclose(ph.exit_coords.t(), (1.49e-4, 1.49e-4, 4.975778), atol=1e-5)
    assert np.allclose(ph.exit_direction.t(), (-6.9e-5, -6.9e-5, 1.0), atol=1e-5)
    ph = orc.photon(o, (0, 0, 0), (0.005, -0.005, 0.1), (0.5, 0.5, 0))
    ix = C.c_int(0)
    r = orc.L.orc_capil_trace(C.byref(o.c), C.byref(ix), C.byref(ph), zeros.ctypes.data_as(harness.c_double_p), zeros.ctypes.data_as(harness.c_double_p), 0)
    assert r == 0
    assert np.allclose(ph.exit_coords.t(), (2.47e-4, -2.47e-4, 0.004948), atol=1e-5)
    ph = orc.photon(o, (0, 0, 0), (0, 0, 1), (0.5, 0.5, 0))
    ix = C.c_int(0)
    assert orc.L.orc_capil_trace(C.byref(o.c), C.byref(ix), C.byref(ph), zeros.ctypes.data_as(harness.c_double_p), zeros.ctypes.data_as(harness.c_double_p), 0) == -2


def test_launch_codes_golden(harness, orc):  # tests/photon.c:238-352
    o = _ellip_optic(harness)
    cases = [((0.21, 0, 0), (0, 0, 1), -2), ((0, 0, 0), (0.005, -0.005, 0.1), 0), ((0, 0, 0), (0, 0, 1.0), 1),
             ((0.15104418, 0.08700043, 0), (0, 0, 1.0), 2)]
    for c, d, code in cases:
        ph = orc.photon(o, c, d, (0.5, 0.5, 0))
        r, _ = orc.launch(o, ph)
        assert r == code, (c, d, r)
        if code == 1:
            assert ph.i_refl == 0


# ------------------------------------------------------------------ random streams
def test_philox_known_answers(orc):  # Random123 kat_vectors, philox4x32 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, expect in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        orc.L.orc_philox4x32_10(c, k, o)
        assert tuple(o) == expect


def test_mt19937_known_answers(harness, orc):
    mt = harness.OrcMT()
    orc.L.orc_mt_seed(C.byref(mt), 5489)
    assert int(orc.L.orc_mt_uniform(C.byref(mt)) * 4294967296.0) == 3499211612  # MT19937 reference output, seed 5489
    orc.L.orc_mt_seed(C.byref(mt), 0)  # GSL: seed 0 means 4357
    a = orc.L.orc_mt_uniform(C.byref(mt))
    orc.L.orc_mt_seed(C.byref(mt), 4357)
    assert a == orc.L.orc_mt_uniform(C.byref(mt))


# ------------------------------------------------------------------ the corpus generated from the reference itself
@pytest.mark.parametrize("name", CORPUS)
def test_oracle_replays_reference_corpus(harness, orc, name):
    g = load_golden(name)
    o = optic_of(harness, g)
    n = g["start"].shape[0]
    worst = 0.0
    for i in range(n):
        s = g["start"][i]
        ph = orc.photon(o, s[0:3], s[3:6], s[6:9])
        r, w = orc.launch(o, ph)
        assert r == g["ret"][i], (name, i)
        assert ph.i_refl == g["i_refl"][i], (name, i)
        got = np.array([*ph.exit_coords.t(), *ph.exit_direction.t(), *ph.exit_electric_vector.t(), ph.d_travel])
        want = np.array([*g["exit_state"][i], g["d_travel"][i]])
        worst = max(worst, rel_err(got, want))
        if r in (0, 1):
            worst = max(worst, rel_err(w, g["weights"][i]))
    assert worst <= 1e-12, worst  # same compiler and libm as the fixture: normally bit-identical


def test_oracle_source_sampling_matches_corpus(harness, orc):
    """orc_source_get_photon over the same MT19937 stream reproduces the reference's start states"""
    for name in ("ellip", "parab", "monocap", "xos1"):
        g = load_golden(name)
        o = optic_of(harness, g)
        s = source_of(harness, g)
        mt = harness.OrcMT()
        orc.L.orc_mt_seed(C.byref(mt), int(g["seed"]))
        uni = C.cast(orc.L.orc_mt_uniform, C.c_void_p)
        for i in range(200):
            ph = harness.OrcPhoton()
            orc.L.orc_source_get_photon(C.byref(s), C.byref(o.c), uni, C.byref(mt), C.byref(ph))
            got = np.array([*ph.start_coords.t(), *ph.start_direction.t(), *ph.start_electric_vector.t()])
            assert rel_err(got, g["start"][i]) <= 1e-13, (name, i)
            assert rel_err([ph.src_start_coords.x, ph.src_start_coords.y], g["src_xy"][i]) <= 1e-13


# ------------------------------------------------------------------ the id-addressed driver
def test_run_ids_consistency(harness, orc):
    g = load_golden("ellip")
    o, s = optic_of(harness, g), source_of(harness, g)
    sw1, c1, st1, rec1 = orc.run_ids(s, o, 20000, 0, 3000, threads=1, records=True)
    sw4, c4, st4, _ = orc.run_ids(s, o, 20000, 0, 3000, threads=4)
    assert c1 == c4 and np.array_equal(st1, st4)
    assert np.allclose(sw1, sw4, rtol=1e-12)
    # additivity over id ranges (what multi-GPU sharding relies on)
    swa, ca, sta, _ = orc.run_ids(s, o, 20000, 0, 1000)
    swb, cb, stb, _ = orc.run_ids(s, o, 20000, 1000, 2000)
    assert np.array_equal(np.concatenate([sta, stb]), st1)
    assert np.allclose(swa + swb, sw1, rtol=1e-12)
    for k in ("n_exit", "n_not_entered", "n_not_transmitted", "n_uncounted", "sum_irefl_exit", "sum_irefl_all", "n_segments"):
        assert ca[k] + cb[k] == c1[k]
    # records: exit z at the end of the optic, weights in [0,1], sum of record weights == tally
    ex = st1 == 1
    assert ex.sum() == c1["n_exit"] > 0
    assert np.all(rec1[ex, 10] == g["z"][-1])
    assert np.allclose(rec1[ex, 17:].sum(axis=0), sw1, rtol=1e-12)
    assert np.all((rec1[ex, 17:] >= 0) & (rec1[ex, 17:] <= 1))


def test_run_ids_efficiencies_match_reference_goldens(harness, orc):
    """tests/source.c:204-222: 0.424/0.349/0.135/0.050/0.022/0.011/0.006 at 1..30 keV (tolerances from that file)"""
    g = load_golden("ellip")
    o, s = optic_of(harness, g), source_of(harness, g)
    sw, c, _, _ = orc.run_ids(s, o, 1234, 0, 30000, threads=8)
    started = c["n_exit"] + c["n_not_entered"] + c["n_not_transmitted"]
    eff = sw / started
    want = [0.424, 0.349, 0.135, 0.050, 0.022, 0.011, 0.006]
    tol = [0.01, 0.01, 0.0075, 0.005, 0.005, 0.005, 0.005]
    for e, w, t in zip(eff, want, tol):
        assert abs(e - w) < 1.5 * t, (eff, want)
    open_area = (c["n_exit"] + c["n_not_transmitted"]) / started
    assert abs(open_area - 0.696365) < 0.02  # tests/source.c:116


# ------------------------------------------------------------------ live cross-check with the compiled reference
def test_oracle_matches_compiled_reference_live(harness, orc, ref):
    """fresh photons (not in the corpus) through oracle/_ref and the restatement, leak_calc off and on"""
    g = load_golden("ellip")
    L = ref.L
    prof = L.polycap_profile_new_from_arrays(int(g["nmax"]), g["ext"].ctypes.data_as(harness.c_double_p), g["cap"].ctypes.data_as(harness.c_double_p),
                                             g["z"].ctypes.data_as(harness.c_double_p), None)
    desc = ref.description(prof, 0.0, 200000, [8, 14], [53.0, 47.0], 2.23)
    energies = np.array([5.0, 20.0, 40.0])
    amu, scatf = ref.optical_constants(desc, energies)
    src = ref.source(desc, 2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5, energies)
    o = harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], 0.0, 200000, 2.23, energies, amu, scatf)
    rng = L.polycap_rng_new_with_seed(777)
    n_leak_seen = 0
    for i in range(120):
        leak = i >= 100  # leak mode is ~1000x slower
        ph = L.polycap_source_get_photon(src, rng, None)
        p = ph.contents
        ophot = orc.photon(o, p.start_coords.t(), p.start_direction.t(), p.start_electric_vector.t())
        r_ref, w_ref = ref.launch(ph, energies, leak=leak)
        r_orc, w_orc = orc.launch(o, ophot, leak=leak)
        assert r_ref == r_orc
        p = ph.contents
        assert p.i_refl == ophot.i_refl
        if r_ref in (0, 1):
            assert rel_err(w_orc, w_ref) <= 1e-12
            assert rel_err(ophot.exit_coords.t(), p.exit_coords.t()) <= 1e-12
        if leak:
            lr, lo = ref.leaks(ph), orc.leaks(ophot, 3)
            for kind in ("ext", "int"):
                assert len(lr[kind]) == len(lo[kind])
                for a, b in zip(lr[kind], lo[kind]):
                    n_leak_seen += 1
                    assert rel_err(b["coords"], a["coords"]) <= 1e-12 and rel_err(b["weight"], a["weight"]) <= 1e-12
                    assert a["n_refl"] == b["n_refl"]
            orc.L.orc_photon_clear_leaks(C.byref(ophot))
        L.polycap_photon_free(ph)
    assert n_leak_seen > 0
    L.polycap_rng_free(rng)
    L.polycap_source_free(src)
    L.polycap_description_free(desc)
    L.polycap_profile_free(prof)
