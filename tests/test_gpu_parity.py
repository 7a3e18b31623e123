"""Parity tests proper: the CUDA path, called through the C-ABI library, against
  (i)  the replay corpus generated from the reference itself (tests/golden, 1e-9 relative as north_star states),
  (ii) the oracle on the same seeded inputs (per-photon, same Philox stream),
  (iii) the golden vectors of the reference's own tests through the TEST_BUILD-private entry points,
  (iv) size-independent properties at larger sizes (determinism, additivity over id ranges, bounds)."""
import ctypes as C

import numpy as np
import pytest

from conftest import CORPUS, load_golden, optic_of, source_of, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-9  # BASELINE.json north_star: per-photon exit position, direction, reflection count and weights within 1e-9 relative


def _description(pc, g):
    prof = pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"])
    comp = {int(z): float(w) for z, w in zip(g["iz"], g["wi"])}
    n_cap = int(g["n_cap"])
    if n_cap <= 1:  # polycap_description_new refuses n_cap <= 1; the .inp parser accepts it (monocap.inp)
        pytest.skip("monocapillary descriptions are built through the .inp parser")
    return pc.Description(prof, float(g["sig_rough"]), n_cap, comp, float(g["density"]))


def _source(pc, g, desc):
    s = [float(v) for v in g["source"]]
    return pc.Source(desc, s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], g["energies"])


@pytest.mark.parametrize("kernel", [2, 1])
@pytest.mark.parametrize("exact", [0, 1])
@pytest.mark.parametrize("name", [n for n in CORPUS if n not in ("monocap", "cone_nan")])
def test_replay_matches_reference_corpus(pc, name, exact, kernel, monkeypatch):
    """exact=0: the default build (trajectory arithmetic bit-faithful, per-energy Fresnel term evaluated cheaply): 1e-9.
    exact=1 (POLYCAP_B200_EXACT_FRESNEL): every operation in the reference's order: weights agree to 1e-13
    (bit-identical where the wall roughness term exp() is not involved)."""
    monkeypatch.setenv("POLYCAP_B200_EXACT_FRESNEL", str(exact))
    monkeypatch.setenv("POLYCAP_B200_KERNEL", str(kernel))  # 2: CTA task-queue tracer (default), 1: warp-voted tracer
    g = load_golden(name)
    desc = _description(pc, g)
    out = pc.launch_batch(desc, g["energies"], g["start"][:, 0:3], g["start"][:, 3:6], g["start"][:, 6:9], amu=g["amu"], scatf=g["scatf"])
    assert np.array_equal(out["ret"], g["ret"]), np.nonzero(out["ret"] != g["ret"])
    assert np.array_equal(out["i_refl"], g["i_refl"])
    # the trajectory is bit-identical in both modes and with both kernels
    assert rel_err(out["exit_coords"], g["exit_state"][:, 0:3]) == 0.0
    assert rel_err(out["exit_direction"], g["exit_state"][:, 3:6]) == 0.0
    assert rel_err(out["d_travel"], g["d_travel"]) == 0.0
    assert rel_err(out["exit_electric_vector"], g["exit_state"][:, 6:9]) <= (1e-13 if exact else RTOL)
    traced = np.isin(g["ret"], (0, 1))
    assert rel_err(out["weights"][traced], g["weights"][traced]) <= (1e-13 if exact else RTOL)


@pytest.mark.parametrize("name", ["monocap", "cone_nan"])
def test_replay_inp_only_configs(pc, name, tmp_path, monkeypatch):
    """mono-capillary (n_cap = 1) and the degenerate shipped cone.inp (NaN photons) via the .inp parser"""
    g = load_golden(name)
    n = int(g["nmax"])
    for ext_, col in (("prf", g["cap"]), ("ext", g["ext"])):
        (tmp_path / ("p." + ext_)).write_text("%d\n" % n + "".join("%.17e %.17e\n" % (a, b) for a, b in zip(g["z"], col)))
    (tmp_path / "p.axs").write_text("%d\n" % n + "".join("%.17e 0. 0.\n" % a for a in g["z"]))
    s = g["source"]
    (tmp_path / "p.inp").write_text("%.17g\n%.17g\n%.17g %.17g\n%.17g %.17g\n%.17g %.17g\n%.17g\n2\n8 53.0\n14 47.0\n2.23\n1.0 30.0 1.0\n100\n-1\np.prf\np.axs\np.ext\n%d.\nout\n"
                                    % (float(g["sig_rough"]), s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7], int(g["n_cap"])))
    monkeypatch.chdir(tmp_path)
    src = pc.Source.new_from_file("p.inp")

    class S(C.Structure):
        _fields_ = [("description", C.c_void_p)]
    desc = pc.Description.__new__(pc.Description)
    desc._description = None
    handle = C.cast(src._source, C.POINTER(S)).contents.description

    class Borrowed:  # launch_batch only needs ._description
        _description = handle
    out = pc.launch_batch(Borrowed, g["energies"], g["start"][:, 0:3], g["start"][:, 3:6], g["start"][:, 6:9], amu=g["amu"], scatf=g["scatf"])
    assert np.array_equal(out["ret"], g["ret"])
    assert np.array_equal(out["i_refl"], g["i_refl"])
    traced = np.isin(g["ret"], (0, 1))
    assert rel_err(out["weights"][traced], g["weights"][traced]) <= RTOL  # NaN == NaN (Appendix A.8)
    assert rel_err(out["exit_coords"], g["exit_state"][:, 0:3]) <= RTOL
    assert rel_err(out["d_travel"], g["d_travel"]) <= RTOL


@pytest.mark.parametrize("name", ["ellip", "parab", "xos1", "conical"])
def test_source_run_matches_oracle_per_photon(pc, harness, orc, name):
    """Source-driven run, photon by photon.  (1) the start states the device draws from each id's Philox stream agree
    with the oracle's source sampling over the same stream to rounding (a direct function evaluation: CUDA libm vs
    glibc differ by a few ulp in tan/atan/acos); (2) the transport is chaotic -- one ulp in a start state grows
    geometrically over ~40 reflections -- so the trace is checked by feeding the DEVICE's start states to the oracle:
    classification identical, exit records and tallies to 1e-9."""
    g = load_golden(name)
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    amu, scatf = desc.optical_constants(g["energies"])
    o = harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], float(g["sig_rough"]), int(g["n_cap"]), float(g["density"]), g["energies"], amu, scatf)
    n, seed, id0 = 6000, 4242, 1000
    starts = src.sample_ids(seed, id0, n)
    s = source_of(harness, g)
    worst = 0.0
    for j in range(0, n, 7):
        st = harness.OrcPhilox(seed, id0 + j, 0)
        ph = harness.OrcPhoton()
        orc.L.orc_source_get_photon(C.byref(s), C.byref(o.c), C.cast(orc.L.orc_philox_uniform, C.c_void_p), C.byref(st), C.byref(ph))
        want = np.array([*ph.start_coords.t(), *ph.start_direction.t(), *ph.start_electric_vector.t(), ph.src_start_coords.x, ph.src_start_coords.y])
        scale = np.array([0.2, 0.2, 1, 1e-3, 1e-3, 1, 1, 1, 1e-3, 0.2, 0.2])  # absolute scale of each component
        worst = max(worst, float(np.max(np.abs(starts[j] - want) / scale)))
    assert worst < 1e-12, worst
    sw, cnt, status, rec = src.run_ids(seed, id0, n, records=True)
    osw, ocnt, ostatus, orec = orc.run_starts(o, starts, threads=8, records=True)
    assert np.array_equal(status, ostatus), np.nonzero(status != ostatus)
    ex = status == 1
    assert ex.sum() > (20 if name == "conical" else 100)
    assert rel_err(rec[ex], orec[ex]) <= RTOL
    assert rel_err(sw, osw) <= RTOL
    assert cnt["exit"] == ocnt["n_exit"] and cnt["not_entered"] == ocnt["n_not_entered"]
    assert cnt["not_transmitted"] == ocnt["n_not_transmitted"] and cnt["uncounted"] == ocnt["n_uncounted"]
    assert cnt["refl_exit"] == ocnt["sum_irefl_exit"] and cnt["refl_all"] == ocnt["sum_irefl_all"]
    assert cnt["ids"] == n


def test_source_run_statistics_match_oracle(pc, harness, orc):
    """north_star check 2: transmission efficiency per energy within 3 sigma binomial error at equal photon counts
    (independent streams: the oracle samples on the CPU, the device on the GPU, different seeds)"""
    g = load_golden("ellip")
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    amu, scatf = desc.optical_constants(g["energies"])
    o = harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], 0.0, int(g["n_cap"]), float(g["density"]), g["energies"], amu, scatf)
    n = 40000
    sw, cnt, _, _ = src.run_ids(1, 0, n, status=False)
    osw, ocnt, _, _ = orc.run_ids(source_of(harness, g), o, 2, 0, n, threads=8)
    started = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]
    ostarted = ocnt["n_exit"] + ocnt["n_not_entered"] + ocnt["n_not_transmitted"]
    eff, oeff = sw / started, osw / ostarted
    for e, oe in zip(eff, oeff):
        sigma = np.sqrt(max(oe * (1 - oe), 1e-6) / started + max(e * (1 - e), 1e-6) / ostarted)
        assert abs(e - oe) < 3 * sigma, (eff, oeff)
    p, op = cnt["exit"] / n, ocnt["n_exit"] / n
    assert abs(p - op) < 3 * np.sqrt(2 * p * (1 - p) / n)


def test_both_tracers_agree_bitwise(pc, monkeypatch):
    """the warp-voted tracer and the task-queue tracer schedule the same arithmetic: identical per-photon results"""
    g = load_golden("xos1")
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    monkeypatch.setenv("POLYCAP_B200_KERNEL", "1")
    a = src.run_ids(11, 0, 30000, records=True)
    monkeypatch.setenv("POLYCAP_B200_KERNEL", "2")
    b = src.run_ids(11, 0, 30000, records=True)
    assert np.array_equal(a[2], b[2])
    ex = a[2] == 1
    assert np.array_equal(a[3][ex], b[3][ex])
    assert np.array_equal(a[0], b[0])
    for k in ("exit", "not_entered", "not_transmitted", "uncounted", "refl_exit", "refl_all", "seg_steps", "solves"):
        assert a[1][k] == b[1][k], k


def test_exact_checks_mode_agrees(pc, monkeypatch):
    """the fast path (interior capillaries skip the hexagon tests) equals the exact-checks path photon by photon"""
    g = load_golden("ellip")
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    a = src.run_ids(7, 0, 20000, records=True)
    monkeypatch.setenv("POLYCAP_B200_EXACT_CHECKS", "1")
    g2 = dict(g)
    desc2 = pc.Description(pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"]), 1e-300, int(g["n_cap"]), {8: 53.0, 14: 47.0}, 2.23)  # new cache key
    src2 = _source(pc, g2, desc2)
    b = src2.run_ids(7, 0, 20000, records=True)
    assert np.array_equal(a[2], b[2])
    ex = a[2] == 1
    assert rel_err(a[3][ex], b[3][ex]) <= 1e-12


# ------------------------------------------------------------------ the reference's unit goldens through the private entry points
class V3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


def test_private_unit_goldens(pc):
    lib = C.CDLL(pc.LIB_PATH)
    lib.polycap_capil_segment.restype = C.c_int
    lib.polycap_capil_segment.argtypes = [V3, V3, C.c_double, C.c_double, V3, V3, V3, C.POINTER(V3), C.POINTER(V3), C.c_void_p]
    hit, sn = V3(0, 0, 0), V3()
    r = lib.polycap_capil_segment(V3(0, 0, 0), V3(0, 0, 0.1), 0.005, 0.005, V3(0, 0, 0), V3(0.005, -0.005, 0.1), V3(0.005, -0.005, 0.1),
                                  C.byref(hit), C.byref(sn), None)  # tests/capil.c:29-65
    assert r == 1
    assert np.allclose((hit.x, hit.y, hit.z), (0.003536, -0.003536, 0.070711), atol=1e-5)
    assert np.allclose((sn.x, sn.y, sn.z), (0.707107, -0.707107, 0.0), atol=1e-5)
    lib.polycap_photon_within_pc_boundary.restype = C.c_int
    lib.polycap_photon_within_pc_boundary.argtypes = [C.c_double, V3, C.c_void_p]
    assert lib.polycap_photon_within_pc_boundary(0.05, V3(0.025, 0.025, 0), None) == 1  # tests/photon.c:166-190
    assert lib.polycap_photon_within_pc_boundary(0.05, V3(0.075, 0.025, 0), None) == 0
    assert lib.polycap_photon_within_pc_boundary(0.05, V3(0.05, 0.05, 0), None) == 0


def test_photon_launch_codes_and_state(pc):
    """tests/photon.c:232-352 and tests/capil.c:405-450 through the public single-photon API"""
    prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
    desc = pc.Description(prof, 0.0, 200000, {8: 53.0, 14: 47.0}, 2.23)
    cases = [((0.21, 0, 0), (0, 0, 1), -2), ((0, 0, 0), (0.005, -0.005, 0.1), 0), ((0, 0, 0), (0, 0, 1.0), 1),
             ((0.15104418, 0.08700043, 0), (0, 0, 1.0), 2), ((-0.192065, -0.022121, 0), (0.000018, 0.000002, 0.2), 0)]
    for c, d, code in cases:
        ph = pc.Photon(desc, c, d, (0.5, 0.5, 0))
        if code == -2:  # the reference sets POLYCAP_ERROR_INVALID_ARGUMENT for a photon outside the optic (photon.c:555)
            with pytest.raises(ValueError, match="photon not within optic boundaries"):
                ph.launch([10.0])
            continue
        w = ph.launch([10.0])
        assert ph.last_return == code, (c, d, ph.last_return)
        if code == 2:
            assert w is None
        if code == 1:
            assert ph.i_refl == 0 and w[0] == 1.0
    ph = pc.Photon(desc, (0, 0, 0), (3e-5, 3e-5, 0.999), (0.5, 0.5, 0))
    w = ph.launch([10.0])
    assert ph.last_return == 1 and abs(w[0] - 0.999585) < 2e-5 or ph.i_refl > 1  # one capil_trace call gives 0.999585 (tests/capil.c:442)
    with pytest.raises(ValueError, match="energies"):
        ph.launch([0.5])


def test_get_photon_matches_reference_stream(pc):
    """polycap_source_get_photon over MT19937 (seed of the corpus) reproduces the reference's start states"""
    for name in ("ellip", "parab"):
        g = load_golden(name)
        desc = _description(pc, g)
        src = _source(pc, g, desc)
        rng = pc.Rng(int(g["seed"]))
        for i in range(40):
            ph = src.get_photon(rng)
            got = np.array([*ph.start_coords, *ph.start_direction, *ph.start_electric_vector])
            assert rel_err(got, g["start"][i]) <= 1e-12, (name, i)


def test_transmission_efficiencies_api(pc):
    """tests/source.c:204-287: efficiencies at 1..30 keV within the reference's own tolerances; record invariants"""
    prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
    desc = pc.Description(prof, 0.0, 200000, {8: 53.0, 14: 47.0}, 2.23)
    src = pc.Source(desc, 2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5, [1, 5, 10, 15, 20, 25, 30])
    with pytest.raises(ValueError):
        src.get_transmission_efficiencies(-1, -1)
    te = src.get_transmission_efficiencies(-1, 30000, False)
    en, eff = te.data
    want = [0.424, 0.349, 0.135, 0.050, 0.022, 0.011, 0.006]
    tol = [0.01, 0.01, 0.0075, 0.005, 0.005, 0.005, 0.005]
    for e, w, t in zip(eff, want, tol):
        assert abs(e - w) < t, (eff, want)
    ex = te.exit_arrays
    st = te.start_arrays
    assert st["n_exit"] == 30000 and ex["coords"].shape == (30000, 3)
    assert np.all(ex["coords"][:, 2] == 9.0)
    assert np.all(ex["n_refl"] >= 0) and 0 < ex["n_refl"][0] < 200 or ex["n_refl"][0] == 0
    assert np.all(ex["d_travel"] >= 9.0)
    assert np.all((ex["weights"] >= 0) & (ex["weights"] <= 1))
    assert np.all(np.hypot(st["coords"][:, 0], st["coords"][:, 1]) <= 0.2065 * 1.0001)
    assert st["n_start"] > 30000
    w0 = next(te.exit_weights)
    assert not w0.flags.writeable and w0.size == 7
    # zero-copy views of the result's own storage (struct _polycap_images) agree with the re-packed getters
    rec = te.records
    assert rec["i_exit"] == 30000 and rec["i_start"] == st["n_start"]
    for k in range(3):
        assert np.array_equal(rec["pc_exit_coords"][k], ex["coords"][:, k])
    for k in range(2):
        assert np.array_equal(rec["pc_exit_dir"][k], ex["direction"][:, k])
        assert np.array_equal(rec["pc_start_coords"][k], st["coords"][:, k])
        assert np.array_equal(rec["src_start_coords"][k], st["src"][:, k])
    assert np.array_equal(rec["pc_exit_nrefl"], ex["n_refl"]) and np.array_equal(rec["pc_exit_dtravel"], ex["d_travel"])
    assert np.array_equal(rec["exit_coord_weights"], ex["weights"])
    assert not rec["exit_coord_weights"].flags.writeable


# ------------------------------------------------------------------ properties at larger sizes
def test_determinism_and_additivity_large(pc):
    g = load_golden("conical")
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    n = 2_000_000
    sw, cnt, _, _ = src.run_ids(99, 0, n, status=False)
    sw2, cnt2, _, _ = src.run_ids(99, 0, n, status=False)
    phys = ("ids", "exit", "not_entered", "not_transmitted", "uncounted", "refl_exit", "refl_all", "seg_steps", "solves", "launched")
    assert np.array_equal(sw, sw2) and all(cnt[k] == cnt2[k] for k in phys)  # fixed-order tally: bitwise reproducible (q_* are scheduler statistics)
    a = src.run_ids(99, 0, n // 2, status=False)
    b = src.run_ids(99, n // 2, n - n // 2, status=False)
    for k in ("exit", "not_entered", "not_transmitted", "uncounted", "refl_exit", "refl_all", "ids"):
        assert a[1][k] + b[1][k] == cnt[k], k
    assert np.allclose(a[0] + b[0], sw, rtol=1e-12)
    assert cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"] + cnt["uncounted"] == n
    started = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]
    eff = sw / started
    assert np.all(eff >= 0) and np.all(eff <= 1) and np.all(np.diff(eff[2:]) <= 1e-3)  # efficiency falls with energy above a few keV
    sw3, cnt3, _, _ = src.run_ids(100, 0, 200000, status=False)  # another seed: statistically compatible open area
    oa = (cnt["exit"] + cnt["not_transmitted"]) / started
    st3 = cnt3["exit"] + cnt3["not_entered"] + cnt3["not_transmitted"]
    oa3 = (cnt3["exit"] + cnt3["not_transmitted"]) / st3
    assert abs(oa - oa3) < 5 * np.sqrt(oa * (1 - oa) / st3)


def test_session_device_tallies(pc):
    """device-resident accumulation used by bench.py and the multi-GPU reduction"""
    import torch
    g = load_golden("ellip")
    desc = _description(pc, g)
    src = _source(pc, g, desc)
    dev = torch.device("cuda:0")
    t_sum = torch.zeros(len(g["energies"]), dtype=torch.float64, device=dev)
    t_cnt = torch.zeros(pc.N_COUNTERS, dtype=torch.int64, device=dev)
    ses = pc.Session(src, device=0, seed=5)
    stream = torch.cuda.current_stream().cuda_stream
    ses.run(0, 50000, t_sum.data_ptr(), t_cnt.data_ptr(), wave_size=20000, stream=stream)
    ses.run(50000, 50000, t_sum.data_ptr(), t_cnt.data_ptr(), wave_size=50000, stream=stream)
    torch.cuda.synchronize()
    sw, cnt, _, _ = src.run_ids(5, 0, 100000, status=False)
    assert np.allclose(t_sum.cpu().numpy(), sw, rtol=1e-12)
    c = t_cnt.cpu().numpy()
    assert c[0] == 100000 and c[1] == cnt["exit"] and c[6] == cnt["refl_all"]


def test_shared_divisor_division_is_ieee(pc):
    """pcb_math.h::pcb_div (one reciprocal refinement per divisor, used by norm3 / segment_solve) against the compiler's
    IEEE `a / b` on the device: ordinary magnitudes, wide exponent ranges, and the full range incl. 0 / subnormal / inf / NaN"""
    for seed, span in ((1, 3), (2, 40), (3, 400), (4, 1000), (5, 1100)):
        assert pc.selftest_division(1 << 25, seed, span) == 0, (seed, span)


def test_repeated_calls_are_identical_whatever_the_wave_partition(pc, monkeypatch):
    """The first call of a source sizes its waves from a probe, later calls from the exit fraction the previous call ended
    with (one wave); photons are addressed by id, so the results must not depend on that -- nor on whether the exit records
    land in a pinned or a pageable host block."""
    monkeypatch.setenv("POLYCAP_B200_SEED", "4242")
    prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
    desc = pc.Description(prof, 0.0, 200000, {8: 53.0, 14: 47.0}, 2.23)
    src = pc.Source(desc, 2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5, [3.0, 9.0, 17.5])
    first = src.get_transmission_efficiencies(-1, 20000, False)
    second = src.get_transmission_efficiencies(-1, 20000, False)
    monkeypatch.setenv("POLYCAP_B200_PAGEABLE_RESULTS", "1")
    third = src.get_transmission_efficiencies(-1, 20000, False)
    a, b, c = first.records, second.records, third.records
    assert a["i_start"] == b["i_start"] == c["i_start"]
    # per-photon records are identical; the tallies are sums over waves, i.e. the same terms in a different grouping
    assert np.allclose(first.data[1], second.data[1], rtol=1e-13, atol=0) and np.array_equal(second.data[1], third.data[1])
    for key in ("pc_exit_coords", "pc_exit_dir", "pc_start_coords", "src_start_coords", "pc_exit_elecv"):
        for x, y, z in zip(a[key], b[key], c[key]):
            assert np.array_equal(x, y) and np.array_equal(x, z), key
    assert np.array_equal(a["exit_coord_weights"], b["exit_coord_weights"]) and np.array_equal(a["exit_coord_weights"], c["exit_coord_weights"])
    assert np.array_equal(a["pc_exit_nrefl"], b["pc_exit_nrefl"]) and np.array_equal(a["pc_exit_nrefl"], c["pc_exit_nrefl"])
    # results stay valid after the objects that shared their pool blocks are gone
    keep = a["exit_coord_weights"].copy()
    del second, third, b, c
    again = src.get_transmission_efficiencies(-1, 20000, False)
    assert np.array_equal(a["exit_coord_weights"], keep)
    assert np.array_equal(again.records["exit_coord_weights"], keep)
