"""Parity where the target runs live: the energy grids and photon counts of the BASELINE workloads.

  * per-photon oracle comparison at nE = 291 (the full example/xos1.inp grid on the measured XOS tables) and nE = 400
    (ellipsoidal, 1-40 keV), >= 1e5 photon ids each: the device's start states go through the CPU oracle
    (reference: /root/reference/src/polycap-capil.c:625-645 is the energy loop these sizes stress);
  * 1e6-id runs of the conical and ellipsoidal geometries with the fast path on (convexity skip, interior-capillary
    margins): classification and reflection counts exact, exit records and tallies to 1e-9.

Tolerance: 1e-9 relative (BASELINE.json north_star); integers (status, i_refl, counters) exact."""
import os

import numpy as np
import pytest

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
RTOL = 1e-9
THREADS = os.cpu_count() or 8
SIO2 = {8: 53.0, 14: 47.0}


def _note(text):
    """measured margins, kept with the GPU run's other outputs when that directory exists"""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_margins.txt"), "a") as f:
            f.write(text + "\n")


def _setup(pc, harness, name, energies):
    g = load_golden(name)
    prof = pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"])
    desc = pc.Description(prof, float(g["sig_rough"]), int(g["n_cap"]), SIO2, 2.23)
    s = [float(v) for v in g["source"]]
    src = pc.Source(desc, *s[:8], energies)
    amu, scatf = desc.optical_constants(energies)
    o = harness.Optic(int(g["nmax"]), g["z"], g["cap"], g["ext"], float(g["sig_rough"]), int(g["n_cap"]), 2.23, energies, amu, scatf)
    return src, o


def _compare(src, orc, o, seed, id0, n, chunk, min_exit):
    """device run of ids [id0, id0+n) vs the oracle fed with the device's own start states, in chunks that bound host memory"""
    worst_rec = worst_sw = 0.0
    tot = dict(exit=0, refl_all=0)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        starts = src.sample_ids(seed, id0 + lo, m)
        sw, cnt, status, rec = src.run_ids(seed, id0 + lo, m, records=True)
        osw, ocnt, ostatus, orec = orc.run_starts(o, starts, threads=THREADS, records=True)
        bad = np.nonzero(status != ostatus)[0]
        assert bad.size == 0, ("classification differs", lo, bad[:10], status[bad[:10]], ostatus[bad[:10]])
        ex = status == 1
        assert np.array_equal(rec[ex][:, 15], orec[ex][:, 15]), "reflection counts differ"
        worst_rec = max(worst_rec, rel_err(rec[ex], orec[ex]))
        worst_sw = max(worst_sw, rel_err(sw, osw))
        assert cnt["exit"] == ocnt["n_exit"] and cnt["not_entered"] == ocnt["n_not_entered"]
        assert cnt["not_transmitted"] == ocnt["n_not_transmitted"] and cnt["uncounted"] == ocnt["n_uncounted"]
        assert cnt["refl_exit"] == ocnt["sum_irefl_exit"] and cnt["refl_all"] == ocnt["sum_irefl_all"]
        assert cnt["ids"] == m
        tot["exit"] += cnt["exit"]
        tot["refl_all"] += cnt["refl_all"]
    _note("per-photon vs oracle: nE=%d ids=%d exits=%d worst record rel err %.3e, worst tally rel err %.3e" % (o.nE, n, tot["exit"], worst_rec, worst_sw))
    assert worst_rec <= RTOL, worst_rec
    assert worst_sw <= RTOL, worst_sw
    assert tot["exit"] >= min_exit, tot
    return worst_rec, worst_sw


def test_xos1_full_grid_291_energies_per_photon(pc, harness, orc):
    """BASELINE configs[3]: measured XOS profile, the 291 energies of example/xos1.inp (1..30 keV step 0.1), 1e5 ids"""
    energies = 1.0 + 0.1 * np.arange(291)
    src, o = _setup(pc, harness, "xos1", energies)
    _compare(src, orc, o, seed=20000, id0=0, n=100_000, chunk=25_000, min_exit=50_000)


def test_ellipsoidal_400_energies_per_photon(pc, harness, orc):
    """BASELINE configs[2] grid: ellipsoidal ellip_l9 geometry, 400 energies 1..40 keV, 1e5 ids"""
    energies = np.linspace(1.0, 40.0, 400)
    src, o = _setup(pc, harness, "ellip", energies)
    _compare(src, orc, o, seed=31337, id0=5_000_000, n=100_000, chunk=25_000, min_exit=50_000)


def test_monocap_grid_80_energies_per_photon(pc, harness, orc):
    """an energy count between the register-dealt and the long-grid regimes (example/monocap.inp has 80 energies), polycapillary XOS tables"""
    energies = np.arange(1.0, 81.0)
    src, o = _setup(pc, harness, "xos1", energies)
    _compare(src, orc, o, seed=7, id0=123, n=40_000, chunk=40_000, min_exit=20_000)


@pytest.mark.parametrize("name,min_exit", [("conical", 20_000), ("ellip", 500_000)])
def test_fast_path_one_million_ids(pc, harness, orc, name, min_exit):
    """1e6 ids with the fast path on (the configuration bench.py times): every id's classification and reflection count equals
    the oracle's, which takes the reference's per-segment checks everywhere; records and tallies to 1e-9"""
    g = load_golden(name)
    energies = np.linspace(1, 30, 10) if name == "conical" else g["energies"]
    src, o = _setup(pc, harness, name, energies)
    _compare(src, orc, o, seed=99, id0=0, n=1_000_000, chunk=250_000, min_exit=min_exit)


@pytest.mark.parametrize("name,nE,hk", [("conical", 64, 0), ("conical", 64, 5), ("xos1", 291, 0), ("xos1", 80, 3), ("dubfoc", 246, 0)])
def test_deferred_weights_equal_immediate_update(pc, monkeypatch, name, nE, hk):
    """Long energy grids evaluate the per-energy products of a photon's buffered reflections in one go (pcb_trace2.cuh, task ENERGY)
    instead of after every reflection.  Same arithmetic in the same order: exit records and tallies are bit-identical to the
    immediate update, including photons that the evaluation finds absorbed after they had been traced on (the conical geometry
    absorbs half of its photons: roll-back of classification and reflection count), histories flushed in mid flight (small hk)
    and rough walls (dub_foc: sig_rough = 5 A)."""
    g = load_golden(name)
    energies = np.linspace(1.0, 30.0, nE)
    out = []
    for mode in ("0", "48"):
        monkeypatch.setenv("POLYCAP_B200_DEFER_MIN", mode)  # read when the device optic is created; part of its cache key
        monkeypatch.setenv("POLYCAP_B200_DEFER_HK", str(hk))
        prof = pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"])
        desc = pc.Description(prof, float(g["sig_rough"]), int(g["n_cap"]), SIO2, 2.23)
        src = pc.Source(desc, *[float(v) for v in g["source"]][:8], energies)
        out.append(src.run_ids(4321, 10_000, 60_000, records=True))
    (sw0, c0, st0, r0), (sw1, c1, st1, r1) = out
    assert np.array_equal(st0, st1), np.nonzero(st0 != st1)[0][:10]
    ex = st0 == 1
    assert ex.sum() > 500
    assert np.array_equal(r0[ex], r1[ex])
    assert np.array_equal(sw0, sw1)
    for k in ("ids", "exit", "not_entered", "not_transmitted", "uncounted", "refl_exit", "refl_all", "launched"):
        assert c0[k] == c1[k], (k, c0[k], c1[k])
    if c0["not_transmitted"] == 0:  # nothing rolled back: no segment was traced in vain
        assert c0["seg_steps"] == c1["seg_steps"] and c0["solves"] == c1["solves"]
    if name == "conical":
        assert c0["not_transmitted"] > 10_000  # the roll-back path was exercised
