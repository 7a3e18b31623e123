"""polycap_b200/csrc/pcb_leak_core.cuh -- the leak / recap state machine the CUDA leak kernel compiles -- built for the HOST
(one lane per "warp") and run on the CPU against the reference's own leak-mode output (tests/golden/replay_leak_*.npz, generated
by tests/golden/make_golden_leaks.py from oracle/_ref) and against itself with the wall march taken step by step.

What this pins without a GPU: the frame stack and the order of the leak records, and the wall march that JUMPS over its
uneventful steps (lk_march_jump): the path length after a jump has to be, bit for bit, what the reference's repeated
floating-point additions reach, or the photon enters the next capillary somewhere else and everything downstream differs.
The host build is test infrastructure; the library has no CPU path (the GPU tests of tests/test_gpu_leaks.py are the parity tests proper)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN, rel_err

RTOL = 1e-9
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def hl(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hl") / "libhostleak.so")
    csrc = os.path.join(ROOT, "polycap_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-I" + csrc, "-x", "c++",
                           os.path.join(ROOT, "tests", "host_leak.cpp"), "-x", "c", os.path.join(csrc, "pcb_econst.c"), "-o", out])
    L = C.CDLL(out)
    L.hl_replay.restype = C.c_long
    L.hl_replay.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_int, dp, dp, dp, C.c_int, C.c_long, dp, C.POINTER(C.c_int), dp, dp,
                            C.POINTER(C.c_long), dp, C.c_long, C.POINTER(C.c_ulonglong)]
    L.hl_trace_wall.restype = C.c_int
    L.hl_trace_wall.argtypes = [C.c_int, dp, dp, dp, C.c_double, C.c_int, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_ulonglong)]
    return L


def _p(a):
    return a.ctypes.data_as(dp)


def replay(L, g, start, flags=0, energies=None, amu=None, scatf=None):
    energies = np.ascontiguousarray(g["energies"] if energies is None else energies, dtype=np.float64)
    amu = np.ascontiguousarray(g["amu"] if amu is None else amu, dtype=np.float64)
    scatf = np.ascontiguousarray(g["scatf"] if scatf is None else scatf, dtype=np.float64)
    nE = len(energies)
    z, cap, ext = (np.ascontiguousarray(g[k], dtype=np.float64) for k in ("z", "cap", "ext"))
    start = np.ascontiguousarray(start, dtype=np.float64)
    n = len(start)
    ret = np.zeros(n, dtype=np.int32)
    state = np.zeros((n, 11))
    w = np.zeros((n, nE))
    ln = np.zeros(n, dtype=np.int64)
    cap_rec = 200000
    recs = np.zeros((cap_rec, 12 + nE))
    work = (C.c_ulonglong * 5)()
    k = L.hl_replay(int(g["nmax"]), _p(z), _p(cap), _p(ext), float(g["n_cap"]), float(g["sig_rough"]), float(g["density"]), nE, _p(energies), _p(amu),
                    _p(scatf), flags, n, _p(start), ret.ctypes.data_as(C.POINTER(C.c_int)), _p(state), _p(w), ln.ctypes.data_as(C.POINTER(C.c_long)),
                    _p(recs), cap_rec, work)
    assert k >= 0, k
    return dict(ret=ret, state=state, weights=w, leak_n=ln, recs=recs[:k].copy(), work=list(work))


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, "replay_leak_%s.npz" % name)))


@pytest.mark.parametrize("name", ["ellip", "conical"])
def test_host_state_machine_matches_reference_corpus(hl, name):
    g = _load(name)
    out = replay(hl, g, g["start"])
    assert np.array_equal(out["ret"], g["ret"]), np.nonzero(out["ret"] != g["ret"])
    assert np.array_equal(out["state"][:, 10].astype(np.int64), g["i_refl"])
    assert np.array_equal(out["leak_n"], g["leak_count"].sum(axis=1))
    ok = np.isin(g["ret"], (0, 1, 2))
    assert rel_err(out["state"][ok, 0:6], g["exit_state"][ok, 0:6]) <= RTOL
    assert rel_err(out["state"][ok, 9], g["d_travel"][ok]) <= RTOL
    traced = np.isin(g["ret"], (0, 1))
    assert rel_err(out["weights"][traced], g["weights"][traced]) <= RTOL
    R, mine = g["leak_records"], out["recs"]
    assert R.shape[0] > 0 and mine.shape == R.shape
    # the reference's records come photon by photon, external list first; within a list the order is the order of creation
    for kind in (0, 1):
        a, b = mine[mine[:, 1] == kind], R[R[:, 1] == kind]
        assert np.array_equal(a[:, 0], b[:, 0]) and np.array_equal(a[:, 11], b[:, 11])
        assert rel_err(a[:, 2:11], b[:, 2:11]) <= RTOL and rel_err(a[:, 12:], b[:, 12:]) <= RTOL
    # the march was mostly jumped over
    assert out["work"][2] < out["work"][1]


@pytest.mark.parametrize("name", ["ellip", "conical"])
def test_jumped_march_is_bit_identical_to_the_walked_march(hl, name):
    """flags 2: every uneventful step recognised one by one; flags 1: the reference's full step everywhere.  Identical photons, leak
    records and reference-equivalent step counts, bit for bit."""
    g = _load(name)
    n = 64 if name == "ellip" else 160
    jump = replay(hl, g, g["start"][:n])
    walk = replay(hl, g, g["start"][:n], flags=2)
    for k in ("ret", "state", "weights", "leak_n", "recs"):
        assert np.array_equal(jump[k], walk[k], equal_nan=True), k
    assert jump["work"][0] == walk["work"][0] and jump["work"][1] == walk["work"][1]  # same reflection events, same number of reference steps
    assert walk["work"][2] == walk["work"][1] and jump["work"][2] * 3 < jump["work"][1]
    full = replay(hl, g, g["start"][:24], flags=1)
    for k in ("ret", "state", "weights", "leak_n"):
        assert np.array_equal(jump[k][:24], full[k], equal_nan=True), k


def test_trace_wall_goldens_and_jump_counts(hl):
    """tests/leaks.c:46-113 (the reference's own trace_wall cases) with and without jumps"""
    g = _load("ellip")
    z, cap, ext = (np.ascontiguousarray(g[k]) for k in ("z", "cap", "ext"))
    cases = [((3.4999972129e-04, 0., 9.9997212889e-06), (0.00333, 0., 1.), 1, 0, 1, 0.029464),
             ((10e-5, 0., 8.9995), (0., 0., 1.), 2, 0, 0, 0.000507),
             ((0.2061, 0., 0.), (1., 0., 1.), 3, 0, 259, 0.012741),
             ((-0.072064, -0.008307, 8.783979), (0.061332, 0.004022, 0.998109), 3, -32, -226, 0.062987)]
    for co, di, code, r_want, q_want, dist in cases:
        res = []
        for flags in (0, 2, 1):
            inp = np.array(co + di, dtype=np.float64)
            dt = C.c_double(0)
            rq = (C.c_int * 2)()
            work = (C.c_ulonglong * 2)()
            rc = hl.hl_trace_wall(int(g["nmax"]), _p(z), _p(cap), _p(ext), float(g["n_cap"]), flags, _p(inp), C.byref(dt), rq, work)
            assert (rc, rq[0], rq[1]) == (code, r_want, q_want), (co, flags, rc, rq[0], rq[1])
            assert abs(dt.value - dist) < 1e-6
            res.append((dt.value, work[0]))
        assert res[0] == res[1] == res[2], res  # same d_travel bits, same number of reference steps


def test_jump_on_long_grid_and_many_photons(hl):
    """more start states than the corpus holds (its photons displaced sideways and tilted), jumped against walked"""
    g = _load("ellip")
    rng = np.random.default_rng(7)
    st = np.repeat(g["start"][:40], 6, axis=0).copy()
    st[:, 0:2] += rng.normal(0, 2e-3, size=(len(st), 2))
    st[:, 3:5] += rng.normal(0, 4e-4, size=(len(st), 2))
    jump = replay(hl, g, st)
    walk = replay(hl, g, st, flags=2)
    for k in ("ret", "state", "weights", "leak_n", "recs"):
        assert np.array_equal(jump[k], walk[k], equal_nan=True), k
    assert jump["work"][1] == walk["work"][1] and jump["recs"].shape[0] > 100


@pytest.mark.parametrize("name,n", [("xos1", 120), ("dubfoc", 120), ("parab", 120), ("conical", 200)])
def test_jump_on_other_geometries(hl, name, n):
    """the measured XOS profile (non-uniform z table), the double-focusing optic with rough walls (the roughness term of the energy loop),
    the paraboloidal and the conical profile with leaks on: the jumped march against the walked one, bit for bit"""
    from conftest import load_golden
    g = load_golden(name)
    st = g["start"][:n]
    jump = replay(hl, g, st)
    walk = replay(hl, g, st, flags=2)
    for k in ("ret", "state", "weights", "leak_n", "recs"):
        assert np.array_equal(jump[k], walk[k], equal_nan=True), k
    assert jump["work"][:2] == walk["work"][:2]
    assert jump["work"][0] > 100  # reflections happened
    full = replay(hl, g, st[:16], flags=1)
    for k in ("ret", "state", "weights", "leak_n"):
        assert np.array_equal(jump[k][:16], full[k], equal_nan=True), k


@pytest.mark.parametrize("name", ["xos1", "dubfoc", "monocap"])
def test_host_state_machine_matches_oracle_live_on_other_geometries(hl, harness, orc, name):
    """a few photons of other geometries through the oracle (the C restatement of the reference, leaks on, full march) and through the
    host build of the leak state machine: codes, reflection counts, leak counts exact, weights and leak records to 1e-9"""
    from conftest import load_golden, optic_of
    g = load_golden(name)
    o = optic_of(harness, g)
    n = 10
    st = g["start"][:n]
    mine = replay(hl, g, st)
    nE = len(g["energies"])
    k = 0
    for i in range(n):
        ph = orc.photon(o, st[i, 0:3], st[i, 3:6], st[i, 6:9])
        r, w = orc.launch(o, ph, leak=True)
        assert r == mine["ret"][i] and ph.i_refl == int(mine["state"][i, 10]), (i, r, mine["ret"][i])
        lk = orc.leaks(ph, nE)
        recs = mine["recs"][mine["recs"][:, 0] == i]
        for kind, key in ((0, "ext"), (1, "int")):
            got = recs[recs[:, 1] == kind]
            assert len(got) == len(lk[key]), (i, key)
            for a, b in zip(got, lk[key]):
                k += 1
                assert rel_err(a[2:5], b["coords"]) <= RTOL and rel_err(a[12:], b["weight"]) <= RTOL and int(a[11]) == b["n_refl"]
        if r in (0, 1):
            assert rel_err(mine["weights"][i], w) <= RTOL
        orc.L.orc_photon_clear_leaks(C.byref(ph))
