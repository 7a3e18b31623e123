"""polycap_source_get_transmission_efficiencies over several devices from ONE process (VERDICT r1 missing 1):
photon ids are dealt to the devices of POLYCAP_B200_DEVICES as contiguous sub-ranges, the "run until n photons reached the
exit" rule (/root/reference/src/polycap-source.c:745-884) cuts the id-ordered stream across devices, tallies are summed
over peer memory.  Photons are addressed by id, so the result must not depend on how many devices shared the work.

On a one-GPU box the same logic runs with the device listed several times (each entry is an independent share with its own
optic, streams and buffers); with two or more GPUs the real devices are used as well."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SIO2 = {8: 53.0, 14: 47.0}


def _source(pc, energies):
    prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
    desc = pc.Description(prof, 0.0, 200000, SIO2, 2.23)
    return pc.Source(desc, 2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5, energies)


def _run(pc, monkeypatch, devices, n_exit, energies, chunks=None):
    monkeypatch.setenv("POLYCAP_B200_SEED", "90210")
    if devices is None:
        monkeypatch.delenv("POLYCAP_B200_DEVICES", raising=False)
    else:
        monkeypatch.setenv("POLYCAP_B200_DEVICES", devices)
    if chunks is None:
        monkeypatch.delenv("POLYCAP_B200_CHUNKS", raising=False)
    else:
        monkeypatch.setenv("POLYCAP_B200_CHUNKS", str(chunks))
    src = _source(pc, energies)
    te = src.get_transmission_efficiencies(-1, n_exit, False)
    rec = te.records
    out = {k: (np.array(v) if not isinstance(v, (list, tuple)) else [np.array(a) for a in v]) for k, v in rec.items() if k not in ("i_start", "i_exit") and not k.startswith("_")}
    return te.data[1].copy(), rec["i_start"], rec["i_exit"], out


def _same(a, b):
    assert a[1] == b[1] and a[2] == b[2]  # photons simulated (the cut fell on the same id), exits
    assert np.allclose(a[0], b[0], rtol=1e-12, atol=0)  # efficiencies: the same terms, summed in another grouping
    for key in a[3]:
        x, y = a[3][key], b[3][key]
        if isinstance(x, list):
            for u, v in zip(x, y):
                assert np.array_equal(u, v), key
        else:
            assert np.array_equal(x, y), key


@pytest.mark.parametrize("devices", ["0,0", "0,0,0,0,0"])
def test_shares_on_one_device_equal_single_device_run(pc, monkeypatch, devices):
    energies = [1.0, 4.5, 9.0, 17.5, 29.0]
    one = _run(pc, monkeypatch, None, 30000, energies)
    many = _run(pc, monkeypatch, devices, 30000, energies)
    _same(one, many)
    again = _run(pc, monkeypatch, devices, 30000, energies)  # second call: waves sized from the first call's exit fraction
    _same(one, again)


def test_chunked_waves_with_overlapped_download_equal_one_wave(pc, monkeypatch):
    """records of a call are downloaded wave by wave while the next wave is traced; 64 energies so that the records are sizeable"""
    energies = np.linspace(1.0, 30.0, 64)
    _run(pc, monkeypatch, None, 40000, energies, chunks=1)  # sets the exit-fraction hint, so the next calls are planned, not probed
    one = _run(pc, monkeypatch, None, 40000, energies, chunks=1)
    eight = _run(pc, monkeypatch, None, 40000, energies, chunks=8)
    _same(one, eight)
    shared = _run(pc, monkeypatch, "0,0,0", 40000, energies, chunks=5)
    _same(one, shared)


def test_real_devices(pc, monkeypatch):
    n = pc.device_count()
    if n < 2:
        pytest.skip("one CUDA device")
    energies = [1.0, 4.5, 9.0, 17.5, 29.0]
    one = _run(pc, monkeypatch, "0", 60000, energies)
    allg = _run(pc, monkeypatch, "all", 60000, energies)
    _same(one, allg)
    monkeypatch.setenv("POLYCAP_B200_NO_PEER", "1")  # tallies staged through the host instead of read over peer memory
    staged = _run(pc, monkeypatch, "all", 60000, energies)
    _same(one, staged)
    assert np.array_equal(allg[0], staged[0])  # same terms in the same (device) order on both paths


@pytest.mark.parametrize("devices", [None, "0,0,0"])
def test_stream_chunks_equal_one_call(pc, monkeypatch, devices):
    """polycap_b200_stream_*: the records of a run delivered in chunks (N3) are the records of one big call, in the same order,
    and the running tallies are those of the ids consumed so far"""
    if devices is None:
        monkeypatch.delenv("POLYCAP_B200_DEVICES", raising=False)
    else:
        monkeypatch.setenv("POLYCAP_B200_DEVICES", devices)
    energies = np.linspace(1.0, 30.0, 12)
    src = _source(pc, energies)
    st = pc.Stream(src, seed=90210)
    chunks = [st.next(7000), st.next(13000), st.next(5000)]
    ids, sw, cnt = st.totals()
    n_tot = sum(f.shape[1] for f, _ in chunks)
    assert n_tot == 25000 and cnt["exit"] == 25000 and cnt["ids"] == ids
    one = _run(pc, monkeypatch, None, 25000, energies)
    rec = one[3]
    F = np.concatenate([f for f, _ in chunks], axis=1)
    Wt = np.concatenate([w for _, w in chunks], axis=0)
    assert np.array_equal(Wt, rec["exit_coord_weights"])
    for k in range(3):
        assert np.array_equal(F[8 + k], rec["pc_exit_coords"][k])
    assert np.array_equal(F[15].astype(np.int64), rec["pc_exit_nrefl"]) and np.array_equal(F[16], rec["pc_exit_dtravel"])
    started = cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]
    assert started == one[1]
    assert np.allclose(sw / started, one[0], rtol=1e-12, atol=0)
    # tally-only continuation with an id budget: no records, ids advance by exactly the budget
    n = st.next(10**9, max_ids=50000, records=False)
    ids2, sw2, cnt2 = st.totals()
    assert ids2 == ids + 50000 and cnt2["exit"] == 25000 + n and np.all(sw2 >= sw)
    sw3, cnt3, _, _ = src.run_ids(90210, 0, ids2, status=False)
    assert cnt3["exit"] == cnt2["exit"] and np.allclose(sw3, sw2, rtol=1e-12)
