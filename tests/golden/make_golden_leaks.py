"""Leak-mode replay corpus tests/golden/replay_leak_*.npz, generated from the REFERENCE ITSELF (oracle/_ref), like
make_golden.py but with leak_calc = true: besides the exit state it stores every external / internal leak record of
every photon (coords, direction, elecv, n_refl, weights) in the reference's order.  Leak mode is ~1000x slower than
plain mode on the CPU, hence the small photon counts.

    python tests/golden/make_golden_leaks.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import harness as H  # noqa: E402
from make_golden import build, OUT  # noqa: E402

CONFIGS = {
    # the reference's statistical test optic (tests/source.c:184-199), hard energies so that leaks occur
    "ellip": dict(profile=(2, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000,
                  source=(2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5), energies=[10., 40., 80.], n=160, seed=30000),
    # short conical optic seen from a close source: many steep photons, external leaks through the side
    "conical": dict(profile=(0, 2., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000,
                    source=(30.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.0), energies=[20., 60., 100.], n=160, seed=30001),
}


def main():
    ref = H.Ref()
    L = ref.L
    for name, cfg in CONFIGS.items():
        src = build(ref, cfg)
        s = src.contents
        d = s.description.contents
        nmax, z, cap, ext = ref.profile_arrays(d.profile)
        nE = s.n_energies
        energies = np.ctypeslib.as_array(s.energies, (nE,)).copy()
        amu, scatf = ref.optical_constants(s.description, energies)
        n = cfg["n"]
        rng = L.polycap_rng_new_with_seed(cfg["seed"])
        start = np.zeros((n, 9))
        ret = np.zeros(n, dtype=np.int32)
        exit_state = np.zeros((n, 9))
        irefl = np.zeros(n, dtype=np.int64)
        dtravel = np.zeros(n)
        weights = np.zeros((n, nE))
        leak_count = np.zeros((n, 2), dtype=np.int64)
        recs = []
        for i in range(n):
            ph = L.polycap_source_get_photon(src, rng, None)
            p = ph.contents
            start[i] = (*p.start_coords.t(), *p.start_direction.t(), *p.start_electric_vector.t())
            r, w = ref.launch(ph, energies, leak=True)
            ret[i] = r
            p = ph.contents
            exit_state[i] = (*p.exit_coords.t(), *p.exit_direction.t(), *p.exit_electric_vector.t())
            irefl[i] = p.i_refl
            dtravel[i] = p.d_travel
            if r in (0, 1):
                weights[i] = w
            lk = ref.leaks(ph)
            for kind, key in ((0, "ext"), (1, "int")):
                leak_count[i, kind] = len(lk[key])
                for rec in lk[key]:
                    recs.append([i, kind, *rec["coords"], *rec["direction"], *rec["elecv"], rec["n_refl"], *rec["weight"]])
            L.polycap_photon_free(ph)
        L.polycap_rng_free(rng)
        recs = np.array(recs, dtype=np.float64).reshape(-1, 12 + nE)
        path = os.path.join(OUT, "replay_leak_%s.npz" % name)
        np.savez_compressed(path, nmax=nmax, z=z, cap=cap, ext=ext, sig_rough=d.sig_rough, n_cap=d.n_cap, density=d.density,
                            iz=np.array([d.iz[k] for k in range(d.nelem)], dtype=np.int32), wi=np.array([d.wi[k] for k in range(d.nelem)]),
                            energies=energies, amu=amu, scatf=scatf, seed=cfg["seed"], start=start, ret=ret, exit_state=exit_state,
                            i_refl=irefl, d_travel=dtravel, weights=weights, leak_count=leak_count, leak_records=recs)
        vals, counts = np.unique(ret, return_counts=True)
        print("%-8s n=%d codes=%s ext leaks %d int leaks %d -> %s (%d KB)" % (name, n, dict(zip(vals.tolist(), counts.tolist())),
              leak_count[:, 0].sum(), leak_count[:, 1].sum(), os.path.basename(path), os.path.getsize(path) // 1024), flush=True)
        L.polycap_source_free(src)


if __name__ == "__main__":
    main()
