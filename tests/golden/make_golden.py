"""Generates the replay corpus tests/golden/*.npz from the REFERENCE ITSELF.

Runs only in the build container: it needs /root/reference (for example/*.prf etc.) and
oracle/_ref/libpolycap_ref.so (`make -C oracle ref`: the reference's own sources compiled in place
against the stand-in headers of oracle/shims).  For every configuration it loops
polycap_source_get_photon + polycap_photon_launch exactly like the reference's own test does
(/root/reference/tests/source.c:313-330) with a seeded MT19937 stream and stores

  inputs : profile tables, material, energies, amu/scatf as that build derives them, and the
           start state (coords, direction, electric vector) of every photon
  outputs: return code, exit coords/direction/electric vector, i_refl, d_travel, weights

so that the CUDA path and the C restatement can be replayed against the reference on a machine
that has neither /root/reference nor the reference build.

    python tests/golden/make_golden.py            # rewrites tests/golden/replay_*.npz
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import harness as H  # noqa: E402

EX = "/root/reference/example"
OUT = os.path.dirname(os.path.abspath(__file__))

SIO2 = ([8, 14], [53.0, 47.0], 2.23)


def build(ref, cfg):
    L = ref.L
    if "inp" in cfg:
        cwd = os.getcwd()
        os.chdir(EX)
        try:
            src = L.polycap_source_new_from_file(cfg["inp"].encode(), None)
        finally:
            os.chdir(cwd)
        assert src, cfg
        if "energies" in cfg:  # thin the .inp energy grid to keep the fixture small
            e = np.ascontiguousarray(cfg["energies"], dtype=np.float64)
            libc = C.CDLL(None)
            libc.malloc.restype = C.c_void_p
            mem = libc.malloc(C.c_size_t(8 * e.size))
            C.memmove(mem, e.ctypes.data, 8 * e.size)
            libc.free.argtypes = [C.c_void_p]
            libc.free(C.cast(src.contents.energies, C.c_void_p))
            src.contents.energies = C.cast(mem, H.c_double_p)
            src.contents.n_energies = e.size
        for k in ("src_shiftx", "src_sigx", "src_sigy"):
            if k in cfg:
                setattr(src.contents, k, cfg[k])
        return src
    prof = L.polycap_profile_new(*cfg["profile"], None)
    assert prof
    iz, wi, rho = SIO2
    desc = ref.description(prof, cfg.get("sig_rough", 0.0), cfg["n_cap"], iz, wi, rho)
    assert desc
    src = ref.source(desc, *cfg["source"], cfg["energies"])
    assert src
    L.polycap_description_free(desc)
    L.polycap_profile_free(prof)
    return src


CONFIGS = {
    # tests/source.c:184-199 (the reference's own statistical test optic)
    "ellip": dict(profile=(2, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000,
                  source=(2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5), energies=[1, 5, 10, 15, 20, 25, 30], n=1500, seed=20000),
    # BASELINE config 2: POLYCAP_PROFILE_CONICAL, 10-energy grid
    "conical": dict(profile=(0, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5), n_cap=200000,
                    source=(100.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.0), energies=np.linspace(1, 30, 10), n=1500, seed=20001),
    # paraboloidal generator, rough walls, off-axis divergent source
    "parab": dict(profile=(1, 6., 0.2883, 0.07, 0.00035, 0.000085, 500., 0.25), n_cap=227701, sig_rough=5e-8,
                  source=(500.0, 0.2, 0.01, 0.00024, 0.00024, 0.05, 0.0, 0.9), energies=np.linspace(1, 50, 12), n=1500, seed=20002),
    # measured profile files (BASELINE config 4)
    "xos1": dict(inp="xos1.inp", energies=np.linspace(1, 30, 9), n=1500, seed=20003),
    # mono-capillary, uniform illumination (BASELINE config 1)
    "monocap": dict(inp="monocap.inp", energies=np.linspace(1, 80, 8), n=1500, seed=20004),
    # the shipped cone.inp: src_y = 0 -> NaN photons
    "cone_nan": dict(inp="cone.inp", energies=np.linspace(1, 30, 4), n=400, seed=20005),
    # double-focusing example as shipped (type 2), polarised, divergent
    "dubfoc": dict(inp="dub_foc.inp", energies=np.linspace(1, 50, 6), n=1500, seed=20006),
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
        srcxy = np.zeros((n, 2))
        ret = np.zeros(n, dtype=np.int32)
        exit_state = np.zeros((n, 9))
        irefl = np.zeros(n, dtype=np.int64)
        dtravel = np.zeros(n)
        weights = np.zeros((n, nE))
        for i in range(n):
            ph = L.polycap_source_get_photon(src, rng, None)
            p = ph.contents
            start[i] = (*p.start_coords.t(), *p.start_direction.t(), *p.start_electric_vector.t())
            srcxy[i] = (p.src_start_coords.x, p.src_start_coords.y)
            r, w = ref.launch(ph, energies)
            ret[i] = r
            p = ph.contents
            exit_state[i] = (*p.exit_coords.t(), *p.exit_direction.t(), *p.exit_electric_vector.t())
            irefl[i] = p.i_refl
            dtravel[i] = p.d_travel
            if r in (0, 1):
                weights[i] = w
            L.polycap_photon_free(ph)
        L.polycap_rng_free(rng)
        iz = np.array([d.iz[k] for k in range(d.nelem)], dtype=np.int32)
        wi = np.array([d.wi[k] for k in range(d.nelem)])
        source = np.array([s.d_source, s.src_x, s.src_y, s.src_sigx, s.src_sigy, s.src_shiftx, s.src_shifty, s.hor_pol])
        path = os.path.join(OUT, "replay_%s.npz" % name)
        np.savez_compressed(path, nmax=nmax, z=z, cap=cap, ext=ext, sig_rough=d.sig_rough, n_cap=d.n_cap, density=d.density,
                            iz=iz, wi=wi, energies=energies, amu=amu, scatf=scatf, source=source, seed=cfg["seed"],
                            start=start, src_xy=srcxy, ret=ret, exit_state=exit_state, i_refl=irefl, d_travel=dtravel, weights=weights)
        vals, counts = np.unique(ret, return_counts=True)
        print("%-9s n=%d nE=%d codes=%s mean i_refl(exit)=%.2f -> %s (%d KB)" % (
            name, n, nE, dict(zip(vals.tolist(), counts.tolist())), irefl[ret == 1].mean() if (ret == 1).any() else 0,
            os.path.basename(path), os.path.getsize(path) // 1024))
        L.polycap_source_free(src)


if __name__ == "__main__":
    main()
