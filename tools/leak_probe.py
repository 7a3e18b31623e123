"""Throughput of polycap_source_get_transmission_efficiencies(leak_calc=True) on the ellipsoidal example (tuning aid).
usage: python tools/leak_probe.py [n_exit] [n_energies]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import polycap_b200 as pc

n_exit = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
nE = int(sys.argv[2]) if len(sys.argv) > 2 else 7
prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
desc = pc.Description(prof, 0.0, 200000, bench.SIO2, 2.23)
src = pc.Source(desc, 2000.0, 0.2065, 0.2065, 0.0, 0.0, 0.0, 0.0, 0.5, np.linspace(1, 40, nE))
os.environ["POLYCAP_B200_SEED"] = "99"
with bench.quiet_stdout():
    for leak in ((False, True) if len(sys.argv) <= 3 else (True,)):
        t0 = time.perf_counter()
        r = src.get_transmission_efficiencies(-1, n_exit, leak)
        dt = time.perf_counter() - t0
        rec = r.records
        print("leak_calc=%s: %d exits of %d started in %.3f s -> %.0f started photons/s" % (leak, rec["i_exit"], rec["i_start"], dt, rec["i_start"] / dt), file=sys.stderr)
