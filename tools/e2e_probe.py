"""Repeated calls of the reference-facing API on the bench's conical workload with wall-clock phases (tuning aid).
usage: POLYCAP_B200_PHASES=1 python tools/e2e_probe.py [n_calls] [n_exit]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import polycap_b200 as pc

n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
n_exit = int(sys.argv[2]) if len(sys.argv) > 2 else 225000
w = bench.WORKLOADS["conical"]
prof = pc.Profile(*w["profile"])
desc = pc.Description(prof, w["sig_rough"], w["n_cap"], bench.SIO2, 2.23)
src = pc.Source(desc, *w["source"], w["energies"])
os.environ["POLYCAP_B200_SEED"] = "777"
os.environ["POLYCAP_B200_REUPLOAD"] = "1"
with bench.quiet_stdout():
    src.get_transmission_efficiencies(-1, n_exit // 10)
    for i in range(n_calls):
        t0 = time.perf_counter()
        r = src.get_transmission_efficiencies(-1, n_exit)
        t1 = time.perf_counter()
        del r
        t2 = time.perf_counter()
        print("call %d: %.1f ms (+ %.1f ms freeing the result)" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), file=sys.stderr)
