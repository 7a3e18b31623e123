"""polycap_source_get_transmission_efficiencies on one device and sharded over all devices by the library itself (POLYCAP_B200_DEVICES),
from ONE plain process: wall clock per call, photons/s.  usage: python tools/multi_device_probe.py [workload] [n_exit_per_device]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("POLYCAP_B200_OPTCONST", "builtin")
os.environ.setdefault("POLYCAP_B200_QUIET", "1")
os.environ["POLYCAP_B200_SEED"] = "777"
import bench  # noqa: E402  (workload table, quiet_stdout, result_counts)
import polycap_b200 as pc  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "conical"
w = bench.WORKLOADS[name]
n_dev = pc.device_count()
per_dev = int(sys.argv[2]) if len(sys.argv) > 2 else (250_000 if name == "conical" else 400_000)
out = {"workload": w["desc"], "devices_visible": n_dev, "runs": []}
for devices in [None] + (["all"] if n_dev > 1 else ["0,0"]):
    if devices is None:
        os.environ.pop("POLYCAP_B200_DEVICES", None)
        g = 1
    else:
        os.environ["POLYCAP_B200_DEVICES"] = devices
        g = n_dev if devices == "all" else len(devices.split(","))
    src = bench.build_source(pc, w)
    n_exit = per_dev * g
    with bench.quiet_stdout():
        warm = [src.get_transmission_efficiencies(-1, n_exit), src.get_transmission_efficiencies(-1, n_exit)]
        del warm
        times = []
        for _ in range(4):
            t0 = time.perf_counter()
            te = src.get_transmission_efficiencies(-1, n_exit)
            times.append(time.perf_counter() - t0)
    started, exits = bench.result_counts(te)
    cnt = pc.last_run_counters()
    out["runs"].append({"devices": devices or "one", "n_devices": g, "n_exit": n_exit, "simulated_photons": started, "seconds_per_call": [round(t, 4) for t in times],
                        "photons_per_s": started / min(times), "reflections_per_s": cnt["refl_all"] / min(times)})
    del te
out["speedup"] = out["runs"][1]["photons_per_s"] / out["runs"][0]["photons_per_s"]
print(json.dumps(out))
