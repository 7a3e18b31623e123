"""BASELINE target run through the library's own multi-device path: 1e9 photon ids of the measured XOS profile (291 energies),
tally-only, from ONE process with POLYCAP_B200_DEVICES=all, via the stream API (include/polycap_b200.h).
usage: python tools/stream_1e9.py [n_ids]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("POLYCAP_B200_OPTCONST", "builtin")
os.environ.setdefault("POLYCAP_B200_QUIET", "1")
os.environ.setdefault("POLYCAP_B200_DEVICES", "all")
import bench  # noqa: E402
import polycap_b200 as pc  # noqa: E402

n_ids = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
w = bench.WORKLOADS["xos1"]
src = bench.build_source(pc, w)
st = pc.Stream(src, seed=20000)
st.next(10**12, max_ids=4_000_000 * pc.device_count(), records=False)  # buffers, module load
ids0, _, cnt0 = st.totals()
t0 = time.perf_counter()
st.next(10**12, max_ids=n_ids, records=False)
ids, sw, cnt = st.totals()
dt = time.perf_counter() - t0
started = (cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"]) - (cnt0["exit"] + cnt0["not_entered"] + cnt0["not_transmitted"])
print(json.dumps({"workload": w["desc"], "devices": os.environ["POLYCAP_B200_DEVICES"], "n_devices": pc.device_count(), "photon_ids": ids - ids0, "seconds": dt,
                  "photons_per_s": started / dt, "reflections_per_s": (cnt["refl_all"] - cnt0["refl_all"]) / dt,
                  "efficiency_first_mid_last": [float((sw / max(cnt["exit"] + cnt["not_entered"] + cnt["not_transmitted"], 1))[i]) for i in (0, 145, 290)],
                  "what": "polycap_b200_stream_next(max_ids=%d, tally only) from one process, wall clock incl. the reduction of the tallies over peer memory" % n_ids}))
