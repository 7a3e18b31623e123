"""Debug helper (GPU box): replay the golden corpus through the C-ABI and list per-photon differences."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("POLYCAP_B200_QUIET", "1")
import polycap_b200 as pc
from conftest import load_golden

def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    d = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
    d[(a == b) | (np.isnan(a) & np.isnan(b))] = 0
    return d

for name in sys.argv[1:] or ["ellip", "conical", "parab", "xos1", "dubfoc"]:
    g = load_golden(name)
    prof = pc.Profile.new_from_arrays(g["ext"], g["cap"], g["z"])
    desc = pc.Description(prof, float(g["sig_rough"]), int(g["n_cap"]), {int(z): float(w) for z, w in zip(g["iz"], g["wi"])}, float(g["density"]))
    out = pc.launch_batch(desc, g["energies"], g["start"][:, 0:3], g["start"][:, 3:6], g["start"][:, 6:9], amu=g["amu"], scatf=g["scatf"])
    n = len(g["ret"])
    bad_ret = np.nonzero(out["ret"] != g["ret"])[0]
    bad_refl = np.nonzero(out["i_refl"] != g["i_refl"])[0]
    traced = np.isin(g["ret"], (0, 1))
    ew = rel(out["weights"], g["weights"]).max(axis=1) * traced
    ec = rel(out["exit_coords"], g["exit_state"][:, 0:3]).max(axis=1)
    ed = rel(out["exit_direction"], g["exit_state"][:, 3:6]).max(axis=1)
    ev = rel(out["exit_electric_vector"], g["exit_state"][:, 6:9]).max(axis=1)
    et = rel(out["d_travel"], g["d_travel"])
    print("== %s n=%d: ret mismatches %d, i_refl mismatches %d; max rel err: weights %.2e coords %.2e dir %.2e ev %.2e dtravel %.2e" % (
        name, n, len(bad_ret), len(bad_refl), ew.max(), ec.max(), ed.max(), ev.max(), et.max()))
    print("   photons beyond 1e-9: weights %d coords %d dir %d ev %d dtravel %d | counters %s" % (
        (ew > 1e-9).sum(), (ec > 1e-9).sum(), (ed > 1e-9).sum(), (ev > 1e-9).sum(), (et > 1e-9).sum(), out["counters"]))
    worst = np.argsort(-np.maximum.reduce([ew, ec, ed, ev, et]))[:6]
    for i in list(bad_ret[:5]) + list(bad_refl[:8]) + list(worst):
        print("   #%d ret %d/%d i_refl %d/%d  err w %.1e c %.1e d %.1e ev %.1e t %.1e  exit z %.6f/%.6f  w0 %.6g/%.6g  ev got %s want %s" % (
            i, out["ret"][i], g["ret"][i], out["i_refl"][i], g["i_refl"][i], ew[i], ec[i], ed[i], ev[i], et[i],
            out["exit_coords"][i, 2], g["exit_state"][i, 2], out["weights"][i, 0], g["weights"][i, 0],
            np.array2string(out["exit_electric_vector"][i], precision=6), np.array2string(g["exit_state"][i, 6:9], precision=6)))
