"""Derive CAL_MU / CAL_F in polycap_b200/csrc/optconst.c.

Compiles optconst.c with both calibration factors forced to 1, evaluates SiO2
(53/47 wt%, rho 2.23) at 10 keV and prints the factors that reproduce the
reference's golden (tests/photon.c:75-76): scatf 0.503696, amu 42.544635.
"""
import ctypes, os, subprocess, tempfile, sys
here = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(here, "..", "polycap_b200", "csrc", "optconst.c")
with tempfile.TemporaryDirectory() as td:
    so = os.path.join(td, "oc.so")
    txt = open(src).read()
    import re
    txt = re.sub(r"#define CAL_MU .*", "#define CAL_MU 1.0", txt)
    txt = re.sub(r"#define CAL_F .*", "#define CAL_F 1.0", txt)
    c = os.path.join(td, "oc.c"); open(c, "w").write(txt)
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-I", os.path.dirname(src), c, "-o", so, "-lm", "-ldl"])
    L = ctypes.CDLL(so)
    for f in ("pcb_builtin_cs_total", "pcb_builtin_fi"):
        getattr(L, f).restype = ctypes.c_double; getattr(L, f).argtypes = [ctypes.c_int, ctypes.c_double]
    L.pcb_builtin_atomic_weight.restype = ctypes.c_double; L.pcb_builtin_atomic_weight.argtypes = [ctypes.c_int]
    w = {8: 0.53, 14: 0.47}
    mu = sum(wi * L.pcb_builtin_cs_total(z, 10.0) for z, wi in w.items())
    print("CAL_MU", 42.544635 / (2.23 * mu))
    base = sum(wi * z / L.pcb_builtin_atomic_weight(z) for z, wi in w.items())
    fcorr = sum(wi * L.pcb_builtin_fi(z, 10.0) / L.pcb_builtin_atomic_weight(z) for z, wi in w.items())
    print("CAL_F", (0.503696 - base) / fcorr, " (f'_O, f'_Si uncalibrated:", L.pcb_builtin_fi(8, 10.0), L.pcb_builtin_fi(14, 10.0), ")")
