"""Debug helper (GPU box): bitwise comparison of single-item device functions against the CPU oracle."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("POLYCAP_B200_QUIET", "1")
import polycap_b200 as pc
from oracle import harness as H
from conftest import load_golden

class V3(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]
lib = C.CDLL(pc.LIB_PATH)
lib.polycap_capil_segment.restype = C.c_int
lib.polycap_capil_segment.argtypes = [V3, V3, C.c_double, C.c_double, V3, V3, V3, C.POINTER(V3), C.POINTER(V3), C.c_void_p]
orc = H.Oracle()
g = load_golden("ellip")
z, cap, ext = g["z"], g["cap"], g["ext"]
rng = np.random.default_rng(1)
nhit = 0; ndiff = np.zeros(6, int); maxulp = np.zeros(6)
hexz = ext / (2 * 0.86602540378443864676 * (258 + 1))
for trial in range(3000):
    i = rng.integers(0, 998)
    A, B = (2 * 37 + 11) * 0.86602540378443864676, 11 * 1.5
    c0 = (A * hexz[i], B * hexz[i], z[i]); c1 = (A * hexz[i + 1], B * hexz[i + 1], z[i + 1])
    # photon just inside the wall at z[i], heading outwards at a grazing angle
    phi = rng.uniform(0, 2 * np.pi); rr = cap[i] * rng.uniform(0.5, 0.999)
    p0 = (c0[0] + rr * np.cos(phi), c0[1] + rr * np.sin(phi), z[i])
    th = rng.uniform(1e-4, 3e-3); psi = phi + rng.uniform(-1.0, 1.0)
    d = (np.sin(th) * np.cos(psi) + (c1[0] - c0[0]) / (z[i + 1] - z[i]), np.sin(th) * np.sin(psi) + (c1[1] - c0[1]) / (z[i + 1] - z[i]), np.cos(th))
    dz = z[i + 1] - z[i]
    p1 = (p0[0] + d[0] * dz / d[2], p0[1] + d[1] * dz / d[2], z[i + 1])
    last = (p0[0], p0[1], z[i] - 0.01)
    hg, ng = V3(*last), V3()
    rg = lib.polycap_capil_segment(V3(*c0), V3(*c1), cap[i], cap[i + 1], V3(*p0), V3(*p1), V3(*d), C.byref(hg), C.byref(ng), None)
    ho, no = H.Vec3(*last), H.Vec3()
    ro = orc.L.orc_capil_segment(H.vec3(*c0), H.vec3(*c1), cap[i], cap[i + 1], H.vec3(*p0), H.vec3(*p1), H.vec3(*d), C.byref(ho), C.byref(no))
    assert rg == ro, (rg, ro)
    if ro != 1:
        continue
    nhit += 1
    a = np.array([hg.x, hg.y, hg.z, ng.x, ng.y, ng.z]); b = np.array([ho.x, ho.y, ho.z, no.x, no.y, no.z])
    ulp = np.abs(a.view(np.int64) - b.view(np.int64))
    ndiff += ulp > 0; maxulp = np.maximum(maxulp, ulp)
print("segment: %d hits; components differing (hit xyz, normal xyz): %s max ulp %s" % (nhit, ndiff, maxulp))
