"""polycap_transmission_efficiencies_write_hdf5 (pcb_hdf5.cpp, no libhdf5): the file is re-read with an independent minimal
parser of the HDF5 File Format Specification (superblock v0, v1 object headers, old-style groups, contiguous layout) and
compared with what the reference writes into its file (src/polycap-transmission-efficiencies.c:345-766: names, shapes,
values, "Units" attributes).  No libhdf5 / h5py exists in this image, so this is a specification-level check."""
import ctypes as C
import struct

import numpy as np
import pytest


class H5:
    """just enough of the format to walk groups and read contiguous f64 datasets"""

    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n"
        assert self.b[8] == 0 and self.b[13] == 8 and self.b[14] == 8  # superblock v0, 8-byte offsets and lengths
        self.leaf_k, self.int_k = struct.unpack_from("<HH", self.b, 16)
        base, _, self.eof, _ = struct.unpack_from("<QQQQ", self.b, 24)
        assert base == 0 and self.eof == len(self.b)
        name_off, ohdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", self.b, 56)
        assert cache == 1
        self.root = (ohdr, btree, heap)

    def messages(self, addr):
        ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", self.b, addr)
        assert ver == 1 and refc == 1 and addr % 8 == 0
        p, out = addr + 16, []
        for _ in range(nmsg):
            t, s, flags = struct.unpack_from("<HHB", self.b, p)
            assert s % 8 == 0
            out.append((t, self.b[p + 8:p + 8 + s]))
            p += 8 + s
        assert p == addr + 16 + size
        return out

    def group(self, ohdr):
        (t, d), = [m for m in self.messages(ohdr) if m[0] == 0x11]
        btree, heap = struct.unpack("<QQ", d)
        assert self.b[heap:heap + 4] == b"HEAP"
        seg_size, free_off, seg = struct.unpack_from("<QQQ", self.b, heap + 8)
        assert self.b[btree:btree + 4] == b"TREE" and self.b[btree + 4] == 0 and self.b[btree + 5] == 0
        used, = struct.unpack_from("<H", self.b, btree + 6)
        members = {}
        p = btree + 24
        prev_key, = struct.unpack_from("<Q", self.b, p)
        for _ in range(used):
            child, key = struct.unpack_from("<QQ", self.b, p + 8)
            assert self.b[child:child + 4] == b"SNOD" and self.b[child + 4] == 1
            n, = struct.unpack_from("<H", self.b, child + 6)
            assert n <= 2 * self.leaf_k
            names = []
            for k in range(n):
                noff, oh, cache = struct.unpack_from("<QQI", self.b, child + 8 + 40 * k)
                end = self.b.index(b"\0", seg + noff)
                name = self.b[seg + noff:end].decode()
                names.append(name)
                members[name] = (oh, cache)
            assert names == sorted(names)
            kend = self.b.index(b"\0", seg + key)
            assert self.b[seg + key:kend].decode() == names[-1]  # right key = largest name of the child
            p += 16
        return members

    def dataset(self, ohdr):
        msgs = dict(self.messages(ohdr))
        sp = msgs[0x01]
        assert sp[0] == 1
        rank = sp[1]
        dims = struct.unpack_from("<%dQ" % rank, sp, 8)
        dt = msgs[0x03]
        assert dt[0] == 0x11 and dt[1] == 0x20 and dt[2] == 0x3f and struct.unpack_from("<I", dt, 4)[0] == 8
        assert struct.unpack_from("<HHBBBBI", dt, 8) == (0, 64, 52, 11, 0, 52, 1023)  # IEEE binary64 little endian
        lay = msgs[0x08]
        assert lay[0] == 3 and lay[1] == 1
        addr, size = struct.unpack_from("<QQ", lay, 2)
        assert size == 8 * int(np.prod(dims)) and addr % 8 == 0 and addr + size <= self.eof
        data = np.frombuffer(self.b, dtype="<f8", count=size // 8, offset=addr).reshape(dims)
        at = msgs[0x0C]
        assert at[0] == 1
        nsz, dtsz, spsz = struct.unpack_from("<HHH", at, 2)
        name = at[8:8 + nsz].rstrip(b"\0").decode()
        q = 8 + (nsz + 7) // 8 * 8
        assert at[q] == 0x13
        slen, = struct.unpack_from("<I", at, q + 4)
        q += (dtsz + 7) // 8 * 8 + (spsz + 7) // 8 * 8
        return data, {name: at[q:q + slen].decode()}


@pytest.fixture
def efficiencies(pc, harness):
    """a polycap_transmission_efficiencies object assembled on the host (struct layouts of src/polycap-private.h:147-181)"""
    prof = pc.Profile(pc.Profile.ELLIPSOIDAL, 9., 0.2065, 0.0585, 0.00035, 9.9153e-5, 1000., 0.5)
    desc = pc.Description(prof, 0.0, 200000, {8: 53.0, 14: 47.0}, 2.23)
    energies = np.array([5., 10., 20.])
    src = pc.Source(desc, 2000., 0.2065, 0.2065, 0., 0., 0., 0., 0.5, energies)
    n, nE, n_ext = 5, 3, 2
    rng = np.random.default_rng(0)
    keep = {}

    def arr(*shape):
        a = np.ascontiguousarray(rng.random(shape))
        keep[len(keep)] = a
        return a

    def ptr(a):
        return a.ctypes.data_as(harness.c_double_p)
    im = harness.RefImages()
    im.i_start, im.i_exit, im.i_extleak, im.i_intleak = 9, n, n_ext, 0
    for field, k in (("src_start_coords", 2), ("pc_start_coords", 2), ("pc_start_dir", 2), ("pc_start_elecv", 2), ("pc_exit_coords", 3),
                     ("pc_exit_dir", 2), ("pc_exit_elecv", 2)):
        rows = [arr(n) for _ in range(k)]
        keep[field] = rows
        setattr(im, field, (harness.c_double_p * k)(*[ptr(r) for r in rows]))
    nrefl = np.arange(n, dtype=np.int64) + 3
    keep["nrefl"] = nrefl
    im.pc_exit_nrefl = nrefl.ctypes.data_as(C.POINTER(C.c_int64))
    keep["dtravel"], keep["w"] = arr(n), arr(n, nE)
    im.pc_exit_dtravel, im.exit_coord_weights = ptr(keep["dtravel"]), ptr(keep["w"])
    keep["xc"] = [arr(n_ext) for _ in range(3)]
    keep["xd"] = [arr(n_ext) for _ in range(2)]
    im.extleak_coords = (harness.c_double_p * 3)(*[ptr(r) for r in keep["xc"]])
    im.extleak_dir = (harness.c_double_p * 2)(*[ptr(r) for r in keep["xd"]])
    keep["xw"] = arr(n_ext, nE)
    im.extleak_coord_weights = ptr(keep["xw"])
    xr = np.array([4, 7], dtype=np.int64)
    keep["xr"] = xr
    im.extleak_n_refl = xr.ctypes.data_as(C.POINTER(C.c_int64))
    eff = harness.RefEfficiencies()
    keep["eff"] = arr(nE)
    keep["en"] = energies
    eff.n_energies, eff.energies, eff.efficiencies = nE, ptr(energies), ptr(keep["eff"])
    eff.images = C.pointer(im)
    eff.source = C.cast(src._source, C.POINTER(harness.RefSource))
    keep["objs"] = (prof, desc, src, im)
    return eff, keep


def test_written_file_parses_and_matches(pc, efficiencies, tmp_path):
    eff, keep = efficiencies
    lib = C.CDLL(pc.LIB_PATH)
    lib.polycap_transmission_efficiencies_write_hdf5.restype = C.c_bool
    lib.polycap_transmission_efficiencies_write_hdf5.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    path = str(tmp_path / "out.h5")
    assert lib.polycap_transmission_efficiencies_write_hdf5(C.byref(eff), path.encode(), None)
    f = H5(path)
    root = f.group(f.root[0])
    assert set(root) == {"Energies", "Transmission_Efficiencies", "Source_Start_Coordinates", "PC_Start", "PC_Exit", "ExternalLeaks", "Input"}
    d, a = f.dataset(root["Energies"][0])
    assert np.array_equal(d, keep["en"]) and a == {"Units": "keV"}
    d, a = f.dataset(root["Transmission_Efficiencies"][0])
    assert np.array_equal(d, keep["eff"]) and a == {"Units": "a.u."}
    ex = f.group(root["PC_Exit"][0])
    assert set(ex) == {"Coordinates", "N_Reflections", "Direction", "Electric_Vector", "Weights", "D_Travel"}
    d, a = f.dataset(ex["Coordinates"][0])
    assert d.shape == (3, 5) and np.array_equal(d, np.stack(keep["pc_exit_coords"])) and a == {"Units": "[cm,cm,cm]"}
    d, _ = f.dataset(ex["Weights"][0])
    assert d.shape == (5, 3) and np.array_equal(d, keep["w"])
    d, _ = f.dataset(ex["N_Reflections"][0])
    assert np.array_equal(d, keep["nrefl"].astype(float))
    st = f.group(root["PC_Start"][0])
    d, _ = f.dataset(st["Electric_Vector"][0])
    assert np.array_equal(d, np.stack(keep["pc_start_elecv"]))
    lk = f.group(root["ExternalLeaks"][0])
    assert set(lk) == {"Coordinates", "Direction", "Weights", "Weight_Total", "N_Reflections"}
    d, _ = f.dataset(lk["Weight_Total"][0])
    assert np.allclose(d, keep["xw"].sum(axis=0) / 9)  # divided by i_start (transmission-efficiencies.c:572)
    inp = f.group(root["Input"][0])
    assert set(inp) == {"PC_Shape", "Cap_Shape", "N_Capillaries", "Surface_Roughness", "Open_Area", "PC_Composition", "PC_Density", "Src_PC_Dist"}
    d, a = f.dataset(inp["PC_Shape"][0])
    assert d.shape == (2, 999) and a == {"Units": "[cm,cm]"} and abs(d[1, 0] - 0.2065) < 1e-9
    d, _ = f.dataset(inp["N_Capillaries"][0])
    assert d[0] == 200000
    d, a = f.dataset(inp["PC_Composition"][0])
    assert d.shape == (2, 2) and list(d[0]) == [8, 14] and a == {"Units": "[Z,w%]"}


def test_write_errors(pc, efficiencies):
    eff, _ = efficiencies
    lib = C.CDLL(pc.LIB_PATH)
    lib.polycap_transmission_efficiencies_write_hdf5.restype = C.c_bool
    lib.polycap_transmission_efficiencies_write_hdf5.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]

    class Err(C.Structure):
        _fields_ = [("code", C.c_int), ("message", C.c_char_p)]
    err = C.POINTER(Err)()
    assert not lib.polycap_transmission_efficiencies_write_hdf5(C.byref(eff), None, C.byref(err))  # tests/source.c:290-292
    assert err.contents.code == 1
    err = C.POINTER(Err)()
    assert not lib.polycap_transmission_efficiencies_write_hdf5(C.byref(eff), b"/hoahohfhwofh/hohadohfowf.h5", C.byref(err))
    assert err.contents.code == 2  # POLYCAP_ERROR_IO, tests/source.c:294-296


def test_h5py_reads_the_file(pc, efficiencies, tmp_path):
    """the writer is restated from the format specification and had only been re-read by this file's own parser.
    Wherever a real HDF5 reader is importable (h5py; none in this image) every dataset and Units attribute is compared through it."""
    h5py = pytest.importorskip("h5py")
    eff, keep = efficiencies
    lib = C.CDLL(pc.LIB_PATH)
    lib.polycap_transmission_efficiencies_write_hdf5.restype = C.c_bool
    lib.polycap_transmission_efficiencies_write_hdf5.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
    path = str(tmp_path / "out.h5")
    assert lib.polycap_transmission_efficiencies_write_hdf5(C.byref(eff), path.encode(), None)
    own = H5(path)

    def walk(group_addr, prefix, out):
        for name, (oh, _) in own.group(group_addr).items():
            try:
                out[prefix + name] = own.dataset(oh)
            except Exception:
                walk(oh, prefix + name + "/", out)
    mine = {}
    walk(own.root[0], "/", mine)
    assert len(mine) > 20
    with h5py.File(path, "r") as f:
        for name, (data, attrs) in mine.items():
            ds = f[name]
            assert np.array_equal(np.asarray(ds, dtype=np.float64).reshape(data.shape), data), name
            for k, v in attrs.items():
                got = ds.attrs[k]
                assert (got.decode() if isinstance(got, bytes) else str(got)) == v, (name, k)
