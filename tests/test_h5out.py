"""HDF5 result writer: files in the layout of the reference's bench scripts, checked by reading
them back with the independent minimal reader that parses the reference's own golden files
(tests/golden/h5mini.py) and by comparing the emitted encodings byte for byte with those libhdf5 wrote into
test/inchworm.h5 / test/topology_eval.h5 (tests/golden/h5_encodings.json, extracted by make_golden.py)."""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from h5mini import H5File  # noqa: E402

from qinchworm_b200 import h5out  # noqa: E402


def read_attrs(h, path):
    """Scalar attributes (message 0x000C, version 1) of the object at `path`."""
    out = {}
    for mtype, body, msize in h.messages(h.lookup(path)):
        if mtype != 0x000C:
            continue
        ver, nlen, tlen, slen = struct.unpack_from("<BxHHH", h.b, body)
        assert ver == 1
        p = body + 8
        name = h.b[p:p + nlen - 1].decode(); p += (nlen + 7) // 8 * 8
        cls = h.b[p] & 0x0F; p += (tlen + 7) // 8 * 8
        assert h.b[p + 1] == 0, "scalar dataspace"
        p += (slen + 7) // 8 * 8
        out[name] = struct.unpack_from("<d" if cls == 1 else "<q", h.b, p)[0]
    return out


def test_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tree = {"data": {"@attrs": {"beta": 10.0, "ntau": 200, "N_samples": 1024, "n_pts_after_max": np.iinfo(np.int64).max},
                     "tau": np.linspace(0, 10, 200), "gf": rng.normal(size=200) + 1j * rng.normal(size=200),
                     "orders": np.arange(5), "P_1": (rng.normal(size=(200, 2, 2)) * 1j)},
            "many": {"d%02d" % i: np.full(3, float(i)) for i in range(37)},     # five symbol nodes
            "empty": {}, "top": np.array([[1, 2, 3], [4, 5, 6]])}
    path = str(tmp_path / "out.h5")
    h5out.write_h5(path, tree)
    h = H5File(path)
    assert len(h.b) == struct.unpack_from("<Q", h.b, 40)[0]          # end-of-file address
    assert sorted(h.children(h.root)) == ["data", "empty", "many", "top"]
    assert h.children(h.lookup("/empty")) == {}
    for k in ("tau", "gf", "orders", "P_1"):
        got = h.read("/data/" + k)
        assert got.dtype == np.asarray(tree["data"][k]).dtype and np.array_equal(got, tree["data"][k])
    assert np.array_equal(h.read("/top"), tree["top"])
    assert sorted(h.children(h.lookup("/many"))) == sorted(tree["many"])
    for k, v in tree["many"].items():
        assert np.array_equal(h.read("/many/" + k), v)
    assert read_attrs(h, "/data") == {"beta": 10.0, "ntau": 200, "N_samples": 1024, "n_pts_after_max": np.iinfo(np.int64).max}


def test_encodings_match_libhdf5(tmp_path):
    enc = json.load(open(os.path.join(HERE, "golden", "h5_encodings.json")))
    path = str(tmp_path / "enc.h5")
    h5out.write_h5(path, {"a": {"c": np.zeros(enc["complex_dataset"]["shape"], dtype=complex)},
                          "r": np.zeros(enc["real_dataset"]["shape"])})
    h = H5File(path)
    for key, p in (("complex_dataset", "/a/c"), ("real_dataset", "/r")):
        mine = {"%04x" % t: h.b[b:b + n].hex() for t, b, n in h.messages(h.lookup(p))}
        ref = enc[key]["messages"]
        assert mine["0001"] == ref["0001"]                       # dataspace
        assert mine["0003"] == ref["0003"]                       # datatype (IEEE double / compound {r, i})
        assert mine["0005"] == ref["0005"]                       # fill value
        assert mine["0008"][:4] == ref["0008"][:4] and mine["0008"][20:] == ref["0008"][20:]   # layout: all but the address
    assert h.b[:32].hex() == enc["superblock_prefix"]
    assert h.b[h.root:h.root + 24].hex() == enc["root_header"]
    btree, heap = struct.unpack_from("<QQ", h.b, h.root + 24)
    assert h.b[heap:heap + 8].hex() == enc["heap_header_prefix"]
    assert h.b[btree:btree + 24].hex() == enc["btree_header"]
    snod = struct.unpack_from("<Q", h.b, btree + 32)[0]
    assert h.b[snod:snod + 6].hex() == enc["snod_header_prefix"]


def test_bench_layouts(tmp_path):
    """The three result layouts of the reference's bench scripts."""
    n_tau = 16
    tau = np.linspace(0, 8, n_tau)
    g = np.exp(-tau) * (1 + 0j)
    h5out.bethe_gf_results(str(tmp_path / "b.h5"), 10.0, n_tau, 2 ** 10, range(0, 4), range(0, 4), range(0, 3), tau, g, -g)
    h = H5File(str(tmp_path / "b.h5"))
    assert sorted(h.children(h.lookup("/data"))) == ["gf", "gf_ref", "orders", "orders_bare", "orders_gf", "tau"]
    assert np.array_equal(h.read("/data/orders_gf"), [0, 1, 2]) and np.array_equal(h.read("/data/gf"), g)
    dims = [1, 1, 1, 1, 2, 2, 2, 2, 4]
    P = np.arange(n_tau * 36).reshape(n_tau, 36) * 1j
    h5out.two_band_results(str(tmp_path / "t.h5"), 8.0, n_tau, 2 ** 10, range(0, 4), range(0, 4), range(0, 3), tau, [g] * 8, -g,
                           P, 2 * P, dims)
    h = H5File(str(tmp_path / "t.h5"))
    kids = h.children(h.lookup("/data"))
    assert len(kids) == 4 + 8 + 18 + 1 and "gf_dn_21" in kids and "Praw_9" in kids
    p9 = h.read("/data/P_9")
    assert p9.shape == (n_tau, 4, 4) and np.array_equal(p9.reshape(n_tau, 16), P[:, 20:])
    h5out.fh_dimer_results(str(tmp_path / "f.h5"), n_tau, 1e-3, range(0, 3), range(0, 3), [2 ** 15, 2 ** 16], [1e-4, 5e-5])
    h = H5File(str(tmp_path / "f.h5"))
    assert np.array_equal(h.read("/data/N_sampless"), [2 ** 15, 2 ** 16]) and read_attrs(h, "/data")["diff_0"] == 1e-3
