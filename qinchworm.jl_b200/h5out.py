"""HDF5 result files in the layout of the reference's bench scripts, written without
libhdf5 (neither h5py nor the C library is in the image).

The reference's drivers store their results with HDF5.jl, e.g.
bench/bethe_gf_convergence/bethe_gf_convergence.jl:72-93 and
bench/two_band_eg_model_discrete_bath/two_band_eg_model_discrete_bath.jl:141-178: one group `data` with
scalar attributes (beta, ntau, N_samples, n_pts_after_max) and datasets (orders, tau, gf, gf_ref, P_s ...).
The plotting scripts next to them read exactly that.  `write_h5` emits the most conservative on-disk
format every HDF5 release reads — superblock version 0, version-1 object headers, old-style groups
(symbol-table message + v1 B-tree + local heap), contiguous little-endian datasets — which is also what
the reference's own golden files test/*.h5 use (checked byte by byte against test/inchworm.h5: same
superblock, object-header, heap, B-tree, symbol-node and datatype encodings, including the compound
{r, i} type HDF5.jl uses for ComplexF64).

    write_h5(path, {"data": {"@attrs": {"beta": 10.0, "ntau": 200}, "tau": tau, "gf": g}})

Arrays are written in C order; a Julia array of size (a, b, c) therefore corresponds to a numpy array of
shape (c, b, a), as in the reference's files (P_s: [n_tau, d, d]).  Supported element types: float64,
int64, complex128 (anything else is converted to one of the three)."""
from __future__ import annotations

import struct

import numpy as np

__all__ = ["write_h5", "bethe_gf_results", "two_band_results", "fh_dimer_results"]

UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16          # library defaults (symbol-node / B-tree fan-out), as in the reference's files

_F64 = bytes.fromhex("11203f0008000000" "00004000340b0034ff030000")     # IEEE double, little endian (datatype v1, class 1)
_I64 = bytes.fromhex("1008000008000000" "00004000")                     # signed 64-bit integer, little endian (class 0)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _complex_type() -> bytes:
    """Compound {r: f64 @0, i: f64 @8} of size 16, datatype version 1 (HDF5.jl's ComplexF64)."""
    out = struct.pack("<BBBBI", 0x16, 2, 0, 0, 16)
    for name, off in ((b"r", 0), (b"i", 8)):
        out += _pad8(name + b"\0") + struct.pack("<IB3xI4x16x", off, 0, 0) + _F64
    return out


def _dtype_of(a: np.ndarray):
    if np.iscomplexobj(a):
        return np.asarray(a, dtype="<c16"), _complex_type()
    if a.dtype.kind in "iub":
        return np.asarray(a, dtype="<i8"), _I64
    return np.asarray(a, dtype="<f8"), _F64


def _dataspace(shape) -> bytes:
    if len(shape) == 0:
        return struct.pack("<BBB5x", 1, 0, 0)                        # scalar
    dims = struct.pack("<%dQ" % len(shape), *shape)
    return struct.pack("<BBB5x", 1, len(shape), 1) + dims + dims     # max dims = dims


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages) -> bytes:
    data = b"".join(messages)
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(data)) + data


def _attribute(name: str, value) -> bytes:
    arr, dt = _dtype_of(np.asarray(value))
    nm = name.encode() + b"\0"
    sp = _dataspace(arr.shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + arr.tobytes()
    return _message(0x000C, body)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)      # the superblock is filled in last

    def put(self, b: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += b
        return addr

    def dataset(self, value, attrs=None) -> int:
        arr, dt = _dtype_of(np.asarray(value))
        raw = arr.tobytes(order="C")
        addr = self.put(raw) if raw else UNDEF
        msgs = [_message(0x0001, _dataspace(arr.shape)),
                _message(0x0003, dt, flags=1),                                   # constant message
                _message(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),          # fill value v2: late alloc, written if set, size 0
                _message(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw)))]    # contiguous layout v3
        msgs += [_attribute(k, v) for k, v in (attrs or {}).items()]
        return self.put(_object_header(msgs))

    def group(self, tree: dict):
        """Old-style group: children first, then local heap, symbol nodes, B-tree, object header.
        Returns (header address, B-tree address, heap address) — the latter two are cached in the parent's entry."""
        attrs = tree.get("@attrs", {})
        entries = []
        for name in sorted((k for k in tree if k != "@attrs"), key=lambda s: s.encode()):
            v = tree[name]
            if isinstance(v, dict):
                entries.append((name, *self.group(v)))
            else:
                entries.append((name, self.dataset(v), None, None))
        n_nodes = (len(entries) + 2 * LEAF_K - 1) // (2 * LEAF_K)
        if n_nodes > 2 * INTERNAL_K:
            raise ValueError("more than %d entries in one group" % (4 * LEAF_K * INTERNAL_K))
        # local heap: empty string at offset 0, then the names, each padded to 8 bytes
        seg = bytearray(8)
        name_off = []
        for name, *_ in entries:
            name_off.append(len(seg))
            seg += _pad8(name.encode() + b"\0")
        seg_addr = self.put(bytes(seg))
        heap = self.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(seg), 1, seg_addr))   # free list head 1 = none
        # symbol nodes of up to 2 * LEAF_K entries, in name order
        keys, children = [0], []
        for i in range(n_nodes):
            chunk = range(i * 2 * LEAF_K, min(len(entries), (i + 1) * 2 * LEAF_K))
            node = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for j in chunk:
                _, hdr, bt, hp = entries[j]
                if bt is None:
                    node += struct.pack("<QQI4x16x", name_off[j], hdr, 0)
                else:
                    node += struct.pack("<QQI4xQQ", name_off[j], hdr, 1, bt, hp)
            node += b"\0" * (8 + 2 * LEAF_K * 40 - len(node))
            children.append(self.put(node))
            keys.append(name_off[chunk[-1]])
        tree_node = b"TREE" + struct.pack("<BBHQQ", 0, 0, n_nodes, UNDEF, UNDEF) + struct.pack("<Q", keys[0])
        for c, k in zip(children, keys[1:]):
            tree_node += struct.pack("<QQ", c, k)
        tree_node += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(tree_node))
        bt = self.put(tree_node)
        msgs = [_message(0x0011, struct.pack("<QQ", bt, heap))] + [_attribute(k, v) for k, v in attrs.items()]
        return self.put(_object_header(msgs)), bt, heap

    def finish(self, root) -> bytes:
        hdr, bt, heap = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, hdr, 1, bt, heap)              # root group symbol-table entry
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path, tree: dict) -> None:
    """Write the nested dict `tree` (dict = group, "@attrs" = its scalar attributes, anything else = dataset)."""
    w = _Writer()
    root = w.group(tree)
    data = w.finish(root)
    with open(path, "wb") as f:
        f.write(data)


# ---- the reference bench scripts' layouts --------------------------------------------------------------

def _common(beta, n_tau, N_samples, n_pts_after_max, orders, orders_bare, orders_gf, tau):
    grp = {"@attrs": {"beta": float(beta), "ntau": int(n_tau), "N_samples": int(N_samples),
                      "n_pts_after_max": int(n_pts_after_max if n_pts_after_max is not None else np.iinfo(np.int64).max)},
           "orders": np.asarray(list(orders), dtype=np.int64), "orders_bare": np.asarray(list(orders_bare), dtype=np.int64),
           "tau": np.asarray(tau, dtype=float)}
    if orders_gf is not None:
        grp["orders_gf"] = np.asarray(list(orders_gf), dtype=np.int64)
    return grp


def bethe_gf_results(path, beta, n_tau, N_samples, orders, orders_bare, orders_gf, tau, gf, gf_ref, n_pts_after_max=None):
    """bench/bethe_gf_convergence/bethe_gf_convergence.jl:72-93."""
    grp = _common(beta, n_tau, N_samples, n_pts_after_max, orders, orders_bare, orders_gf, tau)
    grp["gf"] = np.asarray(gf, dtype=complex)
    grp["gf_ref"] = np.asarray(gf_ref, dtype=complex)
    write_h5(path, {"data": grp})


def two_band_results(path, beta, n_tau, N_samples, orders, orders_bare, orders_gf, tau, gfs, gf_ref, P, P_raw, dims,
                     n_pts_after_max=None):
    """bench/two_band_eg_model_discrete_bath/two_band_eg_model_discrete_bath.jl:141-178.  gfs: the 8 correlators
    in the order (up 11, dn 11, up 22, dn 22, up 12, dn 12, up 21, dn 21); P, P_raw: packed [n_tau, sum d^2] tables."""
    grp = _common(beta, n_tau, N_samples, n_pts_after_max, orders, orders_bare, orders_gf, tau)
    for name, g in zip(("gf_up_11", "gf_dn_11", "gf_up_22", "gf_dn_22", "gf_up_12", "gf_dn_12", "gf_up_21", "gf_dn_21"), gfs):
        grp[name] = np.asarray(g, dtype=complex)
    off = 0
    for s, d in enumerate(dims, start=1):     # P_s: Julia (d, d, n_tau) column-major = C-order [n_tau, d, d]
        grp["P_%d" % s] = np.asarray(P)[:, off:off + d * d].reshape(-1, d, d)
        grp["Praw_%d" % s] = np.asarray(P_raw)[:, off:off + d * d].reshape(-1, d, d)
        off += d * d
    grp["gf_ref"] = np.asarray(gf_ref, dtype=complex)
    write_h5(path, {"data": grp})


def fh_dimer_results(path, n_tau, diff_0, orders, orders_bare, N_sampless, diffs):
    """bench/fermi_hubbard_dimer/benchmark_fh_dimer.jl:111-126."""
    write_h5(path, {"data": {"@attrs": {"ntau": int(n_tau), "diff_0": float(diff_0)},
                             "orders": np.asarray(list(orders), dtype=np.int64),
                             "orders_bare": np.asarray(list(orders_bare), dtype=np.int64),
                             "N_sampless": np.asarray(list(N_sampless), dtype=np.int64),
                             "diffs": np.asarray(diffs, dtype=float)}})
