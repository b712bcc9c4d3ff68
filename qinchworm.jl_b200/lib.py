"""ctypes binding of libqinchworm_cuda.so — the same symbols a Julia `ccall` shim binds
(include/qinchworm.h).  No CPU fallback: if the library is missing, importing the
compute entry points raises; if no CUDA device is present, `Context()` raises."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QIW_LIB", os.path.join(HERE, "libqinchworm_cuda.so"))
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "qinchworm.h")

MODE_BARE, MODE_BOLD, MODE_CORR = 0, 1, 2
DEVICE_CURRENT, DEVICE_NONE = -1, -2
UNIQUE_ID_BYTES = 128
PEER_HANDLE_BYTES = 64


class QiwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libqinchworm_cuda error %d: %s" % (code, msg))
        self.code = code


def build(force=False, verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(HERE, "csrc")
    cmd = ["make", "-C", csrc, "-j%d" % min(4, os.cpu_count() or 1)] + (["-B"] if force else [])
    out = subprocess.run(cmd, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libqinchworm_cuda.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


class _Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("warps_per_block", C.c_int32), ("reserved", C.c_int32 * 6)]


_lib = None

i32p, i64p, u32p, u8p, f64p = (C.POINTER(t) for t in (C.c_int32, C.c_int64, C.c_uint32, C.c_uint8, C.c_double))

_SIGNATURES = {
    "qiw_create": (C.c_int, [C.POINTER(_Opts), C.POINTER(C.c_void_p)]),
    "qiw_destroy": (C.c_int, [C.c_void_p]),
    "qiw_last_error": (C.c_char_p, [C.c_void_p]),
    "qiw_version": (C.c_char_p, []),
    "qiw_set_model": (C.c_int, [C.c_void_p, C.c_int32, i32p, f64p, C.c_int32, i32p, i64p, f64p, C.c_int32,
                                i32p, i32p, i32p, C.c_int32, i32p, i32p]),
    "qiw_set_grid": (C.c_int, [C.c_void_p, C.c_int32, C.c_double]),
    "qiw_set_delta": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, f64p]),
    "qiw_set_P": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, f64p]),
    "qiw_scale_P": (C.c_int, [C.c_void_p, C.c_int32, f64p, C.c_double]),
    "qiw_get_P": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, f64p]),
    "qiw_set_topologies": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, i32p, i32p]),
    "qiw_entry_stats": (C.c_int, [C.c_void_p, C.c_int32, i64p, i64p, i64p, f64p]),
    "qiw_entry_program": (C.c_int, [C.c_void_p, C.c_int32, i64p, C.POINTER(C.c_uint64), i64p, u32p, i64p, f64p,
                                    i64p, i32p, i32p, i32p]),
    "qiw_entry_records": (C.c_int, [C.c_void_p, C.c_int32, i32p, u32p, C.POINTER(C.c_uint16)]),
    "qiw_entry_lane_program": (C.c_int, [C.c_void_p, C.c_int32, i32p, i32p, u32p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]),
    "qiw_entry_walk_units": (C.c_int, [C.c_void_p, C.c_int32, i64p, i64p, u32p, u32p]),
    "qiw_eval": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32, i32p, u32p, u32p,
                           C.c_uint64, f64p]),
    "qiw_eval_batch": (C.c_int, [C.c_void_p, C.c_int32, f64p, C.c_int32, i32p, u32p, u32p, C.c_uint64, f64p]),
    "qiw_eval_seqs": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_int32, i32p, u32p, u32p,
                                C.c_uint64, f64p]),
    "qiw_eval_range": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int32, i32p, u32p, u32p,
                                 C.c_uint64, C.c_uint64, C.c_uint64, f64p]),
    "qiw_eval_at_times": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                    f64p, f64p]),
    "qiw_last_device_ms": (C.c_int, [C.c_void_p, f64p]),
    "qiw_launch_count": (C.c_int, [C.c_void_p, i64p]),
    "qiw_profile_enable": (C.c_int, [C.c_void_p, C.c_int32]),
    "qiw_profile_read": (C.c_int, [C.c_void_p, f64p, i64p, C.c_int32]),
    "qiw_inchworm_run": (C.c_int, [C.c_void_p, C.c_int32, i32p, C.c_int32, i32p, u32p, u32p, C.c_uint64, f64p]),
    "qiw_sobol_direction_numbers": (C.c_int, [C.c_int32, u32p]),
    "qiw_sobol_scramble": (C.c_int, [C.c_int32, u32p, u32p, u8p, u8p]),
    "qiw_sobol_points": (C.c_int, [C.c_void_p, C.c_int32, u32p, u32p, C.c_uint64, C.c_uint64, u32p]),
    "qiw_topologies": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, i64p, i32p, i32p]),
    "qiw_rank_sub_range": (C.c_int, [C.c_uint64, C.c_int32, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "qiw_comm_unique_id": (C.c_int, [u8p]),
    "qiw_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, u8p]),
    "qiw_comm_destroy": (C.c_int, [C.c_void_p]),
    "qiw_peer_handle": (C.c_int, [C.c_void_p, u8p]),
    "qiw_peer_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, u8p]),
    "qiw_measure_fp64_peak": (C.c_int, [C.c_void_p, f64p]),
}


def load():
    """Load the shared library (raises if it has not been built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libqinchworm_cuda.so is not built (run `python -c 'import __graft_entry__ as g; "
                              "g.build()'` or `make -C qinchworm.jl_b200/csrc`); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _cview(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.view(np.float64)


# ---- host-only helpers (no GPU needed) ----------------------------------------------------------

def sobol_direction_numbers(D):
    m = np.zeros((max(D, 1), 32), dtype=np.uint32)
    if load().qiw_sobol_direction_numbers(D, _ptr(m, u32p)):
        raise ValueError("Invalid Sobol dimension %d" % D)
    return m[:D]


def sobol_scramble(m, shift_bits, ltm_bits):
    D = m.shape[0]
    m = np.ascontiguousarray(m, dtype=np.uint32).copy()
    x0 = np.zeros(max(D, 1), dtype=np.uint32)
    sb = np.asfortranarray(shift_bits, dtype=np.uint8)
    lb = np.asfortranarray(ltm_bits, dtype=np.uint8)
    rc = load().qiw_sobol_scramble(D, _ptr(m, u32p), _ptr(x0, u32p), sb.ctypes.data_as(u8p), lb.ctypes.data_as(u8p))
    if rc:
        raise QiwError(rc, "qiw_sobol_scramble")
    return m, x0[:D]


def topologies(order, k=None, with_external_arc=False):
    """get_topologies_at_order(order, k; with_external_arc): (pairs[n, order, 2], parity[n])."""
    kk = -1 if k is None else int(k)
    n = C.c_int64(0)
    L = load()
    rc = L.qiw_topologies(order, kk, int(with_external_arc), C.byref(n), None, None)
    if rc:
        raise QiwError(rc, "qiw_topologies")
    pairs = np.zeros((n.value, order, 2), dtype=np.int32)
    parity = np.zeros(n.value, dtype=np.int32)
    if n.value:
        L.qiw_topologies(order, kk, int(with_external_arc), C.byref(n), _ptr(pairs, i32p), _ptr(parity, i32p))
    return pairs, parity


def rank_sub_range(N, n_ranks, rank):
    s, c = C.c_uint64(0), C.c_uint64(0)
    rc = load().qiw_rank_sub_range(N, n_ranks, rank, C.byref(s), C.byref(c))
    if rc:
        raise QiwError(rc, "qiw_rank_sub_range")
    return s.value, c.value


def comm_unique_id():
    buf = np.zeros(UNIQUE_ID_BYTES, dtype=np.uint8)
    rc = load().qiw_comm_unique_id(_ptr(buf, u8p))
    if rc:
        raise QiwError(rc, "qiw_comm_unique_id")
    return buf


# ---- device context ---------------------------------------------------------------------------------

class Context:
    """One context per process / GPU (qiw_create ... qiw_destroy)."""

    def __init__(self, device=-1, warps_per_block=0):
        self.L = load()
        opts = _Opts(device, warps_per_block, (C.c_int32 * 6)())
        h = C.c_void_p()
        rc = self.L.qiw_create(C.byref(opts), C.byref(h))
        if rc:
            raise QiwError(rc, "qiw_create failed (no CUDA device? this library has no CPU path)")
        self.h = h
        self.device = device
        self.bsize = 0
        self.entry_order = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.qiw_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise QiwError(rc, self.L.qiw_last_error(self.h).decode())

    # -- problem definition
    def set_expansion(self, expansion):
        """Upload model, Delta tables, grid and the current P table of an `Expansion`."""
        pl = expansion.flatten()
        self.set_model(pl)
        if self.device == DEVICE_NONE:
            self.set_grid(pl["n_tau"], pl["beta"])
            return pl
        for t, (kind, data) in enumerate(pl["tables"]):
            self.set_delta(t, kind, data, pl["beta"])
        self.set_grid(pl["n_tau"], pl["beta"])
        self.set_P(0, expansion.P)
        return pl

    def set_model(self, pl):
        pool, poolv = _cview(pl["op_pool"])
        if poolv.size == 0:
            poolv = np.zeros(2)
        arrs = dict(
            dims=np.ascontiguousarray(pl["dims"], dtype=np.int32),
            en=np.ascontiguousarray(pl["energies"], dtype=np.float64),
            tgt=np.ascontiguousarray(pl["op_target"], dtype=np.int32),
            off=np.ascontiguousarray(pl["op_mat_off"], dtype=np.int64),
            pi=np.ascontiguousarray(pl["pair_op_i"], dtype=np.int32),
            pf=np.ascontiguousarray(pl["pair_op_f"], dtype=np.int32),
            pt=np.ascontiguousarray(pl["pair_table"], dtype=np.int32),
            ca=np.ascontiguousarray(np.concatenate([pl["corr_A"], [0]]), dtype=np.int32),
            cb=np.ascontiguousarray(np.concatenate([pl["corr_B"], [0]]), dtype=np.int32))
        self._ck(self.L.qiw_set_model(self.h, pl["S"], _ptr(arrs["dims"], i32p), _ptr(arrs["en"], f64p), pl["n_ops"],
                                      _ptr(arrs["tgt"], i32p), _ptr(arrs["off"], i64p), _ptr(poolv, f64p),
                                      pl["n_pairs"], _ptr(arrs["pi"], i32p), _ptr(arrs["pf"], i32p),
                                      _ptr(arrs["pt"], i32p), pl["n_corr"], _ptr(arrs["ca"], i32p),
                                      _ptr(arrs["cb"], i32p)))
        self.bsize = int(sum(int(d) ** 2 for d in pl["dims"]))
        self.S = int(pl["S"])

    def set_grid(self, n_tau, beta):
        self._ck(self.L.qiw_set_grid(self.h, n_tau, beta))
        self.n_tau = n_tau

    def set_delta(self, table_id, kind, data, beta):
        d, dv = _cview(data)
        self._ck(self.L.qiw_set_delta(self.h, table_id, kind, len(d), beta, _ptr(dv, f64p)))

    def set_P(self, first, rows):
        r, rv = _cview(np.atleast_2d(rows))
        self._ck(self.L.qiw_set_P(self.h, first, r.shape[0], _ptr(rv, f64p)))

    def scale_P(self, k_f, row=None, lam=0.0):
        """qiw_scale_P: set_ppgf! + normalize! across the step seam — row k_f := row (None keeps it), then every stored
        row k times exp(-lam tau_k); one row and lambda travel instead of the whole table."""
        if row is None:
            self._ck(self.L.qiw_scale_P(self.h, k_f, None, float(lam)))
        else:
            r, rv = _cview(np.atleast_2d(row))
            self._ck(self.L.qiw_scale_P(self.h, k_f, _ptr(rv, f64p), float(lam)))

    def scale_P_prepared(self, k_f, row_ptr, lam):
        """qiw_scale_P with a long-lived row buffer."""
        rc = self.L.qiw_scale_P(self.h, k_f, row_ptr, lam)
        if rc:
            self._ck(rc)

    def get_P(self, first=0, count=None):
        count = self.n_tau - first if count is None else count
        out = np.zeros((count, self.bsize), dtype=np.complex128)
        self._ck(self.L.qiw_get_P(self.h, first, count, _ptr(out.view(np.float64), f64p)))
        return out

    def set_topologies(self, entry_id, mode, order, n_pts_after, pairs, parity, corr_idx=0):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        parity = np.ascontiguousarray(parity, dtype=np.int32)
        self._ck(self.L.qiw_set_topologies(self.h, entry_id, mode, order, n_pts_after, corr_idx, len(parity),
                                           _ptr(pairs, i32p), _ptr(parity, i32p)))
        self.entry_order[entry_id] = order

    def entry_stats(self, entry_id):
        a, b, c, f = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_double(0)
        self._ck(self.L.qiw_entry_stats(self.h, entry_id, C.byref(a), C.byref(b), C.byref(c), C.byref(f)))
        return dict(n_top=a.value, n_leaves=b.value, n_edges=c.value, flops_per_sample=f.value)

    def entry_program(self, entry_id):
        """Disassembly of a compiled entry (qiw_entry_program)."""
        nw, nt, nc, nd = (C.c_int64(0) for _ in range(4))
        self._ck(self.L.qiw_entry_program(self.h, entry_id, C.byref(nw), None, C.byref(nt), None, C.byref(nc),
                                          None, C.byref(nd), None, None, None))
        words = np.zeros(nw.value, dtype=np.uint64)
        tree_off = np.zeros(nt.value + 1, dtype=np.uint32)
        coefs = np.zeros(max(nc.value, 1), dtype=np.complex128)
        dslots = np.zeros((max(nd.value, 1), 3), dtype=np.int32)
        pos_src = np.zeros(20, dtype=np.int32)
        info = np.zeros(4, dtype=np.int32)
        self._ck(self.L.qiw_entry_program(self.h, entry_id, C.byref(nw), words.ctypes.data_as(C.POINTER(C.c_uint64)),
                                          C.byref(nt), _ptr(tree_off, u32p), C.byref(nc),
                                          _ptr(coefs.view(np.float64), f64p), C.byref(nd), _ptr(dslots, i32p),
                                          _ptr(pos_src, i32p), _ptr(info, i32p)))
        return dict(words=words, tree_off=tree_off, coefs=coefs[:nc.value], dslots=dslots[:nd.value],
                    pos_src=pos_src, n_nodes=int(info[0]), nP=int(info[1]), S=int(info[2]), scalar=bool(info[3]))

    def entry_records(self, entry_id):
        """Factorised configuration records of a compiled entry (qiw_entry_records)."""
        info = np.zeros(8, dtype=np.int32)
        self._ck(self.L.qiw_entry_records(self.h, entry_id, _ptr(info, i32p), None, None))
        K, L2, nl, nseg, stride, nP, nD = (int(x) for x in info[:7])
        rec2 = np.zeros((max(nl, 1), L2 + 1), dtype=np.uint32)
        segdef = np.zeros((max(nseg, 1), stride), dtype=np.uint16)
        self._ck(self.L.qiw_entry_records(self.h, entry_id, _ptr(info, i32p), _ptr(rec2, u32p),
                                          segdef.ctypes.data_as(C.POINTER(C.c_uint16))))
        return dict(K=K, L2=L2, n_leaves=nl, nSeg=nseg, seg_stride=stride, nP=nP, nD=nD, rec2=rec2[:nl],
                    segdef=segdef[:nseg])

    def entry_lane_program(self, entry_id):
        """Lane program of a compiled entry (qiw_entry_lane_program): what the step kernel executes."""
        info = np.zeros(8, dtype=np.int32)
        self._ck(self.L.qiw_entry_lane_program(self.h, entry_id, _ptr(info, i32p), None, None, None, None))
        nsec, nit, nseg, stride, K, order, seg0, cost = (int(x) for x in info)
        sections = np.zeros((max(nsec, 1), 4), dtype=np.int32)
        items = np.zeros(max(nit, 1), dtype=np.uint32)
        segdef = np.zeros((max(nseg, 1), stride), dtype=np.uint16)
        seg_coef = np.zeros(max(nseg, 1), dtype=np.uint16)
        u16p = C.POINTER(C.c_uint16)
        self._ck(self.L.qiw_entry_lane_program(self.h, entry_id, _ptr(info, i32p), _ptr(sections, i32p), _ptr(items, u32p),
                                               segdef.ctypes.data_as(u16p), seg_coef.ctypes.data_as(u16p)))
        return dict(sections=sections[:nsec], items=items[:nit], segdef=segdef[:nseg], seg_coef=seg_coef[:nseg],
                    K=K, order=order, seg0=seg0, cost=cost, seg_stride=stride)

    def entry_walk_units(self, entry_id):
        """Walk units of a compiled entry of a sector-block model (qiw_entry_walk_units)."""
        nu, nw = C.c_int64(0), C.c_int64(0)
        self._ck(self.L.qiw_entry_walk_units(self.h, entry_id, C.byref(nu), C.byref(nw), None, None))
        off = np.zeros(nu.value + 1, dtype=np.uint32)
        words = np.zeros((max(nw.value, 1), 4), dtype=np.uint32)
        self._ck(self.L.qiw_entry_walk_units(self.h, entry_id, C.byref(nu), C.byref(nw), _ptr(off, u32p), _ptr(words, u32p)))
        return dict(unit_off=off, words=words[:nw.value])

    # -- hot path
    def _sobol_args(self, ids, sobol):
        if sobol is None:
            return None, None, None, None
        ms = [np.asarray(m, dtype=np.uint32).reshape(-1) for m, _ in sobol] + [np.zeros(1, np.uint32)]
        xs = [np.asarray(x, dtype=np.uint32).reshape(-1) for _, x in sobol] + [np.zeros(1, np.uint32)]
        mcat, xcat = np.ascontiguousarray(np.concatenate(ms)), np.ascontiguousarray(np.concatenate(xs))
        return mcat, xcat, _ptr(mcat, u32p), _ptr(xcat, u32p)

    def eval(self, t_i, t_w, t_f, entry_ids, N_total, sobol=None):
        ids = np.ascontiguousarray(entry_ids, dtype=np.int32)
        out = np.zeros((len(ids), self.bsize), dtype=np.complex128)
        _m, _x, pm, px = self._sobol_args(ids, sobol)
        self._ck(self.L.qiw_eval(self.h, t_i, t_w, t_f, len(ids), _ptr(ids, i32p), pm, px, N_total,
                                 _ptr(out.view(np.float64), f64p)))
        return out

    def eval_prepared(self, t_i, t_w, t_f, n, ids_ptr, N_total, out_ptr):
        """qiw_eval with the caller's long-lived id array and result buffer (ctypes pointers made once): the host-stepped
        loop calls this once per time step."""
        rc = self.L.qiw_eval(self.h, t_i, t_w, t_f, n, ids_ptr, None, None, N_total, out_ptr)
        if rc:
            self._ck(rc)

    def eval_batch(self, times, entry_ids, N_total, sobol=None):
        """qiw_eval_batch: times[n_times, 3] = (t_i, t_w, t_f); returns [n_times, n_entries, bsize]."""
        ids = np.ascontiguousarray(entry_ids, dtype=np.int32)
        times = np.ascontiguousarray(times, dtype=np.float64).reshape(-1, 3)
        out = np.zeros((times.shape[0], len(ids), self.bsize), dtype=np.complex128)
        _m, _x, pm, px = self._sobol_args(ids, sobol)
        self._ck(self.L.qiw_eval_batch(self.h, times.shape[0], _ptr(times, f64p), len(ids), _ptr(ids, i32p), pm, px,
                                       N_total, _ptr(out.view(np.float64), f64p)))
        return out

    def eval_seqs(self, t_i, t_w, t_f, entry_ids, N_total, sobol_seqs):
        """qiw_eval_seqs: sobol_seqs[n_seqs][n_entries] = (m, x0); returns [n_seqs, n_entries, bsize]."""
        ids = np.ascontiguousarray(entry_ids, dtype=np.int32)
        ms, xs = [], []
        for seq in sobol_seqs:
            ms += [np.asarray(m, dtype=np.uint32).reshape(-1) for m, _ in seq]
            xs += [np.asarray(x, dtype=np.uint32).reshape(-1) for _, x in seq]
        mcat = np.ascontiguousarray(np.concatenate(ms + [np.zeros(1, np.uint32)]))
        xcat = np.ascontiguousarray(np.concatenate(xs + [np.zeros(1, np.uint32)]))
        out = np.zeros((len(sobol_seqs), len(ids), self.bsize), dtype=np.complex128)
        self._ck(self.L.qiw_eval_seqs(self.h, t_i, t_w, t_f, len(sobol_seqs), len(ids), _ptr(ids, i32p), _ptr(mcat, u32p),
                                      _ptr(xcat, u32p), N_total, _ptr(out.view(np.float64), f64p)))
        return out

    def eval_range(self, t_i, t_w, t_f, entry_ids, N_total, start, count, sobol=None):
        ids = np.ascontiguousarray(entry_ids, dtype=np.int32)
        out = np.zeros((len(ids), self.bsize), dtype=np.complex128)
        _m, _x, pm, px = self._sobol_args(ids, sobol)
        self._ck(self.L.qiw_eval_range(self.h, t_i, t_w, t_f, len(ids), _ptr(ids, i32p), pm, px, start, count,
                                       N_total, _ptr(out.view(np.float64), f64p)))
        return out

    def eval_at_times(self, entry_id, t_i, t_w, t_f, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.zeros((times.shape[0], self.bsize), dtype=np.complex128)
        self._ck(self.L.qiw_eval_at_times(self.h, entry_id, t_i, t_w, t_f, times.shape[0], _ptr(times, f64p),
                                          _ptr(out.view(np.float64), f64p)))
        return out

    def inchworm_run(self, bare_ids, bold_ids, N_total, sobol=None, want_contribs=True):
        """qiw_inchworm_run: the whole inchworm! loop on the device.  Returns the per-entry
        contributions [n_tau, n_bare + n_bold, bsize] (or None); read P with get_P()."""
        b = np.ascontiguousarray(bare_ids, dtype=np.int32)
        d = np.ascontiguousarray(bold_ids, dtype=np.int32)
        _m, _x, pm, px = self._sobol_args(None, sobol)
        hist = np.zeros((self.n_tau, len(b) + len(d), self.bsize), dtype=np.complex128) if want_contribs else None
        self._ck(self.L.qiw_inchworm_run(self.h, len(b), _ptr(b, i32p), len(d), _ptr(d, i32p) if len(d) else None,
                                         pm, px, N_total,
                                         _ptr(hist.view(np.float64), f64p) if want_contribs else None))
        return hist

    PROFILE_CLASSES = ("step_complex", "step_real", "run_kernel", "block_mma", "reduce", "finish_step",
                       "nccl_allreduce", "step_block")

    def profile_enable(self, on=True):
        self._ck(self.L.qiw_profile_enable(self.h, int(on)))

    def profile_read(self, reset=True):
        ms = np.zeros(8, dtype=np.float64)
        n = np.zeros(8, dtype=np.int64)
        self._ck(self.L.qiw_profile_read(self.h, _ptr(ms, f64p), _ptr(n, i64p), int(reset)))
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(self.PROFILE_CLASSES) if n[i]}

    def last_device_ms(self):
        v = C.c_double(0)
        self._ck(self.L.qiw_last_device_ms(self.h, C.byref(v)))
        return v.value

    def launch_count(self):
        v = C.c_int64(0)
        self._ck(self.L.qiw_launch_count(self.h, C.byref(v)))
        return v.value

    def sobol_points(self, m, x0, start, count):
        m = np.ascontiguousarray(m, dtype=np.uint32)
        D = m.shape[0]
        x0 = np.ascontiguousarray(x0 if x0 is not None else np.zeros(D), dtype=np.uint32)
        out = np.zeros((count, D), dtype=np.uint32)
        self._ck(self.L.qiw_sobol_points(self.h, D, _ptr(m, u32p), _ptr(x0, u32p), start, count, _ptr(out, u32p)))
        return out

    def comm_init(self, n_ranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        self._ck(self.L.qiw_comm_init(self.h, n_ranks, rank, _ptr(uid, u8p)))

    def peer_handle(self):
        """Allocate this rank's mailbox and return its CUDA IPC handle (64 bytes)."""
        buf = np.zeros(PEER_HANDLE_BYTES, dtype=np.uint8)
        self._ck(self.L.qiw_peer_handle(self.h, _ptr(buf, u8p)))
        return buf

    def peer_init(self, n_ranks, rank, handles):
        h = np.ascontiguousarray(handles, dtype=np.uint8).reshape(n_ranks * PEER_HANDLE_BYTES)
        self._ck(self.L.qiw_peer_init(self.h, n_ranks, rank, _ptr(h, u8p)))

    def peer_disable(self):
        self._ck(self.L.qiw_peer_init(self.h, 0, 0, None))

    def measure_fp64_peak(self):
        v = C.c_double(0)
        self._ck(self.L.qiw_measure_fp64_peak(self.h, C.byref(v)))
        return v.value
