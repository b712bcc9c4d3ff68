"""Test-only Python interpreter of the traversal programs that libqinchworm_cuda.so compiles
(qiw_entry_program).  It replays the word stream exactly as the CUDA kernel does, with the
per-sample tables computed in numpy, so that the host-side compiler (csrc/qiw_compile.cpp) can be
checked against the oracle on a GPU-less box.  Not part of the product."""
import numpy as np


def grid_interp(D, h, t_f, t_i):
    n = len(D)
    a = min(max(int(np.floor(t_f / h)), 0), n - 2)
    b = min(max(int(np.floor(t_i / h)), 0), n - 2)
    w1, w2 = t_f / h - a, t_i / h - b
    if a == b:
        return D[0] + (w1 - w2) * (D[1] - D[0])
    k = a - b
    return (1 - w1) * (1 - w2) * D[k] + w1 * (1 - w2) * D[k + 1] + (1 - w1) * w2 * D[k - 1] + w1 * w2 * D[k]


def natural_spline(y, h):
    from scipy.interpolate import CubicSpline
    x = np.arange(len(y)) * h
    return CubicSpline(x, y, bc_type="natural")


def sample_table(prog, expansion, payload, mode, t_i, t_w, t_f, times):
    """The per-sample operand table the step kernel builds: i P per (interval, sector), i Delta per slot."""
    S, nP, n_nodes = prog["S"], prog["nP"], prog["n_nodes"]
    beta, n_tau = payload["beta"], payload["n_tau"]
    h = beta / (n_tau - 1)
    # times per position
    t = np.zeros(n_nodes + 1)
    for pos in range(1, n_nodes + 1):
        src = prog["pos_src"][pos]
        t[pos] = {-1: t_i, -2: t_w, -3: t_f}.get(int(src), None) if src < 0 else times[src]
    T = np.zeros(nP + len(prog["dslots"]), dtype=complex)
    for q in range(nP):
        iv, s = divmod(q, S)
        ta, tb = t[iv + 1], max(t[iv + 2], t[iv + 1])
        if mode == 0:
            T[q] = np.exp(-(tb - ta) * payload["energies"][s])
        else:
            T[q] = 1j * grid_interp(expansion.P[:, s], h, tb, ta)
    for j, (pt, ph, tab) in enumerate(prog["dslots"]):
        kind, data = payload["tables"][tab]
        th, tt = t[ph], max(t[pt], t[ph])
        if kind == 1:
            T[nP + j] = 1j * natural_spline(data, beta / (len(data) - 1))(tt - th)
        else:
            T[nP + j] = 1j * grid_interp(data, beta / (len(data) - 1), tt, th)
    return T


def run_records(rec, prog, expansion, payload, mode, t_i, t_w, t_f, times):
    """Replays the factorised configuration records (qiw_entry_records) as the CUDA kernel does:
    segment products first, then K + order operands per configuration."""
    T = sample_table(prog, expansion, payload, mode, t_i, t_w, t_f, times)
    seg = np.ones(rec["nSeg"], dtype=complex)
    for j in range(rec["nSeg"]):
        for q in rec["segdef"][j]:
            if q != 0xFFFF:
                seg[j] *= T[q]
    T = np.concatenate([T, seg])
    assert len(T) == rec["nP"] + rec["nD"] + rec["nSeg"]
    out = np.zeros(prog["S"], dtype=complex)
    for r in rec["rec2"]:
        v = prog["coefs"][int(r[0]) & 0xFFFF]
        for q in r[1:]:
            v = v * T[int(q)]
        out[int(r[0]) >> 16] += v
    return out


def run_lane_program(lp, prog, expansion, payload, mode, t_i, t_w, t_f, times):
    """Replays the lane program (qiw_entry_lane_program) as the step kernel does for one sample: segment products with
    the folded coefficients, then per record prod(Delta operands) * sum over the members of prod(segment products)."""
    T = sample_table(prog, expansion, payload, mode, t_i, t_w, t_f, times)
    assert len(T) == lp["seg0"]
    seg = np.ones(len(lp["segdef"]), dtype=complex)
    for j, d in enumerate(lp["segdef"]):
        assert d[0] != 0xFFFF
        for q in d:
            if q != 0xFFFF:
                seg[j] *= T[q]
        if lp["seg_coef"][j] != 0xFFFF:
            seg[j] *= prog["coefs"][int(lp["seg_coef"][j])]
    T = np.concatenate([T, seg])
    K, n = lp["K"], lp["order"]
    out = np.zeros(prog["S"], dtype=complex)
    n_members = 0
    for s_code, M, n_rec, item0 in lp["sections"]:
        # s_code = sector a | (sector b + 1) << 16: with a second sector, the first M / 2 members of a record belong to
        # sector a and the others to sector b (the pair-interaction operands are shared)
        s_a, s_b = int(s_code) & 0xFFFF, (int(s_code) >> 16) - 1
        ni = (n + M * K + 7) // 8 * 8
        for r in range(n_rec):
            it = lp["items"][item0 + r * ni: item0 + (r + 1) * ni]
            d = np.prod([T[int(q)] for q in it[:n]]) if n else 1.0
            prods = [np.prod([T[int(q)] for q in it[n + m * K: n + (m + 1) * K]]) for m in range(M)]
            if s_b < 0:
                out[s_a] += d * sum(prods)
            else:
                assert M in (2, 4)
                out[s_a] += d * sum(prods[:M // 2])
                out[s_b] += d * sum(prods[M // 2:])
            n_members += M
    return out, n_members


def run_program(prog, expansion, payload, mode, t_i, t_w, t_f, times):
    """Per-sample evaluator value (packed, scalar models): sum over trees of leaf coef * chain."""
    S = prog["S"]
    T = sample_table(prog, expansion, payload, mode, t_i, t_w, t_f, times)
    words = prog["words"]
    out = np.zeros(S, dtype=complex)
    pc = [0]

    def node(vp):
        w = int(words[pc[0]]); pc[0] += 1
        v = vp * T[w & 0xFFF]
        sb = (w >> 12) & 0xFFF
        if sb:
            v = v * T[sb]
        nc = (w >> 24) & 0xFF
        if nc == 0:
            return prog["coefs"][(w >> 32) & 0xFFFF] * v
        return sum(node(v) for _ in range(nc))

    for k in range(len(prog["tree_off"]) - 1):
        pc[0] = int(prog["tree_off"][k])
        root = int(words[pc[0]]); pc[0] += 1
        s_i = (root >> 32) & 0xFFFF
        for _ in range((root >> 24) & 0xFF):
            out[s_i] += node(1.0 + 0j)
        assert pc[0] == int(prog["tree_off"][k + 1]) or k == len(prog["tree_off"]) - 2
    return out


def run_walk_units(units, prog, expansion, payload, mode, t_i, t_w, t_f, times):
    """Replays the walk units of a sector-block model (qiw_entry_walk_units) the way block_walk_kernel does —
    pre-order word stream, running product of at most two columns, branch-point stack — in complex numpy
    arithmetic.  Returns the packed block vector [sum d^2] of one sample."""
    dims = [int(d) for d in payload["dims"]]
    bsize = sum(d * d for d in dims)
    n_nodes = prog["n_nodes"]
    beta, n_tau = payload["beta"], payload["n_tau"]
    h = beta / (n_tau - 1)
    t = np.zeros(n_nodes + 1)
    for pos in range(1, n_nodes + 1):
        src = int(prog["pos_src"][pos])
        t[pos] = {-1: t_i, -2: t_w, -3: t_f}[src] if src < 0 else times[src]
    # tables: TP[interval * bsize + element] = i P (column-major blocks), TD[slot] = i Delta
    nI = n_nodes - 1
    TP = np.zeros(nI * bsize, dtype=complex)
    for q in range(nI):
        ta, tb = t[q + 1], max(t[q + 2], t[q + 1])
        off = 0
        for s, d in enumerate(dims):
            for el in range(d * d):
                r, c = el % d, el // d
                if mode == 0:
                    TP[q * bsize + off + el] = np.exp(-(tb - ta) * payload["energies"][sum(dims[:s]) + r]) if r == c else 0.0
                else:
                    TP[q * bsize + off + el] = 1j * grid_interp(expansion.P[:, off + el], h, tb, ta)
            off += d * d
    TD = np.zeros(len(prog["dslots"]), dtype=complex)
    for j, (pt, ph, tab) in enumerate(prog["dslots"]):
        kind, data = payload["tables"][tab]
        th, tt = t[ph], max(t[pt], t[ph])
        TD[j] = (1j * natural_spline(data, beta / (len(data) - 1))(tt - th)) if kind == 1 else \
            1j * grid_interp(data, beta / (len(data) - 1), tt, th)
    pool = np.asarray(payload["op_pool"])
    words, off = units["words"], units["unit_off"]
    out = np.zeros(bsize, dtype=complex)
    for u in range(len(off) - 1):
        pc = int(off[u])
        x, y, z, w = (int(v) for v in words[pc]); pc += 1
        d0, dcur, has_op = x & 0xF, (x >> 4) & 0xF, (x >> 8) & 1
        nc, c0, nch = (x >> 9) & 3, (x >> 11) & 3, x >> 16
        assert 1 <= nc <= 2 and c0 + nc <= d0
        if has_op:
            V = pool[z:z + dcur * d0].reshape(d0, dcur).T[:, c0:c0 + nc].astype(complex)     # column-major dcur x d0
        else:
            V = np.eye(d0, dtype=complex)[:, c0:c0 + nc]
        acc = np.zeros((d0, nc), dtype=complex)
        dprod = 1.0 + 0j
        stack = []
        while nch > 0:
            if nch > 1:
                stack.append([V.copy(), dprod, nch - 1])
            x, y, z, w = (int(v) for v in words[pc]); pc += 1
            ds, dr, has_op = x & 0xF, (x >> 4) & 0xF, (x >> 8) & 1
            assert V.shape[0] == ds
            Pm = TP[(y & 0xFFFF):(y & 0xFFFF) + ds * ds].reshape(ds, ds).T
            V = Pm @ V
            if has_op:
                V = pool[z:z + dr * ds].reshape(ds, dr).T @ V
            if y >> 16:
                dprod = dprod * TD[(y >> 16) - 1]
            nch = x >> 16
            if nch == 0:
                acc += prog["coefs"][w] * dprod * V
                if not stack:
                    break
                V, dprod, rem = stack[-1]
                V = V.copy()
                if rem > 1:
                    stack[-1][2] = rem - 1
                else:
                    stack.pop()
                nch = 1
        assert pc == int(off[u + 1])
        boff = int(words[int(off[u])][3])
        for j in range(nc):
            out[boff + d0 * (c0 + j): boff + d0 * (c0 + j) + d0] += acc[:, j]
    return out
