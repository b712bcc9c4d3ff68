"""Static consistency of the three views of the C ABI: include/qinchworm.h (the contract), the ctypes
signatures the Python host layer binds (qinchworm.jl_b200/lib.py) and the `ccall` tuples of the Julia shim
(julia/QInchwormCUDA.jl, the binding a maintainer of the reference adds — there is no Julia toolchain in
the image, so it cannot be executed here; this test at least proves every ccall names a declared symbol
with the declared number and classes of arguments)."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s):
    """Split at commas that are not nested in (), [] or {}."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _balanced(s, start):
    """s[start] == '(' -> index one past its closing parenthesis."""
    depth = 0
    for i in range(start, len(s)):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i + 1
    raise ValueError("unbalanced")


def _c_class(t):
    t = t.strip()
    if "*" in t or "[" in t:
        return "ptr"
    t = t.replace("const", "").split()
    t = t[0] if len(t) > 1 else t[0]      # drop the parameter name
    return {"int32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "i64", "double": "f64"}[t]


def header_prototypes():
    hdr = open(os.path.join(ROOT, "include", "qinchworm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", "", hdr)
    protos = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(qiw_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        protos[name] = ("str" if "char" in ret else "i32", [] if args in ("void", "") else [_c_class(a) for a in _split_top(args)])
    return protos


def _julia_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Ctx", "Cstring", "Ptr{Cvoid}"):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "UInt64": "i64", "Float64": "f64", "Cdouble": "f64"}[t]


def julia_ccalls():
    src = open(os.path.join(ROOT, "julia", "QInchwormCUDA.jl")).read()
    src = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in src.splitlines())
    calls = []
    for m in re.finditer(r"ccall\(\(:(qiw_[A-Za-z0-9_]+), lib\)", src):
        end = _balanced(src, m.start() + len("ccall"))
        parts = _split_top(src[m.start() + len("ccall("):end - 1])
        name, ret, types, actual = m.group(1), parts[1], parts[2], parts[3:]
        assert types.startswith("(") and types.endswith(")"), (name, types)
        calls.append((name, ret, _split_top(types[1:-1]), actual))
    return calls


def test_header_is_parsed_completely():
    protos = header_prototypes()
    hdr = open(os.path.join(ROOT, "include", "qinchworm.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    assert set(protos) == set(re.findall(r"\b(qiw_[A-Za-z0-9_]+)\s*\(", hdr))


def test_julia_shim_matches_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 15
    for name, ret, types, actual in calls:
        assert name in protos, "julia shim calls %s which qinchworm.h does not declare" % name
        want_ret, want = protos[name]
        assert (ret == "Cstring") == (want_ret == "str"), (name, ret)
        assert len(types) == len(want), "%s: ccall passes %d argument types, header declares %d" % (name, len(types), len(want))
        assert len(actual) == len(types), "%s: %d values for %d argument types" % (name, len(actual), len(types))
        got = [_julia_class(t) for t in types]
        assert got == want, "%s: argument classes %s, header %s" % (name, got, want)
    # the reference-facing worker path must be covered: model upload, topologies, P, evaluation, device-resident run
    assert {"qiw_create", "qiw_destroy", "qiw_set_model", "qiw_set_delta", "qiw_set_grid", "qiw_set_P", "qiw_set_topologies",
            "qiw_eval_seqs", "qiw_eval_batch", "qiw_inchworm_run", "qiw_get_P", "qiw_comm_init", "qiw_peer_init",
            "qiw_scale_P"} <= {c[0] for c in calls}


def test_julia_shim_session_semantics():
    """What can be checked without a Julia toolchain: the entry cache is keyed on what determines the
    compiled program — (mode, order, n_pts_after, corr_idx), as inchworm.py does — not on object identity; Session maps
    the peer mailboxes; the randomisation loop has the reference's early stop (src/randomization.jl:93-99); the step
    seam sends one row and lambda (qiw_scale_P) instead of re-uploading the table inside inchworm_step."""
    src = open(os.path.join(ROOT, "julia", "QInchwormCUDA.jl")).read()
    code = "\n".join(line.split("#")[0] for line in src.splitlines())
    assert "objectid" not in code
    assert "key = (Int(mode), Int(td.order), Int(td.n_pts_after), Int(corr_idx))" in code
    session_ctor = code[code.index("function Session("):code.index("function upload_P!")]
    assert "init_peer!(s)" in session_ctor
    ev = code[code.index("function eval_entries("):code.index("function mean_std(")]
    assert "rp.target_std" in ev and "break" in ev
    step = code[code.index("function inchworm_step(s::Session"):code.index("function scale_P!")]
    assert "upload_P!" not in step


def _ctypes_class(t):
    if t in (C.c_int32, C.c_int):
        return "i32"
    if t in (C.c_int64, C.c_uint64, C.c_longlong, C.c_ulonglong):
        return "i64"
    if t is C.c_double:
        return "f64"
    return "ptr"


def test_python_signatures_match_header(qlib):
    protos = header_prototypes()
    assert set(qlib._SIGNATURES) == set(protos)
    for name, (res, args) in qlib._SIGNATURES.items():
        want_ret, want = protos[name]
        assert (res is C.c_char_p) == (want_ret == "str"), name
        assert [_ctypes_class(a) for a in args] == want, name
