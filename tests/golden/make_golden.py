#!/usr/bin/env python
"""Extract the reference's golden vectors for the qMC hot path into small JSON fixtures.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Sources (reference file -> fixture):
  test/inchworm.h5       -> inchworm_h5.json       (pins test/inchworm.jl:45-51,206-211)
  test/topology_eval.h5  -> topology_eval_h5.json  (pins test/topology_eval.jl:151)
  test/bethe.h5          -> bethe_h5.json          (rho / g references of test/bethe*.jl)
  test/scrambled_sobol.jl-> sobol_tables.json      (tables at :47-54,75-82,136-143,164-179)
  README.md:178-204      -> readme_counts.json     (topology counts, stale Z / rho)
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from h5mini import H5File  # noqa: E402

REF = os.environ.get("QIW_REFERENCE", "/root/reference")


def enc(a):
    a = np.asarray(a)
    if np.iscomplexobj(a):
        return {"shape": list(a.shape), "re": a.real.ravel().tolist(), "im": a.imag.ravel().tolist()}
    return {"shape": list(a.shape), "re": a.ravel().tolist()}


def dump_h5(name, out):
    h = H5File(os.path.join(REF, "test", name))
    d = {p: enc(h.read(p)) for p in h.tree()}
    with open(os.path.join(HERE, out), "w") as f:
        json.dump(d, f, indent=0)
    print(out, len(d), "datasets")


def sobol_tables():
    src = open(os.path.join(REF, "test", "scrambled_sobol.jl")).read()
    # every "ref = [[...], ...]" literal, in file order: unscr D=1, unscr D=5, scr D=1, scr D=5
    blocks = re.findall(r"ref = (\[\[.*?\]\])\n", src, flags=re.S)
    assert len(blocks) == 4, len(blocks)
    keys = ["unscrambled_D1", "unscrambled_D5", "scrambled_D1", "scrambled_D5"]
    out = {}
    for k, blk in zip(keys, blocks):
        rows = re.findall(r"\[([^\[\]]+)\]", blk)
        out[k] = [[float(x) for x in r.replace("\n", " ").split(",") if x.strip()] for r in rows]
        assert len(out[k]) == 8
    with open(os.path.join(HERE, "sobol_tables.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("sobol_tables.json", {k: np.shape(v) for k, v in out.items()})


def readme_counts():
    src = open(os.path.join(REF, "README.md")).read()
    bare = [int(x) for x in re.findall(r"Bare order \d+, # topologies = (\d+)", src)]
    bold = [[int(a), int(b), int(c)] for a, b, c in
            re.findall(r"Bold order (\d+), n_pts_after (\d+), # topologies = (\d+)", src)]
    z = float(re.search(r"Z = ([0-9.]+)", src).group(1))
    rho = [float(x) for x in re.findall(r"\[\[?([0-9.]+) \+ 0.0im;;\]", src)][:4]
    out = {"bare": bare, "bold": bold, "Z_stale": z, "rho_stale": rho}
    with open(os.path.join(HERE, "readme_counts.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("readme_counts.json", out)


def h5_encodings():
    """On-disk encodings HDF5.jl / libhdf5 produced for the reference's golden files: the header messages of
    one complex dataset, one real dataset, the superblock prefix and the root group's structures.  They pin the
    byte layout the product's HDF5 writer (qinchworm_b200/h5out.py) emits."""
    import struct
    out = {}
    h = H5File(os.path.join(REF, "test", "inchworm.h5"))
    for key, path in (("complex_dataset", "/inchworm/P/1"),):
        out[key] = {"shape": list(h.read(path).shape),
                    "messages": {"%04x" % t: h.b[b:b + n].hex() for t, b, n in h.messages(h.lookup(path))}}
    h2 = H5File(os.path.join(REF, "test", "topology_eval.h5"))
    out["real_dataset"] = {"shape": list(h2.read("/x1_list").shape),
                           "messages": {"%04x" % t: h2.b[b:b + n].hex() for t, b, n in h2.messages(h2.lookup("/x1_list"))}}
    out["superblock_prefix"] = h.b[:32].hex()          # signature, versions, sizes, K values, flags, base address
    out["root_header"] = h.b[h.root:h.root + 24].hex()   # v1 object header prefix + symbol-table message header
    heap = struct.unpack_from("<Q", h.b, h.root + 32)[0]
    out["heap_header_prefix"] = h.b[heap:heap + 8].hex()
    btree = struct.unpack_from("<Q", h.b, h.root + 24)[0]
    out["btree_header"] = h.b[btree:btree + 24].hex()
    snod = struct.unpack_from("<Q", h.b, btree + 32)[0]
    out["snod_header_prefix"] = h.b[snod:snod + 6].hex()
    with open(os.path.join(HERE, "h5_encodings.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("h5_encodings.json", list(out))


if __name__ == "__main__":
    h5_encodings()
    dump_h5("inchworm.h5", "inchworm_h5.json")
    dump_h5("topology_eval.h5", "topology_eval_h5.json")
    dump_h5("bethe.h5", "bethe_h5.json")
    sobol_tables()
    readme_counts()
