"""Minimal pure-Python HDF5 reader (superblock v0, v1 object headers, contiguous datasets).

Just enough to read the reference's golden files test/inchworm.h5, test/topology_eval.h5 and
test/bethe.h5 without h5py (not in this image).  Only used by make_golden.py.
"""
import struct
import numpy as np


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n"
        assert self.b[8] == 0, "superblock version 0 expected"
        self.so, self.sl = self.b[13], self.b[14]
        assert self.so == 8 and self.sl == 8
        # superblock v0: root group symbol table entry starts at byte 56
        # entry: link name offset(8) object header address(8) ...
        self.root = struct.unpack_from("<Q", self.b, 56 + 8)[0]

    # -- object header (version 1) ------------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver = b[addr]
        assert ver == 1, f"object header v{ver} unsupported"
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                body = pos + 8
                if mtype == 0x0010:  # continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    def children(self, addr):
        res = {}
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x0006:  # link message
                b = self.b
                p = body
                ver, flags = b[p], b[p + 1]
                p += 2
                ltype = 0
                if flags & 0x08:
                    ltype = b[p]; p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsz = 1 << (flags & 3)
                nlen = int.from_bytes(b[p:p + lsz], "little"); p += lsz
                name = b[p:p + nlen].decode(); p += nlen
                assert ltype == 0
                res[name] = struct.unpack_from("<Q", b, p)[0]
            elif mtype == 0x0011:  # symbol table message (old-style group)
                btree, heap = struct.unpack_from("<QQ", self.b, body)
                res.update(self._symtab(btree, heap))
        return res

    def _symtab(self, btree, heap):
        b = self.b
        assert b[heap:heap + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0]
        res = {}

        def walk(node):
            assert b[node:node + 4] == b"TREE"
            level = b[node + 5]
            nent = struct.unpack_from("<H", b, node + 6)[0]
            p = node + 8 + 16
            p += 8  # key 0
            for _ in range(nent):
                child = struct.unpack_from("<Q", b, p)[0]
                p += 16
                if level > 0:
                    walk(child)
                else:
                    assert b[child:child + 4] == b"SNOD"
                    ns = struct.unpack_from("<H", b, child + 6)[0]
                    q = child + 8
                    for _ in range(ns):
                        noff, oaddr = struct.unpack_from("<QQ", b, q)
                        s = heap_data + noff
                        e = b.index(b"\0", s)
                        res[b[s:e].decode()] = oaddr
                        q += 40
        walk(btree)
        return res

    def lookup(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            addr = self.children(addr)[part]
        return addr

    def read(self, path):
        addr = self.lookup(path)
        b = self.b
        shape = dtype = data_addr = None
        for mtype, body, msize in self.messages(addr):
            if mtype == 0x0001:
                ver, rank = b[body], b[body + 1]
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from("<" + "Q" * rank, b, off)
            elif mtype == 0x0003:
                cls = b[body] & 0x0F
                size = struct.unpack_from("<I", b, body + 4)[0]
                if cls == 1 and size == 8:
                    dtype = np.float64
                elif cls == 6 and size == 16:
                    dtype = np.complex128
                elif cls == 0 and size == 8:
                    dtype = np.int64
                else:
                    raise NotImplementedError((cls, size))
            elif mtype == 0x0008:
                ver, lcls = b[body], b[body + 1]
                assert ver == 3 and lcls == 1, "contiguous layout v3 expected"
                data_addr = struct.unpack_from("<Q", b, body + 2)[0]
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(b, dtype=dtype, count=n, offset=data_addr)
        return arr.reshape(shape).copy()

    def tree(self, addr=None, prefix=""):
        addr = self.root if addr is None else addr
        out = []
        for name, a in sorted(self.children(addr).items()):
            kids = self.children(a)
            if kids:
                out += self.tree(a, prefix + "/" + name)
            else:
                out.append(prefix + "/" + name)
        return out
