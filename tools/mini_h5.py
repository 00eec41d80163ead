"""Minimal HDF5 reader (superblock v0, symbol-table groups, v1 object headers, contiguous layout).

Dev-time only: converts the reference's mesh fixtures (ressources/meshes/**/*.h5, datasets /Mesh/Nodes f8
and /Mesh/Cells i4, as read by the reference's src/io/HDF5Io.cpp:111-152) into .npz golden files.
h5py is not available in this image.
"""
import struct
import numpy as np


class MiniH5:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        ver = self.b[8]
        assert ver == 0, "superblock version %d unsupported" % ver
        self.so, self.sl = self.b[13], self.b[14]
        assert self.so == 8 and self.sl == 8
        # 8 sig + 8 version bytes + 2+2 (K) + 4 flags = 24 ; then base, freespace, eof, driver (4x8)
        root_entry = 24 + 4 * 8
        self.root = self._sym_entry(root_entry)

    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _sym_entry(self, off):
        name_off = self._u(off, 8)
        ohdr = self._u(off + 8, 8)
        cache = self._u(off + 16, 4)
        btree = heap = None
        if cache == 1:
            btree = self._u(off + 24, 8)
            heap = self._u(off + 32, 8)
        return dict(name_off=name_off, ohdr=ohdr, cache=cache, btree=btree, heap=heap)

    def _messages(self, ohdr):
        assert self.b[ohdr] == 1, "object header v%d unsupported" % self.b[ohdr]
        nmsg = self._u(ohdr + 2, 2)
        size = self._u(ohdr + 8, 4)
        blocks = [(ohdr + 16, size)]
        msgs = []
        while blocks and len(msgs) < nmsg:
            off, sz = blocks.pop(0)
            end = off + sz
            while off + 8 <= end and len(msgs) < nmsg:
                mtype = self._u(off, 2)
                msz = self._u(off + 2, 2)
                data = off + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self._u(data, 8), self._u(data + 8, 8)))
                msgs.append((mtype, data, msz))
                off = data + msz
        return msgs

    def _group_children(self, entry):
        btree, heap = entry["btree"], entry["heap"]
        if btree is None:
            for mtype, data, _ in self._messages(entry["ohdr"]):
                if mtype == 0x11:
                    btree, heap = self._u(data, 8), self._u(data + 8, 8)
        assert self.b[heap:heap + 4] == b"HEAP"
        heap_data = self._u(heap + 8 + 16, 8)
        out = {}

        def walk(node):
            assert self.b[node:node + 4] == b"TREE"
            level = self.b[node + 5]
            nent = self._u(node + 6, 2)
            p = node + 8 + 16  # skip siblings
            for i in range(nent):
                child = self._u(p + 8 + i * 16, 8)
                if level > 0:
                    walk(child)
                else:
                    assert self.b[child:child + 4] == b"SNOD"
                    ns = self._u(child + 6, 2)
                    for k in range(ns):
                        e = self._sym_entry(child + 8 + k * 40)
                        s = heap_data + e["name_off"]
                        name = self.b[s:self.b.index(b"\0", s)].decode()
                        out[name] = e

        walk(btree)
        return out

    def read(self, path):
        ent = self.root
        parts = [p for p in path.split("/") if p]
        for p in parts:
            ent = self._group_children(ent)[p]
        shape = dtype = addr = None
        for mtype, data, msz in self._messages(ent["ohdr"]):
            if mtype == 0x1:
                v = self.b[data]
                rank = self.b[data + 1]
                base = data + 8 if v == 1 else data + 4
                shape = tuple(self._u(base + 8 * i, 8) for i in range(rank))
            elif mtype == 0x3:
                cls = self.b[data] & 0x0F
                size = self._u(data + 4, 4)
                dtype = {(0, 4): "<i4", (0, 8): "<i8", (1, 8): "<f8", (1, 4): "<f4"}[(cls, size)]
            elif mtype == 0x8:
                v = self.b[data]
                assert v == 3 and self.b[data + 1] == 1, "only contiguous layout v3 supported"
                addr = self._u(data + 2, 8)
        n = int(np.prod(shape))
        return np.frombuffer(self.b, dtype=dtype, count=n, offset=addr).reshape(shape).copy()


if __name__ == "__main__":
    import sys
    f = MiniH5(sys.argv[1])
    n, c = f.read("/Mesh/Nodes"), f.read("/Mesh/Cells")
    print(n.shape, n.dtype, c.shape, c.dtype)
    print(n[:4], c[:4])
