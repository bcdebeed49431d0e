"""Independent reader for the subset of the HDF5 file format the reference's output uses (no h5py / libhdf5 in
the image).  Written from the HDF5 File Format Specification ("version 0" superblock family): superblock v0 ->
root symbol-table entry -> v1 B-tree (group nodes) -> symbol-table nodes -> local heap names -> v1 object headers
with Dataspace / Datatype / Fill Value / Layout (contiguous) / Attribute messages -> global heap for vlen strings.
It walks the structures the way libhdf5 does (B-tree keys, binary-searchable sorted SNOD entries) and fails loudly
on anything it does not understand.  TEST INFRASTRUCTURE."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(Exception):
    pass


class Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != SIG:
            raise H5Error("bad signature")
        ver_sb, ver_fs, ver_rg, _, ver_sh, so, sl, _ = struct.unpack_from("<8B", b, 8)
        if (ver_sb, ver_fs, ver_rg, ver_sh, so, sl) != (0, 0, 0, 0, 8, 8):
            raise H5Error("superblock: only v0 with 8-byte offsets/lengths")
        self.leaf_k, self.int_k, flags = struct.unpack_from("<HHI", b, 16)
        base, fsinfo, eof, drv = struct.unpack_from("<4Q", b, 24)
        if base != 0 or fsinfo != UNDEF or drv != UNDEF:
            raise H5Error("superblock addresses")
        if eof != len(b):
            raise H5Error(f"end-of-file address {eof} != file size {len(b)}")
        name_off, hdr, cache, _ = struct.unpack_from("<QQII", b, 56)
        self.root_hdr = hdr
        if cache == 1:
            self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, 80)
        msgs = self.object_header(hdr)
        st = [m for m in msgs if m[0] == 0x0011]
        if len(st) != 1:
            raise H5Error("root group has no symbol table message")
        bt, hp = struct.unpack_from("<QQ", st[0][1], 0)
        if cache == 1 and (bt, hp) != (self.root_btree, self.root_heap):
            raise H5Error("cached root scratch-pad disagrees with the symbol table message")
        self.heap = self.local_heap(hp)
        self.links = {}
        self.walk_btree(bt)

    # ---- low level ------------------------------------------------------------------------------------------------
    def local_heap(self, at):
        b = self.b
        if b[at:at + 4] != b"HEAP" or b[at + 4] != 0:
            raise H5Error("local heap signature/version")
        size, free, data = struct.unpack_from("<QQQ", b, at + 8)
        if free != 1 and free >= size:
            raise H5Error("local heap free list")
        return b[data:data + size]

    def heap_name(self, off):
        end = self.heap.index(b"\0", off)
        return self.heap[off:end].decode()

    def walk_btree(self, at):
        b = self.b
        if b[at:at + 4] != b"TREE":
            raise H5Error("B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", b, at + 4)
        left, right = struct.unpack_from("<QQ", b, at + 8)
        if ntype != 0:
            raise H5Error("not a group B-tree")
        p = at + 24
        keys, kids = [], []
        for n in range(used):
            keys.append(struct.unpack_from("<Q", b, p)[0]); p += 8
            kids.append(struct.unpack_from("<Q", b, p)[0]); p += 8
        keys.append(struct.unpack_from("<Q", b, p)[0])
        for n, child in enumerate(kids):
            if level > 0:
                self.walk_btree(child)
            else:
                names = self.snod(child)
                # B-tree invariant: key[n] < every name in child n <= key[n+1]
                lo, hi = self.heap_name(keys[n]), self.heap_name(keys[n + 1])
                if not all(lo < x <= hi for x in names):
                    raise H5Error(f"B-tree keys ({lo!r},{hi!r}) do not bracket {names}")

    def snod(self, at):
        b = self.b
        if b[at:at + 4] != b"SNOD" or b[at + 4] != 1:
            raise H5Error("symbol table node signature/version")
        n = struct.unpack_from("<H", b, at + 6)[0]
        if n > 2 * self.leaf_k:
            raise H5Error("too many symbols in a node")
        names = []
        for e in range(n):
            off, hdr, cache, _ = struct.unpack_from("<QQII", b, at + 8 + 40 * e)
            name = self.heap_name(off)
            names.append(name)
            self.links[name] = hdr
        if names != sorted(names):
            raise H5Error("symbol table entries are not sorted by name (H5Dopen would fail)")
        return names

    def object_header(self, at):
        b = self.b
        ver, _, nmsg, refc, size = struct.unpack_from("<BBHII", b, at)
        if ver != 1:
            raise H5Error("object header version")
        p, end, out = at + 16, at + 16 + size, []
        for _ in range(nmsg):
            mtype, msize, flags = struct.unpack_from("<HHB", b, p)
            if msize % 8:
                raise H5Error("message size not a multiple of 8")
            out.append((mtype, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        if p != end:
            raise H5Error("object header size does not match its messages")
        return out

    # ---- messages ---------------------------------------------------------------------------------------------------
    @staticmethod
    def dataspace(m):
        ver, rank, flags = struct.unpack_from("<BBB", m, 0)
        if ver != 1 or flags != 0:
            raise H5Error("dataspace message")
        return tuple(struct.unpack_from("<Q", m, 8 + 8 * d)[0] for d in range(rank))

    @staticmethod
    def datatype(m):
        cls, ver = m[0] & 0x0F, m[0] >> 4
        bits = m[1] | (m[2] << 8) | (m[3] << 16)
        size = struct.unpack_from("<I", m, 4)[0]
        if ver != 1:
            raise H5Error("datatype version")
        if cls == 1:  # floating point
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", m, 8)
            if (bits & 1, (bits >> 4) & 3, (bits >> 8) & 0xFF) != (0, 2, 31) or (size, off, prec, eloc, esize, mloc, msize, bias) != (4, 0, 32, 23, 8, 0, 23, 127):
                raise H5Error("not IEEE little-endian binary32")
            return np.dtype("<f4")
        if cls == 0:  # fixed point
            off, prec = struct.unpack_from("<HH", m, 8)
            if bits & 1 or off != 0 or prec != 8 * size:
                raise H5Error("fixed-point layout")
            return np.dtype(("<i" if bits & 8 else "<u") + str(size))
        if cls == 9:  # variable length
            if (bits & 0xF, (bits >> 4) & 0xF, (bits >> 8) & 0xF) != (1, 0, 0) or size != 16:
                raise H5Error("not a null-terminated ASCII vlen string")
            base = Reader.datatype(m[8:])
            if base != np.dtype("u1"):
                raise H5Error("vlen string base type")
            return "vlen_str"
        raise H5Error(f"datatype class {cls}")

    def global_heap_object(self, addr, idx):
        b = self.b
        if b[addr:addr + 4] != b"GCOL" or b[addr + 4] != 1:
            raise H5Error("global heap signature/version")
        total = struct.unpack_from("<Q", b, addr + 8)[0]
        p, end = addr + 16, addr + total
        while p < end:
            i, refc, _, size = struct.unpack_from("<HHIQ", b, p)
            if i == 0:
                if p + size != end:
                    raise H5Error("global heap free-space object does not reach the end of the collection")
                break
            if i == idx:
                return b[p + 16:p + 16 + size]
            p += 16 + (size + 7) // 8 * 8
        raise H5Error(f"global heap object {idx} not found")

    def attribute(self, m):
        ver, _, nlen, dtlen, dslen = struct.unpack_from("<BBHHH", m, 0)
        if ver != 1:
            raise H5Error("attribute version")
        p = 8
        name = m[p:p + nlen].rstrip(b"\0").decode(); p += (nlen + 7) // 8 * 8
        dt = self.datatype(m[p:p + dtlen]); p += (dtlen + 7) // 8 * 8
        shape = self.dataspace(m[p:p + dslen]); p += (dslen + 7) // 8 * 8
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, str):
            vals = []
            for e in range(n):
                length, addr, idx = struct.unpack_from("<IQI", m, p + 16 * e)
                raw = self.global_heap_object(addr, idx)
                if len(raw) != length:
                    raise H5Error("vlen length mismatch")
                vals.append(raw.rstrip(b"\0").decode())
            return name, vals
        val = np.frombuffer(m, dt, n, p)
        return name, (val.reshape(shape) if shape else val[0])

    # ---- API (what H5Dopen + H5Dread + H5Aread would do) -----------------------------------------------------------------
    def names(self):
        return sorted(self.links)

    def dataset(self, name):
        msgs = self.object_header(self.links[name])
        shape = dt = layout = None
        attrs = {}
        for mtype, m in msgs:
            if mtype == 0x0001:
                shape = self.dataspace(m)
            elif mtype == 0x0003:
                dt = self.datatype(m)
            elif mtype == 0x0008:
                ver, cls = m[0], m[1]
                if (ver, cls) != (3, 1):
                    raise H5Error("only contiguous layout v3")
                layout = struct.unpack_from("<QQ", m, 2)
            elif mtype == 0x000C:
                k, v = self.attribute(m)
                attrs[k] = v
            elif mtype == 0x0005:
                if m[0] != 2:
                    raise H5Error("fill value version")
            elif mtype != 0x0000:
                raise H5Error(f"unexpected message type {mtype:#x}")
        addr, size = layout
        n = int(np.prod(shape))
        if size != n * dt.itemsize or addr % 8 or addr + size > len(self.b):
            raise H5Error("layout size/address")
        return np.frombuffer(self.b, dt, n, addr).reshape(shape), attrs
