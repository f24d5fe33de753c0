"""A small pure-Python reader for the HDF5 files pyro2 writes (TEST INFRASTRUCTURE: h5py is not in this image).

pyro's snapshots (pyro/mesh/patch.py:750-788, pyro/simulation_null.py:270-290) are HDF5 superblock version 0 with
old-style groups (B-tree "TREE" nodes, a local "HEAP", symbol-table "SNOD" leaves), version-1 object headers, and
contiguous, uncompressed little-endian datasets; attributes are scalars or short strings.  That subset -- and nothing
else -- is understood here, following the public HDF5 file-format specification (version 1.1), sections III.A-C and IV.A.

    f = h5lite.File(path); f["state/density"] -> numpy array; f.attrs("grid") -> {"nx": 128, ...}; f.keys("state")
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        if self.b[:8] != SIG or self.b[8] != 0:
            raise ValueError("not an HDF5 file with a version-0 superblock")
        if self.b[13] != 8 or self.b[14] != 8:
            raise ValueError("only 8-byte offsets / lengths are handled")
        # superblock v0: 8 sig, 8 versions/sizes, 2+2 group K, 4 flags, then base, free-space, eof, driver addresses and
        # the root group's symbol-table entry
        self.base = self._u64(24)
        root = 24 + 4 * 8
        self.root = self._symtab_entry(root)["header"]

    # ---- primitives -------------------------------------------------------------------------------------------
    def _u64(self, o):
        return struct.unpack_from("<Q", self.b, o)[0]

    def _u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def _u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def _symtab_entry(self, o):
        name_off, header, cache = self._u64(o), self._u64(o + 8), self._u32(o + 16)
        e = {"name_off": name_off, "header": header + self.base}
        if cache == 1:
            e["btree"], e["heap"] = self._u64(o + 24) + self.base, self._u64(o + 32) + self.base
        return e

    # ---- object headers (version 1) -----------------------------------------------------------------------------
    def _messages(self, addr):
        if self.b[addr] != 1:
            raise ValueError("only version-1 object headers are handled")
        nmsg, size = self._u16(addr + 2), self._u32(addr + 8)
        blocks = [(addr + 16, size)]          # 12 bytes of prefix padded to 16
        out = []
        while blocks and len(out) < nmsg:
            o, left = blocks.pop(0)
            end = o + left
            while o + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u16(o), self._u16(o + 2)
                data = o + 8
                if mtype == 0x10:             # continuation: more messages elsewhere
                    blocks.append((self._u64(data) + self.base, self._u64(data + 8)))
                out.append((mtype, data, msize))
                o = data + msize
        return out

    def _group_children(self, addr):
        for mtype, d, _ in self._messages(addr):
            if mtype == 0x11:                 # symbol table message: B-tree and local heap of an old-style group
                return self._walk_btree(self._u64(d) + self.base, self._u64(d + 8) + self.base)
        raise KeyError("object is not a group")

    def _walk_btree(self, node, heap):
        if self.b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = self._u64(heap + 24) + self.base
        out = {}
        if self.b[node:node + 4] == b"SNOD":
            for k in range(self._u16(node + 6)):
                e = self._symtab_entry(node + 8 + 40 * k)
                s = heap_data + e["name_off"]
                out[self.b[s:self.b.index(b"\0", s)].decode()] = e["header"]
            return out
        if self.b[node:node + 4] != b"TREE" or self.b[node + 4] != 0:
            raise ValueError("bad group B-tree node")
        used = self._u16(node + 6)
        o = node + 8 + 16                      # past the sibling pointers
        for k in range(used):
            child = self._u64(o + 8 + 16 * k) + self.base      # key k, child k, key k+1, ...
            out.update(self._walk_btree(child, heap))
        return out

    def _lookup(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            addr = self._group_children(addr)[part]
        return addr

    # ---- datatype / dataspace / layout / attributes ----------------------------------------------------------------
    def _dtype(self, d):
        cls, bits0, size = self.b[d] & 0x0F, self.b[d + 1], self._u32(d + 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1:
            return np.dtype(f"{order}f{size}"), size
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}"), size
        if cls == 3:
            return np.dtype(f"S{size}"), size
        if cls == 9:                           # variable length (h5py's str attributes): a global-heap reference
            return "vlen", size
        if cls == 8:                           # enumeration (h5py's bool): the base integer type follows the 8-byte header
            return self._dtype(d + 8)[0], size
        raise ValueError(f"datatype class {cls} is not handled")

    def _global_heap_object(self, addr, index):
        if self.b[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        end = addr + self._u64(addr + 8)
        o = addr + 16
        while o + 16 <= end:
            idx, size = self._u16(o), self._u64(o + 8)
            if idx == index:
                return self.b[o + 16:o + 16 + size]
            if idx == 0:
                break
            o += 16 + ((size + 7) & ~7)
        raise KeyError("global heap object not found")

    def _dims(self, d):
        ver, rank = self.b[d], self.b[d + 1]
        o = d + (8 if ver == 1 else 4)
        return tuple(self._u64(o + 8 * k) for k in range(rank))

    def __getitem__(self, path):
        addr = self._lookup(path)
        dt = dims = data = None
        for mtype, d, _ in self._messages(addr):
            if mtype == 0x03:
                dt, _ = self._dtype(d)
            elif mtype == 0x01:
                dims = self._dims(d)
            elif mtype == 0x08:
                ver = self.b[d]
                if ver == 3 and self.b[d + 1] == 1:          # contiguous
                    data = (self._u64(d + 2) + self.base, self._u64(d + 10))
                elif ver in (1, 2) and self.b[d + 2] == 1:
                    data = (self._u64(d + 8) + self.base, None)
                else:
                    raise ValueError("only contiguous dataset layouts are handled")
            elif mtype == 0x0B:
                raise ValueError("filtered (compressed) datasets are not handled")
        if dt is None or dims is None or data is None or data[0] == UNDEF + self.base:
            raise KeyError(f"{path}: not a dataset with storage")
        n = int(np.prod(dims)) if dims else 1
        return np.frombuffer(self.b, dtype=dt, count=n, offset=data[0]).reshape(dims).copy()

    def keys(self, path=""):
        return sorted(self._group_children(self._lookup(path)))

    def attrs(self, path=""):
        out = {}
        for mtype, d, _ in self._messages(self._lookup(path)):
            if mtype != 0x0C or self.b[d] != 1:
                continue
            nlen, tlen, slen = self._u16(d + 2), self._u16(d + 4), self._u16(d + 6)
            pad = lambda x: (x + 7) & ~7       # noqa: E731
            o = d + 8
            name = self.b[o:o + nlen].split(b"\0")[0].decode()
            o += pad(nlen)
            dt, size = self._dtype(o)
            o += pad(tlen)
            dims = self._dims(o) if self.b[o + 1] else ()
            o += pad(slen)
            n = int(np.prod(dims)) if dims else 1
            if isinstance(dt, str):             # variable-length string: (length, collection address, object index)
                out[name] = self._global_heap_object(self._u64(o + 4) + self.base, self._u32(o + 12))[:self._u32(o)].decode()
                continue
            val = np.frombuffer(self.b, dtype=dt, count=n, offset=o)
            if dt.kind == "S":
                val = val[0].split(b"\0")[0].decode()
            elif not dims:
                val = val[0].item()
            out[name] = val
        return out


# ---- an h5py-shaped, read-only view (what pyro2_b200.util.io_pyro.read touches: File as a context manager, groups that
#      index / iterate / test membership, .attrs mappings, datasets through np.asarray / [...]) ------------------------------
class Node:
    def __init__(self, f, path):
        self._f, self._path = f, path

    @property
    def attrs(self):
        return self._f.attrs(self._path)

    def __getitem__(self, key):
        if isinstance(key, str):
            if key not in self._f.keys(self._path):
                raise KeyError(key)
            return Node(self._f, f"{self._path}/{key}")
        return self._f[self._path][key]

    def __contains__(self, key):
        return key in self._f.keys(self._path)

    def __iter__(self):
        return iter(self._f.keys(self._path))

    def __len__(self):
        return len(self._f.keys(self._path))

    def __array__(self, dtype=None, copy=None):
        a = self._f[self._path]
        return a if dtype is None else a.astype(dtype)


class H5pyFile(Node):
    """`with h5lite.H5pyFile(name, "r") as f:` -- stands in for h5py.File when READING the reference's files"""

    def __init__(self, filename, mode="r"):
        if mode != "r":
            raise ValueError("h5lite only reads")
        super().__init__(File(filename), "")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
