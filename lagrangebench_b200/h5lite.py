"""Minimal read-only HDF5 reader for the files lagrangebench ships (``train/valid/test.h5``).

The reference reads its datasets with h5py (``lagrangebench/data/data.py:109-113,199-255``); h5py
is not available in this image, so the caller side of the step loop gets a small pure-Python
reader of exactly the subset those files use: superblock version 0/1, old-style groups (symbol
table = version-1 B-tree + local heap), version-1 object headers, contiguous or chunked dataset
layout (version-1 chunk B-tree), the deflate (gzip) and shuffle filters, little-endian
fixed-point and IEEE floating-point datatypes.  Anything else raises ``NotImplementedError``.

    with H5File(path) as f:
        keys = f.keys()                       # ['00000', '00001', ...]
        pos = f["00000/position"]             # H5Dataset-like: .shape, .dtype, [a:b], [:]
        frames = pos[10:26]                   # only the chunks overlapping rows 10..25 are inflated
"""

import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


class H5File:
    def __init__(self, path):
        self.path = path
        self._f = open(path, "rb")
        self._read_superblock()

    # ------------------------------------------------------------------ low level
    def close(self):
        if self._f is not None:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _read(self, addr, n):
        self._f.seek(self._base + addr)
        b = self._f.read(n)
        if len(b) != n:
            raise H5Error(f"short read at {addr} (+{n})")
        return b

    def _u(self, b, off, size):
        return int.from_bytes(b[off:off + size], "little")

    def _read_superblock(self):
        self._base = 0
        head = self._f.read(8)
        if head != _SIG:
            raise H5Error("not an HDF5 file (signature)")
        b = self._f.read(8)
        version = b[0]
        if version not in (0, 1):
            raise NotImplementedError(f"HDF5 superblock version {version} (only 0 and 1: libver='earliest' files)")
        self._so, self._sl = b[5], b[6]  # size of offsets / lengths
        if self._so != 8 or self._sl != 8:
            raise NotImplementedError("only 8-byte offsets and lengths")
        self._f.seek(8 + 8 + 2 + 2 + 4 + (4 if version == 1 else 0))
        o = self._so
        rest = self._f.read(4 * o + 2 * o + 8 + 16)
        self._base = self._u(rest, 0, o)
        ste = rest[4 * o:]
        self._root_header = self._u(ste, o, o)
        cache_type = self._u(ste, 2 * o, 4)
        self._root_cache = None
        if cache_type == 1:
            self._root_cache = (self._u(ste, 2 * o + 8, o), self._u(ste, 2 * o + 8 + o, o))

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr):
        """[(type, data bytes)] of a version-1 object header, continuation blocks followed."""
        h = self._read(addr, 16)
        if h[0] != 1:
            raise NotImplementedError(f"object header version {h[0]} (only version 1)")
        n_msgs = self._u(h, 2, 2)
        size = self._u(h, 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < n_msgs:
            baddr, bsize = blocks.pop(0)
            buf = self._read(baddr, bsize)
            off = 0
            while off + 8 <= bsize and len(out) < n_msgs:
                mtype, msize = self._u(buf, off, 2), self._u(buf, off + 2, 2)
                data = buf[off + 8:off + 8 + msize]
                off += 8 + msize
                if mtype == 0x0010:  # continuation
                    blocks.append((self._u(data, 0, self._so), self._u(data, self._so, self._sl)))
                out.append((mtype, data))
        return out

    # ------------------------------------------------------------------ groups
    def _heap_name(self, heap_addr, offset):
        h = self._read(heap_addr, 8 + 2 * self._sl + self._so)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap signature")
        seg_size = self._u(h, 8, self._sl)
        seg_addr = self._u(h, 8 + 2 * self._sl, self._so)
        data = self._read(seg_addr + offset, min(256, seg_size - offset))
        return data.split(b"\0", 1)[0].decode()

    def _group_entries(self, btree_addr, heap_addr):
        """name -> object header address, walking the group's version-1 B-tree."""
        out = {}
        node = self._read(btree_addr, 8 + 2 * self._so)
        if node[:4] != b"TREE" or node[4] != 0:
            raise H5Error("bad group B-tree node")
        level, used = node[5], self._u(node, 6, 2)
        body = self._read(btree_addr + 8 + 2 * self._so, (2 * used + 1) * 8)
        children = [self._u(body, (2 * i + 1) * 8, 8) for i in range(used)]
        for child in children:
            if level > 0:
                out.update(self._group_entries(child, heap_addr))
                continue
            snod = self._read(child, 8)
            if snod[:4] != b"SNOD":
                raise H5Error("bad symbol table node")
            n = self._u(snod, 6, 2)
            ents = self._read(child + 8, n * (2 * self._so + 24))
            for i in range(n):
                e = ents[i * 40:(i + 1) * 40]
                out[self._heap_name(heap_addr, self._u(e, 0, 8))] = self._u(e, 8, 8)
        return out

    def _children(self, header_addr, cache=None):
        if cache is not None:
            return self._group_entries(*cache)
        for mtype, data in self._messages(header_addr):
            if mtype == 0x0011:  # symbol table message
                return self._group_entries(self._u(data, 0, 8), self._u(data, 8, 8))
            if mtype in (0x0002, 0x0006):
                raise NotImplementedError("new-style (link message) groups; write the file with libver='earliest'")
        return None  # not a group

    def keys(self, group="/"):
        return sorted(self._children(self._resolve(group)[0], self._root_cache if group.strip("/") == "" else None))

    def _resolve(self, path):
        addr, cache = self._root_header, self._root_cache
        for part in [p for p in path.split("/") if p]:
            kids = self._children(addr, cache)
            if kids is None or part not in kids:
                raise KeyError(path)
            addr, cache = kids[part], None
        return addr, cache

    def __getitem__(self, path):
        addr, _ = self._resolve(path)
        msgs = self._messages(addr)
        if any(t == 0x0008 for t, _ in msgs):
            return H5Array(self, msgs, path)
        raise KeyError(f"{path} is a group; use keys('{path}')")

    def __contains__(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False


class H5Array:
    """A dataset: ``shape``, ``dtype``, ``ds[a:b]`` (rows along axis 0), ``ds[:]``, ``ds[i]``."""

    def __init__(self, f, msgs, name):
        self._f, self.name = f, name
        self._filters = []
        self._layout = None
        for mtype, d in msgs:
            if mtype == 0x0001:
                self.shape = self._dataspace(d)
            elif mtype == 0x0003:
                self.dtype = self._datatype(d)
            elif mtype == 0x0008:
                self._layout = self._parse_layout(d)
            elif mtype == 0x000B:
                self._filters = self._pipeline(d)
        if self._layout is None or not hasattr(self, "shape") or not hasattr(self, "dtype"):
            raise H5Error(f"{name}: incomplete dataset header")

    # ---- header messages
    @staticmethod
    def _dataspace(d):
        version, rank, flags = d[0], d[1], d[2]
        off = 8 if version == 1 else 4
        return tuple(int.from_bytes(d[off + 8 * i:off + 8 * i + 8], "little") for i in range(rank))

    @staticmethod
    def _datatype(d):
        cls, bits0 = d[0] & 0x0F, d[1]
        size = int.from_bytes(d[4:8], "little")
        if bits0 & 1:
            raise NotImplementedError("big-endian datatypes")
        if cls == 0:
            return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
        if cls == 1:
            return np.dtype("<f" + str(size))
        raise NotImplementedError(f"HDF5 datatype class {cls}")

    def _parse_layout(self, d):
        version, cls = d[0], d[1]
        if version != 3:
            raise NotImplementedError(f"data layout message version {version}")
        if cls == 1:  # contiguous
            return ("contiguous", int.from_bytes(d[2:10], "little"), int.from_bytes(d[10:18], "little"))
        if cls == 2:  # chunked
            ndim = d[2]
            btree = int.from_bytes(d[3:11], "little")
            dims = tuple(int.from_bytes(d[11 + 4 * i:15 + 4 * i], "little") for i in range(ndim))
            return ("chunked", btree, dims[:-1])
        if cls == 0:
            size = int.from_bytes(d[2:4], "little")
            return ("compact", d[4:4 + size])
        raise NotImplementedError(f"layout class {cls}")

    @staticmethod
    def _pipeline(d):
        version, n = d[0], d[1]
        off = 8 if version == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(d[off:off + 2], "little")
            if version == 1 or fid >= 256:
                name_len = int.from_bytes(d[off + 2:off + 4], "little")
                off += 4
            else:
                name_len = 0
                off += 2
            ncd = int.from_bytes(d[off + 2:off + 4], "little")
            off += 4
            off += (name_len + 7) // 8 * 8 if version == 1 else name_len
            cd = [int.from_bytes(d[off + 4 * i:off + 4 * i + 4], "little") for i in range(ncd)]
            off += 4 * ncd
            if version == 1 and ncd % 2:
                off += 4
            out.append((fid, cd))
        return out

    # ---- chunk index
    def _chunks(self, addr, ndim):
        """[(offsets tuple, file address, stored size, filter mask)] of a version-1 chunk B-tree."""
        f = self._f
        if addr == _UNDEF:
            return []
        node = f._read(addr, 24)
        if node[:4] != b"TREE" or node[4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, used = node[5], int.from_bytes(node[6:8], "little")
        key = 8 + 8 * (ndim + 1)
        body = f._read(addr + 24, used * (key + 8) + key)
        out = []
        for i in range(used):
            k = body[i * (key + 8):i * (key + 8) + key]
            child = int.from_bytes(body[i * (key + 8) + key:(i + 1) * (key + 8)], "little")
            if level > 0:
                out += self._chunks(child, ndim)
            else:
                size, mask = int.from_bytes(k[0:4], "little"), int.from_bytes(k[4:8], "little")
                offs = tuple(int.from_bytes(k[8 + 8 * j:16 + 8 * j], "little") for j in range(ndim))
                out.append((offs, child, size, mask))
        return out

    def _decode(self, raw, mask):
        for i, (fid, cd) in reversed(list(enumerate(self._filters))):
            if mask & (1 << i):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:  # shuffle: bytes of all elements grouped by byte position
                es = cd[0] if cd else self.dtype.itemsize
                n = len(raw) // es
                raw = np.frombuffer(raw[:n * es], np.uint8).reshape(es, n).T.tobytes() + raw[n * es:]
            elif fid == 3:  # fletcher32 checksum: 4 trailing bytes
                raw = raw[:-4]
            else:
                raise NotImplementedError(f"HDF5 filter id {fid}")
        return raw

    # ---- reads
    def read_rows(self, start, stop):
        """Rows ``[start, stop)`` along axis 0 as a numpy array."""
        start, stop = max(0, start), min(self.shape[0] if self.shape else 1, stop)
        shape = (max(0, stop - start),) + tuple(self.shape[1:])
        out = np.zeros(shape, self.dtype)
        if shape[0] == 0:
            return out
        kind = self._layout[0]
        if kind == "contiguous":
            row = int(np.prod(self.shape[1:], dtype=np.int64)) * self.dtype.itemsize
            if self._layout[1] != _UNDEF:
                raw = self._f._read(self._layout[1] + start * row, shape[0] * row)
                out = np.frombuffer(raw, self.dtype).reshape(shape).copy()
            return out
        if kind == "compact":
            full = np.frombuffer(self._layout[1], self.dtype).reshape(self.shape)
            return full[start:stop].copy()
        _, btree, cdims = self._layout
        ndim = len(self.shape)
        if not hasattr(self, "_chunk_index"):
            self._chunk_index = self._chunks(btree, ndim)
        for offs, addr, size, mask in self._chunk_index:
            if offs[0] >= stop or offs[0] + cdims[0] <= start:
                continue
            chunk = np.frombuffer(self._decode(self._f._read(addr, size), mask), self.dtype,
                                  count=int(np.prod(cdims))).reshape(cdims)
            src, dst = [], []
            for ax in range(ndim):
                lo = start if ax == 0 else 0
                hi = stop if ax == 0 else self.shape[ax]
                a, b = max(offs[ax], lo), min(offs[ax] + cdims[ax], hi)
                src.append(slice(a - offs[ax], b - offs[ax]))
                dst.append(slice(a - lo, b - lo))
            out[tuple(dst)] = chunk[tuple(src)]
        return out

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 0 or key is Ellipsis:
            key = slice(None)
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            return self.read_rows(k, k + 1)[0]
        if isinstance(key, slice):
            a, b, step = key.indices(self.shape[0])
            rows = self.read_rows(a, b) if step > 0 else self.read_rows(b + 1, a + 1)[::-1]
            return rows[::abs(step)] if abs(step) != 1 else rows
        raise NotImplementedError("only integer and slice indexing along axis 0")
