"""Test helper: writes a tiny HDF5 file (superblock v0, old-style groups, contiguous or gzip-chunked
datasets) byte by byte, so that the reader of ``lagrangebench_b200/h5lite.py`` can be exercised on
machines where neither h5py nor the reference's fixtures exist (the GPU box)."""

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def align(self, n=8):
        self.b += b"\0" * ((-len(self.b)) % n)

    def put(self, data):
        self.align()
        at = len(self.b)
        self.b += data
        return at

    def patch(self, at, data):
        self.b[at:at + len(data)] = data


def _msg(mtype, data):
    data = data + b"\0" * ((-len(data)) % 8)
    return struct.pack("<HHB3x", mtype, len(data), 0) + data


def _header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            bits = (0x20, 31, 0)
        else:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
            bits = (0x20, 63, 0)
        return _msg(0x0003, bytes([0x11, bits[0], bits[1], bits[2]]) + struct.pack("<I", dt.itemsize) + props)
    signed = 0x08 if dt.kind == "i" else 0
    return _msg(0x0003, bytes([0x10, signed, 0, 0]) + struct.pack("<I", dt.itemsize) +
                struct.pack("<HH", 0, 8 * dt.itemsize))


def _dataset(buf, arr, chunk_rows=None):
    arr = np.ascontiguousarray(arr)
    space = _msg(0x0001, struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", s) for s in arr.shape))
    msgs = [space, _dtype_msg(arr.dtype)]
    if chunk_rows is None:
        at = buf.put(arr.tobytes())
        msgs.append(_msg(0x0008, struct.pack("<BBQQ", 3, 1, at, arr.nbytes)))
    else:  # gzip-chunked along axis 0, one leaf B-tree node
        cdims = (chunk_rows,) + arr.shape[1:]
        keys = []
        for r0 in range(0, arr.shape[0], chunk_rows):
            chunk = np.zeros(cdims, arr.dtype)
            part = arr[r0:r0 + chunk_rows]
            chunk[:part.shape[0]] = part
            raw = zlib.compress(chunk.tobytes(), 4)
            keys.append((len(raw), r0, buf.put(raw)))
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
        for size, r0, at in keys:
            node += struct.pack("<II", size, 0) + struct.pack("<Q", r0) + b"\0" * (8 * arr.ndim) + struct.pack("<Q", at)
        node += struct.pack("<II", 0, 0) + struct.pack("<Q", arr.shape[0]) + b"\0" * (8 * arr.ndim)
        btree = buf.put(node)
        msgs.append(_msg(0x000B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" +
                         struct.pack("<II", 4, 0)))
        msgs.append(_msg(0x0008, struct.pack("<BBBQ", 3, 2, arr.ndim + 1, btree) +
                         b"".join(struct.pack("<I", c) for c in cdims) + struct.pack("<I", arr.dtype.itemsize)))
    return buf.put(_header(msgs))


def _group(buf, entries):
    """entries: {name: object header address} (at most 8).  Returns (header addr, btree addr, heap addr)."""
    assert 0 < len(entries) <= 8
    names = sorted(entries)
    heap_data = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        heap_data += n.encode() + b"\0"
        heap_data += b"\0" * ((-len(heap_data)) % 8)
    seg = buf.put(bytes(heap_data))
    heap = buf.put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, seg))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
    snod += b"\0" * (40 * (8 - len(names)))
    snod_at = buf.put(snod)
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_at, offs[names[-1]])
    btree = buf.put(tree)
    hdr = buf.put(_header([_msg(0x0011, struct.pack("<QQ", btree, heap))]))
    return hdr, btree, heap


def write_h5(path, groups, chunk_rows=None):
    """``groups = {"00000": {"position": array, "particle_type": array}, ...}``; datasets named
    ``position`` are gzip-chunked along axis 0 when ``chunk_rows`` is given."""
    buf = _Buf()
    buf.b += b"\0" * 96  # superblock placeholder (v0, 8-byte offsets: 56 bytes + 40-byte root entry)
    top = {}
    for gname, dsets in groups.items():
        ents = {name: _dataset(buf, arr, chunk_rows if name == "position" else None) for name, arr in dsets.items()}
        top[gname] = _group(buf, ents)[0]
    root_hdr, root_tree, root_heap = _group(buf, top)
    buf.align()
    eof = buf.tell()
    sb = b"\x89HDF\r\n\x1a\n" + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", root_tree, root_heap)
    assert len(sb) == 96
    buf.patch(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(buf.b))
