"""Test-only writer of the HDF5 subset autolabel's `features.hdf` uses (scripts/compute_feature_maps.py:82-118): one
group, one chunked float16 dataset with the lzf filter, attributes `pca` (opaque bytes), `min`, `range`.  Written
independently of autolabel_b200/hdf5_lite.py from the HDF5 File Format Specification (superblock 0, version-1 object
headers, symbol-table groups, v1 B-trees); used to exercise the reader — h5py itself is not available in this image.
"""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def lzf_compress(data):
    """A small greedy liblzf-format compressor (hash of 3-byte prefixes, window 8191, match <= 264)."""
    n = len(data)
    out = bytearray()
    lit = bytearray()
    table = {}
    i = 0

    def flush():
        k = 0
        while k < len(lit):
            run = lit[k:k + 32]
            out.append(len(run) - 1)
            out.extend(run)
            k += 32
        lit.clear()
    while i < n:
        if i + 2 < n:
            key = bytes(data[i:i + 3])
            ref = table.get(key)
            table[key] = i
            if ref is not None and 0 < i - ref <= 8191:
                length = 3
                while i + length < n and length < 264 and data[ref + length] == data[i + length]:
                    length += 1
                flush()
                off = i - ref - 1
                l2 = length - 2
                if l2 < 7:
                    out.append((l2 << 5) | (off >> 8))
                else:
                    out.append((7 << 5) | (off >> 8))
                    out.append(l2 - 7)
                out.append(off & 0xFF)
                i += length
                continue
        lit.append(data[i])
        i += 1
    flush()
    return bytes(out)


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def tell(self):
        return len(self.b)

    def align(self, a=8):
        while len(self.b) % a:
            self.b.append(0)

    def write(self, data):
        at = len(self.b)
        self.b.extend(data)
        return at

    def patch(self, at, data):
        self.b[at:at + len(data)] = data


def _msg(mtype, body, flags=0):
    body = bytes(body)
    pad = (-len(body)) % 8
    return struct.pack("<HHB3x", mtype, len(body) + pad, flags) + body + b"\0" * pad


def _object_header(msgs):
    body = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body


def _dataspace(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _datatype(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        exp_bits, man_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[size]
        bias = (1 << (exp_bits - 1)) - 1
        # class 1 version 1; bit field: little endian, mantissa normalisation 2 (implied msb), sign at the top bit
        head = struct.pack("<BBBBI", 0x11, 0x20, size * 8 - 1, 0, size)
        props = struct.pack("<HHBBBBI", 0, size * 8, man_bits, exp_bits, 0, man_bits, bias)
        return head + props
    if dt.kind in "iu":
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "V":
        tag = b"NUMPY:V\0"                               # ascii tag, padded to a multiple of 8
        return struct.pack("<BBBBI", 0x15, len(tag), 0, 0, dt.itemsize) + tag
    raise NotImplementedError(dt)


def _attribute(name, value):
    value = np.asarray(value)
    nm = name.encode() + b"\0"
    dt = _datatype(value.dtype)
    ds = _dataspace(value.shape) if value.shape else struct.pack("<BBB5x", 1, 0, 0)

    def pad(b):
        return b + b"\0" * ((-len(b)) % 8)
    return struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + pad(nm) + pad(dt) + pad(ds) + value.tobytes()


def _group(buf, entries):
    """entries: {name: object header address}.  Writes heap + SNOD + B-tree; returns the symbol-table message body."""
    names = sorted(entries)
    heap_data = bytearray(b"\0" * 8)
    offs = {}
    for n in names:
        offs[n] = len(heap_data)
        heap_data.extend(n.encode() + b"\0")
        while len(heap_data) % 8:
            heap_data.append(0)
    buf.align()
    data_at = buf.write(heap_data)
    buf.align()
    heap_at = buf.write(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, data_at))
    buf.align()
    snod = b"SNOD" + struct.pack("<BxH", 1, len(names))
    for n in names:
        snod += struct.pack("<QQII16x", offs[n], entries[n], 0, 0)
    snod_at = buf.write(snod)
    buf.align()
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_at, offs[names[-1]] if names else 0)
    tree_at = buf.write(tree)
    return struct.pack("<QQ", tree_at, heap_at)


def write_features_hdf(path, name, features, chunks, attrs, compress=True):
    """features: float16 [N, H, W, C]; chunks: chunk shape; attrs: {name: array | bytes}."""
    features = np.ascontiguousarray(features, dtype=np.float16)
    rank = features.ndim
    buf = _Buf()
    buf.write(b"\0" * 96)                                # superblock placeholder
    # ---- chunks
    records = []
    grid = [range(0, s, c) for s, c in zip(features.shape, chunks)]
    cbytes = int(np.prod(chunks)) * 2
    import itertools
    for k, offs in enumerate(itertools.product(*grid)):
        block = np.zeros(chunks, np.float16)
        sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunks, features.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sel)] = features[sel]
        raw = block.tobytes()
        mask = 0
        if compress:
            z = lzf_compress(raw)
            if len(z) < len(raw) and k % 5 != 4:         # every fifth chunk is stored raw (filter skipped, mask bit set)
                raw = z
            else:
                mask = 1
        else:
            mask = 0
        buf.align()
        records.append((len(raw), mask, offs, buf.write(raw)))
    # ---- chunk B-tree: one leaf node
    buf.align()
    tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(records), UNDEF, UNDEF)
    for (size, mask, offs, addr) in records:
        tree += struct.pack("<II", size, mask) + b"".join(struct.pack("<Q", o) for o in offs) + struct.pack("<Q", 0)
        tree += struct.pack("<Q", addr)
    tree += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in features.shape) + struct.pack("<Q", 0)
    tree_at = buf.write(tree)
    # ---- dataset object header
    layout = struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", tree_at) + b"".join(struct.pack("<I", c) for c in chunks) + struct.pack("<I", 2)
    msgs = [_msg(0x0001, _dataspace(features.shape)), _msg(0x0003, _datatype(np.float16), flags=1), _msg(0x0008, layout)]
    if compress:
        fname = b"lzf\0" + b"\0" * 4
        filt = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 32000, len(fname), 1, 3) + fname + struct.pack("<III", 4, 0x105, cbytes) + b"\0" * 4
        msgs.append(_msg(0x000B, filt))
    for k, v in attrs.items():
        if isinstance(v, (bytes, bytearray)):
            v = np.void(bytes(v))
        msgs.append(_msg(0x000C, _attribute(k, v)))
    buf.align()
    dset_at = buf.write(_object_header(msgs))
    # ---- groups
    st = _group(buf, {name: dset_at})
    buf.align()
    feat_at = buf.write(_object_header([_msg(0x0011, st)]))
    st = _group(buf, {"features": feat_at})
    buf.align()
    root_at = buf.write(_object_header([_msg(0x0011, st)]))
    # ---- superblock 0
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, buf.tell(), UNDEF)
    tree_addr, heap_addr = struct.unpack("<QQ", st)
    sb += struct.pack("<QQII", 0, root_at, 1, 0) + struct.pack("<QQ", tree_addr, heap_addr)
    buf.patch(0, sb)
    with open(path, "wb") as f:
        f.write(buf.b)
