"""Read-only HDF5 subset reader, pure Python + numpy: what is needed to load a scene's ``features.hdf`` the way
``autolabel/dataset.py:438-449`` does (``hdf['features/<name>'][:]`` plus the attributes ``pca``, ``min``, ``range``
written by ``scripts/compute_feature_maps.py:82-118``) on a machine without h5py / libhdf5.

Covers the file layout h5py writes by default for such a file (HDF5 File Format Specification 2.0/3.0):
superblock versions 0-3, old-style groups (symbol-table message -> v1 B-tree + local heap + SNOD nodes) and compact
new-style groups (link messages), version-1 and version-2 object headers with continuation blocks, dataspace v1/v2,
fixed-point / floating-point / opaque / string datatypes, contiguous, compact and chunked (v1 B-tree index, layout
message v3) storage, the filter pipeline with ``lzf`` (h5py's filter 32000 — the one ``compute_feature_maps.py`` asks
for), ``deflate`` (1), ``shuffle`` (2) and ``fletcher32`` (3, checksum stripped), and attribute messages v1-v3.
Not covered (raises ``NotImplementedError``): dense (fractal-heap) groups / attributes, v2 B-tree chunk indexes of
layout v4 (``libver='latest'``), variable-length data, external storage.

PROVENANCE / PINNING.  h5py is absent from this image and from the GPU box, so this reader is written against the
published format specification and validated against files produced by ``tests/hdf5_writer.py`` (an independent
writer of the same subset, test infrastructure) — NOT against a file written by h5py itself.  ``load_features`` uses
h5py when it is importable and falls back to this reader otherwise.
"""
import struct
import zlib

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


def lzf_decompress(src, out_len):
    """liblzf stream -> bytes (the format h5py's filter 32000 stores)."""
    src = memoryview(src)
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:                                   # literal run of ctrl + 1 bytes
            ctrl += 1
            out[op:op + ctrl] = src[ip:ip + ctrl]
            ip += ctrl
            op += ctrl
        else:                                           # back reference
            length = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if length == 7:
                length += src[ip]
                ip += 1
            ref -= src[ip]
            ip += 1
            length += 2
            if ref < 0 or op + length > out_len:
                raise ValueError("corrupt lzf stream")
            if ref + length <= op:
                out[op:op + length] = out[ref:ref + length]
            else:                                       # overlapping copy: byte by byte
                for _ in range(length):
                    out[op] = out[ref]
                    op += 1
                    ref += 1
                continue
            op += length
    if op != out_len:
        raise ValueError(f"lzf stream decoded to {op} bytes, expected {out_len}")
    return bytes(out)


def _unshuffle(buf, itemsize):
    a = np.frombuffer(buf, dtype=np.uint8)
    n = a.size // itemsize
    body = a[:n * itemsize].reshape(itemsize, n).T.reshape(-1)
    return body.tobytes() + a[n * itemsize:].tobytes()


class _Reader:
    def __init__(self, data):
        self.d = data
        self.O = 8
        self.L = 8

    def u(self, off, n):
        return int.from_bytes(self.d[off:off + n], "little")

    def off(self, o):
        return self.u(o, self.O)

    def length(self, o):
        return self.u(o, self.L)


class _Object:
    """A parsed object header: list of (type, flags, payload offset, payload size)."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.msgs = []
        r = f.r
        if r.d[addr:addr + 4] == b"OHDR":
            self._parse_v2(addr)
        else:
            self._parse_v1(addr)

    def _parse_v1(self, addr):
        r = self.f.r
        if r.u(addr, 1) != 1:
            raise ValueError("unsupported object header version")
        nmsg = r.u(addr + 2, 2)
        size = r.u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        count = 0
        while blocks and count < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and count < nmsg:
                mtype, msize, flags = r.u(p, 2), r.u(p + 2, 2), r.u(p + 4, 1)
                body = p + 8
                if mtype == 0x10:
                    blocks.append((r.off(body), r.length(body + r.O)))
                else:
                    self.msgs.append((mtype, flags, body, msize))
                p = body + msize
                count += 1

    def _parse_v2(self, addr):
        r = self.f.r
        flags = r.u(addr + 5, 1)
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szbytes = 1 << (flags & 3)
        chunk0 = r.u(p, szbytes)
        p += szbytes
        track = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = r.u(p, 1), r.u(p + 1, 2), r.u(p + 3, 1)
                body = p + 4 + (2 if track else 0)
                if mtype == 0x10:
                    caddr, clen = r.off(body), r.length(body + r.O)
                    blocks.append((caddr + 4, clen - 8))        # skip "OCHK", drop the checksum
                elif mtype != 0:
                    self.msgs.append((mtype, mflags, body, msize))
                p = body + msize

    def find(self, mtype):
        return [m for m in self.msgs if m[0] == mtype]


def _parse_dataspace(r, p):
    ver, rank, flags = r.u(p, 1), r.u(p + 1, 1), r.u(p + 2, 1)
    q = p + (8 if ver == 1 else 4)
    return tuple(r.length(q + i * r.L) for i in range(rank))


def _parse_datatype(r, p):
    """-> (numpy dtype, encoded size in the message)."""
    cv = r.u(p, 1)
    cls, ver = cv & 0x0F, cv >> 4
    bits0 = r.u(p + 1, 1)
    size = r.u(p + 4, 4)
    order = ">" if (bits0 & 1) else "<"
    if cls == 0:                                        # fixed point
        signed = bool(bits0 & 0x08)
        return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), 8 + 4
    if cls == 1:                                        # floating point
        return np.dtype(f"{order}f{size}"), 8 + 12
    if cls == 3:                                        # string (fixed length)
        return np.dtype(f"S{size}"), 8
    if cls == 5:                                        # opaque: tag of ascii bytes, padded to 8
        taglen = bits0
        return np.dtype(f"V{size}"), 8 + (taglen + 7) // 8 * 8
    raise NotImplementedError(f"HDF5 datatype class {cls} (version {ver})")


class Dataset:
    def __init__(self, f, obj, name):
        self.f, self.obj, self.name = f, obj, name
        r = f.r
        (_, _, p, _), = obj.find(0x0001)[:1]
        self.shape = _parse_dataspace(r, p)
        (_, _, p, _), = obj.find(0x0003)[:1]
        self.dtype, _ = _parse_datatype(r, p)
        self.filters = []
        for (_, _, p, _) in obj.find(0x000B):
            self.filters = self._parse_filters(p)
        (_, _, p, _), = obj.find(0x0008)[:1]
        self._layout = p
        self.attrs = _attributes(f, obj)

    def _parse_filters(self, p):
        r = self.f.r
        ver, n = r.u(p, 1), r.u(p + 1, 1)
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(n):
            fid = r.u(q, 2)
            q += 2
            namelen = 0
            if ver == 1 or fid >= 256:
                namelen = r.u(q, 2)
                q += 2
            q += 2                                       # flags
            ncd = r.u(q, 2)
            q += 2
            q += (namelen + 7) // 8 * 8 if ver == 1 else namelen
            cd = [r.u(q + 4 * i, 4) for i in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _defilter(self, buf, mask, nbytes):
        for i, (fid, cd) in reversed(list(enumerate(self.filters))):
            if mask & (1 << i):
                continue                                 # this filter was skipped for this chunk
            if fid == 32000:
                # h5py's lzf filter: cd_values = (filter revision, liblzf version, chunk size in bytes); a chunk that
                # does not shrink is stored with this filter's bit set in the mask (handled above)
                buf = lzf_decompress(buf, cd[2] if len(cd) > 2 and cd[2] else nbytes)
            elif fid == 1:
                buf = zlib.decompress(buf)
            elif fid == 2:
                buf = _unshuffle(buf, cd[0] if cd else self.dtype.itemsize)
            elif fid == 3:
                buf = buf[:-4]
            else:
                raise NotImplementedError(f"HDF5 filter {fid}")
        return buf

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        r = self.f.r
        p = self._layout
        ver, cls = r.u(p, 1), r.u(p + 1, 1)
        if ver != 3:
            raise NotImplementedError(f"data layout message version {ver} (write the file with h5py's default libver)")
        n = int(np.prod(self.shape)) if self.shape else 1
        if cls == 1:                                     # contiguous
            addr, size = r.off(p + 2), r.length(p + 2 + r.O)
            if addr == _UNDEF:
                return np.zeros(self.shape, self.dtype)
            return np.frombuffer(r.d, self.dtype, n, addr).reshape(self.shape).copy()
        if cls == 0:                                     # compact
            size = r.u(p + 2, 2)
            return np.frombuffer(r.d, self.dtype, n, p + 4).reshape(self.shape).copy()
        if cls != 2:
            raise NotImplementedError(f"layout class {cls}")
        rank1 = r.u(p + 2, 1)
        btree = r.off(p + 3)
        cdims = tuple(r.u(p + 3 + r.O + 4 * i, 4) for i in range(rank1))[:-1]
        out = np.zeros(self.shape, self.dtype)
        if btree == _UNDEF:
            return out
        cbytes = int(np.prod(cdims)) * self.dtype.itemsize
        for (size, mask, offs, addr) in self._chunks(btree, rank1):
            raw = self._defilter(bytes(r.d[addr:addr + size]), mask, cbytes)
            chunk = np.frombuffer(raw, self.dtype, int(np.prod(cdims))).reshape(cdims)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, self.shape))
            sel_in = tuple(slice(0, s.stop - s.start) for s in sel_out)
            out[sel_out] = chunk[sel_in]
        return out

    def _chunks(self, addr, rank1):
        r = self.f.r
        if r.d[addr:addr + 4] != b"TREE":
            raise ValueError("bad chunk B-tree node")
        level, used = r.u(addr + 5, 1), r.u(addr + 6, 2)
        p = addr + 8 + 2 * r.O
        keysz = 8 + 8 * rank1
        for i in range(used):
            k = p + i * (keysz + r.O)
            size, mask = r.u(k, 4), r.u(k + 4, 4)
            offs = tuple(r.u(k + 8 + 8 * j, 8) for j in range(rank1 - 1))
            child = r.off(k + keysz)
            if level == 0:
                yield size, mask, offs, child
            else:
                yield from self._chunks(child, rank1)


def _attributes(f, obj):
    r = f.r
    out = {}
    if obj.find(0x0015) and any(r.off(p + 2 + (2 if (r.u(p + 1, 1) & 1) else 0)) != _UNDEF for (_, _, p, _) in obj.find(0x0015)):
        raise NotImplementedError("dense attribute storage (fractal heap)")
    for (_, _, p, _) in obj.find(0x000C):
        ver = r.u(p, 1)
        nsz, tsz, ssz = r.u(p + 2, 2), r.u(p + 4, 2), r.u(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda v: (v + 7) // 8 * 8) if ver == 1 else (lambda v: v)
        name = bytes(r.d[q:q + nsz]).split(b"\0")[0].decode()
        q += pad(nsz)
        dt, _ = _parse_datatype(r, q)
        q += pad(tsz)
        shape = _parse_dataspace(r, q) if ssz else ()
        q += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(r.d, dt, n, q).reshape(shape).copy()
        out[name] = val if shape else val.reshape(())[()]
    return out


class Group:
    def __init__(self, f, obj, name):
        self.f, self.obj, self.name = f, obj, name
        self._links = None
        self.attrs = _attributes(f, obj)

    def _load(self):
        if self._links is not None:
            return self._links
        r = self.f.r
        links = {}
        for (_, _, p, _) in self.obj.find(0x0011):       # symbol table message: v1 B-tree + local heap
            btree, heap = r.off(p), r.off(p + r.O)
            if r.d[heap:heap + 4] != b"HEAP":
                raise ValueError("bad local heap")
            data = r.off(heap + 8 + 2 * r.L)
            self._walk(btree, data, links)
        for (_, _, p, _) in self.obj.find(0x0006):       # link message (compact new-style group)
            flags = r.u(p + 1, 1)
            q = p + 2
            ltype = 0
            if flags & 0x08:
                ltype = r.u(q, 1)
                q += 1
            if flags & 0x04:
                q += 8
            if flags & 0x10:
                q += 1
            lsz = 1 << (flags & 3)
            nlen = r.u(q, lsz)
            q += lsz
            name = bytes(r.d[q:q + nlen]).decode()
            q += nlen
            if ltype == 0:
                links[name] = r.off(q)
        if self.obj.find(0x0002):
            for (_, _, p, _) in self.obj.find(0x0002):   # link info: dense storage if a fractal heap is attached
                flags = r.u(p + 1, 1)
                q = p + 2 + (8 if flags & 1 else 0)
                if r.off(q) != _UNDEF:
                    raise NotImplementedError("dense group storage (fractal heap)")
        self._links = links
        return links

    def _walk(self, addr, heap_data, links):
        r = self.f.r
        sig = bytes(r.d[addr:addr + 4])
        if sig == b"TREE":
            used = r.u(addr + 6, 2)
            p = addr + 8 + 2 * r.O
            for i in range(used):
                child = r.off(p + r.L + i * (r.L + r.O))
                self._walk(child, heap_data, links)
        elif sig == b"SNOD":
            n = r.u(addr + 6, 2)
            p = addr + 8
            for i in range(n):
                e = p + i * (2 * r.O + 8 + 16)
                noff, oaddr = r.off(e), r.off(e + r.O)
                s = heap_data + noff
                links[r.d[s:r.d.find(b"\0", s)].decode()] = oaddr
        else:
            raise ValueError("bad group B-tree node")

    def keys(self):
        return list(self._load())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            links = node._load()
            if part not in links:
                raise KeyError(path)
            obj = _Object(self.f, links[part])
            node = Dataset(self.f, obj, part) if obj.find(0x0008) else Group(self.f, obj, part)
        return node


class File(Group):
    """``with File(path) as f: f['features/dino'][:]`` — the h5py calls of autolabel/dataset.py:438-449."""

    def __init__(self, path, mode="r"):
        if mode != "r":
            raise ValueError("hdf5_lite is read-only")
        with open(path, "rb") as fh:
            data = fh.read()
        if data[:8] != _SIG:
            raise ValueError(f"{path}: not an HDF5 file")
        self.r = _Reader(data)
        r = self.r
        ver = data[8]
        if ver in (0, 1):
            r.O, r.L = data[13], data[14]
            p = 24 + (4 if ver == 1 else 0)
            p += 4 * r.O                                 # base, free space, end of file, driver info
            root = r.off(p + r.O)                        # root symbol table entry: link name offset, header address
        elif ver in (2, 3):
            r.O, r.L = data[9], data[10]
            root = r.off(12 + 3 * r.O)
        else:
            raise NotImplementedError(f"superblock version {ver}")
        if (r.O, r.L) != (8, 8):
            raise NotImplementedError("only 8-byte offsets / lengths")
        super().__init__(self, _Object(self, root), "/")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def load_features(scene_path, name):
    """autolabel/dataset.py:438-449 `_load_features`: -> (features [N, H*W, C] float16, W, H, C, attrs)."""
    import os
    path = os.path.join(scene_path, "features.hdf")
    try:
        import h5py
        with h5py.File(path, "r") as hdf:
            ds = hdf[f"features/{name}"]
            arr, attrs = ds[:], {k: ds.attrs[k] for k in ds.attrs}
    except ImportError:
        with File(path) as hdf:
            ds = hdf[f"features/{name}"]
            arr, attrs = ds[:], dict(ds.attrs)
    N, H, W, C = arr.shape
    return arr.reshape(N, H * W, C), W, H, C, attrs
