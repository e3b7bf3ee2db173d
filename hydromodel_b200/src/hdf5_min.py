"""
A small HDF5 writer and reader in NumPy + zlib, for the result files of the
reference (code/src/simulation.py:697-707: root-level datasets, one per key of
``Simulation.output``, gzip-compressed; ``loadResults``: simulation.py:712-746).

h5py / libhdf5 are not part of this image, and SURVEY 8(f).3 asks that the result
files do not depend on them.  The writer emits the oldest, most widely readable
on-disk structures of the HDF5 File Format Specification (the ones libhdf5 1.6+
and h5py with ``libver='earliest'`` use):

    superblock version 0                       (spec III.A)
    root group: version-1 object header with a Symbol Table message, one
        version-1 B-tree node (type 0), one symbol-table node, a local heap
                                                (spec III.B-D, IV.A.2.r)
    per dataset: version-1 object header with Dataspace (v1), Datatype (v1, fixed
        point or IEEE floating point, little endian), Fill Value (v2), Filter
        Pipeline (v1, deflate) and Data Layout (v3) messages; the raw data is one
        chunk indexed by a version-1 B-tree node (type 1), or contiguous when the
        array is empty or compression is off        (spec IV.A.2)

The reader is independent of the writer (it walks the file through the addresses
it finds): superblock 0/1, symbol-table groups (nested groups are flattened to
``a/b`` names), object-header continuation blocks, contiguous / chunked / compact
layouts, multi-level chunk B-trees, deflate and shuffle filters, little- and
big-endian integers and floats.  That covers files written here and the files
h5py writes for the reference with its default (``earliest``) format bounds.

Verification status: round trips and structural checks (tests/test_host_objects.py);
NOT checked against libhdf5 itself, which does not exist in this image.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
GROUP_LEAF_K = 64          # symbol-table node holds up to 2K entries
GROUP_INTERNAL_K = 16
CHUNK_BTREE_K = 32         # default of superblock version 0
HEAP_FREE_NULL = 1         # H5HL_FREE_NULL


class HDF5FormatError(RuntimeError):
    pass


def _pad8(b):
    return b + b"\x00" * (-len(b) % 8)


# =============================================================================== writer
def _datatype_message(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        else:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        head = struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize)
        prop = struct.pack("<HHBBBBI", 0, 8 * dt.itemsize, eloc, esize, 0, msize, bias)
        return head + prop
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10 | 0, 0x08 if dt.kind == "i" else 0x00, 0, 0, dt.itemsize)
        return head + struct.pack("<HH", 0, 8 * dt.itemsize)
    raise HDF5FormatError(f"unsupported dtype {dt}")


def _message(mtype, data, flags=0):
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _normalise(arr):
    arr = np.asarray(arr)
    if arr.dtype == np.bool_:
        arr = arr.astype(np.uint8)
    if arr.dtype.kind not in "fiu":
        raise HDF5FormatError(f"dataset of dtype {arr.dtype} cannot be stored")
    if arr.dtype.kind == "f" and arr.dtype.itemsize == 2:
        arr = arr.astype(np.float32)
    return np.require(arr.astype(arr.dtype.newbyteorder("<"), copy=False), requirements="C")


def check_writable(shapes, itemsize=8):
    """Raise HDF5FormatError now if a dataset of one of ``shapes`` could not be stored by
    ``write_file`` (each dataset is cut into at most 64 chunks along its first axis and a chunk
    must stay below 4 GiB) -- so that a long run learns it before it starts, not at save time."""
    for shape in shapes:
        shape = tuple(int(d) for d in shape)
        if not shape or 0 in shape:
            continue
        rows = -(-shape[0] // (2 * CHUNK_BTREE_K))
        cbytes = rows * int(np.prod(shape[1:], dtype=np.int64)) * itemsize
        if any(d >= 2 ** 32 for d in shape) or cbytes >= 2 ** 32:
            raise HDF5FormatError(f"a dataset of shape {shape} is too large for the built-in HDF5 writer "
                                  f"(64 chunks of < 4 GiB); decimate the output (out_stride) or install h5py")


def write_file(path, datasets, compression_level=4, chunk_rows=None):
    """Write ``{name: ndarray}`` as root-level datasets (gzip level 4 = h5py's default
    for ``compression='gzip'``).  ``compression_level=None`` stores contiguous raw data.
    ``chunk_rows``: rows of the first axis per chunk (default: the whole array is one
    chunk; raised if needed so that a dataset has at most 64 chunks, one B-tree leaf)."""
    names = sorted(datasets, key=lambda s: s.encode("utf-8"))
    if len(names) > 2 * GROUP_LEAF_K:
        raise HDF5FormatError("too many datasets for one symbol-table node")
    for n in names:
        if not n or "/" in n or "\x00" in n:
            raise HDF5FormatError(f"bad dataset name {n!r}")

    # ---- local heap data segment: "" at offset 0, then the names
    heap = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode("utf-8") + b"\x00")
    free_off = len(heap)
    heap += struct.pack("<QQ", HEAP_FREE_NULL, 16)          # one free block closes the segment

    # ---- fixed part of the file
    sb_size, root_oh_size = 96, 16 + 24
    btree_size = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8
    snod_size = 8 + 2 * GROUP_LEAF_K * 40
    a_root = sb_size
    a_btree = a_root + root_oh_size
    a_heap = a_btree + btree_size
    a_heap_data = a_heap + 32
    a_snod = a_heap_data + len(heap)
    pos = a_snod + snod_size
    pos += -pos % 8

    blobs, oh_addr = [], {}
    for n in names:
        arr = _normalise(datasets[n])
        rank, esize = arr.ndim, arr.dtype.itemsize
        msgs = [
            _message(0x0001, struct.pack("<BBB5x", 1, rank, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)),
            _message(0x0003, _datatype_message(arr.dtype), flags=1),
        ]
        chunked = compression_level is not None and arr.size > 0 and rank > 0
        # fill value, version 2: allocation time (3 incremental / 2 late), write time 2 (if set),
        # defined = 1 with size 0 (library default fill)
        msgs.append(_message(0x0005, struct.pack("<BBBBI", 2, 3 if chunked else 2, 2, 1, 0)))
        raw = arr.tobytes()
        if chunked:
            rows = arr.shape[0] if not chunk_rows else max(1, min(int(chunk_rows), arr.shape[0]))
            rows = max(rows, -(-arr.shape[0] // (2 * CHUNK_BTREE_K)))      # one B-tree leaf: <= 64 chunks
            n_chunks = -(-arr.shape[0] // rows)
            cshape = (rows,) + arr.shape[1:]
            cbytes = int(np.prod(cshape, dtype=np.int64)) * esize
            if any(d >= 2 ** 32 for d in arr.shape) or cbytes >= 2 ** 32:
                raise HDF5FormatError(f"dataset {n!r}: chunk too large")
            if n_chunks > 2 * CHUNK_BTREE_K:
                raise HDF5FormatError(f"dataset {n!r}: more than {2 * CHUNK_BTREE_K} chunks")
            comps = []
            for k in range(n_chunks):                      # edge chunks are stored at full size
                part = arr[k * rows:(k + 1) * rows]
                if part.shape[0] < rows:
                    part = np.concatenate([part, np.zeros((rows - part.shape[0],) + arr.shape[1:], arr.dtype)])
                comps.append(zlib.compress(part.tobytes(), compression_level))
            msgs.append(_message(0x000B, struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", 1, 0, 1, 1)
                                 + struct.pack("<I4x", compression_level)))
            key_size = 8 + 8 * (rank + 1)
            node_size = 24 + (2 * CHUNK_BTREE_K + 1) * key_size + 2 * CHUNK_BTREE_K * 8
            layout_len = 3 + 8 + 4 * (rank + 1)
            oh_len = 16 + sum(len(m) for m in msgs) + 8 + layout_len + (-layout_len % 8)
            a_oh = pos
            a_node = a_oh + oh_len
            a_node += -a_node % 8
            msgs.append(_message(0x0008, struct.pack("<BBBQ", 3, 2, rank + 1, a_node)
                                 + b"".join(struct.pack("<I", d) for d in cshape) + struct.pack("<I", esize)))
            oh = _object_header(msgs)
            assert len(oh) == oh_len
            node = bytearray(struct.pack("<4sBBHQQ", b"TREE", 1, 0, n_chunks, UNDEF, UNDEF))
            a_chunk = a_node + node_size
            data = bytearray()
            for k, comp in enumerate(comps):               # key k, child k
                node += struct.pack("<II", len(comp), 0) + struct.pack("<Q", k * rows) + b"\x00" * (8 * rank)
                node += struct.pack("<Q", a_chunk + len(data))
                data += _pad8(comp)
            node += struct.pack("<II", 0, 0) + struct.pack("<Q", n_chunks * rows) + b"\x00" * (8 * rank)   # last key
            node += b"\x00" * (node_size - len(node))
            blob = oh + b"\x00" * (a_node - a_oh - len(oh)) + bytes(node) + bytes(data)
        else:
            a_oh = pos
            layout_len = 2 + 16
            oh_len = 16 + sum(len(m) for m in msgs) + 8 + layout_len + (-layout_len % 8)
            a_data = a_oh + oh_len if raw else UNDEF
            msgs.append(_message(0x0008, struct.pack("<BBQQ", 3, 1, a_data, len(raw))))
            oh = _object_header(msgs)
            assert len(oh) == oh_len
            blob = oh + raw
        blob = _pad8(blob)
        oh_addr[n] = a_oh
        blobs.append(blob)
        pos += len(blob)
    eof = pos

    out = bytearray()
    # superblock, version 0
    out += SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0)
    out += struct.pack("<HHI", GROUP_LEAF_K, GROUP_INTERNAL_K, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    out += struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)    # root symbol-table entry
    assert len(out) == sb_size
    # root group object header: Symbol Table message
    out += _object_header([_message(0x0011, struct.pack("<QQ", a_btree, a_heap))])
    assert len(out) == a_btree
    # group B-tree node: one child (the symbol-table node); keys are heap offsets of names
    node = bytearray(struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1 if names else 0, UNDEF, UNDEF))
    if names:
        node += struct.pack("<QQQ", 0, a_snod, name_off[names[-1]])
    node += b"\x00" * (btree_size - len(node))
    out += node
    # local heap
    out += struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), free_off, a_heap_data) + heap
    assert len(out) == a_snod
    # symbol-table node, entries sorted by name
    snod = bytearray(struct.pack("<4sBBH", b"SNOD", 1, 0, len(names)))
    for n in names:
        snod += struct.pack("<QQII16x", name_off[n], oh_addr[n], 0, 0)
    snod += b"\x00" * (snod_size - len(snod))
    out += snod
    out += b"\x00" * (-len(out) % 8)
    for blob in blobs:
        out += blob
    assert len(out) == eof
    with open(path, "wb") as fh:
        fh.write(out)
    return str(path)


# =============================================================================== reader
class _Reader:
    def __init__(self, buf):
        # the superblock sits at byte 0 or, behind a user block, at 512, 1024, 2048, ... (spec
        # III.A); every address in the file is relative to the superblock's base address
        sb = 0
        while buf[sb:sb + 8] != SIGNATURE:
            sb = 512 if sb == 0 else 2 * sb
            if sb + 8 > len(buf):
                raise HDF5FormatError("not an HDF5 file (signature)")
        whole = buf
        buf = buf[sb:]
        ver = buf[8]
        if ver not in (0, 1):
            raise HDF5FormatError(f"superblock version {ver} is not supported (0 and 1 are)")
        if buf[13] != 8 or buf[14] != 8:
            raise HDF5FormatError("only 8-byte offsets and lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", buf, off)
        off += 32
        _, self.root_oh, cache, _ = struct.unpack_from("<QQII", buf, off)
        if self.base not in (0, sb):
            raise HDF5FormatError("base address differs from the superblock's position")
        if self.eof > len(whole):
            raise HDF5FormatError("file is truncated (end-of-file address beyond the file)")
        self.b = buf                                   # addresses below are offsets into this view
        self.base = 0

    # ---- object headers
    def messages(self, addr):
        b = self.b
        if b[addr] != 1:
            raise HDF5FormatError("only version-1 object headers are supported")
        nmsg, = struct.unpack_from("<H", b, addr + 2)
        size, = struct.unpack_from("<I", b, addr + 8)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                data = bytes(b[p + 8:p + 8 + msize])
                p += 8 + msize
                if mtype == 0x0010:                       # continuation
                    a, ln = struct.unpack("<QQ", data[:16])
                    blocks.append((a, ln))
                out.append((mtype, data))
        return out

    # ---- groups
    def group_entries(self, oh_addr, prefix=""):
        for mtype, data in self.messages(oh_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack("<QQ", data[:16])
                hsize, hfree, hdata = struct.unpack_from("<QQQ", self.b, heap + 8)
                if bytes(self.b[heap:heap + 4]) != b"HEAP":
                    raise HDF5FormatError("bad local heap")
                yield from self._group_btree(btree, hdata, prefix)

    def _group_btree(self, addr, hdata, prefix):
        b = self.b
        if bytes(b[addr:addr + 4]) == b"SNOD":
            n, = struct.unpack_from("<H", b, addr + 6)
            for k in range(n):
                noff, oh, cache = struct.unpack_from("<QQI", b, addr + 8 + 40 * k)
                end = b.index(b"\x00", hdata + noff)
                name = bytes(b[hdata + noff:end]).decode("utf-8")
                if any(t == 0x0011 for t, _ in self.messages(oh)):
                    yield from self.group_entries(oh, prefix + name + "/")
                else:
                    yield prefix + name, oh
            return
        if bytes(b[addr:addr + 4]) != b"TREE" or b[addr + 4] != 0:
            raise HDF5FormatError("bad group B-tree node")
        used, = struct.unpack_from("<H", b, addr + 6)
        for k in range(used):
            child, = struct.unpack_from("<Q", b, addr + 24 + 8 + 16 * k)
            yield from self._group_btree(child, hdata, prefix)

    # ---- datasets
    def dataset(self, oh_addr):
        shape = dtype = layout = None
        filters = []
        for mtype, d in self.messages(oh_addr):
            if mtype == 0x0001:
                ver, rank, flags = d[0], d[1], d[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, d, p)
            elif mtype == 0x0003:
                dtype = self._dtype(d)
            elif mtype == 0x000B:
                filters = self._filters(d)
            elif mtype == 0x0008:
                layout = d
        if shape is None or dtype is None or layout is None:
            raise HDF5FormatError("dataset without dataspace / datatype / layout")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] in (1, 2):
            # versions 1 and 2 (libhdf5 <= 1.6): dimensionality, class, 5 reserved bytes, the
            # address (absent for compact data), 4-byte dimensions, then the element size
            # (chunked) or the compact data; rewritten here in the version-3 form used below
            nd, cls = layout[1], layout[2]
            if cls == 0:
                size, = struct.unpack_from("<I", layout, 8 + 4 * nd)
                layout = bytes([3, 0]) + struct.pack("<H", size) + layout[12 + 4 * nd:12 + 4 * nd + size]
            elif cls == 1:
                addr, = struct.unpack_from("<Q", layout, 8)
                layout = bytes([3, 1]) + struct.pack("<QQ", addr, n * dtype.itemsize)
            elif cls == 2:
                layout = bytes([3, 2, nd]) + layout[8:16] + layout[16:16 + 4 * nd]
        if layout[0] != 3:
            raise HDF5FormatError(f"data layout message version {layout[0]} is not supported")
        cls = layout[1]
        if cls == 0:                                       # compact
            size, = struct.unpack_from("<H", layout, 2)
            raw = layout[4:4 + size]
            return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
        if cls == 1:                                       # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF or n == 0:
                return np.zeros(shape, dtype=dtype)
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=addr + self.base).reshape(shape).copy()
        if cls == 2:                                       # chunked
            nd = layout[2]
            btree, = struct.unpack_from("<Q", layout, 3)
            cdims = struct.unpack_from("<" + "I" * nd, layout, 11)
            out = np.zeros(shape, dtype=dtype)
            if btree != UNDEF:
                for offs, mask, addr, nbytes in self._chunks(btree, nd):
                    raw = bytes(self.b[addr:addr + nbytes])
                    for k, (fid, cd) in reversed(list(enumerate(filters))):
                        if mask & (1 << k):
                            continue
                        if fid == 1:
                            raw = zlib.decompress(raw)
                        elif fid == 2:                     # shuffle
                            es = cd[0] if cd else dtype.itemsize
                            raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                        else:
                            raise HDF5FormatError(f"filter {fid} is not supported")
                    chunk = np.frombuffer(raw, dtype=dtype, count=int(np.prod(cdims[:-1]))).reshape(cdims[:-1])
                    sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                    out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
            return out
        raise HDF5FormatError(f"layout class {cls}")

    def _chunks(self, addr, nd):
        b = self.b
        if bytes(b[addr:addr + 4]) != b"TREE" or b[addr + 4] != 1:
            raise HDF5FormatError("bad chunk B-tree node")
        level, used = b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        ks = 8 + 8 * nd
        p = addr + 24
        for _ in range(used):
            nbytes, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<" + "Q" * nd, b, p + 8)
            child, = struct.unpack_from("<Q", b, p + ks)
            if level == 0:
                yield offs[:-1], mask, child, nbytes
            else:
                yield from self._chunks(child, nd)
            p += ks + 8

    @staticmethod
    def _dtype(d):
        cls, ver = d[0] & 0x0F, d[0] >> 4
        size, = struct.unpack_from("<I", d, 4)
        order = ">" if d[1] & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if d[1] & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        raise HDF5FormatError(f"datatype class {cls} is not supported")

    @staticmethod
    def _filters(d):
        ver, n = d[0], d[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", d, p)
            if ver == 1 or fid >= 256:
                nlen, flags, ncd = struct.unpack_from("<HHH", d, p + 2)
                p += 8 + nlen + (-nlen % 8 if ver == 1 else 0)
            else:
                flags, ncd = struct.unpack_from("<HH", d, p + 2)
                p += 6
            cd = struct.unpack_from("<" + "I" * ncd, d, p)
            p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
            out.append((fid, cd))
        return out


def read_file(path):
    """All datasets of the file as ``{name: ndarray}`` (nested groups: ``a/b``)."""
    with open(path, "rb") as fh:
        buf = fh.read()
    rd = _Reader(buf)
    return {name: rd.dataset(oh) for name, oh in rd.group_entries(rd.root_oh)}
