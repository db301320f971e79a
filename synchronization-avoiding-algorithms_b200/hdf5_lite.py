"""Dependency-free HDF5 for the result files of the time loop.

The reference stores its per-rank results with h5py: `Results/Dynamics/Local-rank-r.hdf5` and
`Modeled_Local-rank-r.hdf5`, dataset `Displacement` of shape (3n, n_saved) (Data_prepare.py:243-246,
Online_predictor.py:321-324), and `Results/sol_on_shared/rank=r-shared_dof.hdf5` (Shared_extraction.py:38-40); it
reads them back with `h5py.File(path, 'r')['Displacement']` (Shared_extraction.py:32-34, Tools/DNN_tools.py:286-287,
Results/plotter.py:34-44).  Where h5py is installed the package uses it.  Where it is not (this image ships no HDF5
library) the `h5py` stand-in in compat/ writes and reads the SAME files through this module, so what lands in
`Results/` is genuine HDF5 that h5py / h5dump / MATLAB / ParaView open elsewhere.

Written (write_file): the layout every HDF5 library since 1.6 reads and h5py's default (`libver='earliest'`)
produces — superblock version 0, root group as a symbol table (local heap + version-1 B-tree + symbol-table nodes),
version-1 object headers, one contiguous little-endian dataset per array (IEEE floats and fixed-point integers).  The
`compression=` argument of `create_dataset` is a storage option, not part of the data model: arrays are stored
uncompressed.
Read (read_file / File): the above plus what real h5py writes by default for these files — object-header
continuation blocks, data layouts version 1-3 (compact / contiguous / chunked through the version-1 chunk B-tree), the
deflate, shuffle and fletcher32 filters, big- or little-endian floats and integers, a user block in front of the
superblock.  Anything else (superblock 2 / 3 of `libver='latest'`, compound or string types, ...) raises `Hdf5Error`
naming the feature.

Format reference: "HDF5 File Format Specification Version 3.0".  The reader is pinned against a file written by the
HDF5 library itself (a MATLAB 7.3 test file shipped with scipy, tests/test_hdf5_lite.py), the writer against the reader
and, structure by structure, against the bytes of that file.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16            # library defaults: <= 2*LEAF_K symbols per node, <= 2*INTERNAL_K children per B-tree node
CHUNK_K = 32                          # default K of the chunk index (indexed-storage internal node K)


class Hdf5Error(OSError):
    pass


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _datatype_message(dt: np.dtype) -> bytes:
    """Datatype message (type 0x0003) of a little-endian numpy dtype (IV.A.2.d of the specification)."""
    dt = np.dtype(dt)
    n = dt.itemsize
    if dt.kind == "f" and n in (2, 4, 8):
        exp_bits, mant_bits = {2: (5, 10), 4: (8, 23), 8: (11, 52)}[n]
        # class 1 (floating point), version 1; bit field: byte order LE (bit 0 = 0), mantissa normalisation 2 = "msb
        # implied" (bits 4-5), sign bit location in the second byte
        head = struct.pack("<BBBBI", 0x11, 0x20, 8 * n - 1, 0, n)
        prop = struct.pack("<HHBBBBI", 0, 8 * n, mant_bits, exp_bits, 0, mant_bits, (1 << (exp_bits - 1)) - 1)
        return head + prop
    if dt.kind in "iu" and n in (1, 2, 4, 8):
        head = struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, n)   # class 0 (fixed point), bit 3 = signed
        return head + struct.pack("<HH", 0, 8 * n)
    raise Hdf5Error(f"hdf5_lite: dtype {dt} cannot be written (IEEE floats and integers only)")


def _storable(a) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype.kind == "b":
        a = a.astype(np.int8)
    if a.dtype.kind not in "fiu":
        raise Hdf5Error(f"hdf5_lite: dtype {a.dtype} cannot be written (IEEE floats and integers only)")
    return np.ascontiguousarray(a, dtype=a.dtype.newbyteorder("<")).reshape(a.shape)    # (ascontiguousarray makes 0-d arrays 1-d)


def _dataset_header(a: np.ndarray, data_address: int) -> bytes:
    """Version-1 object header of a contiguous dataset: dataspace, datatype, fill value, data layout."""
    space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", int(s)) for s in a.shape)
    fill = struct.pack("<BBBBI", 1, 2, 2, 1, 0)          # version 1, allocate late, write if set, defined, size 0 (library default)
    layout = struct.pack("<BBQQ", 3, 1, data_address if a.nbytes else UNDEF, a.nbytes)   # version 3, class 1 = contiguous
    msgs = [_message(0x0001, space), _message(0x0003, _datatype_message(a.dtype), 1), _message(0x0005, fill, 1),
            _message(0x0008, layout)]
    body = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


def write_file(path, datasets: dict) -> None:
    """Write `datasets` (name -> array) as the members of the root group of a new HDF5 file."""
    names = sorted(datasets, key=lambda s: s.encode("utf-8"))           # symbol tables are ordered by strcmp
    for nm in names:
        if not nm or "/" in nm or "\0" in nm:
            raise Hdf5Error(f"hdf5_lite: dataset name {nm!r}: only members of the root group are supported")
    arrays = [_storable(datasets[nm]) for nm in names]
    per_node = 2 * LEAF_K
    n_nodes = -(-len(names) // per_node)
    if n_nodes > 2 * INTERNAL_K:
        raise Hdf5Error(f"hdf5_lite: more than {2 * INTERNAL_K * per_node} datasets in one file are not supported")
    # ---- local heap: the link names; offset 0 holds the empty string, the unused tail is one free block -------------
    heap_data, name_off = bytearray(8), []
    for nm in names:
        name_off.append(len(heap_data))
        heap_data += _pad8(nm.encode("utf-8") + b"\0")
    free_off = len(heap_data)
    seg_size = max(256, -(-(free_off + 16) // 8) * 8)
    heap_data += struct.pack("<QQ", 1, seg_size - free_off)             # free block: next = 1 (end of list), its size
    heap_data += b"\0" * (seg_size - len(heap_data))
    # ---- addresses ----------------------------------------------------------------------------------------------
    root_body = _message(0x0011, b"\0" * 16, 1) + _message(0x0000, b"")    # symbol-table message (patched below) + NIL
    root_hdr_size = 16 + len(root_body)
    snod_size = 8 + per_node * 40
    tree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    a_root = 96
    a_heap = a_root + root_hdr_size
    a_heap_data = a_heap + 32
    a_tree = a_heap_data + seg_size
    a_snod = a_tree + tree_size
    a_hdr = a_snod + n_nodes * snod_size
    hdr_addr, pos = [], a_hdr
    hdr_sizes = [len(_dataset_header(a, 0)) for a in arrays]
    for sz in hdr_sizes:
        hdr_addr.append(pos)
        pos += sz
    data_addr = []
    for a in arrays:
        pos = -(-pos // 8) * 8
        data_addr.append(pos)
        pos += a.nbytes
    eof = pos
    # ---- structures -----------------------------------------------------------------------------------------------
    superblock = (SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
                  + struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
                  + struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_tree, a_heap))   # root entry, cached B-tree / heap
    assert len(superblock) == 96
    root = struct.pack("<BBHII4x", 1, 0, 2, 1, len(root_body)) + _message(0x0011, struct.pack("<QQ", a_tree, a_heap), 1) + _message(0x0000, b"")
    heap = b"HEAP" + struct.pack("<B3xQQQ", 0, seg_size, free_off, a_heap_data)
    groups = [list(range(i, min(len(names), i + per_node))) for i in range(0, len(names), per_node)]
    tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, len(groups), UNDEF, UNDEF))
    tree += struct.pack("<Q", 0)                                          # key 0: the empty string
    for i, g in enumerate(groups):
        tree += struct.pack("<QQ", a_snod + i * snod_size, name_off[g[-1]])   # child i, key i+1 = the largest name in it
    tree += b"\0" * (tree_size - len(tree))
    snods = bytearray()
    for g in groups:
        node = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(g)))
        for j in g:
            node += struct.pack("<QQII16x", name_off[j], hdr_addr[j], 0, 0)
        node += b"\0" * (snod_size - len(node))
        snods += node
    with open(path, "wb") as f:
        f.write(superblock + root + heap + bytes(heap_data) + bytes(tree) + bytes(snods))
        assert f.tell() == a_hdr
        for a, da in zip(arrays, data_addr):
            f.write(_dataset_header(a, da))
        for a, da in zip(arrays, data_addr):
            f.write(b"\0" * (da - f.tell()))
            if a.nbytes:
                f.write(memoryview(a.reshape(-1)).cast("B"))                # no copy of the (possibly large) history
        assert f.tell() == eof


# ------------------------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes, path=""):
        self.b, self.path = buf, path
        at = 0
        while True:                                   # the superblock sits at 0, 512, 1024, ... (user block in front)
            if buf[at:at + 8] == SIGNATURE:
                break
            at = 512 if at == 0 else at * 2
            if at + 8 > len(buf):
                raise Hdf5Error(f"{path}: not an HDF5 file (no signature)")
        self.sb = at
        ver = buf[at + 8]
        if ver in (0, 1):
            if buf[at + 13] != 8 or buf[at + 14] != 8:
                raise Hdf5Error(f"{path}: offsets / lengths of {buf[at + 13]} / {buf[at + 14]} bytes are not supported")
            o = at + 24 + (4 if ver == 1 else 0)
            self.base = struct.unpack_from("<Q", buf, o)[0]
            _, root_hdr, cache, _ = struct.unpack_from("<QQII", buf, o + 32)
            self.root = ("v1", root_hdr)
        else:
            raise Hdf5Error(f"{path}: superblock version {ver} is not supported (file written with libver='latest'?)")

    # -- primitives -------------------------------------------------------------------------------------------------
    def _at(self, addr):
        return self.base + addr

    def _messages(self, addr):
        """(type, flags, data bytes) of every message of the version-1 object header at `addr`, continuation blocks
        followed."""
        b, o = self.b, self._at(addr)
        out = []
        if b[o:o + 4] == b"OHDR":
            raise Hdf5Error(f"{self.path}: version-2 object headers are not supported (file written with libver='latest'?)")
        ver, _, n_msg, _, size = struct.unpack_from("<BBHII", b, o)
        if ver != 1:
            raise Hdf5Error(f"{self.path}: object header version {ver} at {addr} is not supported")
        blocks = [(o + 16, o + 16 + size)]
        while blocks and len(out) < n_msg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                t, s, fl = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + s]
                p += 8 + s
                if t == 0x10:
                    ca, cl = struct.unpack_from("<QQ", data)
                    blocks.append((self._at(ca), self._at(ca) + cl))
                elif t != 0:
                    out.append((t, fl, data))
        return out

    def _heap_string(self, heap_addr, off):
        b, o = self.b, self._at(heap_addr)
        if b[o:o + 4] != b"HEAP":
            raise Hdf5Error(f"{self.path}: local heap expected at {heap_addr}")
        seg = self._at(struct.unpack_from("<Q", b, o + 24)[0])
        end = b.find(b"\0", seg + off)
        return b[seg + off:end].decode("utf-8")

    def _group_btree(self, addr, heap, out):
        b, o = self.b, self._at(addr)
        if b[o:o + 4] == b"SNOD":
            n = struct.unpack_from("<H", b, o + 6)[0]
            for i in range(n):
                name_off, hdr = struct.unpack_from("<QQ", b, o + 8 + 40 * i)
                out[self._heap_string(heap, name_off)] = hdr
            return
        if b[o:o + 4] != b"TREE":
            raise Hdf5Error(f"{self.path}: B-tree node expected at {addr}")
        ntype, level, used = struct.unpack_from("<BBH", b, o + 4)
        if ntype != 0:
            raise Hdf5Error(f"{self.path}: group B-tree of type {ntype}")
        for i in range(used):
            child = struct.unpack_from("<Q", b, o + 24 + 8 + 16 * i)[0]
            self._group_btree(child, heap, out)

    def members(self, where=None):
        """name -> object header address of the members of a group (default: the root group)."""
        kind, addr = where or self.root
        out = {}
        for t, fl, data in self._messages(addr):
            if t == 0x11:                                                   # symbol table: B-tree + heap
                tree, heap = struct.unpack_from("<QQ", data)
                self._group_btree(tree, heap, out)
            elif t in (0x02, 0x06):
                raise Hdf5Error(f"{self.path}: new-style groups (link messages) are not supported")
        return out

    # -- datasets ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _dtype(data):
        cls, ver = data[0] & 0x0F, data[0] >> 4
        bits0, size = data[1], struct.unpack_from("<I", data, 4)[0]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            if size not in (2, 4, 8):
                raise Hdf5Error(f"floating-point type of {size} bytes is not supported")
            return np.dtype(f"{order}f{size}")
        raise Hdf5Error(f"datatype class {cls} is not supported (fixed-point and floating-point only)")

    @staticmethod
    def _shape(data):
        ver, rank = data[0], data[1]
        if ver == 1:
            return tuple(struct.unpack_from(f"<{rank}Q", data, 8))
        if ver == 2:
            if data[3] == 2:
                raise Hdf5Error("null dataspace")
            return tuple(struct.unpack_from(f"<{rank}Q", data, 4))
        raise Hdf5Error(f"dataspace version {ver} is not supported")

    @staticmethod
    def _filters(data):
        ver, n = data[0], data[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = struct.unpack_from("<H", data, p)[0]
            if ver == 1 or fid >= 256:
                name_len, flags, n_cd = struct.unpack_from("<HHH", data, p + 2)
                p += 8 + (-(-name_len // 8) * 8 if ver == 1 else name_len)
            else:
                flags, n_cd = struct.unpack_from("<HH", data, p + 2)
                p += 6
            cd = struct.unpack_from(f"<{n_cd}I", data, p)
            p += 4 * n_cd
            if ver == 1 and n_cd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def _chunks(self, addr, rank, out):
        """(offsets, filter mask, address, stored size) of every chunk under the version-1 chunk B-tree node at addr."""
        b, o = self.b, self._at(addr)
        if b[o:o + 4] != b"TREE":
            raise Hdf5Error(f"{self.path}: chunk B-tree node expected at {addr}")
        ntype, level, used = struct.unpack_from("<BBH", b, o + 4)
        if ntype != 1:
            raise Hdf5Error(f"{self.path}: chunk B-tree of type {ntype}")
        key = 8 + 8 * (rank + 1)
        p = o + 24
        for _ in range(used):
            size, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from(f"<{rank}Q", b, p + 8)
            child = struct.unpack_from("<Q", b, p + key)[0]
            if level == 0:
                out.append((offs, mask, child, size))
            else:
                self._chunks(child, rank, out)
            p += key + 8

    def dataset(self, addr):
        shape = dtype = layout = None
        filters = []
        for t, fl, data in self._messages(addr):
            if t == 0x01:
                shape = self._shape(data)
            elif t == 0x03:
                dtype = self._dtype(data)
            elif t == 0x08:
                layout = data
            elif t == 0x0B:
                filters = self._filters(data)
        if shape is None or dtype is None or layout is None:
            raise Hdf5Error(f"{self.path}: object at {addr} is not a dataset")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        ver = layout[0]
        if ver in (1, 2):
            rank1, cls = layout[1], layout[2]
            p = 8
            a = UNDEF
            if cls != 0:
                a = struct.unpack_from("<Q", layout, p)[0]
                p += 8
            dims = struct.unpack_from(f"<{rank1}I", layout, p)
            p += 4 * rank1
            if cls == 0:
                size = struct.unpack_from("<I", layout, p)[0]
                raw = layout[p + 4:p + 4 + size]
                return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()
            chunk = dims[:-1]
        elif ver == 3:
            cls = layout[1]
            if cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                return np.frombuffer(layout[4:4 + size], dtype=dtype, count=n).reshape(shape).copy()
            if cls == 1:
                a = struct.unpack_from("<Q", layout, 2)[0]
            elif cls == 2:
                rank1 = layout[2]
                a = struct.unpack_from("<Q", layout, 3)[0]
                chunk = struct.unpack_from(f"<{rank1}I", layout, 11)[:-1]
            else:
                raise Hdf5Error(f"{self.path}: data layout class {cls} is not supported")
        else:
            raise Hdf5Error(f"{self.path}: data layout version {ver} is not supported (written with libver='latest'?)")
        if cls == 1:
            if a == UNDEF or n == 0:
                return np.zeros(shape, dtype=dtype.newbyteorder("="))
            return np.frombuffer(self.b, dtype=dtype, count=n, offset=self._at(a)).reshape(shape).astype(dtype.newbyteorder("="))
        out = np.zeros(shape, dtype=dtype.newbyteorder("="))
        if a == UNDEF:
            return out
        chunks = []
        self._chunks(a, len(shape), chunks)
        for offs, mask, ca, size in chunks:
            raw = self.b[self._at(ca):self._at(ca) + size]
            for i in reversed(range(len(filters))):
                fid, cd = filters[i]
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    raw = np.frombuffer(raw, dtype=np.uint8).reshape(es, -1).T.tobytes() if len(raw) % es == 0 else raw
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise Hdf5Error(f"{self.path}: filter {fid} is not supported (deflate, shuffle, fletcher32 only)")
            block = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out


def _map(path):
    """The file's bytes without reading them: a read-only memory map (datasets are copied out one by one)."""
    import mmap
    with open(path, "rb") as f:
        try:
            return mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:                                                  # empty file
            return b""


def read_file(path) -> dict:
    """name -> array of every dataset in the root group of an HDF5 file."""
    r = _Reader(_map(path), str(path))
    out = {}
    for name, addr in r.members().items():
        try:
            out[name] = r.dataset(addr)
        except Hdf5Error as e:
            if "not a dataset" not in str(e):
                raise
    return out


class File:
    """The slice of h5py.File the reference uses: File(path, 'w').create_dataset(name, data=...), File(path, 'r')[name]
    (array-like; `np.array(f[name])`, `f[name][:]`, `.shape`), keys(), get(), close(), context manager."""

    def __init__(self, name, mode="r", **_):
        self.filename, self.mode = str(name), mode
        self._d, self._reader, self._addr = {}, None, {}
        if mode in ("r", "r+", "a"):
            try:
                self._reader = _Reader(_map(self.filename), self.filename)
                self._addr = self._reader.members()
            except FileNotFoundError:
                if mode != "a":
                    raise
        elif mode not in ("w", "w-", "x"):
            raise ValueError(f"invalid mode {mode!r}")

    def create_dataset(self, name, shape=None, dtype=None, data=None, **_):
        if self.mode == "r":
            raise OSError("file is open read-only")
        name = name.lstrip("/")
        a = np.zeros(shape, dtype=dtype or np.float32) if data is None else np.asarray(data, dtype=dtype)
        self._d[name] = _storable(a).astype(a.dtype.newbyteorder("=")) if a.dtype.kind != "b" else _storable(a)
        return self._d[name]

    def __getitem__(self, name):
        name = name.lstrip("/")
        if name not in self._d:
            if name not in self._addr:
                raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
            self._d[name] = self._reader.dataset(self._addr[name])
        return self._d[name]

    def get(self, name, default=None):
        try:
            return self[name]
        except KeyError:
            return default

    def __contains__(self, name):
        return name.lstrip("/") in self._d or name.lstrip("/") in self._addr

    def keys(self):
        return sorted(set(self._d) | set(self._addr))

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def flush(self):
        if self.mode != "r" and self._d is not None:
            import os
            for k in self._addr:                                     # r+ / a: keep what the file already held
                self[k]
            os.makedirs(os.path.dirname(self.filename) or ".", exist_ok=True)
            write_file(self.filename, self._d)

    def close(self):
        if self._d is not None:
            self.flush()
            self._d = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
