"""Minimal pure-Python HDF5 reader and writer (h5py is not available in this image).

Covers exactly what the CLOUDSC2 data files need (the reference's `data/reference_*.h5`
and an `input.h5` written the same way): superblock version 0, group symbol tables (v1
B-tree + local heap), version-1 object headers (with continuation blocks), simple
dataspaces, fixed-point / IEEE floating-point datatypes and **contiguous, unfiltered**
data layout (layout message v3 class 1).  Anything else raises `NotImplementedError`.

Replaces, for this path only, the h5py calls behind `ifs_physics_common.iox.HDF5Operator`
used at reference `iox.py:212-244`, `setup.py:47-70` and `nonlinear/reference.py:28-55`.

`write_file(filename, datasets)` writes the same subset (one root group of contiguous datasets,
little-endian int32 / int64 / float32 / float64): used to produce `input.h5`-style files with the
reference's dataset names from synthetic or user data (`synthetic.write_input_h5`), so that the
whole file path -- HDF5Operator, HDF5GridOperator.get_field, get_state, `run_nonlinear
--input-file` -- is exercised although the reference's own `data/input.h5` is not shipped.
"""
from __future__ import annotations

import struct
from typing import Dict, Tuple

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class File:
    def __init__(self, filename: str) -> None:
        self.filename = filename
        with open(filename, "rb") as fh:
            self._buf = fh.read()
        if self._buf[:8] != _SIG:
            raise ValueError(f"{filename}: not an HDF5 file")
        version = self._buf[8]
        if version != 0:
            raise NotImplementedError(f"{filename}: superblock version {version} not supported")
        size_offsets, size_lengths = self._buf[13], self._buf[14]
        if (size_offsets, size_lengths) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths are supported")
        # base(8) freespace(8) eof(8) driver(8) then root symbol table entry
        root_entry = 24 + 4 * 8
        _, obj_header, cache_type = struct.unpack_from("<QQI", self._buf, root_entry)
        if cache_type == 1:
            btree, heap = struct.unpack_from("<QQ", self._buf, root_entry + 24)
        else:
            btree, heap = self._group_addresses(obj_header)
        self._datasets: Dict[str, int] = {}
        self._walk_btree(btree, heap)
        self._cache: Dict[str, np.ndarray] = {}

    # ---------------------------------------------------------------- group traversal
    def _group_addresses(self, header_addr: int) -> Tuple[int, int]:
        for mtype, body in self._messages(header_addr):
            if mtype == 0x11:  # symbol table message
                return struct.unpack_from("<QQ", body, 0)
        raise NotImplementedError("group without symbol table message")

    def _heap_data(self, heap_addr: int) -> int:
        if self._buf[heap_addr : heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        return struct.unpack_from("<Q", self._buf, heap_addr + 8 + 16)[0]

    def _walk_btree(self, addr: int, heap_addr: int) -> None:
        buf = self._buf
        if buf[addr : addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        node_type, level, nused = struct.unpack_from("<BBH", buf, addr + 4)
        if node_type != 0:
            raise NotImplementedError("only group B-trees are supported")
        pos = addr + 8 + 16  # skip siblings
        heap_data = self._heap_data(heap_addr)
        for i in range(nused):
            child = struct.unpack_from("<Q", buf, pos + 8 + i * 16)[0]
            if level > 0:
                self._walk_btree(child, heap_addr)
            else:
                self._read_snod(child, heap_data)

    def _read_snod(self, addr: int, heap_data: int) -> None:
        buf = self._buf
        if buf[addr : addr + 4] != b"SNOD":
            raise ValueError("bad symbol table node signature")
        nsym = struct.unpack_from("<H", buf, addr + 6)[0]
        for i in range(nsym):
            entry = addr + 8 + i * 40
            name_off, obj_header = struct.unpack_from("<QQ", buf, entry)
            start = heap_data + name_off
            end = buf.index(b"\x00", start)
            self._datasets[buf[start:end].decode("ascii")] = obj_header

    # ---------------------------------------------------------------- object headers
    def _messages(self, addr: int):
        buf = self._buf
        version, _, nmsg = struct.unpack_from("<BBH", buf, addr)
        if version != 1:
            raise NotImplementedError(f"object header version {version} not supported")
        header_size = struct.unpack_from("<I", buf, addr + 8)[0]
        blocks = [(addr + 16, header_size)]
        count = 0
        while blocks and count < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and count < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", buf, pos)
                body = buf[pos + 8 : pos + 8 + msize]
                count += 1
                if mtype == 0x10:  # continuation
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                else:
                    yield mtype, body
                pos += 8 + msize

    def _dataset(self, name: str) -> np.ndarray:
        shape = dtype = data_addr = data_size = None
        for mtype, body in self._messages(self._datasets[name]):
            if mtype == 0x01:  # dataspace
                version, rank, flags = struct.unpack_from("<BBB", body, 0)
                off = 8 if version == 1 else 4
                shape = struct.unpack_from(f"<{rank}Q", body, off)
            elif mtype == 0x03:  # datatype
                cls = body[0] & 0x0F
                bits0 = body[1]
                size = struct.unpack_from("<I", body, 4)[0]
                order = ">" if bits0 & 1 else "<"
                if cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
                elif cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                elif cls == 8:  # enumeration (e.g. h5py's bool): values are stored in the base integer type
                    base_bits0 = body[8 + 1]
                    base_size = struct.unpack_from("<I", body, 8 + 4)[0]
                    border = ">" if base_bits0 & 1 else "<"
                    dtype = np.dtype(f"{border}{'i' if (base_bits0 >> 3) & 1 else 'u'}{base_size}")
                else:
                    raise NotImplementedError(f"{name}: datatype class {cls} not supported")
            elif mtype == 0x08:  # layout
                version, lclass = body[0], body[1]
                if version != 3 or lclass != 1:
                    raise NotImplementedError(f"{name}: only contiguous layout v3 is supported")
                data_addr, data_size = struct.unpack_from("<QQ", body, 2)
            elif mtype == 0x0B:
                raise NotImplementedError(f"{name}: filtered datasets are not supported")
        if shape is None or dtype is None or data_addr is None:
            raise ValueError(f"{name}: incomplete dataset header")
        count = int(np.prod(shape)) if len(shape) else 1
        if data_addr == _UNDEF:
            return np.zeros(shape, dtype=dtype.newbyteorder("="))
        arr = np.frombuffer(self._buf, dtype=dtype, count=count, offset=data_addr).reshape(shape)
        return arr.astype(dtype.newbyteorder("="))

    # ---------------------------------------------------------------- mapping protocol
    def keys(self):
        return self._datasets.keys()

    def __contains__(self, name: str) -> bool:
        return name in self._datasets

    def __getitem__(self, name: str) -> np.ndarray:
        if name not in self._cache:
            self._cache[name] = self._dataset(name)
        return self._cache[name]

    def get(self, name: str, default=None):
        return self[name] if name in self else default


# ------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _message(mtype: int, body: bytes) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _datatype_message(dtype: np.dtype) -> bytes:
    size = dtype.itemsize
    if dtype.kind == "f":
        sign, exp_loc, exp_size, man_size, bias = (63, 52, 11, 52, 1023) if size == 8 else (31, 23, 8, 23, 127)
        head = struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, size)  # v1, class 1; LE, mantissa normalisation 2 (implied msb)
        props = struct.pack("<HHBBBBI", 0, 8 * size, exp_loc, exp_size, 0, man_size, bias)
    elif dtype.kind in "iu":
        head = struct.pack("<BBBBI", 0x10 | 0, 0x08 if dtype.kind == "i" else 0x00, 0, 0, size)  # LE, two's complement
        props = struct.pack("<HH", 0, 8 * size)
    else:
        raise NotImplementedError(f"dtype {dtype} not supported")
    return _message(0x03, head + props)


def write_file(filename: str, datasets: Dict[str, np.ndarray], leaf_k: int = 64) -> None:
    """Write `datasets` (name -> array; scalars become shape-(1,) datasets, bools int32) as an HDF5 file: superblock
    v0, root group = v1 B-tree (one level-0 node) + local heap + symbol-table nodes of up to 2 * leaf_k entries, v1
    object headers, simple dataspaces, contiguous layout v3."""
    items = []
    for name in sorted(datasets, key=lambda n: n.encode("ascii")):
        a = np.asarray(datasets[name])
        if a.dtype == np.bool_:
            a = a.astype(np.int32)
        if a.dtype.kind not in "fiu" or a.dtype.itemsize not in (4, 8):
            raise NotImplementedError(f"{name}: dtype {a.dtype}")
        a = np.ascontiguousarray(a.reshape(1) if a.ndim == 0 else a).astype(a.dtype.newbyteorder("<"))
        items.append((name, a))
    cap = 2 * leaf_k
    groups = [items[i : i + cap] for i in range(0, len(items), cap)] or [[]]
    if len(groups) > 32:
        raise NotImplementedError("too many datasets for a single B-tree node")

    # ---- layout: superblock | root object header | B-tree node | heap header | heap data | SNODs | (header, data) per dataset
    SUPER, ROOT_HDR = 96, 16 + 24
    names_blob, name_off = bytearray(b"\x00" * 8), {}
    for name, _ in items:
        name_off[name] = len(names_blob)
        names_blob += _pad8(name.encode("ascii") + b"\x00")
    heap_data = bytes(names_blob)
    btree_size = 8 + 16 + (2 * len(groups) + 1) * 8
    snod_size = 8 + cap * 40
    addr_root = SUPER
    addr_btree = addr_root + ROOT_HDR
    addr_heap = addr_btree + btree_size
    addr_heap_data = addr_heap + 32
    addr_snod0 = addr_heap_data + len(heap_data)
    pos = addr_snod0 + snod_size * len(groups)
    headers, addr_hdr, addr_data = {}, {}, {}
    for name, a in items:
        rank = a.ndim
        msgs = _message(0x01, struct.pack("<BBB5x", 1, rank, 0) + struct.pack(f"<{rank}Q", *a.shape))
        msgs += _datatype_message(a.dtype)
        layout_pos = len(msgs)
        msgs += _message(0x08, struct.pack("<BBQQ", 3, 1, 0, a.nbytes))  # address patched below
        hdr_size = 16 + len(msgs)
        addr_hdr[name] = pos
        addr_data[name] = pos + hdr_size
        msgs = bytearray(msgs)
        struct.pack_into("<Q", msgs, layout_pos + 8 + 2, addr_data[name])
        headers[name] = struct.pack("<BBHII4x", 1, 0, 3, 1, len(msgs)) + bytes(msgs)
        pos = addr_data[name] + a.nbytes + (-a.nbytes % 8)
    eof = pos

    out = bytearray(eof)
    # superblock v0
    struct.pack_into("<8sBBBBBBBBHHI", out, 0, _SIG, 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, 16, 0)
    struct.pack_into("<QQQQ", out, 24, 0, _UNDEF, eof, _UNDEF)
    struct.pack_into("<QQII", out, 56, 0, addr_root, 1, 0)  # root symbol table entry, cache type 1
    struct.pack_into("<QQ", out, 56 + 24, addr_btree, addr_heap)
    # root group object header: one symbol-table message
    root_msgs = _message(0x11, struct.pack("<QQ", addr_btree, addr_heap))
    out[addr_root : addr_root + ROOT_HDR] = struct.pack("<BBHII4x", 1, 0, 1, 1, len(root_msgs)) + root_msgs
    # B-tree node (group node, level 0): key0 = "", key i = largest name of child i-1
    struct.pack_into("<4sBBHQQ", out, addr_btree, b"TREE", 0, 0, len(groups) if items else 0, _UNDEF, _UNDEF)
    p_ = addr_btree + 24
    struct.pack_into("<Q", out, p_, 0)
    for g, grp in enumerate(groups):
        struct.pack_into("<Q", out, p_ + 8 + g * 16, addr_snod0 + g * snod_size)
        struct.pack_into("<Q", out, p_ + 16 + g * 16, name_off[grp[-1][0]] if grp else 0)
    # local heap
    struct.pack_into("<4sB3xQQQ", out, addr_heap, b"HEAP", 0, len(heap_data), 1, addr_heap_data)  # free list: none (1)
    out[addr_heap_data : addr_heap_data + len(heap_data)] = heap_data
    # symbol table nodes
    for g, grp in enumerate(groups):
        a0 = addr_snod0 + g * snod_size
        struct.pack_into("<4sBBH", out, a0, b"SNOD", 1, 0, len(grp))
        for i, (name, _) in enumerate(grp):
            struct.pack_into("<QQII", out, a0 + 8 + i * 40, name_off[name], addr_hdr[name], 0, 0)
    for name, a in items:
        h = headers[name]
        out[addr_hdr[name] : addr_hdr[name] + len(h)] = h
        out[addr_data[name] : addr_data[name] + a.nbytes] = a.tobytes()
    with open(filename, "wb") as fh:
        fh.write(bytes(out))
