"""Minimal pure-Python HDF5 reader (h5py is not available in this image).

Covers exactly what the CLOUDSC2 data files need (the reference's `data/reference_*.h5`
and an `input.h5` written the same way): superblock version 0, group symbol tables (v1
B-tree + local heap), version-1 object headers (with continuation blocks), simple
dataspaces, fixed-point / IEEE floating-point datatypes and **contiguous, unfiltered**
data layout (layout message v3 class 1).  Anything else raises `NotImplementedError`.

Replaces, for this path only, the h5py calls behind `ifs_physics_common.iox.HDF5Operator`
used at reference `iox.py:212-244`, `setup.py:47-70` and `nonlinear/reference.py:28-55`.
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
