"""TEST INFRASTRUCTURE: a minimal HDF5 *writer* (old-style format: superblock 0, symbol-table
groups, object headers v1, layout v3, chunk B-trees v1, shuffle + deflate) used to build 10x-style
``.h5`` fixtures for tests/test_h5lite.py -- h5py does not exist in this image.  The reader it
exercises (nanoranger_b200/h5lite.py) is additionally pinned to a file written by libhdf5 itself
(scipy's MATLAB v7.3 test file)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class Writer:
    def __init__(self, userblock: int = 0):
        self.ub = userblock
        self.buf = bytearray(96)              # superblock + root symbol table entry, filled at the end

    def alloc(self, data: bytes, align: int = 8) -> int:
        while len(self.buf) % align:
            self.buf.append(0)
        a = len(self.buf)
        self.buf += data
        return a

    # ---- object headers -------------------------------------------------------------------------
    @staticmethod
    def _msg(t: int, data: bytes) -> bytes:
        data += b"\0" * (-len(data) % 8)
        return struct.pack("<HHB3x", t, len(data), 0) + data

    def _ohdr(self, msgs: list[bytes]) -> int:
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    # ---- datasets -------------------------------------------------------------------------------
    @staticmethod
    def _dtype_msg(dt: np.dtype) -> bytes:
        if dt.kind in "iu":
            return struct.pack("<BBBBIHH", 0x10 | 0, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
        if dt.kind == "f":
            return struct.pack("<BBBBI", 0x10 | 1, 0x20, 0x3F if dt.itemsize == 8 else 0x1F, 0, dt.itemsize) + b"\0" * 12
        if dt.kind == "S":
            return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, dt.itemsize)
        raise ValueError(dt)

    def dataset(self, arr: np.ndarray, chunks: int | None = None, shuffle: bool = True, level: int = 4,
                two_level: bool = False) -> int:
        arr = np.ascontiguousarray(arr)
        assert arr.ndim == 1
        dt = arr.dtype
        # chunked datasets are extendible, as PyTables / Cell Ranger write them: maximum dimensions present
        space = struct.pack("<BBBx4x", 1, 1, 1 if chunks else 0) + struct.pack("<Q", len(arr))
        if chunks:
            space += struct.pack("<Q", UNDEF)
        msgs = [self._msg(1, space), self._msg(3, self._dtype_msg(dt))]
        if chunks is None:
            addr = self.alloc(arr.tobytes()) if len(arr) else UNDEF
            msgs.append(self._msg(8, struct.pack("<BBQQ", 3, 1, addr, arr.nbytes)))
            return self._ohdr(msgs)
        filt = b""
        nf = 0
        # filter descriptions as libhdf5 writes them: with the filter's name, padded to 8 bytes
        if shuffle:
            filt += struct.pack("<HHHH8sI4x", 2, 8, 1, 1, b"shuffle\0", dt.itemsize); nf += 1
        if level:
            filt += struct.pack("<HHHH8sI4x", 1, 8, 1, 1, b"deflate\0", level); nf += 1
        if nf:
            msgs.append(self._msg(0x0B, struct.pack("<BB6x", 1, nf) + filt))
        entries = []
        for off in range(0, len(arr), chunks):
            block = np.zeros(chunks, dt)
            part = arr[off:off + chunks]
            block[:len(part)] = part
            raw = block.tobytes()
            if shuffle:
                raw = np.frombuffer(raw, np.uint8).reshape(chunks, dt.itemsize).T.tobytes()
            if level:
                raw = zlib.compress(raw, level)
            entries.append((self.alloc(raw), len(raw), off))

        def node(level_, items, last_off):
            # items: (child address, chunk size, first offset)
            body = struct.pack("<4sBBHQQ", b"TREE", 1, level_, len(items), UNDEF, UNDEF)
            for child, size, off in items:
                body += struct.pack("<IIQQ", size, 0, off, 0) + struct.pack("<Q", child)
            body += struct.pack("<IIQQ", 0, 0, last_off, 0)
            return self.alloc(body)

        end = (len(arr) + chunks - 1) // chunks * chunks
        if two_level and len(entries) >= 2:
            h = len(entries) // 2
            a = node(0, entries[:h], entries[h][2])
            b = node(0, entries[h:], end)
            root = node(1, [(a, 0, entries[0][2]), (b, 0, entries[h][2])], end)
        else:
            root = node(0, entries, end) if entries else UNDEF
        msgs.append(self._msg(8, struct.pack("<BBBQII", 3, 2, 2, root, chunks, dt.itemsize)))
        return self._ohdr(msgs)

    # ---- groups ---------------------------------------------------------------------------------
    def group(self, members: dict[str, int]) -> tuple[int, int, int]:
        """-> (object header, btree, heap) of a group linking `members` (name -> object header)."""
        names = sorted(members)
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += n.encode() + b"\0"
            heap += b"\0" * (-len(heap) % 8)
        hdata = self.alloc(bytes(heap))
        haddr = self.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), UNDEF, hdata))
        snod = struct.pack("<4sBxH", b"SNOD", 1, len(names))
        for n in names:
            snod += struct.pack("<QQII16x", offs[n], members[n], 0, 0)
        saddr = self.alloc(snod)
        tree = struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, saddr, offs[names[-1]] if names else 0)
        taddr = self.alloc(tree)
        oh = self._ohdr([self._msg(0x11, struct.pack("<QQ", taddr, haddr))])
        return oh, taddr, haddr

    def group2(self, members: dict[str, int]) -> tuple[int, int, int]:
        """the same group in the HDF5 1.8 style: object header v2 with a link-info message (no
        fractal heap: compact storage) and one link message per member"""
        body = b"\x02" + struct.pack("<HB", 18, 0) + struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)
        for n in sorted(members):
            nb = n.encode()
            link = struct.pack("<BBB", 1, 0, len(nb)) + nb + struct.pack("<Q", members[n])
            body += b"\x06" + struct.pack("<HB", len(link), 0) + link
        hdr = b"OHDR" + struct.pack("<BBH", 2, 0x01, len(body)) + body + b"\0\0\0\0"      # checksum not verified by the reader
        return self.alloc(hdr), UNDEF, UNDEF

    def finish(self, root: tuple[int, int, int]) -> bytes:
        oh, taddr, haddr = root
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, 4, 16, 0)
        sb += struct.pack("<QQQQ", self.ub, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", taddr, haddr)
        assert len(sb) == 96, len(sb)
        self.buf[:96] = sb
        return b"\0" * self.ub + bytes(self.buf)


def write_10x(path: str, barcodes: list[str], n_genes_per_cell: list[int], n_features: int = 50, version: int = 3,
              userblock: int = 0, explicit_zeros: int = 0, chunks: int = 7, seed: int = 0, new_groups: bool = False):
    """A Cell Ranger style matrix file: CSC over barcodes under /matrix (version 3) or /GRCh38
    (version 2).  Cell j gets n_genes_per_cell[j] entries > 0 plus `explicit_zeros` stored zeros."""
    rng = np.random.default_rng(seed)
    data, indices, indptr = [], [], [0]
    for g in n_genes_per_cell:
        rows = np.sort(rng.choice(n_features, g + explicit_zeros, replace=False))
        vals = rng.integers(1, 9, g + explicit_zeros)
        if explicit_zeros:
            vals[rng.choice(g + explicit_zeros, explicit_zeros, replace=False)] = 0
        data += list(vals); indices += list(rows); indptr.append(len(data))
    w = Writer(userblock)
    width = max(len(b) for b in barcodes)
    ds = {
        "barcodes": w.dataset(np.array([b.encode() for b in barcodes], f"S{width}"), chunks=chunks, two_level=True),
        "data": w.dataset(np.array(data, np.int32), chunks=chunks * 3),
        "indices": w.dataset(np.array(indices, np.int64), chunks=chunks * 3, shuffle=False),
        "indptr": w.dataset(np.array(indptr, np.int64), chunks=chunks, level=0),
        "shape": w.dataset(np.array([n_features, len(barcodes)], np.int32)),
    }
    mk = w.group2 if new_groups else w.group
    grp = mk(ds)
    root = mk({"matrix" if version == 3 else "GRCh38": grp[0]})
    if new_groups:
        # superblock 2: signature, version, sizes, flags, base, extension, end of file, root header, checksum
        sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBB", 2, 8, 8, 0) + struct.pack("<QQQQ", userblock, UNDEF, len(w.buf), root[0]) + b"\0" * 4
        w.buf[:len(sb)] = sb
        open(path, "wb").write(b"\0" * userblock + bytes(w.buf))
        return
    open(path, "wb").write(w.finish(root))
