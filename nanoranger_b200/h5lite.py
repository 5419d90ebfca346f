"""Minimal pure-Python HDF5 reader: just enough to read the datasets of a 10x Genomics
``*_feature_bc_matrix.h5`` (``barcodes``, ``indptr``, ``data``) without h5py / scanpy, which the
reference's ``.h5`` whitelist inputs need (``utils.py:606-610`` ``write_bc_5p10X``,
``utils.py:1116-1132`` ``write_bc_3p10XTCR_nuc``: ``sc.read_10x_h5`` + ``sc.pp.filter_cells``).

Covers what libhdf5 writes with its default ("earliest") format bounds, which is what Cell Ranger
(and PyTables before it) produce: superblock 0 / 1 (2 / 3 accepted), a user block in front, old
style groups (symbol table: B-tree v1 + local heap + SNOD), object headers v1 (v2 with compact
link messages accepted), dataspace v1 / v2, fixed-point / floating-point / fixed-length string
datatypes, data layout v1-v3 (compact, contiguous, chunked through a B-tree v1), and the deflate,
shuffle and fletcher32 filters.  Anything else (dense groups in fractal heaps, layout v4 chunk
indexes, variable-length strings, other filters) raises ``H5Unsupported`` naming the feature.
Host-side input parsing only: nothing on the matcher's path depends on it.
"""
from __future__ import annotations

import mmap
import zlib

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"


class H5Unsupported(NotImplementedError):
    pass


class H5Error(ValueError):
    pass


class H5Lite:
    def __init__(self, path: str):
        self._fh = open(path, "rb")
        try:
            self.buf = mmap.mmap(self._fh.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError:
            self._fh.close()
            raise H5Error(f"{path}: empty file")
        self.path = path
        self._superblock()

    def close(self):
        try:
            self.buf.close()
        finally:
            self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- primitives -------------------------------------------------------------------------
    def _u(self, off: int, size: int) -> int:
        if off < 0 or off + size > len(self.buf):
            raise H5Error(f"{self.path}: read of {size} bytes at {off} beyond the end of the file")
        return int.from_bytes(self.buf[off:off + size], "little")

    def _addr(self, off: int) -> int | None:
        v = self._u(off, self.O)
        return None if v == (1 << (8 * self.O)) - 1 else v + self.base

    # ---- superblock -------------------------------------------------------------------------
    def _superblock(self):
        off, n = 0, len(self.buf)
        while off + 8 <= n and self.buf[off:off + 8] != SIG:
            off = 512 if off == 0 else off * 2          # 0, 512, 1024, ... (user block sizes)
        if off + 8 > n:
            raise H5Error(f"{self.path}: not an HDF5 file (no signature)")
        ver = self.buf[off + 8]
        if ver in (0, 1):
            self.O, self.L = self.buf[off + 13], self.buf[off + 14]
            p = off + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.O)               # = size of the user block, if there is one
            p += 4 * self.O                              # base, free space, end of file, driver info
            # root group symbol table entry
            self.root = self._u(p + self.O, self.O) + self.base
        elif ver in (2, 3):
            self.O, self.L = self.buf[off + 9], self.buf[off + 10]
            p = off + 12
            self.base = self._u(p, self.O)
            self.root = self._u(p + 3 * self.O, self.O) + self.base
        else:
            raise H5Unsupported(f"{self.path}: superblock version {ver}")
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5Unsupported(f"{self.path}: offsets of {self.O} / lengths of {self.L} bytes")

    # ---- object headers ---------------------------------------------------------------------
    def _messages(self, addr: int):
        """[(type, data offset, size)] of the object header at addr (continuations followed)."""
        out = []
        if self.buf[addr:addr + 4] == b"OHDR":
            flags = self.buf[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            csz = 1 << (flags & 3)
            size0 = self._u(p, csz)
            p += csz
            blocks = [(p, size0)]
            track = bool(flags & 0x04)
            while blocks:
                b, sz = blocks.pop(0)
                end = b + sz
                while b + 4 <= end:
                    t, s = self.buf[b], self._u(b + 1, 2)
                    b += 4 + (2 if track else 0)
                    if t == 0x10:
                        coff, clen = self._addr(b), self._u(b + self.O, self.L)
                        if self.buf[coff:coff + 4] != b"OCHK":
                            raise H5Error(f"{self.path}: bad object header continuation")
                        blocks.append((coff + 4, clen - 8))
                    elif t != 0:
                        out.append((t, b, s))
                    b += s
            return out
        if self.buf[addr] != 1:
            raise H5Unsupported(f"{self.path}: object header version {self.buf[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        blocks = [(addr + 16, self._u(addr + 8, 4))]
        while blocks and len(out) < nmsg + 64:
            b, sz = blocks.pop(0)
            end = b + sz
            while b + 8 <= end:
                t, s = self._u(b, 2), self._u(b + 2, 2)
                if (self.buf[b + 4] & 0x02) and t in (0x01, 0x03, 0x08, 0x0B):
                    raise H5Unsupported(f"{self.path}: shared header message (type {t:#x})")
                b += 8
                if t == 0x10:
                    blocks.append((self._addr(b), self._u(b + self.O, self.L)))
                elif t != 0:
                    out.append((t, b, s))
                b += (s + 7) & ~7
        return out

    # ---- groups -----------------------------------------------------------------------------
    def members(self, addr: int | None = None) -> dict[str, int]:
        """name -> object header address of the links of the group at addr (root by default)."""
        addr = self.root if addr is None else addr
        out: dict[str, int] = {}
        for t, p, s in self._messages(addr):
            if t == 0x11:                                 # symbol table: B-tree v1 + local heap
                btree, heap = self._addr(p), self._addr(p + self.O)
                if self.buf[heap:heap + 4] != b"HEAP":
                    raise H5Error(f"{self.path}: bad local heap")
                hdata = self._addr(heap + 8 + 2 * self.L)
                self._group_btree(btree, hdata, out)
            elif t == 0x06:                               # link message (new-style compact group)
                fl = self.buf[p + 1]
                q = p + 2
                ltype = 0
                if fl & 0x08:
                    ltype = self.buf[q]; q += 1
                if fl & 0x04:
                    q += 8
                if fl & 0x10:
                    q += 1
                nsz = 1 << (fl & 3)
                nlen = self._u(q, nsz); q += nsz
                name = bytes(self.buf[q:q + nlen]).decode("utf-8", "replace"); q += nlen
                if ltype == 0:
                    out[name] = self._addr(q)
            elif t == 0x02:                               # link info: dense storage?
                fl = self.buf[p + 1]
                q = p + 2 + (8 if fl & 1 else 0)
                if self._addr(q) is not None:
                    raise H5Unsupported(f"{self.path}: group with links in a fractal heap (dense storage)")
        return out

    def _group_btree(self, addr: int, hdata: int, out: dict):
        if self.buf[addr:addr + 4] != b"TREE" or self.buf[addr + 4] != 0:
            raise H5Error(f"{self.path}: bad group B-tree node at {addr}")
        level, n = self.buf[addr + 5], self._u(addr + 6, 2)
        p = addr + 8 + 2 * self.O
        for k in range(n):
            child = self._addr(p + self.L + k * (self.L + self.O))
            if level > 0:
                self._group_btree(child, hdata, out)
                continue
            if self.buf[child:child + 4] != b"SNOD":
                raise H5Error(f"{self.path}: bad symbol table node at {child}")
            ns = self._u(child + 6, 2)
            esz = 2 * self.O + 24
            for e in range(ns):
                q = child + 8 + e * esz
                noff = hdata + self._u(q, self.O)
                end = self.buf.find(b"\0", noff)
                out[bytes(self.buf[noff:end]).decode("utf-8", "replace")] = self._u(q + self.O, self.O) + self.base

    def lookup(self, path: str) -> int:
        addr = self.root
        for part in [x for x in path.split("/") if x]:
            m = self.members(addr)
            if part not in m:
                raise KeyError(f"{self.path}: no object '{part}' on the way to '{path}' (have {sorted(m)})")
            addr = m[part]
        return addr

    # ---- datasets ---------------------------------------------------------------------------
    def _dtype(self, p: int) -> np.dtype:
        cls, bits0 = self.buf[p] & 0x0F, self.buf[p + 1]
        size = self._u(p + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        if cls == 1:
            return np.dtype(f"{order}f{size}")
        if cls == 3:
            return np.dtype(f"S{size}")
        raise H5Unsupported(f"{self.path}: datatype class {cls}" +
                            (" (variable-length)" if cls == 9 else ""))

    def read(self, path: str) -> np.ndarray:
        """The whole dataset at `path` as a numpy array (C order)."""
        addr = self.lookup(path)
        shape = dtype = layout = None
        filters: list[tuple[int, list[int]]] = []
        for t, p, s in self._messages(addr):
            if t == 0x01:
                ver, rank = self.buf[p], self.buf[p + 1]
                q = p + (8 if ver == 1 else 4)
                shape = tuple(self._u(q + k * self.L, self.L) for k in range(rank))
            elif t == 0x03:
                dtype = self._dtype(p)
            elif t == 0x08:
                layout = p
            elif t == 0x0B:
                filters = self._filters(p)
        if shape is None or dtype is None or layout is None:
            raise H5Error(f"{self.path}: '{path}' is not a dataset")
        n = int(np.prod(shape)) if shape else 1
        ver = self.buf[layout]
        if ver in (1, 2):
            ndim, cls = self.buf[layout + 1], self.buf[layout + 2]
            q = layout + 8
            daddr = None
            if cls != 0:
                daddr = self._addr(q); q += self.O
            dims = [self._u(q + 4 * k, 4) for k in range(ndim)]
            q += 4 * ndim
            if cls == 0:
                csize = self._u(q, 4)
                return np.frombuffer(self.buf[q + 4:q + 4 + csize], dtype, n).reshape(shape).copy()
            if cls == 1:
                return self._contiguous(daddr, dtype, shape, n)
            return self._chunked(daddr, dims, dtype, shape, filters)
        if ver == 3:
            cls = self.buf[layout + 1]
            if cls == 0:
                csize = self._u(layout + 2, 2)
                return np.frombuffer(self.buf[layout + 4:layout + 4 + csize], dtype, n).reshape(shape).copy()
            if cls == 1:
                return self._contiguous(self._addr(layout + 2), dtype, shape, n)
            if cls == 2:
                ndim = self.buf[layout + 2]
                q = layout + 3 + self.O
                dims = [self._u(q + 4 * k, 4) for k in range(ndim)]
                return self._chunked(self._addr(layout + 3), dims, dtype, shape, filters)
            raise H5Unsupported(f"{self.path}: data layout class {cls}")
        raise H5Unsupported(f"{self.path}: data layout message version {ver} (written with libver='latest'?)")

    def _contiguous(self, daddr, dtype, shape, n):
        if daddr is None:
            return np.zeros(shape, dtype)
        return np.frombuffer(self.buf[daddr:daddr + n * dtype.itemsize], dtype, n).reshape(shape).copy()

    def _filters(self, p: int):
        ver, nf = self.buf[p], self.buf[p + 1]
        q = p + (8 if ver == 1 else 2)
        out = []
        for _ in range(nf):
            fid = self._u(q, 2); q += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = self._u(q, 2); q += 2
            q += 2                                        # flags
            ncd = self._u(q, 2); q += 2
            q += (nlen + 7) & ~7 if ver == 1 else nlen
            cd = [self._u(q + 4 * k, 4) for k in range(ncd)]
            q += 4 * ncd
            if ver == 1 and ncd % 2:
                q += 4
            out.append((fid, cd))
        return out

    def _chunked(self, btree, dims, dtype, shape, filters):
        rank = len(shape)
        if len(dims) != rank + 1:
            raise H5Error(f"{self.path}: chunk dimensionality {len(dims)} for rank {rank}")
        cshape = tuple(dims[:rank])
        out = np.zeros(shape, dtype)
        if btree is None:
            return out
        chunks: list[tuple[int, int, int, tuple]] = []
        self._chunk_btree(btree, rank, chunks)
        cbytes = int(np.prod(cshape)) * dtype.itemsize
        for caddr, csize, fmask, coff in chunks:
            raw = bytes(self.buf[caddr:caddr + csize])
            for k in range(len(filters) - 1, -1, -1):
                if fmask & (1 << k):
                    continue
                fid, cd = filters[k]
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    m = len(raw) // es
                    raw = np.frombuffer(raw[:m * es], np.uint8).reshape(es, m).T.tobytes() + raw[m * es:]
                elif fid == 3:
                    raw = raw[:-4]
                else:
                    raise H5Unsupported(f"{self.path}: filter id {fid}" + (" (blosc)" if fid == 32001 else ""))
            if len(raw) < cbytes:
                raise H5Error(f"{self.path}: chunk of {len(raw)} bytes, expected {cbytes}")
            block = np.frombuffer(raw, dtype, int(np.prod(cshape))).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(coff, cshape, shape))
            out[sl] = block[tuple(slice(0, x.stop - x.start) for x in sl)]
        return out

    def _chunk_btree(self, addr, rank, chunks):
        if self.buf[addr:addr + 4] != b"TREE" or self.buf[addr + 4] != 1:
            raise H5Error(f"{self.path}: bad chunk B-tree node at {addr}")
        level, n = self.buf[addr + 5], self._u(addr + 6, 2)
        ksz = 8 + 8 * (rank + 1)
        p = addr + 8 + 2 * self.O
        for k in range(n):
            q = p + k * (ksz + self.O)
            child = self._addr(q + ksz)
            if level > 0:
                self._chunk_btree(child, rank, chunks)
            else:
                coff = tuple(self._u(q + 8 + 8 * d, 8) for d in range(rank))
                chunks.append((child, self._u(q, 4), self._u(q + 4, 4), coff))


def read_10x_h5_barcodes(path: str, min_genes: int) -> list[str]:
    """Barcodes of the cells with at least `min_genes` genes detected (entries > 0 in the cell's
    column), in file order: what ``sc.pp.filter_cells(sc.read_10x_h5(path), min_genes=...)`` leaves
    in ``adata.obs.index`` (reference ``utils.py:607-610, 1119-1123``).  Cell Ranger >= 3 keeps the
    matrix under ``/matrix``, Cell Ranger 2 under ``/<genome>``."""
    with H5Lite(path) as f:
        top = f.members()
        grp = "matrix" if "matrix" in top else None
        if grp is None:
            for name, addr in top.items():
                try:
                    if {"barcodes", "indptr", "data"} <= set(f.members(addr)):
                        grp = name
                        break
                except (H5Error, H5Unsupported):
                    continue
        if grp is None:
            raise H5Error(f"{path}: no 10x matrix group (barcodes / indptr / data) found")
        barcodes = f.read(f"{grp}/barcodes")
        indptr = f.read(f"{grp}/indptr").astype(np.int64)
        data = f.read(f"{grp}/data")
    if len(indptr) != len(barcodes) + 1:
        raise H5Error(f"{path}: indptr has {len(indptr)} entries for {len(barcodes)} barcodes")
    cs = np.concatenate([[0], np.cumsum(data > 0, dtype=np.int64)])
    n_genes = cs[indptr[1:]] - cs[indptr[:-1]]
    keep = n_genes >= min_genes
    return [b.decode("ascii", "replace") for b in barcodes[keep]]
