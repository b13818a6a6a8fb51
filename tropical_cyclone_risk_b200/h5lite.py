"""A small read-only HDF5 reader -- enough of the format to open the NetCDF-4 files the reference reads
and writes, in an image that has no netCDF4 / h5py / libhdf5:

  * the static inputs shipped with the reference: intensity/data/{bathymetry,land,mld_climatology,
    strat_climatology}.nc (intensity/geo.py:9-33, intensity/ocean.py:11-60) -- superblock version 0,
    symbol-table groups, version-1 object headers, chunked + shuffle + deflate datasets;
  * the files xarray writes through netCDF4 (the env_wnd_* / thermo_* caches of track/env_wind.py:154-158 and
    thermo/calc_thermo.py:103-116, the track files of util/compute.py:244-268, the samples under notebooks/data):
    superblock version 2, version-2 object headers, links stored compactly (link messages) or densely (fractal
    heap), contiguous or chunked data.

Implemented from the published HDF5 File Format Specification (version 3.0); nothing here is derived from the
reference.  Supported: fixed-point and IEEE floating-point datasets of any rank (little / big endian), fixed-length
strings, variable-length strings through the global heap; data layouts compact / contiguous / chunked (version-1
chunk B-tree) with the deflate, shuffle and fletcher32 filters; attributes of those types, compact or dense (fractal heap).  Anything else raises NotImplementedError rather than guessing.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


class Dataset:
    def __init__(self, f, name, msgs):
        self.file, self.name, self._msgs = f, name, msgs
        self.shape = self.dtype = None
        self._layout = self._filters = None
        self._fill = None
        self.attrs = {}
        self._vlen_str = False
        self._parse()

    # -- header messages ----------------------------------------------------------------------------
    def _parse(self):
        f = self.file
        for mtype, body in self._msgs:
            if mtype == 0x01:
                self.shape = f._dataspace(body)
            elif mtype == 0x03:
                self.dtype, self._vlen_str = f._datatype(body)
            elif mtype == 0x08:
                self._layout = body
            elif mtype == 0x05:
                self._fill = self._fill_value(body)
            elif mtype == 0x0B:
                self._filters = f._filters(body)
            elif mtype == 0x0C:
                k, v = f._attribute(body)
                self.attrs[k] = v
            elif mtype == 0x15:
                self.attrs.update(f._dense_attributes(body, self.name))

    @property
    def is_dataset(self):
        return self.shape is not None and self.dtype is not None and self._layout is not None

    # -- data ---------------------------------------------------------------------------------------
    def read(self):
        """The whole dataset as a NumPy array (native byte order)."""
        f, b = self.file, self._layout
        ver = b[0]
        n = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        isz = self.dtype.itemsize
        if ver == 3:
            cls = b[1]
            if cls == 0:                                           # compact
                size = _u(b, 2, 2)
                raw = bytes(b[4:4 + size])
                return self._finish(np.frombuffer(raw, dtype=self.dtype, count=n))
            if cls == 1:                                           # contiguous
                addr, size = _u(b, 2, 8), _u(b, 10, 8)
                if addr == UNDEF:
                    return self._finish(self._filled(n))
                return self._finish(np.frombuffer(f.buf, dtype=self.dtype, count=n, offset=addr))
            if cls == 2:                                           # chunked, version-1 B-tree
                rank1 = b[2]
                btree = _u(b, 3, 8)
                cdims = [_u(b, 11 + 4 * i, 4) for i in range(rank1)]
                return self._finish(self._read_chunked(btree, cdims[:-1], n))
            raise NotImplementedError("data layout class %d" % cls)
        if ver in (1, 2):
            rank1, cls = b[1], b[2]
            off = 8
            addr = None
            if cls != 0:
                addr = _u(b, off, 8)
                off += 8
            dims = [_u(b, off + 4 * i, 4) for i in range(rank1)]
            if cls == 1:
                if addr == UNDEF:
                    return self._finish(self._filled(n))
                return self._finish(np.frombuffer(f.buf, dtype=self.dtype, count=n, offset=addr))
            if cls == 2:
                return self._finish(self._read_chunked(addr, dims[:-1], n))
            raise NotImplementedError("version-%d layout class %d" % (ver, cls))
        raise NotImplementedError("data layout message version %d (HDF5 1.10 chunk indexes)" % ver)

    @staticmethod
    def _fill_value(body):
        """Raw bytes of the dataset's fill value (fill value message, versions 1-3), or None."""
        ver = body[0]
        if ver in (1, 2):
            if ver == 2 and not body[3]:
                return None
            size = _u(body, 4, 4)
            return bytes(body[8:8 + size]) if size else None
        if ver == 3:
            if not (body[1] & 0x20):
                return None
            size = _u(body, 2, 4)
            return bytes(body[6:6 + size]) if size else None
        return None

    def _filled(self, n):
        """What unwritten storage reads as: the fill value if the dataset defines one, else zeros."""
        out = np.zeros(n, dtype=self.dtype)
        if self._fill is not None and len(self._fill) == self.dtype.itemsize and not self._vlen_str:
            out[:] = np.frombuffer(self._fill, dtype=self.dtype, count=1)[0]
        return out

    def _finish(self, flat):
        a = np.array(flat).reshape(self.shape if self.shape else ())
        if self._vlen_str:
            return self.file._vlen_strings(a)
        if a.dtype.byteorder == ">":
            a = a.astype(a.dtype.newbyteorder("="))
        return a

    def _read_chunked(self, btree, cdims, n):
        f = self.file
        shape = tuple(self.shape)
        rank = len(shape)
        if len(cdims) != rank:
            raise H5Error("chunk rank mismatch in %s" % self.name)
        out = self._filled(int(np.prod(shape, dtype=np.int64))).reshape(shape)
        if btree == UNDEF:
            return out.reshape(-1)
        csize = int(np.prod(cdims)) * self.dtype.itemsize
        for size, mask, offs, addr in f._chunk_btree(btree, rank):
            raw = f.buf[addr:addr + size]
            if self._filters:
                for k in range(len(self._filters) - 1, -1, -1):      # undo the pipeline back to front
                    fid, cd = self._filters[k]
                    if mask & (1 << k):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else self.dtype.itemsize
                        m = len(raw) // es
                        raw = np.frombuffer(raw, np.uint8, m * es).reshape(es, m).T.tobytes() + bytes(raw[m * es:])
                    elif fid == 3:
                        raw = raw[:-4]                               # fletcher32 checksum trailer
                    else:
                        raise NotImplementedError("HDF5 filter %d" % fid)
            if len(raw) < csize:
                raise H5Error("short chunk in %s" % self.name)
            chunk = np.frombuffer(raw, dtype=self.dtype, count=csize // self.dtype.itemsize).reshape(cdims)
            sel_out = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
            sel_in = tuple(slice(0, so.stop - so.start) for so in sel_out)
            out[sel_out] = chunk[sel_in]
        return out.reshape(-1)

    def __repr__(self):
        return "<h5lite.Dataset %s %s %s>" % (self.name, self.shape, self.dtype)


class File:
    """f = File(path); f.keys(); f[name] -> Dataset (.shape, .dtype, .attrs, .read()); f.groups[name] -> File-like group."""

    def __init__(self, path=None, _parent=None, _links=None, _attrs=None):
        if _parent is not None:
            self.buf, self.osz, self.lsz = _parent.buf, _parent.osz, _parent.lsz
            self._links, self.attrs = _links, _attrs
            return
        with open(path, "rb") as fh:
            self.buf = fh.read()
        if self.buf[:8] != SIGNATURE:
            raise H5Error("%s is not an HDF5 file (NetCDF-3 classic files are read with scipy.io.netcdf_file)" % path)
        ver = self.buf[8]
        if ver in (0, 1):
            self.osz, self.lsz = self.buf[13], self.buf[14]
            off = 24 + (4 if ver == 1 else 0)
            off += 4 * self.osz                                     # base, free-space, end-of-file, driver-info addresses
            root_hdr = _u(self.buf, off + self.osz, self.osz)       # root symbol-table entry: name offset, header address
        elif ver in (2, 3):
            self.osz, self.lsz = self.buf[9], self.buf[10]
            root_hdr = _u(self.buf, 12 + 3 * self.osz, self.osz)
        else:
            raise NotImplementedError("superblock version %d" % ver)
        if self.osz != 8 or self.lsz != 8:
            raise NotImplementedError("offset / length sizes other than 8 bytes")
        msgs = self._object_header(root_hdr)
        self._links, self.attrs = self._group_links(msgs)

    # -- mapping interface --------------------------------------------------------------------------
    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        return name in self._links

    def __getitem__(self, name):
        msgs = self._object_header(self._links[name])
        ds = Dataset(self, name, msgs)
        if ds.is_dataset:
            return ds
        links, attrs = self._group_links(msgs)
        return File(_parent=self, _links=links, _attrs=attrs)

    def variables(self):
        """{name: Dataset} of every dataset in this group."""
        out = {}
        for k in self._links:
            o = self[k]
            if isinstance(o, Dataset):
                out[k] = o
        return out

    # -- object headers -----------------------------------------------------------------------------
    def _object_header(self, addr):
        b = self.buf
        msgs = []
        if b[addr:addr + 4] == b"OHDR":                              # version 2
            flags = b[addr + 5]
            off = addr + 6
            if flags & 0x20:
                off += 16
            if flags & 0x10:
                off += 4
            nsz = 1 << (flags & 3)
            size0 = _u(b, off, nsz)
            off += nsz
            blocks = [(off, size0)]
            has_order = bool(flags & 0x04)
            while blocks:
                start, size = blocks.pop(0)
                p, end = start, start + size
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], _u(b, p + 1, 2), b[p + 3]
                    p += 4 + (2 if has_order else 0)
                    body = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        caddr, clen = _u(body, 0, 8), _u(body, 8, 8)
                        if b[caddr:caddr + 4] != b"OCHK":
                            raise H5Error("bad object header continuation")
                        blocks.append((caddr + 4, clen - 8))            # signature in front, checksum behind
                    elif mtype != 0:
                        msgs.append((mtype, body))
            return msgs
        if b[addr] != 1:
            raise H5Error("unknown object header at %d" % addr)
        nmsg, hsize = _u(b, addr + 2, 2), _u(b, addr + 8, 4)
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg + 64:
            start, size = blocks.pop(0)
            p, end = start, start + size
            while p + 8 <= end:
                mtype, msize = _u(b, p, 2), _u(b, p + 2, 2)
                body = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    blocks.append((_u(body, 0, 8), _u(body, 8, 8)))
                elif mtype != 0:
                    msgs.append((mtype, body))
        return msgs

    # -- groups -------------------------------------------------------------------------------------
    def _group_links(self, msgs):
        links, attrs = {}, {}
        for mtype, body in msgs:
            if mtype == 0x11:                                        # symbol table: B-tree + local heap
                self._symtab(_u(body, 0, 8), _u(body, 8, 8), links)
            elif mtype == 0x06:                                      # link message (compact storage)
                name, addr = self._link_message(body)
                if addr is not None:
                    links[name] = addr
            elif mtype == 0x02:                                      # link info: dense storage in a fractal heap
                flags = body[1]
                off = 2 + (8 if flags & 1 else 0)
                heap = _u(body, off, 8)
                if heap != UNDEF:
                    for obj in self._fractal_heap_objects(heap):
                        name, addr = self._link_message(obj)
                        if addr is not None:
                            links[name] = addr
            elif mtype == 0x0C:
                k, v = self._attribute(body)
                attrs[k] = v
            elif mtype == 0x15:
                attrs.update(self._dense_attributes(body, "group"))
        return links, attrs

    def _dense_attributes(self, body, what):
        """Attribute-info message: more than eight attributes live in a fractal heap.  Returns {name: value}; if the
        heap cannot be walked the object is refused (dropping attributes silently could lose a scale_factor)."""
        flags = body[1]
        heap = _u(body, 2 + (2 if flags & 1 else 0), 8)
        out = {}
        if heap == UNDEF:
            return out
        try:
            for obj in self._fractal_heap_objects(heap, "attr"):
                k, v = self._attribute(obj)
                out[k] = v
        except (H5Error, NotImplementedError, IndexError, ValueError, UnicodeDecodeError) as e:
            raise NotImplementedError("%s keeps its attributes in dense storage that this reader cannot walk (%s)" % (what, e))
        return out

    def _symtab(self, btree, heap, links):
        b = self.buf
        if b[heap:heap + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = _u(b, heap + 24, 8)

        def walk(node):
            if b[node:node + 4] == b"TREE":
                level, used = b[node + 5], _u(b, node + 6, 2)
                p = node + 24
                for i in range(used):
                    child = _u(b, p + 8, 8)
                    p += 16
                    walk(child)
            elif b[node:node + 4] == b"SNOD":
                nsym = _u(b, node + 6, 2)
                p = node + 8
                for i in range(nsym):
                    noff, haddr = _u(b, p, 8), _u(b, p + 8, 8)
                    s = data + noff
                    e = b.index(b"\0", s)
                    links[b[s:e].decode("utf-8")] = haddr
                    p += 40
            else:
                raise H5Error("bad group B-tree node")

        walk(btree)

    def _link_message(self, body):
        flags = body[1]
        off = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[off]
            off += 1
        if flags & 0x04:
            off += 8
        if flags & 0x10:
            off += 1
        nlen_sz = 1 << (flags & 3)
        nlen = _u(body, off, nlen_sz)
        off += nlen_sz
        name = bytes(body[off:off + nlen]).decode("utf-8")
        off += nlen
        if ltype != 0:
            return name, None                                        # soft / external links are not followed
        return name, _u(body, off, 8)

    # -- fractal heap (dense link / attribute storage) ------------------------------------------------
    def _fractal_heap_objects(self, addr, kind="link"):
        """Every managed object of a fractal heap, in storage order: the link messages of a dense group (kind "link")
        or the attribute messages of an object with more than eight attributes (kind "attr")."""
        span = self._link_span if kind == "link" else self._attribute_span
        starts = (1,) if kind == "link" else (1, 2, 3)                 # message version bytes; free space is zero-filled
        b = self.buf
        if b[addr:addr + 4] != b"FRHP":
            raise H5Error("bad fractal heap header")
        p = addr + 5
        heap_id_len, io_filter_len, flags = _u(b, p, 2), _u(b, p + 2, 2), b[p + 4]
        p += 5
        p += 4                                                         # maximum size of managed objects
        p += 8 + 8                                                     # next huge id, huge-object B-tree
        p += 8 + 8                                                     # free space, free-space manager
        p += 8 + 8 + 8 + 8                                             # managed space, allocated, iterator offset, number of managed objects
        n_managed = _u(b, p - 8, 8)
        p += 8 + 8 + 8 + 8                                             # huge size / count, tiny size / count
        table_width = _u(b, p, 2)
        start_block = _u(b, p + 2, 8)
        max_direct = _u(b, p + 10, 8)
        max_heap_bits = _u(b, p + 18, 2)
        start_rows = _u(b, p + 20, 2)
        root = _u(b, p + 22, 8)
        cur_rows = _u(b, p + 30, 2)
        if io_filter_len:
            raise NotImplementedError("filtered fractal heap")
        off_bytes = (max_heap_bits + 7) // 8
        checksummed = bool(flags & 2)
        objs = []

        def direct(baddr, bsize):
            if b[baddr:baddr + 4] != b"FHDB":
                raise H5Error("bad fractal heap direct block")
            q = baddr + 5 + 8 + off_bytes + (4 if checksummed else 0)
            end = baddr + bsize
            # managed objects are packed back to back from the start of the block; free space is zero-filled
            while q < end and b[q] in starts:
                used = span(q)[-1]
                if used <= 0 or q + used > end:
                    raise H5Error("fractal heap object overruns its block")
                objs.append(b[q:q + used])
                q += used

        def indirect(iaddr, nrows):
            if b[iaddr:iaddr + 4] != b"FHIB":
                raise H5Error("bad fractal heap indirect block")
            q = iaddr + 5 + 8 + off_bytes
            max_direct_rows = (max_direct // start_block).bit_length() + 1
            for r in range(nrows):
                bsize = start_block * (1 if r < 2 else 1 << (r - 1))
                for c in range(table_width):
                    child = _u(b, q, 8)
                    q += 8
                    if child == UNDEF:
                        continue
                    if r < max_direct_rows:
                        direct(child, bsize)
                    else:
                        indirect(child, (bsize // start_block // table_width).bit_length())

        if root != UNDEF:
            if cur_rows == 0:
                direct(root, start_block)
            else:
                indirect(root, cur_rows)
        return objs

    def _link_span(self, q):
        """(name, address, bytes used) of the link message starting at buffer offset q."""
        b = self.buf
        flags = b[q + 1]
        off = q + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[off]
            off += 1
        if flags & 0x04:
            off += 8
        if flags & 0x10:
            off += 1
        nlen_sz = 1 << (flags & 3)
        nlen = _u(b, off, nlen_sz)
        off += nlen_sz
        name = b[off:off + nlen].decode("utf-8")
        off += nlen
        addr = None
        if ltype == 0:
            addr = _u(b, off, 8)
            off += 8
        elif ltype == 1:
            off += 2 + _u(b, off, 2)
        else:
            off += 2 + _u(b, off, 2)
        return name, addr, off - q

    def _attribute_span(self, q):
        """(bytes used,) of the attribute message starting at buffer offset q (dense attribute storage)."""
        b = self.buf
        ver = b[q]
        nsz, tsz, ssz = _u(b, q + 2, 2), _u(b, q + 4, 2), _u(b, q + 6, 2)
        off = q + 8 + (1 if ver == 3 else 0)
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        if nsz == 0 or nsz > 1024 or tsz == 0 or tsz > 4096 or ssz > 4096:
            raise H5Error("implausible attribute message in a fractal heap")
        t0 = off + pad(nsz)
        esize = _u(b, t0 + 4, 4)                                        # element size: every datatype class keeps it here
        shape = self._dataspace(b[t0 + pad(tsz):t0 + pad(tsz) + ssz]) if ssz else ()
        n = int(np.prod(shape, dtype=np.int64)) if shape else (0 if shape is None else 1)
        end = t0 + pad(tsz) + pad(ssz) + n * esize
        return (end - q,)

    # -- message decoders -----------------------------------------------------------------------------
    def _dataspace(self, body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            if body[3] == 2:
                return None                                           # null dataspace
            off = 4
        else:
            raise NotImplementedError("dataspace message version %d" % ver)
        return tuple(_u(body, off + 8 * i, 8) for i in range(rank))

    def _datatype(self, body):
        cls, ver = body[0] & 0x0F, body[0] >> 4
        bits0 = body[1]
        size = _u(body, 4, 4)
        order = ">" if (bits0 & 1) else "<"
        if cls == 0:
            kind = "i" if (bits0 & 0x08) else "u"
            return np.dtype("%s%s%d" % (order, kind, size)), False
        if cls == 1:
            if size not in (2, 4, 8):
                raise NotImplementedError("%d-byte floating point" % size)
            return np.dtype("%sf%d" % (order, size)), False
        if cls == 3:
            return np.dtype("S%d" % size), False
        if cls == 9:
            if (bits0 & 0x0F) == 1:                                   # variable-length string
                return np.dtype("V%d" % size), True
            raise NotImplementedError("variable-length sequence datatype")
        raise NotImplementedError("HDF5 datatype class %d" % cls)

    def _filters(self, body):
        ver, n = body[0], body[1]
        off = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = _u(body, off, 2)
            off += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = _u(body, off, 2)
                off += 2
            off += 2                                                   # flags
            ncd = _u(body, off, 2)
            off += 2
            if nlen:
                off += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [_u(body, off + 4 * i, 4) for i in range(ncd)]
            off += 4 * ncd
            if ver == 1 and ncd % 2:
                off += 4
            out.append((fid, cd))
        return out

    def _attribute(self, body):
        ver = body[0]
        nsz, tsz, ssz = _u(body, 2, 2), _u(body, 4, 2), _u(body, 6, 2)
        off = 8 + (1 if ver == 3 else 0)
        pad = (lambda x: (x + 7) // 8 * 8) if ver == 1 else (lambda x: x)
        name = bytes(body[off:off + nsz]).split(b"\0")[0].decode("utf-8")
        off += pad(nsz)
        try:
            dtype, vlen = self._datatype(body[off:off + tsz])
            shape = self._dataspace(body[off + pad(tsz):off + pad(tsz) + ssz])
        except NotImplementedError:
            return name, None
        off += pad(tsz) + pad(ssz)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if shape is None:
            return name, None
        raw = bytes(body[off:off + n * dtype.itemsize])
        if len(raw) < n * dtype.itemsize:
            return name, None
        a = np.frombuffer(raw, dtype=dtype, count=n)
        if vlen:
            a = self._vlen_strings(a)
        elif dtype.kind == "S":
            a = np.array([x.split(b"\0")[0].decode("utf-8", "replace") for x in a])
        elif dtype.byteorder == ">":
            a = a.astype(dtype.newbyteorder("="))
        if not shape:
            return name, a[0]
        return name, a.reshape(shape)

    def _vlen_strings(self, a):
        """Variable-length strings: each element is {length u4, global heap collection address u8, object index u4}."""
        b = self.buf
        flat = a.reshape(-1)
        out = []
        for rec in flat:
            r = rec.tobytes()
            ln, gaddr, idx = _u(r, 0, 4), _u(r, 4, 8), _u(r, 12, 4)
            out.append(self._global_heap_object(gaddr, idx)[:ln].decode("utf-8", "replace") if gaddr not in (0, UNDEF) else "")
        return np.array(out, dtype=object).reshape(a.shape)

    def _global_heap_object(self, gaddr, idx):
        b = self.buf
        if b[gaddr:gaddr + 4] != b"GCOL":
            raise H5Error("bad global heap collection")
        size = _u(b, gaddr + 8, 8)
        p, end = gaddr + 16, gaddr + size
        while p + 16 <= end:
            oid, osize = _u(b, p, 2), _u(b, p + 8, 8)
            if oid == 0:
                break
            if oid == idx:
                return bytes(b[p + 16:p + 16 + osize])
            p += 16 + (osize + 7) // 8 * 8
        raise H5Error("global heap object %d not found" % idx)

    # -- chunk B-tree (version 1, node type 1) ----------------------------------------------------------
    def _chunk_btree(self, node, rank):
        b = self.buf
        if b[node:node + 4] != b"TREE" or b[node + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level, used = b[node + 5], _u(b, node + 6, 2)
        p = node + 24
        ksz = 8 + 8 * (rank + 1)
        for i in range(used):
            size, mask = _u(b, p, 4), _u(b, p + 4, 4)
            offs = [_u(b, p + 8 + 8 * d, 8) for d in range(rank)]
            child = _u(b, p + ksz, 8)
            p += ksz + 8
            if level == 0:
                yield size, mask, offs, child
            else:
                yield from self._chunk_btree(child, rank)


def read_netcdf4(path, names=None):
    """{variable: ndarray} of a NetCDF-4 file's root group (all variables, or `names`)."""
    f = File(path)
    out = {}
    for k in (names or f.keys()):
        o = f[k]
        if isinstance(o, Dataset):
            out[k] = o.read()
    return out
