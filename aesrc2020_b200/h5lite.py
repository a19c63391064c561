"""h5lite -- a small pure-Python reader (and a matching minimal writer) for the HDF5 subset Keras weight files use.

Why: the reference checkpoints with `model.save("%s/%03d.h5")` (train.py:35) and warm-starts with
`load_weights(raw, by_name=True, skip_mismatch=True)` (model.py:183, 416-417); h5py is not in this image, so the
container is parsed here.

Subset read (what h5py/libhdf5 >= 1.8 write with the default `libver='earliest'`, i.e. every Keras 2.x file):
  superblock v0 / v1; "old style" groups (symbol-table message -> B-tree v1 + local heap + SNOD nodes, any depth);
  object headers v1 with continuation blocks (v2 "OHDR" headers with link messages are read too: files written with
  libver='latest'); dataspace v1 / v2; datatypes: IEEE float 16/32/64, fixed-point 8..64 bit, fixed-length strings;
  data layout v3 (and the older v1 / v2 encodings) contiguous or compact; chunked layout WITHOUT filters (B-tree v1
  chunk index); attribute messages v1 / v2 / v3 holding scalars or arrays of the types above.
Not read (raises H5Error with the feature's name): filtered / compressed chunks, variable-length data, dense attribute
storage, virtual / external storage.  Variable-length *attributes* (e.g. Keras' `backend`, `keras_version` written
from Python str) are skipped, not errors: the weights do not depend on them.

The writer emits superblock v0, old-style groups (one B-tree level), object headers v1, contiguous datasets and v1
attribute messages -- the same structures the reader has to understand in a real Keras file -- so a round trip
exercises every parser above except continuation blocks and multi-level B-trees, which have their own unit tests
(tests/test_h5lite.py builds those byte layouts by hand).

Keras layout (keras/engine/saving.py `save_weights_to_hdf5_group`): the weights group -- the file root for
`model.save_weights`, `/model_weights` for `model.save` -- has attribute `layer_names` (array of fixed-length byte
strings, split into `layer_names0`, `layer_names1`, ... when larger than 64 KB); every layer is a sub-group with
attribute `weight_names`; each weight is the dataset `<layer>/<weight_name>` (weight names contain '/', i.e. nested
groups: `conv2d_1/conv2d_1/kernel:0`).
"""
from __future__ import annotations

import struct
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ====================================================================================== reader
class _Datatype:
    def __init__(self, np_dtype: Optional[np.dtype], size: int, vlen: bool = False, desc: str = ""):
        self.np_dtype, self.size, self.vlen, self.desc = np_dtype, size, vlen, desc


def _parse_datatype(b: bytes, off: int = 0) -> Tuple[_Datatype, int]:
    """Datatype message body -> (_Datatype, bytes consumed)."""
    cv = b[off]
    cls, ver = cv & 0x0F, cv >> 4
    bits0, bits1, bits2 = b[off + 1], b[off + 2], b[off + 3]
    size = struct.unpack_from("<I", b, off + 4)[0]
    p = off + 8
    if cls == 0:                                      # fixed-point
        order = ">" if bits0 & 1 else "<"
        signed = bool(bits0 & 0x08)
        p += 4
        if size not in (1, 2, 4, 8):
            raise H5Error("fixed-point size %d" % size)
        return _Datatype(np.dtype("%s%s%d" % (order, "i" if signed else "u", size)), size), p - off
    if cls == 1:                                      # floating point
        order = ">" if bits0 & 1 else "<"
        p += 12
        if size not in (2, 4, 8):
            raise H5Error("float size %d" % size)
        return _Datatype(np.dtype("%sf%d" % (order, size)), size), p - off
    if cls == 3:                                      # fixed-length string
        return _Datatype(np.dtype("S%d" % size), size), p - off
    if cls == 9:                                      # variable length: skipped by the callers
        base, used = _parse_datatype(b, p)
        return _Datatype(None, size, vlen=True, desc="vlen"), p + used - off
    raise H5Error("datatype class %d (version %d) is not supported" % (cls, ver))


def _parse_dataspace(b: bytes, off: int = 0) -> Tuple[Tuple[int, ...], int]:
    ver, rank, flags = b[off], b[off + 1], b[off + 2]
    if ver == 1:
        p = off + 8
    elif ver == 2:
        if b[off + 3] == 2:                           # null dataspace
            return (0,), 4
        p = off + 4
    else:
        raise H5Error("dataspace version %d" % ver)
    dims = struct.unpack_from("<%dQ" % rank, b, p) if rank else ()
    p += 8 * rank
    if flags & 1:
        p += 8 * rank
    if ver == 1 and flags & 2:
        p += 8 * rank
    return tuple(int(d) for d in dims), p - off


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class H5Object:
    """A group or a dataset.  Groups: `keys()`, `obj[name]` (paths with '/' descend), `attrs`.  Datasets: `shape`,
    `dtype`, `read()` -> numpy array."""

    def __init__(self, f: "H5File", addr: int, name: str = "/"):
        self.file, self.addr, self.name = f, addr, name
        self._msgs = f._object_messages(addr)
        self._links: Optional["OrderedDict[str, int]"] = None
        self._attrs: Optional[Dict[str, np.ndarray]] = None

    # ---- attributes
    @property
    def attrs(self) -> Dict[str, np.ndarray]:
        if self._attrs is None:
            self._attrs = {}
            for t, body in self._msgs:
                if t == 0x000C:
                    kv = self.file._parse_attribute(body)
                    if kv is not None:
                        self._attrs[kv[0]] = kv[1]
                elif t == 0x0015:
                    fh, bt = struct.unpack_from("<QQ", body, 4 if not (body[1] & 1) else 6)
                    if fh != UNDEF:
                        raise H5Error("dense attribute storage (fractal heap) is not supported")
        return self._attrs

    # ---- groups
    @property
    def is_group(self) -> bool:
        return any(t in (0x0011, 0x0002, 0x0006) for t, _ in self._msgs) or not self.is_dataset

    @property
    def is_dataset(self) -> bool:
        return any(t == 0x0008 for t, _ in self._msgs)

    def _load_links(self):
        if self._links is not None:
            return
        links: "OrderedDict[str, int]" = OrderedDict()
        for t, body in self._msgs:
            if t == 0x0011:                            # symbol table message: old-style group
                btree, heap = struct.unpack_from("<QQ", body, 0)
                for nm, a in self.file._walk_group_btree(btree, heap):
                    links[nm] = a
            elif t == 0x0006:                          # link message (new-style compact group)
                nm, a = self.file._parse_link(body)
                if a is not None:
                    links[nm] = a
            elif t == 0x0002:                          # link info: dense storage?
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                fh = struct.unpack_from("<Q", body, p)[0]
                if fh != UNDEF:
                    raise H5Error("dense link storage (fractal heap) is not supported")
        self._links = links

    def keys(self) -> List[str]:
        self._load_links()
        return list(self._links)

    def __contains__(self, path: str) -> bool:
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str) -> "H5Object":
        obj = self
        for part in [p for p in path.split("/") if p]:
            obj._load_links()
            if part not in obj._links:
                raise KeyError("%s: no member %r" % (obj.name, part))
            obj = H5Object(self.file, obj._links[part], obj.name.rstrip("/") + "/" + part)
        return obj

    def visit_datasets(self, prefix: str = ""):
        """Yield (path relative to this group, dataset object) depth-first."""
        for k in self.keys():
            o = self[k]
            p = prefix + k
            if o.is_dataset:
                yield p, o
            else:
                yield from o.visit_datasets(p + "/")

    # ---- datasets
    def _dataset_info(self):
        dt = shape = layout = None
        for t, body in self._msgs:
            if t == 0x0003:
                dt, _ = _parse_datatype(body)
            elif t == 0x0001:
                shape, _ = _parse_dataspace(body)
            elif t == 0x0008:
                layout = body
            elif t == 0x000B and len(body) >= 2 and body[1] > 0:
                raise H5Error("%s: filtered (compressed) datasets are not supported" % self.name)
        if dt is None or shape is None or layout is None:
            raise H5Error("%s is not a dataset" % self.name)
        return dt, shape, layout

    @property
    def shape(self):
        return self._dataset_info()[1]

    @property
    def dtype(self):
        return self._dataset_info()[0].np_dtype

    def read(self) -> np.ndarray:
        dt, shape, lay = self._dataset_info()
        if dt.vlen:
            raise H5Error("%s: variable-length data is not supported" % self.name)
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * dt.size
        f = self.file
        ver = lay[0]
        if ver == 3:
            cls = lay[1]
            if cls == 0:
                sz = struct.unpack_from("<H", lay, 2)[0]
                raw = lay[4:4 + sz]
            elif cls == 1:
                addr, sz = struct.unpack_from("<QQ", lay, 2)
                raw = b"\0" * nbytes if addr == UNDEF else f._read(addr, nbytes)
            elif cls == 2:
                rank1 = lay[2]
                btree = struct.unpack_from("<Q", lay, 3)[0]
                cdims = struct.unpack_from("<%dI" % rank1, lay, 11)
                raw = f._read_chunked(btree, shape, cdims[:-1], dt.size)
            else:
                raise H5Error("%s: data layout class %d is not supported" % (self.name, cls))
        elif ver in (1, 2):
            rank, cls = lay[1], lay[2]
            p = 8
            addr = None
            if cls != 0:
                addr = struct.unpack_from("<Q", lay, p)[0]
                p += 8
            dims = struct.unpack_from("<%dI" % rank, lay, p)
            p += 4 * rank
            if cls == 1:
                raw = f._read(addr, nbytes)
            elif cls == 0:
                sz = struct.unpack_from("<I", lay, p)[0]
                raw = lay[p + 4:p + 4 + sz]
            else:
                raw = f._read_chunked(addr, shape, dims[:-1], dt.size)
        else:
            raise H5Error("%s: data layout version %d is not supported" % (self.name, ver))
        if len(raw) < nbytes:
            raise H5Error("%s: truncated data (%d of %d bytes)" % (self.name, len(raw), nbytes))
        a = np.frombuffer(raw[:nbytes], dtype=dt.np_dtype).reshape(shape)
        return a.astype(dt.np_dtype.newbyteorder("="), copy=True)


class H5File(H5Object):
    def __init__(self, path_or_bytes: Union[str, bytes, bytearray]):
        if isinstance(path_or_bytes, (bytes, bytearray)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self.buf = fh.read()
        base = None
        for off in [0] + [512 << i for i in range(12)]:   # the superblock may sit at 0, 512, 1024, ...
            if self.buf[off:off + 8] == SIG:
                base = off
                break
        if base is None:
            raise H5Error("not an HDF5 file (no superblock signature)")
        b = self.buf
        ver = b[base + 8]
        self.sb_version = ver
        if ver in (0, 1):
            so, sl = b[base + 13], b[base + 14]
            if (so, sl) != (8, 8):
                raise H5Error("only 8-byte offsets/lengths are supported (file has %d/%d)" % (so, sl))
            self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, base + 16)
            p = base + 24 + (4 if ver == 1 else 0)
            self.base_addr = struct.unpack_from("<Q", b, p)[0]
            root_entry = p + 32
            root_addr = struct.unpack_from("<Q", b, root_entry + 8)[0]
        elif ver in (2, 3):
            so, sl = b[base + 9], b[base + 10]
            if (so, sl) != (8, 8):
                raise H5Error("only 8-byte offsets/lengths are supported (file has %d/%d)" % (so, sl))
            self.base_addr = struct.unpack_from("<Q", b, base + 12)[0]
            root_addr = struct.unpack_from("<Q", b, base + 36)[0]
            self.leaf_k = self.internal_k = 0
        else:
            raise H5Error("superblock version %d is not supported" % ver)
        if self.base_addr == UNDEF:
            self.base_addr = 0
        self.base_addr += 0 if ver >= 2 else 0
        self._sb_off = base
        super().__init__(self, root_addr, "/")

    # ---- raw access (addresses are relative to the base address)
    def _read(self, addr: int, n: int) -> bytes:
        a = addr + self.base_addr
        if a < 0 or a + n > len(self.buf):
            raise H5Error("address %d (+%d) is outside the file" % (addr, n))
        return self.buf[a:a + n]

    # ---- object headers
    def _object_messages(self, addr: int) -> List[Tuple[int, bytes]]:
        head = self._read(addr, 16)
        msgs: List[Tuple[int, bytes]] = []
        if head[:4] == b"OHDR":
            return self._object_messages_v2(addr)
        if head[0] != 1:
            raise H5Error("object header version %d at %d is not supported" % (head[0], addr))
        nmsg = struct.unpack_from("<H", head, 2)[0]
        hsize = struct.unpack_from("<I", head, 8)[0]
        blocks = [(addr + 16, hsize)]
        while blocks and len(msgs) < nmsg:
            a, sz = blocks.pop(0)
            blk = self._read(a, sz)
            p = 0
            while p + 8 <= sz and len(msgs) < nmsg:
                t, s, fl = struct.unpack_from("<HHB", blk, p)
                body = blk[p + 8:p + 8 + s]
                p += 8 + s
                if t == 0x0010:
                    ca, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cl))
                if fl & 2 and t not in (0x0010,):
                    raise H5Error("shared header messages are not supported (type 0x%04x)" % t)
                msgs.append((t, body))
        return msgs

    def _object_messages_v2(self, addr: int) -> List[Tuple[int, bytes]]:
        b = self._read(addr, 6)
        flags = b[5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szb = 1 << (flags & 3)
        size0 = int.from_bytes(self._read(p, szb), "little")
        p += szb
        track = bool(flags & 0x04)
        msgs: List[Tuple[int, bytes]] = []
        blocks = [(p, size0)]
        while blocks:
            a, sz = blocks.pop(0)
            blk = self._read(a, sz)
            q = 0
            hdr = 4 + (2 if track else 0)
            while q + hdr <= sz:
                t = blk[q]
                s = struct.unpack_from("<H", blk, q + 1)[0]
                fl = blk[q + 3]
                body = blk[q + hdr:q + hdr + s]
                q += hdr + s
                if t == 0x10:
                    ca, cl = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca + 4, cl - 8))           # skip "OCHK", drop the checksum
                    continue
                if t == 0 and s == 0:
                    continue
                if fl & 2:
                    raise H5Error("shared header messages are not supported (type 0x%02x)" % t)
                msgs.append((t, body))
        return msgs

    def _parse_link(self, body: bytes):
        ver, flags = body[0], body[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = body[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        lsz = 1 << (flags & 3)
        n = int.from_bytes(body[p:p + lsz], "little"); p += lsz
        name = body[p:p + n].decode("utf8"); p += n
        if ltype != 0:
            return name, None                          # soft / external links: ignored
        return name, struct.unpack_from("<Q", body, p)[0]

    # ---- old-style groups
    def _heap_string(self, heap_addr: int, off: int) -> str:
        h = self._read(heap_addr, 32)
        if h[:4] != b"HEAP":
            raise H5Error("bad local heap signature at %d" % heap_addr)
        dsize, _free, daddr = struct.unpack_from("<QQQ", h, 8)
        seg = self._read(daddr, dsize)
        end = seg.index(b"\0", off)
        return seg[off:end].decode("utf8")

    def _walk_group_btree(self, addr: int, heap: int):
        node = self._read(addr, 24)
        if node[:4] != b"TREE":
            raise H5Error("bad B-tree signature at %d" % addr)
        ntype, level, used = node[4], node[5], struct.unpack_from("<H", node, 6)[0]
        if ntype != 0:
            raise H5Error("expected a group B-tree node at %d" % addr)
        body = self._read(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            child = struct.unpack_from("<Q", body, (2 * i + 1) * 8)[0]
            if level > 0:
                yield from self._walk_group_btree(child, heap)
            else:
                sn = self._read(child, 8)
                if sn[:4] != b"SNOD":
                    raise H5Error("bad symbol-table node signature at %d" % child)
                nsym = struct.unpack_from("<H", sn, 6)[0]
                ents = self._read(child + 8, 40 * nsym)
                for j in range(nsym):
                    noff, oaddr = struct.unpack_from("<QQ", ents, 40 * j)
                    yield self._heap_string(heap, noff), oaddr

    # ---- chunked datasets without filters
    def _read_chunked(self, btree: int, shape, cdims, esize: int) -> bytes:
        out = np.zeros(shape, dtype=np.uint8).reshape(-1) if False else None
        rank = len(shape)
        arr = np.zeros(tuple(shape) + (esize,), dtype=np.uint8)

        def walk(addr):
            node = self._read(addr, 24)
            if node[:4] != b"TREE" or node[4] != 1:
                raise H5Error("bad chunk B-tree node at %d" % addr)
            level, used = node[5], struct.unpack_from("<H", node, 6)[0]
            ksz = 8 + 8 * (rank + 1)
            body = self._read(addr + 24, used * (ksz + 8) + ksz)
            for i in range(used):
                kp = i * (ksz + 8)
                csize, fmask = struct.unpack_from("<II", body, kp)
                offs = struct.unpack_from("<%dQ" % (rank + 1), body, kp + 8)
                child = struct.unpack_from("<Q", body, kp + ksz)[0]
                if level > 0:
                    walk(child)
                    continue
                if csize != int(np.prod(cdims)) * esize:
                    raise H5Error("filtered (compressed) chunks are not supported")
                chunk = np.frombuffer(self._read(child, csize), dtype=np.uint8).reshape(tuple(cdims) + (esize,))
                sl_dst, sl_src = [], []
                for d in range(rank):
                    lo = offs[d]
                    hi = min(lo + cdims[d], shape[d])
                    sl_dst.append(slice(lo, hi)); sl_src.append(slice(0, hi - lo))
                arr[tuple(sl_dst)] = chunk[tuple(sl_src)]
        if btree != UNDEF:
            walk(btree)
        return arr.tobytes()

    # ---- attributes
    def _parse_attribute(self, body: bytes):
        ver = body[0]
        nsz, dsz, ssz = struct.unpack_from("<HHH", body, 2)
        if ver == 1:
            p = 8
            name = body[p:p + nsz].split(b"\0")[0].decode("utf8"); p += _pad8(nsz)
            dtb = body[p:p + dsz]; p += _pad8(dsz)
            spb = body[p:p + ssz]; p += _pad8(ssz)
        elif ver in (2, 3):
            if body[1] & 3:
                raise H5Error("shared attribute datatype/dataspace is not supported")
            p = 8 + (1 if ver == 3 else 0)
            name = body[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz
            dtb = body[p:p + dsz]; p += dsz
            spb = body[p:p + ssz]; p += ssz
        else:
            raise H5Error("attribute message version %d" % ver)
        try:
            dt, _ = _parse_datatype(dtb)
        except H5Error:
            return None
        if dt.vlen:
            return None
        shape, _ = _parse_dataspace(spb)
        n = int(np.prod(shape)) if shape else 1
        raw = body[p:p + n * dt.size]
        a = np.frombuffer(raw, dtype=dt.np_dtype).reshape(shape)
        return name, a.copy()


# ====================================================================================== writer
class _W:
    """Bump allocator over a growing bytearray (8-byte aligned objects)."""

    def __init__(self):
        self.b = bytearray()

    def alloc(self, n: int) -> int:
        a = _pad8(len(self.b))
        self.b.extend(b"\0" * (a - len(self.b) + n))
        return a

    def put(self, addr: int, data: bytes):
        self.b[addr:addr + len(data)] = data


def _dt_message(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        props = {2: (0, 16, 10, 5, 0, 10, 15), 4: (0, 32, 23, 8, 0, 23, 127), 8: (0, 64, 52, 11, 0, 52, 1023)}[dt.itemsize]
        bits0 = 0x20                                   # little endian, mantissa normalisation = implied msb
        sign_loc = dt.itemsize * 8 - 1
        return (bytes([0x11, bits0, sign_loc, 0]) + struct.pack("<I", dt.itemsize) +
                struct.pack("<HHBBBBI", *props))
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0
        return bytes([0x10, bits0, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return bytes([0x13, 0x00, 0, 0]) + struct.pack("<I", dt.itemsize)      # null-terminated/padded ASCII
    raise H5Error("cannot write dtype %s" % dt)


def _ds_message(shape) -> bytes:
    shape = tuple(int(s) for s in shape)
    return bytes([1, len(shape), 0, 0, 0, 0, 0, 0]) + b"".join(struct.pack("<Q", s) for s in shape)


def _msg(t: int, body: bytes) -> bytes:
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    if len(body) > 0xFFF8:
        raise H5Error("header message of %d bytes exceeds the 64 KB object-header limit" % len(body))
    return struct.pack("<HHBBBB", t, len(body), 0, 0, 0, 0) + body


def _attr_message(name: str, value) -> bytes:
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    if a.dtype.kind == "S" and a.dtype.itemsize == 0:
        a = a.astype("S1")
    nm = name.encode("utf8") + b"\0"
    dtb, spb = _dt_message(a.dtype), _ds_message(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtb), len(spb))
    for part in (nm, dtb, spb):
        body += part + b"\0" * (_pad8(len(part)) - len(part))
    body += np.require(a, requirements="C").tobytes()
    return _msg(0x000C, body)


def _object_header(w: _W, messages: List[bytes]) -> int:
    blob = b"".join(messages)
    addr = w.alloc(16 + len(blob))
    w.put(addr, struct.pack("<BBHII", 1, 0, len(messages), 1, len(blob)) + b"\0\0\0\0" + blob)
    return addr


class H5Group(dict):
    """In-memory group for the writer: members are H5Group or numpy arrays; `.attrs` is a dict."""

    def __init__(self):
        super().__init__()
        self.attrs: Dict[str, np.ndarray] = OrderedDict()

    def require_group(self, path: str) -> "H5Group":
        g = self
        for part in [p for p in path.split("/") if p]:
            if part not in g:
                g[part] = H5Group()
            g = g[part]
        return g

    def create_dataset(self, path: str, data):
        parts = [p for p in path.split("/") if p]
        self.require_group("/".join(parts[:-1]))[parts[-1]] = np.require(np.asarray(data), requirements="C")


LEAF_K, INTERNAL_K = 16, 16


def _write_dataset(w: _W, a: np.ndarray, attrs) -> int:
    a = np.require(np.asarray(a), requirements="C")     # (np.ascontiguousarray would turn a scalar into shape (1,))
    daddr = w.alloc(max(a.nbytes, 1))
    w.put(daddr, a.tobytes())
    msgs = [_msg(0x0001, _ds_message(a.shape)), _msg(0x0003, _dt_message(a.dtype)),
            _msg(0x0008, bytes([3, 1]) + struct.pack("<QQ", daddr, a.nbytes))]
    msgs += [_attr_message(k, v) for k, v in (attrs or {}).items()]
    return _object_header(w, msgs)


def _write_group(w: _W, g: H5Group) -> Tuple[int, int, int]:
    """-> (object header address, B-tree address, local heap address)."""
    names = sorted(g.keys(), key=lambda s: s.encode("utf8"))
    child_addr = {}
    for nm in names:
        v = g[nm]
        child_addr[nm] = _write_group(w, v)[0] if isinstance(v, H5Group) else _write_dataset(w, v, None)
    # local heap: offset 0 holds the empty string (the B-tree's first key)
    heap_data = bytearray(b"\0" * 8)
    name_off = {}
    for nm in names:
        name_off[nm] = len(heap_data)
        e = nm.encode("utf8") + b"\0"
        heap_data += e + b"\0" * (_pad8(len(e)) - len(e))
    free_off = len(heap_data)
    heap_data += struct.pack("<QQ", 1, 16)            # one free block: (next = 1 "none", size = 16)
    hdata_addr = w.alloc(len(heap_data))
    w.put(hdata_addr, bytes(heap_data))
    heap_addr = w.alloc(32)
    w.put(heap_addr, b"HEAP" + bytes([0, 0, 0, 0]) + struct.pack("<QQQ", len(heap_data), free_off, hdata_addr))
    # symbol-table nodes of up to 2*LEAF_K entries, one level-0 B-tree node over them
    per = 2 * LEAF_K
    groups = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
    if len(groups) > 2 * INTERNAL_K:
        raise H5Error("group with %d members exceeds this writer's single-level B-tree (%d)" % (len(names), per * 2 * INTERNAL_K))
    snods = []
    for chunk in groups:
        sa = w.alloc(8 + 40 * per)
        ents = b""
        for nm in chunk:
            ents += struct.pack("<QQII", name_off[nm], child_addr[nm], 0, 0) + b"\0" * 16
        w.put(sa, b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(chunk)) + ents)
        snods.append(sa)
    bt_addr = w.alloc(24 + (2 * 2 * INTERNAL_K + 1) * 8)
    body = struct.pack("<Q", 0)
    for chunk, sa in zip(groups, snods):
        body += struct.pack("<QQ", sa, name_off[chunk[-1]] if chunk else 0)
    used = len(snods) if names else 0
    w.put(bt_addr, b"TREE" + bytes([0, 0]) + struct.pack("<H", used) + struct.pack("<QQ", UNDEF, UNDEF) + body)
    msgs = [_msg(0x0011, struct.pack("<QQ", bt_addr, heap_addr))]
    msgs += [_attr_message(k, v) for k, v in g.attrs.items()]
    return _object_header(w, msgs), bt_addr, heap_addr


def write_h5(path, root: H5Group):
    """Serialise `root` (H5Group tree of numpy arrays) as an HDF5 file (superblock v0, old-style groups)."""
    w = _W()
    w.alloc(96)                                        # superblock v0 (56 bytes) + root symbol-table entry (40)
    oh, bt, heap = _write_group(w, root)
    eof = _pad8(len(w.b))
    w.b.extend(b"\0" * (eof - len(w.b)))
    sb = SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, heap)
    assert len(sb) == 96
    w.put(0, sb)
    data = bytes(w.b)
    if hasattr(path, "write"):
        path.write(data)
    else:
        with open(path, "wb") as fh:
            fh.write(data)


# ====================================================================================== Keras weight files
def _chunked_attr(group: H5Object, name: str) -> List[str]:
    """keras.engine.saving.load_attributes_from_hdf5_group: `name`, or the pieces `name0`, `name1`, ..."""
    at = group.attrs
    if name in at:
        vals = list(np.atleast_1d(at[name]))
    else:
        vals, i = [], 0
        while "%s%d" % (name, i) in at:
            vals += list(np.atleast_1d(at["%s%d" % (name, i)]))
            i += 1
    return [v.decode("utf8") if isinstance(v, (bytes, np.bytes_)) else str(v) for v in vals]


def read_keras_weights(path) -> "OrderedDict[str, np.ndarray]":
    """`model.save_weights(path)` / `model.save(path)` file -> {Keras weight name ('conv2d_1/kernel:0'): array}, in the
    file's layer order.  Falls back to walking every dataset when the `layer_names` attribute is absent."""
    f = H5File(path)
    g = f["model_weights"] if "model_weights" in f.keys() else f
    out: "OrderedDict[str, np.ndarray]" = OrderedDict()
    layers = _chunked_attr(g, "layer_names")
    if layers:
        for ln in layers:
            lg = g[ln]
            for wn in _chunked_attr(lg, "weight_names"):
                out[wn] = lg[wn].read()
    else:
        for ln in g.keys():
            if g[ln].is_dataset:
                continue
            for p, ds in g[ln].visit_datasets():
                out[p] = ds.read()
    return out


def write_keras_weights(path, layers: "OrderedDict[str, OrderedDict[str, np.ndarray]]", full_model: bool = False,
                        extra_layers: Optional[List[str]] = None):
    """Write `{layer name: {weight name: array}}` the way keras.engine.saving.save_weights_to_hdf5_group does.
    `full_model`: under `/model_weights` like `model.save()`.  `extra_layers`: weightless layer names (Activation,
    Add, ...) a real Keras file also lists."""
    root = H5Group()
    g = root.require_group("model_weights") if full_model else root
    names = list(layers) + list(extra_layers or [])
    g.attrs["layer_names"] = np.array([n.encode("utf8") for n in names]) if names else np.zeros((0,), "S1")
    g.attrs["backend"] = np.bytes_(b"tensorflow")
    g.attrs["keras_version"] = np.bytes_(b"2.2.4")
    for ln in names:
        lg = g.require_group(ln)
        ws = layers.get(ln, {})
        lg.attrs["weight_names"] = np.array([n.encode("utf8") for n in ws]) if ws else np.zeros((0,), "S1")
        for wn, a in ws.items():
            lg.create_dataset(wn, np.asarray(a, dtype=np.float32))
    write_h5(path, root)
