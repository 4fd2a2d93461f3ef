"""A dependency-free subset of HDF5 — exactly what Keras model files need (h5py / libhdf5 are not installed here).

Writes and reads the "earliest" file layout that h5py (and therefore Keras' `model.save('*.h5')` and the
`model.weights.h5` member of `*.keras` archives) produces by default:
  superblock v0 - old-style groups (symbol-table message -> B-tree v1 + SNOD nodes + local heap) - object headers v1
  (with continuation blocks) - dataspace v1/v2 - datatypes: IEEE float32/float64, (u)int8..64 little-endian,
  fixed-length strings, variable-length strings (global heap, read only) - contiguous or compact dataset layout -
  attribute messages v1..v3.
Not supported (raises): chunked / filtered datasets, superblock >= 2, dense attribute storage.

PARITY UNPINNED: the reference ships no model file (its .gitignore drops models/), so the reader is validated against
files produced by this writer and against the published HDF5 File Format Specification (v1.1 structures), not against a
Keras-written file.
"""
from __future__ import annotations

import io
import struct
from typing import BinaryIO, Dict, List, Optional, Union

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 32, 16          # symbols per SNOD = 2*LEAF_K; children per B-tree node = 2*INTERNAL_K


class Dataset:
    def __init__(self, data: np.ndarray, attrs: Optional[dict] = None):
        self.data = np.ascontiguousarray(data)
        self.attrs = dict(attrs or {})

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    @property
    def shape(self):
        return self.data.shape


class Group:
    def __init__(self):
        self.groups: Dict[str, "Group"] = {}
        self.datasets: Dict[str, Dataset] = {}
        self.attrs: Dict[str, object] = {}

    def group(self, name: str) -> "Group":
        return self.groups.setdefault(name, Group())

    def dataset(self, name: str, data, attrs: Optional[dict] = None) -> Dataset:
        d = self.datasets[name] = Dataset(np.asarray(data), attrs)
        return d

    def __getitem__(self, path: str):
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.groups[part] if part in node.groups else node.datasets[part]
        return node

    def __contains__(self, name: str) -> bool:
        return name in self.groups or name in self.datasets

    def keys(self):
        return list(self.groups) + list(self.datasets)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ====================================================================================================== datatype / dataspace
def _dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize in (4, 8):
        if dt.itemsize == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
            return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + props
        props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + props
    if dt.kind in "iu":
        flags = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, flags, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, max(dt.itemsize, 1))       # null-padded, ASCII
    raise TypeError(f"h5lite cannot store dtype {dt}")


def _space_msg(shape) -> bytes:
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _as_storable(value) -> np.ndarray:
    if isinstance(value, str):
        value = value.encode("utf8")
    if isinstance(value, bytes):
        return np.array(value, dtype=f"S{max(len(value), 1)}")
    if isinstance(value, (list, tuple)) and value and isinstance(value[0], (bytes, str)):
        vals = [v.encode("utf8") if isinstance(v, str) else v for v in value]
        return np.array(vals, dtype=f"S{max(1, max(len(v) for v in vals))}")
    if isinstance(value, (list, tuple)) and not value:
        return np.zeros((0,), dtype="S1")
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf8")
    return a


def _attr_msg(name: str, value) -> bytes:
    a = _as_storable(value)
    nm = name.encode("utf8") + b"\0"
    dt, sp = _dtype_msg(a.dtype), _space_msg(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + a.tobytes()
    return body


# ====================================================================================================== writer
class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\0" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def reserve(self, n: int) -> int:
        return self.alloc(b"\0" * n)

    def object_header(self, messages: List[tuple]) -> int:
        body = b""
        for mtype, data, flags in messages:
            data = _pad8(data)
            if len(data) > 0xFFF8:
                raise ValueError(f"HDF5 object-header message of {len(data)} bytes exceeds the 64 KiB limit "
                                 "(store large strings as datasets or split them, as Keras does with layer_names)")
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        hdr = struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\0" * 4
        return self.alloc(hdr + body)

    def write_dataset(self, ds: Dataset) -> int:
        a = ds.data
        if a.dtype.kind == "U":
            a = np.char.encode(a, "utf8")
        raw = a.tobytes()
        addr = self.alloc(raw) if raw else UNDEF
        msgs = [(0x0001, _space_msg(a.shape), 0), (0x0003, _dtype_msg(a.dtype), 1),
                (0x0005, struct.pack("<BBBB", 2, 2, 2, 0), 1),
                (0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw)), 0)]
        msgs += [(0x000C, _attr_msg(k, v), 0) for k, v in ds.attrs.items()]
        return self.object_header(msgs)

    def write_group(self, g: Group):
        """-> (object header address, btree address, heap address)"""
        children = {}
        for name, sub in g.groups.items():
            children[name] = ("g",) + self.write_group(sub)
        for name, ds in g.datasets.items():
            children[name] = ("d", self.write_dataset(ds))
        names = sorted(children, key=lambda s: s.encode("utf8"))
        # local heap: offset 0 is the empty string
        heap = bytearray(b"\0" * 8)
        offs = {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf8") + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, heap_data))
        # symbol-table nodes
        per = 2 * LEAF_K
        chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(chunks) > 2 * INTERNAL_K:
            raise ValueError("too many links in one group for a single-level B-tree")
        snods = []
        for ch in chunks:
            ent = b""
            for n in ch:
                c = children[n]
                if c[0] == "g":
                    ent += struct.pack("<QQII", offs[n], c[1], 1, 0) + struct.pack("<QQ", c[2], c[3])
                else:
                    ent += struct.pack("<QQII", offs[n], c[1], 0, 0) + b"\0" * 16
            ent += b"\0" * (40 * (per - len(ch)))
            snods.append(self.alloc(b"SNOD" + struct.pack("<BBH", 1, 0, len(ch)) + ent))
        # B-tree v1, level 0: key[0] = "" ; key[i+1] = last name of child i
        body = b""
        for i, ch in enumerate(chunks):
            body += struct.pack("<Q", 0 if i == 0 else offs[chunks[i - 1][-1]]) + struct.pack("<Q", snods[i])
        body += struct.pack("<Q", offs[chunks[-1][-1]] if chunks[-1] else 0)
        used = len(chunks) if names else 0
        if not names:
            body = b""
        full = (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
        body += b"\0" * (full - len(body))
        btree = self.alloc(b"TREE" + struct.pack("<BBHQQ", 0, 0, used, UNDEF, UNDEF) + body)
        msgs = [(0x0011, struct.pack("<QQ", btree, heap_addr), 0)]
        msgs += [(0x000C, _attr_msg(k, v), 0) for k, v in g.attrs.items()]
        return self.object_header(msgs), btree, heap_addr


def write(f: Union[BinaryIO, io.BytesIO], root: Group) -> None:
    w = _Writer()
    w.reserve(96)                                   # superblock v0
    ohdr, btree, heap = w.write_group(root)
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, ohdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == 96
    w.buf[0:96] = sb
    f.write(bytes(w.buf))


# ====================================================================================================== reader
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file (bad signature)")
        ver = data[8]
        if ver not in (0, 1):
            raise ValueError(f"HDF5 superblock version {ver} is not supported (only the h5py/Keras default 'earliest' layout)")
        self.so, self.sl = data[13], data[14]
        if (self.so, self.sl) != (8, 8):
            raise ValueError("only 8-byte offsets/lengths are supported")
        pos = 24 if ver == 0 else 28
        self.base, _, self.eof, _ = struct.unpack_from("<QQQQ", data, pos)
        pos += 32
        _, self.root_ohdr, _, _ = struct.unpack_from("<QQII", data, pos)
        self._gcol: Dict[int, Dict[int, bytes]] = {}

    # ---- object headers
    def messages(self, addr: int):
        d = self.d
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", d, addr)
        if ver != 1:
            raise ValueError(f"object header version {ver} at {addr} is not supported")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            pos, left = blocks.pop(0)
            end = pos + left
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", d, pos)
                body = d[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, ln))
                out.append((mtype, body, flags))
        return out

    # ---- groups
    def _heap_name(self, heap_addr: int, off: int) -> str:
        d = self.d
        if d[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        seg = struct.unpack_from("<Q", d, heap_addr + 24)[0]
        end = d.index(b"\0", seg + off)
        return d[seg + off:end].decode("utf8")

    def _btree_entries(self, addr: int, heap: int, out: list):
        d = self.d
        if addr == UNDEF:
            return
        if d[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        ntype, level, used = struct.unpack_from("<BBH", d, addr + 4)
        if ntype != 0:
            raise ValueError("expected a group B-tree node")
        pos = addr + 24
        for i in range(used):
            child = struct.unpack_from("<Q", d, pos + 8)[0]
            pos += 16
            if level > 0:
                self._btree_entries(child, heap, out)
            else:
                if d[child:child + 4] != b"SNOD":
                    raise ValueError("bad symbol-table node signature")
                n = struct.unpack_from("<H", d, child + 6)[0]
                for k in range(n):
                    noff, ohdr, ctype = struct.unpack_from("<QQI", d, child + 8 + 40 * k)
                    out.append((self._heap_name(heap, noff), ohdr))

    def read_object(self, addr: int):
        msgs = self.messages(addr)
        attrs = {}
        sym = space = dtype = layout = None
        for mtype, body, flags in msgs:
            if mtype == 0x0011:
                sym = struct.unpack_from("<QQ", body, 0)
            elif mtype == 0x0001:
                space = self._space(body)
            elif mtype == 0x0003:
                dtype = self._dtype(body)
            elif mtype == 0x0008:
                layout = body
            elif mtype == 0x000C:
                k, v = self._attr(body)
                attrs[k] = v
        if sym is not None:
            g = Group()
            g.attrs = attrs
            entries: list = []
            self._btree_entries(sym[0], sym[1], entries)
            for name, ohdr in entries:
                obj = self.read_object(ohdr)
                (g.groups if isinstance(obj, Group) else g.datasets)[name] = obj
            return g
        if space is None or dtype is None or layout is None:
            raise ValueError(f"object at {addr} is neither an old-style group nor a simple dataset")
        return Dataset(self._data(layout, dtype, space), attrs)

    # ---- pieces
    @staticmethod
    def _space(b: bytes):
        ver, rank, flags = b[0], b[1], b[2]
        if ver == 1:
            off = 8
        elif ver == 2:
            if b[3] == 2:        # null dataspace
                return (0,)
            off = 4
        else:
            raise ValueError(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", b, off + 8 * i)[0] for i in range(rank))

    @staticmethod
    def _dtype(b: bytes):
        cls, ver = b[0] & 0x0F, b[0] >> 4
        bits0, size = b[1], struct.unpack_from("<I", b, 4)[0]
        if bits0 & 1 and cls in (0, 1):
            raise ValueError("big-endian data is not supported")
        if cls == 1:
            return np.dtype(f"<f{size}")
        if cls == 0:
            return np.dtype(("<i" if bits0 & 0x08 else "<u") + str(size))
        if cls == 3:
            return np.dtype(f"S{size}")
        if cls == 9:
            base = _Reader._dtype(b[8:])
            if (b[1] & 0x0F) == 1:
                return ("vlen_str", size)
            return ("vlen", base, size)
        raise ValueError(f"HDF5 datatype class {cls} is not supported")

    def _global_heap_obj(self, addr: int, index: int) -> bytes:
        col = self._gcol.get(addr)
        if col is None:
            d = self.d
            if d[addr:addr + 4] != b"GCOL":
                raise ValueError("bad global heap signature")
            size = struct.unpack_from("<Q", d, addr + 8)[0]
            col, pos = {}, addr + 16
            while pos + 16 <= addr + size:
                idx, _, _, osz = struct.unpack_from("<HHIQ", d, pos)
                if idx == 0:
                    break
                col[idx] = d[pos + 16:pos + 16 + osz]
                pos += 16 + osz + (-osz % 8)
            self._gcol[addr] = col
        return col[index]

    def _decode(self, raw: bytes, dtype, shape):
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dtype, tuple):
            if dtype[0] != "vlen_str":
                raise ValueError("variable-length sequences other than strings are not supported")
            out = []
            for i in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                out.append(self._global_heap_obj(gaddr, gidx)[:ln] if ln else b"")
            return out[0] if not shape else np.array(out, dtype=object).reshape(shape)
        a = np.frombuffer(raw, dtype=dtype, count=n).reshape(shape)
        if not shape:
            return bytes(a.reshape(-1)[0]).rstrip(b"\0") if dtype.kind == "S" else a.reshape(-1)[0]
        return a.copy()

    def _attr(self, b: bytes):
        ver = b[0]
        if ver == 1:
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            p = 8
            name = b[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz + (-nsz % 8)
            dt = self._dtype(b[p:p + dsz]); p += dsz + (-dsz % 8)
            sp = self._space(b[p:p + ssz]); p += ssz + (-ssz % 8)
        elif ver in (2, 3):
            nsz, dsz, ssz = struct.unpack_from("<HHH", b, 2)
            p = 8 if ver == 2 else 9
            name = b[p:p + nsz].split(b"\0")[0].decode("utf8"); p += nsz
            dt = self._dtype(b[p:p + dsz]); p += dsz
            sp = self._space(b[p:p + ssz]); p += ssz
        else:
            raise ValueError(f"attribute message version {ver}")
        return name, self._decode(b[p:], dt, sp)

    def _data(self, layout: bytes, dtype, shape):
        ver = layout[0]
        n = int(np.prod(shape)) if shape else 1
        isz = 16 if isinstance(dtype, tuple) else dtype.itemsize
        if ver == 3:
            cls = layout[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", layout, 2)
                raw = b"" if addr == UNDEF else self.d[addr:addr + n * isz]
            elif cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                raw = layout[4:4 + size]
            else:
                raise ValueError("chunked datasets are not supported (Keras stores weights contiguously)")
        elif ver in (1, 2):
            rank, cls = layout[1], layout[2]
            if cls != 1:
                raise ValueError("only contiguous layout is supported for layout message v1/v2")
            addr = struct.unpack_from("<Q", layout, 8)[0]
            raw = self.d[addr:addr + n * isz]
        else:
            raise ValueError(f"data layout message version {ver}")
        if n == 0:
            return np.zeros(shape, dtype=dtype if not isinstance(dtype, tuple) else object)
        out = self._decode(raw, dtype, shape if shape else ())
        return np.asarray(out)


def read(data: Union[bytes, BinaryIO]) -> Group:
    if not isinstance(data, (bytes, bytearray)):
        data = data.read()
    r = _Reader(bytes(data))
    root = r.read_object(r.root_ohdr)
    if not isinstance(root, Group):
        raise ValueError("root object is not a group")
    return root
