"""Minimal HDF5 reader / writer for Keras weight files -- SURVEY row f2 (keras_nerf/model/nerf/nerf.py:63-64,
132-136: `save_weights('coarse.h5')` / `load_weights('coarse.h5')`).

h5py / libhdf5 are not available in this image, so the subset of the HDF5 File Format Specification (version 1.x
structures, what libhdf5 writes with `libver='earliest'`, h5py's default) that Keras weight files use is
implemented here in pure Python:

  superblock v0/v1, object headers v1 (+ continuation blocks), old-style groups (symbol-table message, B-tree v1
  group nodes, SNOD symbol nodes, local heaps), dataspace v1/v2, datatypes (IEEE floats, integers, fixed-length
  and variable-length strings via global heap collections), data layout v3 (compact and contiguous storage; chunked
  datasets are rejected -- Keras weight files do not chunk), attribute messages v1-v3.

Keras layout (`save_weights_to_hdf5_group`): root attributes `layer_names`, `backend`, `keras_version`; one group per
layer with attribute `weight_names`; each weight a dataset at `<layer>/<weight_name>` (weight names contain '/', so
they sit in nested groups).

FORMAT PIN STATUS: written from the published specification; no libhdf5 is available offline to cross-check, so
compatibility with files written by real h5py is UNVERIFIED (tests round-trip this writer through this reader and
check the byte-level structures named above).  Host-side file I/O, not part of the per-ray hot path.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF

MSG_NIL, MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_ATTRIBUTE, MSG_CONTINUATION, \
    MSG_SYMBOL_TABLE = 0x0, 0x1, 0x3, 0x4, 0x5, 0x8, 0xC, 0x10, 0x11


class Hdf5Error(ValueError):
    pass


class Node:
    """A group (children) or a dataset (data), with attributes."""

    def __init__(self):
        self.attrs: Dict[str, object] = {}
        self.children: Dict[str, "Node"] = {}
        self.data: Optional[np.ndarray] = None

    @property
    def is_dataset(self):
        return self.data is not None

    def __getitem__(self, path: str) -> "Node":
        node = self
        for part in [p for p in path.split("/") if p]:
            node = node.children[part]
        return node

    def datasets(self, prefix=""):
        """[(path, array)] of every dataset below this node, depth first in name order"""
        out = []
        for name in sorted(self.children):
            c = self.children[name]
            p = f"{prefix}/{name}" if prefix else name
            out += [(p, c.data)] if c.is_dataset else c.datasets(p)
        return out


def _pad8(n: int) -> int:
    return (n + 7) & ~7


# ================================================ reader =========================================================
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise Hdf5Error("not an HDF5 file (bad signature)")
        ver = buf[8]
        if ver not in (0, 1):
            raise Hdf5Error(f"superblock version {ver} not supported (only the v0/v1 'earliest' format Keras/h5py write)")
        if buf[13] != 8 or buf[14] != 8:
            raise Hdf5Error("only 8-byte offsets and lengths are supported")
        off = 24 + (4 if ver == 1 else 0)
        self.base = struct.unpack_from("<Q", buf, off)[0]
        root_entry = off + 32
        self.root_header = struct.unpack_from("<Q", buf, root_entry + 8)[0]
        self._gcol: Dict[int, Dict[int, bytes]] = {}

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    # ---- object headers -------------------------------------------------------------------------
    def messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """[(type, data offset, data size)] of a version-1 object header at `addr`, continuations followed"""
        addr += self.base
        ver, _, nmsg, _refs, hsize = self.u("BBHII", addr)
        if ver != 1:
            if self.b[addr:addr + 4] == b"OHDR":
                raise Hdf5Error("version-2 object headers (libver='latest') are not supported")
            raise Hdf5Error(f"bad object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = self.u("HHB", pos)
                data = pos + 8
                if mtype == MSG_CONTINUATION:
                    caddr, clen = self.u("QQ", data)
                    blocks.append((caddr + self.base, clen))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    # ---- groups ---------------------------------------------------------------------------------
    def heap_string(self, heap_addr: int, offset: int) -> str:
        h = heap_addr + self.base
        if self.b[h:h + 4] != b"HEAP":
            raise Hdf5Error("bad local heap signature")
        data_addr = self.u("Q", h + 24)[0] + self.base
        s = data_addr + offset
        return self.b[s:self.b.index(b"\0", s)].decode("utf-8")

    def group_entries(self, btree: int, heap: int) -> List[Tuple[str, int]]:
        b = btree + self.base
        if self.b[b:b + 4] != b"TREE":
            raise Hdf5Error("bad B-tree signature")
        ntype, level, used = self.u("BBH", b + 4)
        if ntype != 0:
            raise Hdf5Error("not a group B-tree")
        out = []
        for i in range(used):
            child = self.u("Q", b + 24 + 8 + i * 16)[0]
            if level > 0:
                out += self.group_entries(child, heap)
                continue
            s = child + self.base
            if self.b[s:s + 4] != b"SNOD":
                raise Hdf5Error("bad symbol node signature")
            nsym = self.u("H", s + 6)[0]
            for k in range(nsym):
                name_off, ohdr = self.u("QQ", s + 8 + 40 * k)
                out.append((self.heap_string(heap, name_off), ohdr))
        return out

    # ---- datatypes / dataspaces -----------------------------------------------------------------
    def datatype(self, off: int):
        """-> (kind, numpy dtype or None, element size, total message size)"""
        cv, b0, b1, _b2, size = self.u("BBBBI", off)
        cls, ver = cv & 0x0F, cv >> 4
        if ver not in (1, 2, 3):
            raise Hdf5Error(f"datatype version {ver}")
        order = ">" if (b0 & 1) else "<"
        if cls == 0:
            signed = bool(b0 & 0x08)
            return "num", np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size, 8 + 4
        if cls == 1:
            return "num", np.dtype(f"{order}f{size}"), size, 8 + 12
        if cls == 3:
            return "str", None, size, 8
        if cls == 9:
            is_str = (b0 & 0x0F) == 1
            _, _, _, base_len = self.datatype(off + 8)
            if not is_str:
                raise Hdf5Error("variable-length sequences are not supported")
            return "vstr", None, size, 8 + base_len
        raise Hdf5Error(f"datatype class {cls} not supported")

    def dataspace(self, off: int) -> Tuple[int, ...]:
        ver, rank, flags = self.u("BBB", off)
        if ver == 1:
            dims = off + 8
        elif ver == 2:
            if self.b[off + 3] == 2:
                return (0,)                                   # null dataspace
            dims = off + 4
        else:
            raise Hdf5Error(f"dataspace version {ver}")
        return tuple(self.u("Q", dims + 8 * i)[0] for i in range(rank))

    def global_heap_object(self, addr: int, index: int) -> bytes:
        if addr not in self._gcol:
            g = addr + self.base
            if self.b[g:g + 4] != b"GCOL":
                raise Hdf5Error("bad global heap signature")
            size = self.u("Q", g + 8)[0]
            pos, objs = g + 16, {}
            while pos + 16 <= g + size:
                idx, _rc, _r, osize = self.u("HHIQ", pos)
                if idx == 0:
                    break
                objs[idx] = self.b[pos + 16:pos + 16 + osize]
                pos += 16 + _pad8(osize)
            self._gcol[addr] = objs
        return self._gcol[addr][index]

    def decode(self, kind, dtype, esize, shape, raw_off: int):
        n = int(np.prod(shape)) if shape else 1
        if kind == "num":
            a = np.frombuffer(self.b, dtype=dtype, count=n, offset=raw_off).astype(dtype.newbyteorder("="))
            return a.reshape(shape) if shape else a.reshape(())
        if kind == "str":
            items = [self.b[raw_off + i * esize:raw_off + (i + 1) * esize].split(b"\0")[0] for i in range(n)]
        else:  # vstr: length(4) heap address(8) object index(4)
            items = []
            for i in range(n):
                ln, gaddr, idx = self.u("IQI", raw_off + i * 16)
                items.append(self.global_heap_object(gaddr, idx)[:ln] if gaddr not in (0, UNDEF) else b"")
        return np.array(items, dtype=object).reshape(shape) if shape else items[0]

    # ---- objects --------------------------------------------------------------------------------
    def attribute(self, off: int):
        ver = self.b[off]
        if ver == 1:
            nlen, tlen, slen = self.u("HHH", off + 2)
            p = off + 8
            name = self.b[p:p + nlen].split(b"\0")[0].decode()
            p += _pad8(nlen)
            t, s = p, p + _pad8(tlen)
            d = s + _pad8(slen)
        elif ver in (2, 3):
            nlen, tlen, slen = self.u("HHH", off + 2)
            p = off + 8 + (1 if ver == 3 else 0)
            name = self.b[p:p + nlen].split(b"\0")[0].decode()
            t = p + nlen
            s = t + tlen
            d = s + slen
        else:
            raise Hdf5Error(f"attribute version {ver}")
        kind, dtype, esize, _ = self.datatype(t)
        return name, self.decode(kind, dtype, esize, self.dataspace(s), d)

    def node(self, header_addr: int) -> Node:
        node = Node()
        msgs = self.messages(header_addr)
        kinds = {m[0]: m for m in msgs}
        for mtype, off, _size in msgs:
            if mtype == MSG_ATTRIBUTE:
                k, v = self.attribute(off)
                node.attrs[k] = v
        if MSG_SYMBOL_TABLE in kinds:
            btree, heap = self.u("QQ", kinds[MSG_SYMBOL_TABLE][1])
            for name, ohdr in self.group_entries(btree, heap):
                node.children[name] = self.node(ohdr)
            return node
        if MSG_LAYOUT not in kinds:
            return node                                        # an empty new-style group or a committed datatype
        kind, dtype, esize, _ = self.datatype(kinds[MSG_DATATYPE][1])
        shape = self.dataspace(kinds[MSG_DATASPACE][1])
        lo = kinds[MSG_LAYOUT][1]
        lver, lclass = self.u("BB", lo)
        if lver != 3:
            raise Hdf5Error(f"data layout version {lver} not supported")
        if lclass == 0:
            raw = lo + 4
        elif lclass == 1:
            addr = self.u("Q", lo + 2)[0]
            raw = None if addr == UNDEF else addr + self.base
        else:
            raise Hdf5Error("chunked datasets are not supported (Keras weight files are contiguous)")
        if raw is None:
            node.data = np.zeros(shape, dtype=dtype if kind == "num" else object)
        else:
            node.data = self.decode(kind, dtype, esize, shape, raw)
        return node


def read_h5(path: str) -> Node:
    with open(path, "rb") as f:
        r = _Reader(f.read())
    return r.node(r.root_header)


# ================================================ writer =========================================================
_LEAF_K = 32      # a symbol node holds 2K = 64 links: one node per group is enough for Keras files
_INTERNAL_K = 16


def _dt_message(a) -> bytes:
    if isinstance(a, np.ndarray) and a.dtype.kind == "f":
        s = a.dtype.itemsize
        exp_bits, man_bits, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[s]
        return struct.pack("<BBBBI", 0x11, 0x20, 8 * s - 1, 0, s) + \
            struct.pack("<HHBBBBI", 0, 8 * s, man_bits, exp_bits, 0, man_bits, bias)
    if isinstance(a, np.ndarray) and a.dtype.kind in "iu":
        s = a.dtype.itemsize
        return struct.pack("<BBBBI", 0x10, 0x08 if a.dtype.kind == "i" else 0, 0, 0, s) + struct.pack("<HH", 0, 8 * s)
    if isinstance(a, np.ndarray) and a.dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, a.dtype.itemsize)   # null-padded ASCII, as h5py maps numpy 'S'
    raise Hdf5Error(f"cannot store {type(a)} {getattr(a, 'dtype', '')}")


def _ds_message(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", d) for d in shape)


def _as_storable(v) -> np.ndarray:
    if isinstance(v, str):
        v = v.encode("utf-8")
    if isinstance(v, bytes):
        return np.array(v, dtype=f"S{max(len(v), 1)}")
    if isinstance(v, (list, tuple)) and v and isinstance(v[0], (str, bytes)):
        v = [x.encode("utf-8") if isinstance(x, str) else x for x in v]
        return np.array(v, dtype=f"S{max(max(len(x) for x in v), 1)}")
    a = np.asarray(v)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    return np.asarray(a.astype(a.dtype.newbyteorder("<")) if a.dtype.kind in "fiu" else a, order="C")


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\0" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _attr_message(name: str, value) -> bytes:
    a = _as_storable(value)
    nm = name.encode("utf-8") + b"\0"
    dt, ds = _dt_message(a), _ds_message(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    body += nm + b"\0" * (_pad8(len(nm)) - len(nm))
    body += dt + b"\0" * (_pad8(len(dt)) - len(dt))
    body += ds + b"\0" * (_pad8(len(ds)) - len(ds))
    return _message(MSG_ATTRIBUTE, body + a.tobytes())


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)                               # superblock placeholder

    def alloc(self, data: bytes) -> int:
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, messages: List[bytes]) -> int:
        body = b"".join(messages)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)

    def dataset(self, node: Node) -> int:
        a = _as_storable(node.data)
        raw = self.alloc(a.tobytes()) if a.size else UNDEF
        msgs = [_message(MSG_DATASPACE, _ds_message(a.shape)),
                _message(MSG_DATATYPE, _dt_message(a), flags=1),               # constant message
                _message(MSG_FILL, struct.pack("<BBBB", 2, 2, 2, 0)),          # late allocation, fill if set, undefined
                _message(MSG_LAYOUT, struct.pack("<BBQQ", 3, 1, raw, a.nbytes))]
        msgs += [_attr_message(k, v) for k, v in node.attrs.items()]
        return self.object_header(msgs)

    def group(self, node: Node) -> Tuple[int, int, int]:
        """-> (object header, B-tree, local heap) addresses"""
        names = sorted(node.children, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * _LEAF_K:
            raise Hdf5Error(f"more than {2 * _LEAF_K} links in one group")
        child_info = []
        for n in names:
            c = node.children[n]
            child_info.append((self.dataset(c), None, None) if c.is_dataset else self.group(c))
        # local heap: offset 0 = "", then the link names, then one free block
        heap_data, offsets = bytearray(8), []
        for n in names:
            offsets.append(len(heap_data))
            e = n.encode("utf-8") + b"\0"
            heap_data += e + b"\0" * (_pad8(len(e)) - len(e))
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 32) + b"\0" * 16                    # free block: next = H5HL_FREE_NULL, size
        data_addr = self.alloc(bytes(heap_data))
        heap = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, data_addr))
        snod = bytearray(b"SNOD" + struct.pack("<BBH", 1, 0, len(names)))
        for off, (ohdr, bt, hp) in zip(offsets, child_info):
            if bt is None:
                snod += struct.pack("<QQII16x", off, ohdr, 0, 0)
            else:
                snod += struct.pack("<QQIIQQ", off, ohdr, 1, 0, bt, hp)
        snod += b"\0" * (8 + 2 * _LEAF_K * 40 - len(snod))
        tree = bytearray(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF))
        if names:
            snod_addr = self.alloc(bytes(snod))
            tree += struct.pack("<QQQ", 0, snod_addr, offsets[-1])             # key0 = "", child, key1 = largest name
        tree += b"\0" * (24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(tree))
        btree = self.alloc(bytes(tree))
        msgs = [_message(MSG_SYMBOL_TABLE, struct.pack("<QQ", btree, heap))]
        msgs += [_attr_message(k, v) for k, v in node.attrs.items()]
        return self.object_header(msgs), btree, heap

    def finish(self, root: Node) -> bytes:
        ohdr, btree, heap = self.group(root)
        while len(self.buf) % 8:
            self.buf.append(0)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQIIQQ", 0, ohdr, 1, 0, btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path: str, root: Node) -> None:
    data = _Writer().finish(root)
    with open(path, "wb") as f:
        f.write(data)


# ================================================ Keras weight files ==============================================
def save_keras_weights(path: str, model_name: str, layer_names: List[str], weights: List[List[np.ndarray]],
                       keras_version: str = "2.9.0") -> None:
    """`Model.save_weights(path)` in Keras' legacy HDF5 layout (hdf5_format.save_weights_to_hdf5_group)."""
    root = Node()
    root.attrs["layer_names"] = [n.encode() for n in layer_names]
    root.attrs["backend"] = b"tensorflow"
    root.attrs["keras_version"] = keras_version.encode()
    for lname, (kernel, bias) in zip(layer_names, weights):
        g = Node()
        wnames = [f"{model_name}/{lname}/kernel:0", f"{model_name}/{lname}/bias:0"]
        g.attrs["weight_names"] = [w.encode() for w in wnames]
        for wname, arr in zip(wnames, (kernel, bias)):
            cur = g
            parts = wname.split("/")
            for p in parts[:-1]:
                cur = cur.children.setdefault(p, Node())
            leaf = Node()
            leaf.data = np.ascontiguousarray(arr, dtype=np.float32)
            cur.children[parts[-1]] = leaf
        root.children[lname] = g
    write_h5(path, root)


def _names(v) -> List[str]:
    arr = np.asarray(v, dtype=object).reshape(-1)
    return [x.decode("utf-8") if isinstance(x, bytes) else str(x) for x in arr]


def load_keras_weights(path: str, layer_names: List[str]) -> List[np.ndarray]:
    """[kernel, bias] arrays per layer of `layer_names` (hdf5_format.load_weights_from_hdf5_group): follows the
    `weight_names` attribute of each layer group; falls back to the datasets found below the group."""
    root = read_h5(path)
    if "model_weights" in root.children:                      # a full `model.save()` file keeps them one level down
        root = root.children["model_weights"]
    out = []
    for lname in layer_names:
        if lname not in root.children:
            raise Hdf5Error(f"{path}: no layer group '{lname}' (file has {sorted(root.children)})")
        g = root.children[lname]
        found = {}
        if "weight_names" in g.attrs:
            for w in _names(g.attrs["weight_names"]):
                found[w.rsplit("/", 1)[-1]] = np.asarray(g[w].data, dtype=np.float32)
        else:
            for p, a in g.datasets():
                found[p.rsplit("/", 1)[-1]] = np.asarray(a, dtype=np.float32)
        if "kernel:0" not in found or "bias:0" not in found:
            raise Hdf5Error(f"{path}: layer '{lname}' has no kernel:0 / bias:0 ({sorted(found)})")
        out += [found["kernel:0"], found["bias:0"]]
    return out
