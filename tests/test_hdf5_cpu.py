"""f2: Keras-HDF5 weight files without h5py (keras_nerf/model/nerf/nerf.py:63-64,132-136).  The writer is checked
structure by structure against the HDF5 File Format Specification and round-tripped through the reader."""
import struct

import numpy as np
import pytest

from keras_nerf_b200.utils import hdf5 as H

LAYERS = [f"layer_{i}" for i in range(8)] + ["sigma", "features", "rgb_features", "rgb"]


def _weights(rng):
    shapes = [(63, 256)] + [(256, 256)] * 4 + [(319, 256)] + [(256, 256)] * 2 + [(256, 1), (256, 256), (283, 128), (128, 3)]
    return [[rng.standard_normal(s).astype(np.float32), rng.standard_normal(s[1]).astype(np.float32)] for s in shapes]


def test_keras_weight_file_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    w = _weights(rng)
    p = str(tmp_path / "coarse.h5")
    H.save_keras_weights(p, "coarse_nerf", LAYERS, w)
    back = H.load_keras_weights(p, LAYERS)
    assert len(back) == 24
    for (k, b), k2, b2 in zip(w, back[0::2], back[1::2]):
        assert k2.dtype == np.float32 and np.array_equal(k, k2) and np.array_equal(b, b2)
    root = H.read_h5(p)
    assert [x.decode() for x in root.attrs["layer_names"]] == LAYERS
    assert root.attrs["backend"] == b"tensorflow" and root.attrs["keras_version"] == b"2.9.0"
    g = root["layer_5"]
    assert [x.decode() for x in g.attrs["weight_names"]] == ["coarse_nerf/layer_5/kernel:0", "coarse_nerf/layer_5/bias:0"]
    assert g["coarse_nerf/layer_5/kernel:0"].data.shape == (319, 256)
    assert [n for n, _ in root.datasets()][:2] == ["features/coarse_nerf/features/bias:0",
                                                   "features/coarse_nerf/features/kernel:0"]
    with pytest.raises(H.Hdf5Error):
        H.load_keras_weights(p, ["no_such_layer"])


def test_file_structures_follow_the_specification(tmp_path):
    p = str(tmp_path / "t.h5")
    root = H.Node()
    root.attrs["ints"] = np.arange(5, dtype=np.int64)
    root.attrs["scalar"] = np.float64(2.5)
    ds = H.Node()
    ds.data = np.arange(12, dtype=np.float32).reshape(3, 4)
    ds.attrs["note"] = "hello"
    root.children["g"] = H.Node()
    root.children["g"].children["x"] = ds
    root.children["empty"] = H.Node()
    H.write_h5(p, root)
    b = open(p, "rb").read()
    # superblock v0: signature, versions, 8-byte offsets/lengths, EOF address, root symbol-table entry
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8:13] == bytes(5) and b[13] == 8 and b[14] == 8
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert base == 0 and free == H.UNDEF and eof == len(b) and drv == H.UNDEF and len(b) % 8 == 0
    name_off, ohdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and cache == 1 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    # root object header v1: version, message count, reference count, first message = symbol table (0x0011)
    ver, _, nmsg, refs, hsize = struct.unpack_from("<BBHII", b, ohdr)
    assert (ver, nmsg, refs) == (1, 3, 1) and hsize % 8 == 0
    mtype, msize = struct.unpack_from("<HH", b, ohdr + 16)
    assert mtype == 0x11 and msize == 16 and struct.unpack_from("<QQ", b, ohdr + 24) == (btree, heap)
    # group B-tree: leaf level, one child, keys are local-heap offsets; symbol node sorted by name
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, btree + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, H.UNDEF, H.UNDEF)
    key0, snod, key1 = struct.unpack_from("<QQQ", b, btree + 24)
    assert key0 == 0 and b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<H", b, snod + 6)[0] == 2
    seg_size, free_off, seg = struct.unpack_from("<QQQ", b, heap + 8)
    names = [b[seg + struct.unpack_from("<Q", b, snod + 8 + 40 * i)[0]:].split(b"\0")[0] for i in range(2)]
    assert names == [b"empty", b"g"] and b[seg + key1:].split(b"\0")[0] == b"g" and free_off + 32 == seg_size
    back = H.read_h5(p)
    assert np.array_equal(back.attrs["ints"], np.arange(5)) and float(back.attrs["scalar"]) == 2.5
    x = back["g/x"]
    assert x.data.dtype == np.float32 and np.array_equal(x.data, ds.data) and x.attrs["note"] == b"hello"
    assert back["empty"].children == {} and not back["empty"].is_dataset


def test_reader_rejects_what_it_does_not_implement(tmp_path):
    p = tmp_path / "bad.h5"
    p.write_bytes(b"not hdf5 at all")
    with pytest.raises(H.Hdf5Error):
        H.read_h5(str(p))
    good = tmp_path / "ok.h5"
    H.write_h5(str(good), H.Node())
    data = bytearray(good.read_bytes())
    data[8] = 2                                               # superblock v2 (libver='latest')
    p.write_bytes(bytes(data))
    with pytest.raises(H.Hdf5Error):
        H.read_h5(str(p))


def test_reader_handles_variable_length_strings_and_continuations():
    """Newer Keras versions store attribute strings as variable-length (global heap) strings and libhdf5 moves
    attributes into continuation blocks: build such a header by hand and read it."""
    buf = bytearray(H.SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0) + bytes(72))
    # global heap collection with two strings
    gcol = len(buf)
    objs = b""
    for i, s in enumerate([b"layer_0", b"sigma"], start=1):
        objs += struct.pack("<HHIQ", i, 1, 0, len(s)) + s + b"\0" * (H._pad8(len(s)) - len(s))
    body = objs + struct.pack("<HHIQ", 0, 0, 0, 0)
    buf += b"GCOL" + struct.pack("<B3xQ", 1, 16 + len(body)) + body
    # attribute message (v1) 'layer_names': vlen string type (class 9, base = 1-byte char), shape (2,)
    vlen_dt = struct.pack("<BBBBI", 0x19, 0x01, 0x00, 0, 16) + struct.pack("<BBBBI", 0x10, 0, 0, 0, 1) + struct.pack("<HH", 0, 8)
    ds = H._ds_message((2,))
    name = b"layer_names\0"
    attr = struct.pack("<BBHHH", 1, 0, len(name), len(vlen_dt), len(ds))
    attr += name + b"\0" * (H._pad8(len(name)) - len(name)) + vlen_dt + b"\0" * (H._pad8(len(vlen_dt)) - len(vlen_dt)) + ds
    attr += struct.pack("<IQI", 7, gcol, 1) + struct.pack("<IQI", 5, gcol, 2)
    cont_block = H._message(H.MSG_ATTRIBUTE, attr)
    while len(buf) % 8:
        buf.append(0)
    cont = len(buf)
    buf += cont_block
    # an (empty) group: local heap, B-tree without entries, header = symbol table + continuation -> attribute
    seg = len(buf)
    buf += bytes(8) + struct.pack("<QQ", 1, 24) + bytes(8)
    heap = len(buf)
    buf += b"HEAP" + struct.pack("<B3xQQQ", 0, 32, 8, seg)
    btree = len(buf)
    buf += b"TREE" + struct.pack("<BBHQQ", 0, 0, 0, H.UNDEF, H.UNDEF) + bytes(8 * 17)
    msgs = H._message(H.MSG_SYMBOL_TABLE, struct.pack("<QQ", btree, heap)) + \
        H._message(H.MSG_CONTINUATION, struct.pack("<QQ", cont, len(cont_block)))
    ohdr = len(buf)
    buf += struct.pack("<BBHII4x", 1, 0, 3, 1, len(msgs)) + msgs
    struct.pack_into("<QQQQ", buf, 24, 0, H.UNDEF, len(buf), H.UNDEF)
    struct.pack_into("<QQIIQQ", buf, 56, 0, ohdr, 1, 0, btree, heap)
    r = H._Reader(bytes(buf))
    node = r.node(r.root_header)
    assert H._names(node.attrs["layer_names"]) == ["layer_0", "sigma"] and node.children == {}
