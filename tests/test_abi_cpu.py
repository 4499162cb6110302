"""CPU (-m "not gpu"): the C-ABI library loads, exports every symbol include/knerf.h declares, and its
host-side logic (model geometry, sizing, argument validation) behaves -- no compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from keras_nerf_b200 import _lib
    return _lib


def test_header_symbols_exported(lib):
    header = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("knerf.h", "knerf_debug.h"))
    declared = set(re.findall(r"\b(knerf_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    raw = C.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in knerf.h but not exported"
    assert declared == set(lib.SIGNATURES), "ctypes table out of sync with knerf.h"
    assert lib.load().knerf_abi_version() == 2


def test_no_torch_types_in_abi():
    header = open(os.path.join(ROOT, "include", "knerf.h")).read()
    assert "torch" not in header.lower().replace("pytorch", "") and "at::" not in header


@pytest.mark.parametrize("cfgkw,dx,dd", [({}, 0, 0), ({}, 99, 99), (dict(n_layers=5, dense_units=64), 15, 9),
                                         (dict(n_layers=3, dense_units=32, skip_layer=1), 0, 0)])
def test_layer_table_matches_oracle(lib, cfgkw, dx, dd):
    oc = O.NerfConfig(**cfgkw)
    cfg = lib.Config(oc.n_coarse, oc.n_fine, oc.pos_emb_xyz, oc.pos_emb_dir, oc.n_layers, oc.dense_units,
                     oc.skip_layer, dx, dd)
    shapes = O.layer_shapes(oc, dx or None, dd or None)
    L = lib.load()
    assert L.knerf_param_count(C.byref(cfg)) == O.param_count(oc, dx or None, dd or None)
    n = len(shapes)
    ko, bo = (C.c_int64 * n)(), (C.c_int64 * n)()
    fi, fo = (C.c_int32 * n)(), (C.c_int32 * n)()
    assert L.knerf_layer_table(C.byref(cfg), n, ko, bo, fi, fo) == n
    off = 0
    for i, (_, a, b) in enumerate(shapes):
        assert (fi[i], fo[i]) == (a, b)
        assert ko[i] == off and bo[i] == off + a * b
        off += a * b + b


def test_flagship_param_count(lib):
    cfg = lib.Config(64, 128, 10, 4, 8, 256, 4, 0, 0)
    assert lib.load().knerf_param_count(C.byref(cfg)) == 595844      # 24 tensors per net


def test_workspace_sizing_and_errors(lib):
    L = lib.load()
    cfg = lib.Config(64, 128, 10, 4, 8, 256, 4, 0, 0)
    small = L.knerf_workspace_bytes(C.byref(cfg), 1024 * 192, 0, 0)
    train = L.knerf_workspace_bytes(C.byref(cfg), 1024 * 192, 0, 1)
    assert 0 < small < train
    assert L.knerf_workspace_bytes(C.byref(cfg), 2048 * 192, 0, 1) > train
    assert L.knerf_workspace_bytes(C.byref(cfg), 10, 7, 0) < 0                      # unknown precision
    bad = lib.Config(64, 128, 10, 4, 0, 256, 4, 0, 0)
    assert L.knerf_param_count(C.byref(bad)) < 0
    assert b"n_layers" in L.knerf_last_error()
    # argument validation happens before any CUDA call: safe without a GPU
    assert L.knerf_generate_rays(None, 4, 4, 1.0, 2.0, 6.0, 8, None, 0, None, None, None, None) == -1
    assert b"null" in L.knerf_last_error()
    assert L.knerf_sample_fine(None, None, None, None, 0, None, 1, 64, 128, 0, None, None, None, None, None, None) == -1
    assert L.knerf_composite_forward(None, None, None, None, 1, 64, 1, 1, 1e-10, None, None, None, None, None) == -1
    with pytest.raises(lib.KnerfError):
        lib.call("knerf_adam_step", None, None, None, None, 10, 1e-3, 0.9, 0.999, 1e-7, 1, 1, None)


def test_comm_entry_points_validate_arguments(lib):
    """a14 behind the C ABI (knerf_comm_*, knerf_allreduce_grads, knerf_train_chunk_dp): argument validation happens
    before NCCL or CUDA is touched, so it is checkable here; the collective itself is tests/test_gpu_multi.py"""
    L = lib.load()
    comm = C.c_void_p()
    assert L.knerf_comm_unique_id(None) == -1
    assert L.knerf_comm_create(None, 0, 2, C.byref(comm)) == -1
    ident = (C.c_ubyte * lib.COMM_ID_BYTES)()
    assert L.knerf_comm_create(ident, 2, 2, C.byref(comm)) == -1 and b"rank 2 of 2" in L.knerf_last_error()
    assert L.knerf_comm_adopt(None, 0, 1, C.byref(comm)) == -1
    assert L.knerf_allreduce_grads(None, None, 10, None) == -1
    assert L.knerf_comm_rank(None, None, None) == -1
    assert L.knerf_comm_destroy(None) == 0                      # destroying nothing is fine
    # the option bits live in the precision argument and do not disturb the sizing entry point
    cfg = lib.Config(64, 128, 10, 4, 8, 256, 4, 0, 0)
    base = L.knerf_workspace_bytes(C.byref(cfg), 4096, lib.FP32, 1)
    assert L.knerf_workspace_bytes(C.byref(cfg), 4096, lib.FP32 | lib.TC_ORDERED | lib.BWD_DGRAD_ONLY, 1) == base
    assert L.knerf_workspace_bytes(C.byref(cfg), 4096, lib.FP32_TC, 1) > base       # + the split weight operand blobs
    # ... except KNERF_REC_FP8, which halves the training records of the bf16 mode (and nothing else)
    rows = 128 * 1000
    bf = L.knerf_workspace_bytes(C.byref(cfg), rows, lib.BF16, 1)
    f8 = L.knerf_workspace_bytes(C.byref(cfg), rows, lib.BF16 | lib.REC_FP8, 1)
    assert (bf - f8) == 1000 * ((576 + 548) - (304 + 258)) * 1024                    # per 128-sample tile, tc_layout.cuh
    assert L.knerf_workspace_bytes(C.byref(cfg), rows, lib.BF16 | lib.REC_FP8, 0) == L.knerf_workspace_bytes(
        C.byref(cfg), rows, lib.BF16, 0)
    assert L.knerf_workspace_bytes(C.byref(cfg), 4096, lib.FP32 | lib.REC_FP8, 1) == base
    # shapes the fused bf16 kernels take (csrc/api.cu tc_chain_map): <= 8 layers of <= 256 units, at most one skip concat,
    # not into the heads, no more than 10 / 4 encoding frequencies
    def packed(n_layers=8, units=256, skip=4, lx=10, ld=4):
        c = lib.Config(64, 128, lx, ld, n_layers, units, skip, 0, 0)
        return L.knerf_packed_weight_bytes(C.byref(c))
    for nl, sk in ((8, 4), (6, 4), (6, 3), (4, 4), (7, 4), (8, 8), (2, 4), (1, 4)):
        assert packed(nl, 256, sk) == packed(), (nl, sk)
    for nl, sk in ((8, 3), (8, 5), (5, 4), (8, 2), (9, 4), (7, 3)):
        assert packed(nl, 256, sk) < 0, (nl, sk)
    assert packed(units=128) == packed() and packed(units=64, n_layers=4, skip=2) == packed()   # zero-padded to 256
    assert packed(units=512) < 0 and packed(units=258) < 0 and packed(lx=11) < 0 and packed(ld=5) < 0
    assert packed(lx=6, ld=2) == packed() and b"dense_units <= 256" in L.knerf_last_error()

    # the whole (n_layers, skip_layer) grid against the rule stated in DESIGN.md §2: layers taking the concat are i + 1
    # for every skip point 0 < i < n_layers with i % skip == 0 (mlp.py:36-38); at most one of them, not the heads (i =
    # n_layers - 1), it must fit at chain layer 5 (<= 5 layers in front of it, <= 2 behind)
    def embeds(nl, sk):
        if nl > 8:
            return False
        pts = [i for i in range(1, nl) if i % sk == 0]
        if any(i == nl - 1 for i in pts):
            return False
        cons = [i + 1 for i in pts]
        if len(cons) > 1:
            return False
        return not cons or (cons[0] <= 5 and nl - 1 - cons[0] <= 2)
    for nl in range(1, 11):
        for sk in range(1, 11):
            assert (packed(nl, 256, sk) > 0) == embeds(nl, sk), (nl, sk)
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "knerf.h")).read()
    for name, val in (("KNERF_TC_ORDERED", lib.TC_ORDERED), ("KNERF_BWD_DGRAD_ONLY", lib.BWD_DGRAD_ONLY),
                      ("KNERF_BWD_WGRAD_ONLY", lib.BWD_WGRAD_ONLY), ("KNERF_REC_FP8", lib.REC_FP8)):
        assert int(re.search(rf"#define {name} (0x[0-9a-fA-F]+)", hdr).group(1), 16) == val


def test_product_path_has_no_cpu_fallback(lib):
    import torch
    from keras_nerf_b200 import NeRFUtils
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.KnerfError):
        NeRFUtils(1, 2, 2, 4, 10, 4).positional_encoding(np.zeros((4, 3), np.float32), 10)
    # and nothing under keras_nerf_b200/ imports the oracle
    for dp, _, files in os.walk(os.path.join(ROOT, "keras_nerf_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_camera_helpers_vs_golden():
    from conftest import load_golden
    from keras_nerf_b200 import get_focal_from_fov, pose_spherical
    g = load_golden("camera")
    assert get_focal_from_fov(0.6911112070083618, 100) == pytest.approx(138.88887889922103)
    for th, pose in zip(g["thetas"], g["poses"]):
        np.testing.assert_allclose(pose_spherical(float(th), float(g["phi"]), float(g["radius"])), pose, atol=1e-6)
