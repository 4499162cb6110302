"""NeRFMLP (reference: keras_nerf/model/nerf/mlp.py).  Weights live in ONE flat fp32 CUDA buffer in Keras
variable order (layer_0..layer_{n-1}, sigma, features, rgb_features, rgb; kernel[in,out] then bias)."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np
import torch

from ... import _lib

_init_rng = np.random.default_rng(42)   # the reference scripts call tf.random.set_seed(42) (train.py:10)


def set_seed(seed: int) -> None:
    global _init_rng
    _init_rng = np.random.default_rng(seed)


def _truncated_normal(rng, stddev, shape):
    """Keras' VarianceScaling(distribution='truncated_normal'): samples within two standard deviations, the
    standard deviation corrected by 0.87962566103423978 so that the truncated distribution has the asked variance"""
    sd = stddev / 0.87962566103423978
    out = rng.normal(0.0, sd, size=shape)
    bad = np.abs(out) > 2 * sd
    while bad.any():
        out[bad] = rng.normal(0.0, sd, size=int(bad.sum()))
        bad = np.abs(out) > 2 * sd
    return out


# Keras initializer identifiers accepted by Dense(kernel_initializer=...) (keras_nerf/model/nerf/mlp.py:5,13-27):
# name -> f(rng, fan_in, fan_out) -> [fan_in, fan_out] float64
INITIALIZERS = {
    "glorot_uniform": lambda r, fi, fo: r.uniform(-math.sqrt(6.0 / (fi + fo)), math.sqrt(6.0 / (fi + fo)), (fi, fo)),
    "glorot_normal": lambda r, fi, fo: _truncated_normal(r, math.sqrt(2.0 / (fi + fo)), (fi, fo)),
    "he_uniform": lambda r, fi, fo: r.uniform(-math.sqrt(6.0 / fi), math.sqrt(6.0 / fi), (fi, fo)),
    "he_normal": lambda r, fi, fo: _truncated_normal(r, math.sqrt(2.0 / fi), (fi, fo)),
    "lecun_uniform": lambda r, fi, fo: r.uniform(-math.sqrt(3.0 / fi), math.sqrt(3.0 / fi), (fi, fo)),
    "lecun_normal": lambda r, fi, fo: _truncated_normal(r, math.sqrt(1.0 / fi), (fi, fo)),
    "random_uniform": lambda r, fi, fo: r.uniform(-0.05, 0.05, (fi, fo)),
    "random_normal": lambda r, fi, fo: r.normal(0.0, 0.05, (fi, fo)),
    "zeros": lambda r, fi, fo: np.zeros((fi, fo)),
    "ones": lambda r, fi, fo: np.ones((fi, fo)),
}


class NeRFMLP:
    def __init__(self, n_layers: int = 8, dense_units: int = 256, skip_layer=4, initializer='glorot_uniform',
                 name=None, device=None, **kwargs):
        # keras_nerf/model/nerf/mlp.py:5-27
        if not callable(initializer) and str(initializer).lower() not in INITIALIZERS:
            raise NotImplementedError(f"initializer {initializer!r}: one of {sorted(INITIALIZERS)} or a callable "
                                      "f(rng, fan_in, fan_out) -> [fan_in, fan_out] array")
        self.initializer = initializer
        self.n_layers = int(n_layers)
        self.dense_units = int(dense_units)
        self.skip_layer = int(skip_layer)
        self.name = name
        self.device = torch.device(device) if device is not None else None
        self.params = None          # flat fp32 CUDA tensor
        self.cfg = None
        self._ws = None

    # ---- construction -------------------------------------------------------------------------
    @property
    def built(self):
        return self.params is not None

    def build(self, dx: int, dd: int, pos_emb_xyz: int = 0, pos_emb_dir: int = 0, n_coarse=0, n_fine=0):
        """Create glorot-uniform kernels / zero biases (Keras Dense defaults) for encodings of width dx, dd."""
        self.device = self.device or _lib.default_device()
        self.cfg = _lib.Config(n_coarse, n_fine, pos_emb_xyz, pos_emb_dir, self.n_layers, self.dense_units,
                               self.skip_layer, dx, dd)
        lib = _lib.load()
        n = lib.knerf_param_count(C.byref(self.cfg))
        if n < 0:
            raise _lib.KnerfError(lib.knerf_last_error().decode())
        nl = self.n_layers + 4
        k_off, b_off = (C.c_int64 * nl)(), (C.c_int64 * nl)()
        fin, fout = (C.c_int32 * nl)(), (C.c_int32 * nl)()
        got = lib.knerf_layer_table(C.byref(self.cfg), nl, k_off, b_off, fin, fout)
        if got != nl:
            raise _lib.KnerfError(lib.knerf_last_error().decode())
        self.layers = [(int(k_off[i]), int(b_off[i]), int(fin[i]), int(fout[i])) for i in range(nl)]
        host = np.zeros(n, dtype=np.float32)
        init = self.initializer if callable(self.initializer) else INITIALIZERS[str(self.initializer).lower()]
        for ko, bo, fi, fo in self.layers:    # kernels in Keras variable order; biases stay zero (Dense default)
            host[ko:ko + fi * fo] = np.asarray(init(_init_rng, fi, fo)).astype(np.float32).reshape(-1)
        self.params = torch.from_numpy(host).to(self.device)
        return self

    @property
    def layer_names(self):
        return [f"layer_{i}" for i in range(self.n_layers)] + ["sigma", "features", "rgb_features", "rgb"]

    @property
    def trainable_variables(self):
        """24 views (kernel[in,out], bias[out] per Dense) into the flat buffer, Keras order."""
        out = []
        for ko, bo, fi, fo in self.layers:
            out.append(self.params[ko:ko + fi * fo].view(fi, fo))
            out.append(self.params[bo:bo + fo])
        return out

    # ---- call ---------------------------------------------------------------------------------
    def __call__(self, inputs):
        return self.call(inputs)

    def call(self, inputs):
        """keras_nerf/model/nerf/mlp.py:29-50: (xyz_enc[...,dx], dir_enc[...,dd]) -> (rgb[...,3], sigma[...,1])."""
        xyz, dirs = inputs
        device = self.device or _lib.default_device()
        xyz, dirs = _lib.dev(xyz, device), _lib.dev(dirs, device)
        dx, dd = xyz.shape[-1], dirs.shape[-1]
        if not self.built:
            self.build(dx, dd)
        if (dx, dd) != (self.cfg.dx, self.cfg.dd):
            raise ValueError(f"NeRFMLP was built for inputs of width ({self.cfg.dx}, {self.cfg.dd}), got ({dx}, {dd})")
        lead = xyz.shape[:-1]
        rows = xyz.numel() // dx
        rgb = torch.empty(lead + (3,), dtype=torch.float32, device=device)
        sigma = torch.empty(lead + (1,), dtype=torch.float32, device=device)
        lib = _lib.load()
        # bound the scratch: process in slabs of rows
        slab = min(rows, 1 << 18)
        need = lib.knerf_workspace_bytes(C.byref(self.cfg), slab, _lib.FP32, 0)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        x2, d2, rgb2, sig2 = xyz.view(rows, dx), dirs.view(rows, dd), rgb.view(rows, 3), sigma.view(rows, 1)
        with torch.cuda.device(device):
            for s in range(0, rows, slab):
                e = min(rows, s + slab)
                _lib.call("knerf_mlp_forward_encoded", C.byref(self.cfg), _lib.ptr(self.params),
                          _lib.ptr(x2[s:e]), dx, _lib.ptr(d2[s:e]), dd, e - s, _lib.ptr(rgb2[s:e]),
                          _lib.ptr(sig2[s:e]), self._ws.data_ptr(), self._ws.numel(), _lib.stream())
        return rgb, sigma

    def get_config(self):
        # keras_nerf/model/nerf/mlp.py:52-59
        return {'name': self.name, 'n_layers': self.n_layers, 'dense_units': self.dense_units,
                'skip_layer': self.skip_layer}

    def count_params(self):
        return 0 if not self.built else int(self.params.numel())

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        if self.built:
            for name, (ko, bo, fi, fo) in zip(self.layer_names, self.layers):
                print_fn(f"  {name:<14} Dense  ({fi} -> {fo})  params {fi * fo + fo}")
        print_fn(f"Total params: {self.count_params()}")

    # ---- weight I/O (Keras [in,out] layout) ----------------------------------------------------
    # `*.h5` paths are Keras' legacy HDF5 weight files (nerf.py:63-64,134-136) through utils/hdf5.py (h5py is not
    # a dependency); any other path is an .npz archive with the same `<layer>/<kernel:0|bias:0>` names.
    def get_weights(self):
        return [v.detach().cpu().numpy().copy() for v in self.trainable_variables]

    def set_weights(self, weights):
        with torch.no_grad():
            for v, w in zip(self.trainable_variables, weights):
                v.copy_(torch.as_tensor(np.asarray(w, dtype=np.float32)).reshape(v.shape))

    def save_weights(self, path):
        w = self.get_weights()
        if path.endswith(".h5") or path.endswith(".hdf5"):
            from ...utils.hdf5 import save_keras_weights
            save_keras_weights(path, self.name, list(self.layer_names), [w[i:i + 2] for i in range(0, len(w), 2)])
            return
        np.savez(path, **{f"{n}/{k}": a for n, (k, a) in
                          zip(np.repeat(self.layer_names, 2), zip(["kernel:0", "bias:0"] * len(self.layer_names), w))})

    def load_weights(self, path):
        if (path.endswith(".h5") or path.endswith(".hdf5")) and os.path.exists(path):
            from ...utils.hdf5 import load_keras_weights
            self.set_weights(load_keras_weights(path, list(self.layer_names)))
            return
        if path.endswith(".h5"):
            path = path[:-3] + ".npz"                          # checkpoints written before HDF5 support
        if not os.path.exists(path) and os.path.exists(path + ".npz"):
            path = path + ".npz"
        z = np.load(path)
        self.set_weights([z[f"{n}/{k}"] for n in self.layer_names for k in ("kernel:0", "bias:0")])
