#!/usr/bin/env python
"""MLP forward / backward of the two fp32 modes timed alone (fine network, R rays x 192 samples, CUDA events):
    python benchmarks/fp32_kernels.py [R=4096] [fp32] [fp32_tc]
SIMT FFMA (`precision="fp32"`) against the 3-way-bf16-split tcgen05 GEMMs (`precision="fp32_tc"`); numbers in
profiles/r02_fp32_tc.md."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from keras_nerf_b200 import NeRF, _lib
from keras_nerf_b200.model.nerf import mlp as mlp_mod
dev = torch.device("cuda", 0)
R, S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 192
g = torch.Generator().manual_seed(0)
o = torch.zeros(R, 3, device=dev); o[:, 2] = 4.0
d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
t = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, dim=-1).values.to(dev).contiguous()
dpre = (torch.randn(R, S, 4, generator=g) * 1e-4).to(dev)
out = {}
for prec in (sys.argv[2:] or ("fp32", "fp32_tc")):
    mlp_mod.set_seed(42)
    m = NeRF(precision=prec, device=dev)
    m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=R // 64, image_width=64, ray_chunks=R, white_background=True)
    rgbs = torch.empty(R, S, 4, device=dev); grads = torch.zeros_like(m.fine.params)
    def fwd():
        _lib.call("knerf_mlp_forward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(o), _lib.ptr(d), _lib.ptr(t), R, S, m._prec, 1, _lib.ptr(rgbs), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
    def bwd():
        _lib.call("knerf_mlp_backward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(dpre), R, S, m._prec, _lib.ptr(grads), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
    def tm(fn, it=3):
        fn(); fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it
    f, b = tm(fwd), tm(bwd)
    rows = R * S
    out[prec] = {"fwd_ms": f, "bwd_ms": b, "fwd_tflops": 1186816 * rows / f / 1e9, "bwd_tflops": (3489024 - 1186816) * rows / b / 1e9,
                 "train_rays_per_s_mlp_only(fine)": R / ((f + b) * 1e-3)}
print(json.dumps(out))
