#!/usr/bin/env python
"""Time the tcgen05 MLP kernels alone (CUDA events on the launching stream): forward (inference / training),
dgrad, wgrad.  usage: python benchmarks/tc_kernels.py [rays=8192] [samples_per_ray=192] [iters=10] [records=bf16|fp8]"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from benchmarks.roofline import (FLOP_DGRAD_PER_SAMPLE, FLOP_FWD_INFER_EXECUTED, FLOP_FWD_PER_SAMPLE,  # noqa: E402
                                 FLOP_WGRAD_PER_SAMPLE, _time_ms)


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 192
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    records = sys.argv[4] if len(sys.argv) > 4 else "bf16"
    from keras_nerf_b200 import NeRF, _lib
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    dev = torch.device("cuda", 0)
    mlp_mod.set_seed(42)
    model = NeRF(precision="bf16", device=dev, n_coarse=64, n_fine=S - 64, records=records)
    model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=R // 64, image_width=64, ray_chunks=R,
                  white_background=True)
    rows = R * S
    g = torch.Generator(device="cpu").manual_seed(0)
    o = torch.zeros(R, 3, device=dev)
    o[:, 2] = 4.0
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    t = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, dim=-1).values.to(dev).contiguous()
    rgbs = torch.empty(R, S, 4, device=dev)
    dpre = (torch.randn(R, S, 4, generator=g) * 1e-4).to(dev)
    grads = torch.zeros_like(model.fine.params)
    prec = model._prec_train
    packed = model._packed_ptr("fine")
    ws, wsn = model._ws.data_ptr(), model._ws.numel()
    lib = _lib.load()

    def fwd(train):
        _lib.call("knerf_mlp_forward", C.byref(model.cfg), _lib.ptr(model.fine.params), packed, _lib.ptr(o), _lib.ptr(d),
                  _lib.ptr(t), R, S, prec, train, _lib.ptr(rgbs), ws, wsn, _lib.stream())

    def bwd(flags=0):   # flags: _lib.BWD_DGRAD_ONLY / _lib.BWD_WGRAD_ONLY time the two backward kernels apart
        _lib.call("knerf_mlp_backward", C.byref(model.cfg), _lib.ptr(model.fine.params), packed, _lib.ptr(dpre), R, S,
                  prec | flags, _lib.ptr(grads), ws, wsn, _lib.stream())

    out = {"rays": R, "samples_per_ray": S, "rows": rows, "records": records}

    def rec(name, ms, flop):
        out[name] = {"ms": round(ms, 4), "tflops": round(flop * rows / ms / 1e9, 1), "ns_per_sample": round(ms * 1e6 / rows, 4)}

    rec("fwd_infer", _time_ms(lambda: fwd(0), iters), FLOP_FWD_PER_SAMPLE)
    out["fwd_infer"]["tflops_executed"] = round(out["fwd_infer"]["tflops"] * FLOP_FWD_INFER_EXECUTED / FLOP_FWD_PER_SAMPLE, 1)
    rec("fwd_train", _time_ms(lambda: fwd(1), iters), FLOP_FWD_PER_SAMPLE)
    for name, flag, flop in (("dgrad", _lib.BWD_DGRAD_ONLY, FLOP_DGRAD_PER_SAMPLE),
                             ("wgrad", _lib.BWD_WGRAD_ONLY, FLOP_WGRAD_PER_SAMPLE)):
        rec(name, _time_ms(lambda: bwd(flag), iters), flop)
    out["train_ns_per_sample"] = round(sum(out[k]["ns_per_sample"] for k in ("fwd_train", "dgrad", "wgrad")), 4)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
