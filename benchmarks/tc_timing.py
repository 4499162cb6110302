#!/usr/bin/env python
"""In-kernel cycle breakdown of the tcgen05 forward kernel (needs a -DKNERF_TC_TIMING build:
KNERF_EXTRA_NVCC_FLAGS=-DKNERF_TC_TIMING python -c 'import __graft_entry__ as g; g.build(force=True)')."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    S = 192
    from keras_nerf_b200 import NeRF, _lib
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    dev = torch.device("cuda", 0)
    mlp_mod.set_seed(42)
    records = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    model = NeRF(precision="bf16", device=dev, records=records)
    model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=R // 64, image_width=64, ray_chunks=R,
                  white_background=True)
    g = torch.Generator(device="cpu").manual_seed(0)
    o = torch.zeros(R, 3, device=dev)
    o[:, 2] = 4.0
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    t = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, dim=-1).values.to(dev).contiguous()
    rgbs = torch.empty(R, S, 4, device=dev)
    packed = model._packed_ptr("fine")
    ws, wsn = model._ws.data_ptr(), model._ws.numel()
    lib = _lib.load()
    buf = (C.c_ulonglong * (160 * 40))()
    for train in (0, 1):
        for it in range(3):
            _lib.call("knerf_mlp_forward", C.byref(model.cfg), _lib.ptr(model.fine.params), packed, _lib.ptr(o),
                      _lib.ptr(d), _lib.ptr(t), R, S, model._prec_train, train, _lib.ptr(rgbs), ws, wsn, _lib.stream())
            torch.cuda.synchronize()
            n = lib.knerf_debug_tc_timing(buf, 160 * 40)
        if n == 0:
            print("library built without -DKNERF_TC_TIMING")
            return
        a = np.frombuffer(buf, dtype=np.uint64).reshape(160, 40)[:148].astype(np.float64)
        tiles = R * S / 128 / 148
        m = a.mean(0)
        print(f"train={train}: tiles/CTA {tiles:.1f}; per CTA kcycles: producer-wait-empty {m[0]/1e3:.0f}, "
              f"mma-wait-A {m[1]/1e3:.0f}, mma-wait-stage {m[2]/1e3:.0f}, mma-total {m[3]/1e3:.0f}, "
              f"compute-wait-acc {m[4]/1e3:.0f}, compute-total {m[5]/1e3:.0f}, mma-wait-peer-stage {m[38]/1e3:.0f}")
        units = tiles / 2
        for tl in range(2):
            print(f"  epilogue cycles per step, tile slot {tl}: " + " ".join(f"{m[6 + tl*16 + s]/units:.0f}" for s in range(10)))


if __name__ == "__main__":
    main()
