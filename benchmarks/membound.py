#!/usr/bin/env python
"""HBM roofline of the memory-bound per-ray kernels (timed alone, CUDA events, inputs >> L2, burst peak).
usage: python benchmarks/membound.py [R]   -> one JSON object per kernel"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from keras_nerf_b200 import _lib  # noqa: E402


def time_ms(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 640000
    peak = 6549.1
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(0)
    out = []

    def report(name, bytes_per_ray, ms, rays=R):
        gbs = bytes_per_ray * rays / (ms * 1e-3) / 1e9
        out.append({"kernel": name, "rays": rays, "ms": round(ms, 4), "algorithmic_bytes_per_ray": bytes_per_ray,
                    "achieved_GBps": round(gbs, 1), "peak_GBps": peak, "frac": round(gbs / peak, 3)})

    for S in (64, 192):
        rgbs = torch.rand(R, S, 4, device=dev)
        rgbs[..., 3] *= 5
        t = torch.sort(torch.rand(R, S, device=dev) * 4 + 2, dim=-1).values.contiguous()
        img, dep, w, acc = (torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, S, device=dev),
                            torch.empty(R, device=dev))
        ms = time_ms(lambda: _lib.call("knerf_composite_forward", _lib.ptr(rgbs), None, None, _lib.ptr(t), R, S, 1, 1,
                                       1e-10, _lib.ptr(img), _lib.ptr(dep), _lib.ptr(w), _lib.ptr(acc), _lib.stream()))
        report(f"composite_fwd S={S}", 24 * S + 20, ms)
        tgt = torch.rand(R, 3, device=dev)
        dout, sq = torch.empty(R, S, 4, device=dev), torch.empty(R, device=dev)
        ms = time_ms(lambda: _lib.call("knerf_composite_backward", _lib.ptr(rgbs), _lib.ptr(t), R, S, 1, 1, 1e-10, None,
                                       _lib.ptr(tgt), 1e-5, 1, _lib.ptr(dout), _lib.ptr(sq), _lib.stream()))
        report(f"composite_bwd S={S}", 36 * S + 24, ms)
        del rgbs, t, w, dout
    for Nf, seq in ((128, 0), (128, 1), (256, 0)):
        Nc = 64
        tc = torch.sort(torch.rand(R, Nc, device=dev) * 4 + 2, dim=-1).values.contiguous()
        w = torch.rand(R, Nc, device=dev) ** 4
        u = torch.rand(R, Nf, device=dev)
        ts = torch.empty(R, Nc + Nf, device=dev)
        flags = _lib.OOB_ZERO | (_lib.SCAN_SEQUENTIAL if seq else 0)
        ms = time_ms(lambda: _lib.call("knerf_sample_fine", _lib.ptr(tc), None, _lib.ptr(w), _lib.ptr(u), 0, None, R, Nc,
                                       Nf, flags, _lib.ptr(ts), None, None, None, None, _lib.stream()))
        report(f"sample_fine Nf={Nf} {'sequential' if seq else 'warp'}-scan", 8 * Nc + 4 * Nf + 4 * (Nc + Nf), ms)
        del tc, w, u, ts
    H = W = 800
    import numpy as np
    from keras_nerf_b200 import pose_spherical
    c2w = np.ascontiguousarray(pose_spherical(30.0, -30.0, 4.0))
    o, d, t = (torch.empty(H * W, 3, device=dev), torch.empty(H * W, 3, device=dev), torch.empty(H * W, 64, device=dev))
    ms = time_ms(lambda: _lib.call("knerf_generate_rays", c2w.ctypes.data, H, W, 1111.0, 2.0, 6.0, 64, None, 7,
                                   _lib.ptr(o), _lib.ptr(d), _lib.ptr(t), _lib.stream()))
    report("generate_rays 800x800 N=64 (Philox)", 24 + 4 * 64, ms, rays=H * W)
    for (ih, iw), (oh, ow) in (((800, 800), (400, 400)), ((800, 800), (800, 800)), ((4096, 4096), (2048, 2048))):
        src = torch.randint(0, 256, (ih, iw, 4), dtype=torch.uint8, device=dev)
        dst = torch.empty(oh, ow, 4, device=dev)
        ms = time_ms(lambda: _lib.call("knerf_image_prepare", _lib.ptr(src, torch.uint8), ih, iw, oh, ow, 1,
                                       _lib.ptr(dst), _lib.stream()))
        report(f"image_prepare {ih}x{iw} -> {oh}x{ow} (per output pixel)", 16 + 4.0 * ih * iw / (oh * ow), ms,
               rays=oh * ow)
    for B_, H_, W_ in ((1, 128, 256), (4, 800, 800)):
        x = torch.rand(B_, H_, W_, 3, device=dev)
        y = torch.rand(B_, H_, W_, 3, device=dev)
        nws = _lib.load().knerf_image_metrics_workspace_floats(B_, H_, W_, 3)
        ws, res = torch.empty(nws, device=dev), torch.empty(2, B_, device=dev)
        ms = time_ms(lambda: _lib.call("knerf_image_metrics", _lib.ptr(x), _lib.ptr(y), B_, H_, W_, 3, 1.0, _lib.ptr(res[0]),
                                       _lib.ptr(res[1]), _lib.ptr(ws), nws, _lib.stream()))
        report(f"image_metrics (MSE + SSIM) {B_}x{H_}x{W_}x3 (per pixel)", 24, ms, rays=B_ * H_ * W_)
    n = 2 * 595844
    pbuf, gbuf, mbuf, vbuf = (torch.zeros(n, device=dev) for _ in range(4))
    ms = time_ms(lambda: _lib.call("knerf_adam_step", _lib.ptr(pbuf), _lib.ptr(gbuf), _lib.ptr(mbuf), _lib.ptr(vbuf), n, 1e-3,
                                   0.9, 0.999, 1e-7, 1, 1, _lib.stream()))
    out.append({"kernel": "adam (1.19M params, L2-resident)", "ms": round(ms, 4)})
    for o_ in out:
        print(json.dumps(o_))


if __name__ == "__main__":
    main()
