"""Roofline numbers for bench.py: the dominant kernel timed alone with CUDA events on the launching stream
(burst peaks apply), and the secondary BASELINE metric (800x800 render ms/frame)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_FWD_PER_SAMPLE = 1_186_816
FLOP_TRAIN_PER_SAMPLE = 3_489_024


def _time_ms(fn, iters=5, warmup=2):
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


def dominant_kernel_roofline(model, precision, peaks):
    """MLP stage of the fine network (the > 95 % FLOP stage): forward and backward timed separately on the
    training shapes of the bench (R = ray_chunks rays x 192 samples)."""
    from keras_nerf_b200 import _lib
    dev = model.device
    R, S = model.ray_chunks, model.n_coarse + model.n_fine
    rows = R * S
    g = torch.Generator(device="cpu").manual_seed(0)
    o = torch.zeros(R, 3, device=dev)
    o[:, 2] = 4.0
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(dev)
    t = torch.sort(torch.rand(R, S, generator=g) * 4 + 2, dim=-1).values.to(dev).contiguous()
    rgbs = torch.empty(R, S, 4, device=dev)
    dpre = (torch.randn(R, S, 4, generator=g) * 1e-4).to(dev)
    grads = torch.zeros_like(model.fine.params)
    prec = model._prec
    packed = model._packed_ptr("fine")
    ws, wsn = model._ws.data_ptr(), model._ws.numel()
    lib = _lib.load()

    def fwd():
        _lib.call("knerf_mlp_forward", C.byref(model.cfg), _lib.ptr(model.fine.params), packed, _lib.ptr(o), _lib.ptr(d),
                  _lib.ptr(t), R, S, prec, 1, _lib.ptr(rgbs), ws, wsn, _lib.stream())

    def bwd():
        _lib.call("knerf_mlp_backward", C.byref(model.cfg), _lib.ptr(model.fine.params), packed, _lib.ptr(dpre), R, S,
                  prec, _lib.ptr(grads), ws, wsn, _lib.stream())

    l0 = lib.knerf_launch_count()
    fwd()
    l1 = lib.knerf_launch_count()
    bwd()
    l2 = lib.knerf_launch_count()
    ms_f = _time_ms(fwd)
    ms_b = _time_ms(bwd)
    tf_f = FLOP_FWD_PER_SAMPLE * rows / (ms_f * 1e-3) / 1e12
    tf_b = (FLOP_TRAIN_PER_SAMPLE - FLOP_FWD_PER_SAMPLE) * rows / (ms_b * 1e-3) / 1e12
    peak = peaks["bf16_tflops"]
    dominant = "mlp_backward" if ms_b > ms_f else "mlp_forward"
    ach = tf_b if ms_b > ms_f else tf_f
    kernel = ("sgemm_kernel/wgrad_kernel (fp32 SIMT FFMA; measured against the bf16 tensor peak)"
              if precision == "fp32" else "tc_mlp (tcgen05 bf16)")
    return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
            "peak_source": f"{peaks['source']} burst bf16 (MEASURED_PEAKS.json)", "kernel": kernel, "stage": dominant,
            "algorithmic_flop_per_sample": {"forward": FLOP_FWD_PER_SAMPLE,
                                            "backward": FLOP_TRAIN_PER_SAMPLE - FLOP_FWD_PER_SAMPLE},
            "samples_per_launch": rows,
            "forward": {"ms": ms_f, "tflops": tf_f, "frac": tf_f / peak, "launches": int(l1 - l0)},
            "backward": {"ms": ms_b, "tflops": tf_b, "frac": tf_b / peak, "launches": int(l2 - l1)}}


def render_ms_per_frame(precision, dev, wh=800, frames=2):
    """BASELINE config[2]: 800x800 render, 64 coarse + 128 fine, white background (inference.py path)."""
    from keras_nerf_b200 import NeRF
    from keras_nerf_b200.data.synthetic import SyntheticScene
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(42)
    model = NeRF(precision=precision, device=dev)
    model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=wh, image_width=wh, ray_chunks=32000,
                  white_background=True, is_training=False)
    scene = SyntheticScene(wh, model.n_coarse, n_views=40, device=dev)
    views = [scene.view(k, seed=k)[1] for k in range(2)]

    def one(k=[0]):
        o, d, t = views[k[0] % 2]
        k[0] += 1
        model.predict_and_render_images((o[None], d[None], t[None]), seed=k[0])

    ms = _time_ms(one, iters=frames, warmup=1)
    samples = wh * wh * (model.n_coarse + model.n_coarse + model.n_fine)
    return {"metric": "render_ms_per_frame_800x800", "value": ms, "unit": "ms", "n_gpus": 1,
            "tflops": FLOP_FWD_PER_SAMPLE * samples / (ms * 1e-3) / 1e12, "ray_chunks": 32000,
            "precision_mode": precision}
