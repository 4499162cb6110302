"""Roofline numbers for bench.py: each MLP kernel of the training step timed alone with CUDA events on the
launching stream (burst peaks apply), and the secondary BASELINE metric (800x800 render ms/frame)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work per sample (unpadded shapes)
FLOP_FWD_PER_SAMPLE = 1_186_816            # 593,408 MAC
# bf16 INFERENCE executes fewer: `features` is folded into `rgb_features` (no activation in between, mlp.py:42-46)
# and sigma becomes output column 128 of that step: 593,408 - 256*256 + 288*16 MAC.  Throughput figures keep the
# unfolded count and carry the executed one next to it.
FLOP_FWD_INFER_EXECUTED = 2 * (593_408 - 256 * 256 + 288 * 16)
FLOP_TRAIN_PER_SAMPLE = 3_489_024          # fwd + dgrad + wgrad
FLOP_WGRAD_PER_SAMPLE = 1_186_816          # every weight once more
FLOP_DGRAD_PER_SAMPLE = FLOP_TRAIN_PER_SAMPLE - FLOP_FWD_PER_SAMPLE - FLOP_WGRAD_PER_SAMPLE
# bf16 mode, algorithmic HBM bytes per sample (records of 608 KB + 548 KB per 128 samples)
BYTES_FWD_SAVE = 600 * 1024 / 128                       # forward writes the activation record (+ ReLU' bits) once
BYTES_DGRAD = 8 * 32 + 548 * 1024 / 128                 # reads the 8 x 1-bit ReLU' tiles, writes the dZ record
BYTES_WGRAD = 1196 * 1024 / 128                         # operand units of the 10 weight-gradient tasks
# measured DRAM traffic per sample from `ncu --set full` (profiles/r01_bf16_ncu_full.md, 393,216-sample launches)
NCU_TRAFFIC_PER_SAMPLE = {"tc_mlp_fwd_kernel<train>": (0.002823 + 1.835313) * 1e9 / 393216,
                          "tc_mlp_dgrad_kernel": (0.108100 + 1.666665) * 1e9 / 393216,
                          "tc_wgrad_kernel": (3.750000 + 0.006960) * 1e9 / 393216}


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
    """The MLP kernels of the fine network (> 95 % of the step) on the bench's training shapes
    (R = ray_chunks rays x 192 samples), each timed alone; `roofline` names the one with the largest time."""
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

    bf16_peak, hbm_peak = peaks["bf16_tflops"], peaks["hbm_gbs"]
    src = f"{peaks['source']} (MEASURED_PEAKS.json: cuBLAS bf16 burst, copy bandwidth)"
    kernels = {}

    def add(name, ms, flop_ps, bytes_ps, launches):
        tf = flop_ps * rows / (ms * 1e-3) / 1e12
        k = {"ms": ms, "launches": launches, "tflops": tf, "tensor_frac": tf / bf16_peak}
        if bytes_ps:
            gb = bytes_ps * rows / (ms * 1e-3) / 1e9
            k.update({"hbm_GBps": gb, "hbm_frac": gb / hbm_peak, "algorithmic_bytes_per_sample": bytes_ps})
        if name in NCU_TRAFFIC_PER_SAMPLE:
            k["ncu_dram_bytes_per_sample"] = NCU_TRAFFIC_PER_SAMPLE[name]
        k["algorithmic_flop_per_sample"] = flop_ps
        kernels[name] = k

    l0 = lib.knerf_launch_count()
    fwd()
    n_f = int(lib.knerf_launch_count() - l0)
    ms_f = _time_ms(fwd)
    if precision == "bf16":
        add("tc_mlp_fwd_kernel<train>", ms_f, FLOP_FWD_PER_SAMPLE, BYTES_FWD_SAVE, n_f)
        for name, mask, flop, byt in (("tc_mlp_dgrad_kernel", 1, FLOP_DGRAD_PER_SAMPLE, BYTES_DGRAD),
                                      ("tc_wgrad_kernel", 2, FLOP_WGRAD_PER_SAMPLE, BYTES_WGRAD)):
            lib.knerf_debug_backward_parts(mask)
            try:
                add(name, _time_ms(bwd), flop, byt, 1)
            finally:
                lib.knerf_debug_backward_parts(3)
    else:
        add("fp32 forward (13 launches: encode + 12 sgemm_kernel)", ms_f, FLOP_FWD_PER_SAMPLE, None, n_f)
        l0 = lib.knerf_launch_count()
        bwd()
        n_b = int(lib.knerf_launch_count() - l0)
        add("fp32 backward (sgemm_kernel<T> + wgrad_kernel + colsum_kernel)", _time_ms(bwd),
            FLOP_TRAIN_PER_SAMPLE - FLOP_FWD_PER_SAMPLE, None, n_b)

    name = max(kernels, key=lambda k: kernels[k]["ms"])
    k = kernels[name]
    hbm_bound = precision == "bf16" and k.get("hbm_frac", 0) > k["tensor_frac"]
    roof = {"kernel": name, "samples_per_launch": rows, "peak_source": src, "kernels": kernels}
    if hbm_bound:
        roof.update({"bound": "hbm", "achieved": k["hbm_GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": k["hbm_frac"],
                     "traffic": k.get("ncu_dram_bytes_per_sample", 0) * rows or None})
    else:
        roof.update({"bound": "tensor", "achieved": k["tflops"], "peak": bf16_peak, "unit": "TFLOP/s",
                     "frac": k["tensor_frac"],
                     "traffic": k.get("ncu_dram_bytes_per_sample", 0) * rows or None})
    if precision != "bf16":
        roof["note"] = "fp32 SIMT FFMA parity mode measured against the bf16 tensor peak"
    return roof


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
    tf = FLOP_FWD_PER_SAMPLE * samples / (ms * 1e-3) / 1e12
    out = {"metric": "render_ms_per_frame_800x800", "value": ms, "unit": "ms", "n_gpus": 1, "tflops": tf,
           "ray_chunks": 32000, "precision_mode": precision}
    if precision == "bf16":
        out["tflops_executed"] = FLOP_FWD_INFER_EXECUTED * samples / (ms * 1e-3) / 1e12
        out["note"] = ("tflops counts the unfolded 593,408 MAC/sample; the inference kernel folds features into "
                       "rgb_features and executes 532,480 MAC/sample (tflops_executed)")
    return out
