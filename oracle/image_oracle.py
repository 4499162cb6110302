"""CPU oracle (TEST INFRASTRUCTURE ONLY) for ImageLoader.__call__ -- keras_nerf/data/image.py:17-35 -- and for the
per-image PSNR / SSIM of NeRF.update_and_return_metrics -- keras_nerf/model/nerf/nerf.py:306-330 (bottom of file).

The arithmetic lives in TensorFlow (>= 2.9, requirements.txt:2, absent offline):
  * tf.image.convert_image_dtype(uint8 -> float32): cast, then multiply by float32(1/255);
  * tf.image.resize(method='bilinear', antialias=True): python `scale = cast(size, f32) / cast(shape, f32)`, then
    the ScaleAndTranslate op (tensorflow/core/kernels/image/scale_and_translate_op.cc) with the triangle
    kernel (radius 1): ComputeSpansCore builds, per output coordinate, a span of input coordinates and
    normalised weights; GatherSpans resizes rows first ([out_h, in_w] intermediate), columns second, each output
    the fp32 sum `sum_k w_k * in[start + k]` accumulated in span order.
This file restates that published algorithm in NumPy float32.  PARITY UNPINNED against TF itself (TF is not
installed and the reference's image tests need the nerf_synthetic dataset, tests/data/test_image.py:12-20);
the structure (span ends, weights, normalisation) is cross-checked against Pillow's independent
antialiased-bilinear resampler in tests/test_io_cpu.py.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def compute_spans(out_size: int, in_size: int, antialias: bool = True):
    """ComputeSpansCore (scale_and_translate_op.cc) for translate = 0, triangle kernel.
    Returns starts[out] (int), weights[out, span] (float32, zero padded)."""
    scale = F(out_size) / F(in_size)                       # image_ops_impl.py: cast(size) / cast(shape)
    inv_scale = F(1.0 / float(scale))                      # `const float inv_scale = 1.0 / scale;`
    kernel_scale = max(inv_scale, F(1.0)) if antialias else F(1.0)
    radius = F(1.0)
    span_size = min(2 * int(np.ceil(radius * kernel_scale)) + 1, in_size)
    inv_kernel_scale = F(1.0) / kernel_scale
    starts = np.zeros(out_size, dtype=np.int64)
    weights = np.zeros((out_size, span_size), dtype=F)
    for x in range(out_size):
        sample = F(F(x) + F(0.5)) * inv_scale
        if sample < 0 or sample > F(in_size):
            continue
        s0 = int(np.ceil(F(F(sample - radius * kernel_scale) - F(0.5))))
        s1 = int(np.floor(F(F(sample + radius * kernel_scale) - F(0.5))))
        s0 = min(max(s0, 0), in_size - 1)
        s1 = min(max(s1, 0), in_size - 1) + 1
        assert s1 - s0 <= span_size
        w = np.zeros(s1 - s0, dtype=F)
        total = F(0.0)
        for k, src in enumerate(range(s0, s1)):
            pos = F(F(F(src) + F(0.5)) - sample)
            a = np.abs(F(pos * inv_kernel_scale))
            w[k] = F(1.0) - a if a < 1 else F(0.0)
            total = F(total + w[k])
        if abs(total) >= F(1000.0) * np.finfo(F).tiny:
            weights[x, :len(w)] = w * (F(1.0) / total)
        starts[x] = s0
    return starts, weights


def gather_spans(img: np.ndarray, starts, weights, axis: int) -> np.ndarray:
    """GatherRows / GatherColumns: out = sum_k w[k] * in[start + k] along `axis`, fp32, span order."""
    img = np.moveaxis(img, axis, 0).astype(F)
    out = np.zeros((len(starts),) + img.shape[1:], dtype=F)
    n = img.shape[0]
    for k in range(weights.shape[1]):
        idx = np.minimum(starts + k, n - 1)                # padded taps carry weight 0
        w = weights[:, k].reshape((-1,) + (1,) * (img.ndim - 1))
        out = (out + (w * img[idx]).astype(F)).astype(F)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_antialias(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """tf.image.resize(img[H,W,C] float32, (out_h, out_w), antialias=True)."""
    sy, wy = compute_spans(out_h, img.shape[0])
    sx, wx = compute_spans(out_w, img.shape[1])
    rows = gather_spans(img, sy, wy, axis=0)               # [out_h, in_w, C]
    return gather_spans(rows, sx, wx, axis=1)              # [out_h, out_w, C]


def image_loader(rgba_u8: np.ndarray, image_width: int, image_height: int, white_background: bool) -> np.ndarray:
    """image.py:19-35 from the decoded [H,W,4] uint8 array on."""
    img = rgba_u8.astype(F) * F(1.0 / 255.0)               # convert_image_dtype
    img = resize_bilinear_antialias(img, image_width, image_height)   # (sic) image.py:22-23 passes (width, height)
    alpha = img[..., 3:4]
    bg = np.ones_like(img[..., :3]) if white_background else np.zeros_like(img[..., :3])
    rgb = ((alpha * img[..., :3]).astype(F) + ((F(1.0) - alpha).astype(F) * bg).astype(F)).astype(F)
    return np.clip(np.concatenate([rgb, alpha], axis=-1), 0.0, 1.0).astype(F)


# ---- tf.image.psnr / tf.image.ssim (keras_nerf/model/nerf/nerf.py:306-330) ----------------------------------------
def psnr(a: np.ndarray, b: np.ndarray, max_val: float = 1.0) -> np.ndarray:
    """tf.image.psnr: 20 log10(max_val) - 10 log10(mean over [H,W,C] of (a-b)^2), per image."""
    mse = ((a.astype(F) - b.astype(F)) ** 2).reshape(a.shape[0], -1).mean(axis=1, dtype=np.float64)
    return (20.0 * np.log10(max_val) - 10.0 * np.log10(mse)).astype(F)


def ssim(a: np.ndarray, b: np.ndarray, max_val: float = 1.0, filter_size: int = 11, filter_sigma: float = 1.5,
         k1: float = 0.01, k2: float = 0.03) -> np.ndarray:
    """tf.image.ssim (image_ops_impl.py: _fspecial_gauss, _ssim_per_channel, _ssim_helper) on [B,H,W,C]:
    the window is softmax(-(x^2 + y^2) / (2 sigma^2)) over the 11x11 grid, applied as a VALID depthwise
    convolution to x, y, x*y and x^2 + y^2; luminance * cs averaged over positions, then channels."""
    coords = np.arange(filter_size, dtype=F) - F(filter_size - 1) / F(2.0)
    g = np.square(coords) * F(-0.5 / np.square(F(filter_sigma)))
    g2 = (g.reshape(1, -1) + g.reshape(-1, 1)).astype(np.float64)
    win = np.exp(g2 - g2.max())
    win = (win / win.sum()).astype(F)                                      # tf.nn.softmax
    a, b = a.astype(F), b.astype(F)
    Ho, Wo = a.shape[1] - filter_size + 1, a.shape[2] - filter_size + 1

    def reducer(x):
        out = np.zeros((x.shape[0], Ho, Wo, x.shape[3]), dtype=F)
        for dy in range(filter_size):
            for dx in range(filter_size):
                out += win[dy, dx] * x[:, dy:dy + Ho, dx:dx + Wo, :]
        return out

    c1, c2 = F((k1 * max_val) ** 2), F((k2 * max_val) ** 2)
    mean0, mean1 = reducer(a), reducer(b)
    num0 = mean0 * mean1 * F(2.0)
    den0 = np.square(mean0) + np.square(mean1)
    luminance = (num0 + c1) / (den0 + c1)
    num1 = reducer(a * b) * F(2.0)
    den1 = reducer(np.square(a) + np.square(b))
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (luminance * cs).mean(axis=(1, 2), dtype=np.float64).mean(axis=-1).astype(F)
