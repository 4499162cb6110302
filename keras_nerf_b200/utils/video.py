"""Output side of inference.py:112-122: `imageio.mimwrite(path.gif, frames, fps=20)` without imageio (Pillow)."""
from __future__ import annotations

import numpy as np
import torch


def frames_to_uint8(frames) -> list:
    """imageio's float -> uint8 rule for GIF: values in [0,1] are scaled by 255 and rounded."""
    out = []
    for f in frames:
        a = f.detach().float().cpu().numpy() if torch.is_tensor(f) else np.asarray(f)
        if a.dtype != np.uint8:
            a = (np.clip(a.astype(np.float64), 0.0, 1.0) * 255.0 + 0.4999999999).astype(np.uint8)
        out.append(np.ascontiguousarray(a))
    return out


def mimwrite(path: str, frames, fps: int = 20, loop: int = 0) -> None:
    """Animated GIF of [H,W,3] float/uint8 frames at `fps` (frame duration rounded to GIF's 10 ms ticks)."""
    from PIL import Image
    imgs = [Image.fromarray(a, mode="RGB") for a in frames_to_uint8(frames)]
    if not imgs:
        raise ValueError("mimwrite: no frames")
    duration_ms = max(10, int(round(100.0 / fps)) * 10)
    imgs[0].save(path, format="GIF", save_all=True, append_images=imgs[1:], duration=duration_ms, loop=loop,
                 optimize=False, disposal=1)
