"""CPU restatement of keras_nerf's per-ray hot path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Every function cites the reference file:line it follows (paths relative to
/root/reference).  All arithmetic is float32 on the CPU (torch-CPU, scans done
sequentially in fp32 like TF's CPU kernels).  TF op semantics that could not
be executed offline are marked [TF-sem].

Tensors are torch CPU float32; helpers accept numpy arrays too.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

F32 = torch.float32

OOB_ZERO = "zero"    # TF-GPU tf.gather: out-of-range index -> 0   (parity default)
OOB_CLAMP = "clamp"  # XLA-style clamp
OOB_RAISE = "raise"  # TF-CPU: InvalidArgumentError

__all__ = [
    "OOB_ZERO", "OOB_CLAMP", "OOB_RAISE", "NerfConfig", "as_f32",
    "get_focal_from_fov", "pose_spherical", "linspace_tf", "generate_rays",
    "positional_encoding", "encode_position_and_directions", "layer_shapes", "param_count",
    "init_params", "flatten_params", "unflatten_params", "mlp_forward",
    "cumprod_exclusive_seq", "cumsum_seq", "reduce_sum_seq", "render_image_depth_chunk", "fine_cdf",
    "fine_hierarchical_sampling_chunk", "predict_and_render_chunk_single", "predict_and_render_chunk", "render_given_points",
    "predict_and_render_images", "mse", "psnr", "AdamState", "adam_apply", "train_step",
    "composite_backward_analytic", "uniform24",
]


def as_f32(x) -> torch.Tensor:
    if torch.is_tensor(x):
        return x.detach().to(device="cpu", dtype=F32)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))


def uniform24(rng: np.random.Generator, shape) -> np.ndarray:
    """Uniform draws in [0,1) that are exact float32 multiples of 2^-24."""
    return (rng.integers(0, 1 << 24, size=shape, dtype=np.int64).astype(np.float32)
            * np.float32(2.0 ** -24))


@dataclass
class NerfConfig:
    """Mirror of the 7 ints in model_config.json (keras_nerf/model/nerf/nerf.py:47-55)."""
    n_coarse: int = 64
    n_fine: int = 128
    pos_emb_xyz: int = 10
    pos_emb_dir: int = 4
    n_layers: int = 8
    dense_units: int = 256
    skip_layer: int = 4

    @property
    def dx(self) -> int:
        return 3 + 6 * self.pos_emb_xyz

    @property
    def dd(self) -> int:
        return 3 + 6 * self.pos_emb_dir


# ----------------------------------------------------------------------------------------------
# a1 / a2  host-side camera helpers
# ----------------------------------------------------------------------------------------------
def get_focal_from_fov(field_of_view: float, width: int) -> float:
    """keras_nerf/data/utils.py:5-16 -- 0.5*width/tan(0.5*fov), fp32 (width cast to fp32)."""
    w = torch.tensor(float(width), dtype=F32)
    half = torch.tensor(0.5 * float(field_of_view), dtype=F32)   # python-float product, then fp32
    return float((torch.tensor(0.5, dtype=F32) * w) / torch.tan(half))


def pose_spherical(theta: float, phi: float, t: float) -> torch.Tensor:
    """keras_nerf/data/utils.py:19-63 -- c2w = F @ R_theta @ R_phi @ T(t), fp32 4x4.

    Angles are degrees; the deg->rad product is done in python float64 and then fed to
    fp32 cos/sin (tf.cos of a python float converts to float32 first) [TF-sem]."""
    def f32(v):
        return torch.tensor(v, dtype=F32)

    def trans(tt):
        m = torch.eye(4, dtype=F32)
        m[2, 3] = f32(tt)
        return m

    def rot_phi(p):
        c, s = torch.cos(f32(p)), torch.sin(f32(p))
        m = torch.eye(4, dtype=F32)
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
        return m

    def rot_theta(th):
        c, s = torch.cos(f32(th)), torch.sin(f32(th))
        m = torch.eye(4, dtype=F32)
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
        return m

    c2w = trans(t)
    c2w = rot_phi(phi / 180.0 * np.pi) @ c2w
    c2w = rot_theta(theta / 180.0 * np.pi) @ c2w
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=F32)
    return flip @ c2w


# ----------------------------------------------------------------------------------------------
# a3  ray generation + stratified coarse samples
# ----------------------------------------------------------------------------------------------
def linspace_tf(start: float, stop: float, num: int) -> torch.Tensor:
    """tf.linspace [TF-sem]: lin_0=start, lin_{n-1}=stop exactly, lin_i = start + ((stop-start)/(n-1))*i."""
    s, e = torch.tensor(start, dtype=F32), torch.tensor(stop, dtype=F32)
    if num == 1:
        return s.reshape(1)
    delta = (e - s) / torch.tensor(float(num - 1), dtype=F32)
    idx = torch.arange(num, dtype=F32)
    lin = s + delta * idx
    lin[0] = s
    lin[-1] = e
    return lin


def generate_rays(c2w, H: int, W: int, focal: float, near: float, far: float, n_sample: int,
                  u) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """keras_nerf/data/rays.py:69-130.  `u` are the uniform draws, shape [H,W,N].

    The reference draws noise as [W,H,N] (rays.py:122-123) and adds it to a [N] vector, so the
    result is only a valid [H,W,N] tensor for square images; we take u in ray-major [H,W,N]
    (identical flat layout when H == W)."""
    c2w = as_f32(c2w)
    u = as_f32(u).reshape(H, W, n_sample)
    f = torch.tensor(focal, dtype=F32)
    Wf, Hf = torch.tensor(float(W), dtype=F32), torch.tensor(float(H), dtype=F32)
    x = torch.arange(W, dtype=F32).reshape(1, W).expand(H, W)      # meshgrid 'xy': x = column
    y = torch.arange(H, dtype=F32).reshape(H, 1).expand(H, W)      #                y = row
    xc = (x - Wf * 0.5) / f                                        # :89
    yc = (y - Hf * 0.5) / f                                        # :90
    cam = torch.stack([xc, -yc, -torch.ones_like(xc)], dim=-1)     # :93-94  [H,W,3]
    rot = c2w[:3, :3]
    trans = c2w[:3, 3]
    world = cam[..., None, :] * rot                                # :103-104 [H,W,3,3]
    d = world[..., 0] + world[..., 1] + world[..., 2]              # :107 reduce_sum(axis=-1), j=0,1,2
    norm = torch.sqrt((d * d).sum(dim=-1, keepdim=True))           # tf.norm
    d = d / norm                                                   # :108-109
    o = trans.expand(H, W, 3).contiguous()                         # :112-113
    lin = linspace_tf(near, far, n_sample)                         # :116-117
    nearf, farf = torch.tensor(near, dtype=F32), torch.tensor(far, dtype=F32)
    interval = (farf - nearf) / torch.tensor(float(n_sample), dtype=F32)   # :120  (N, not N-1)
    noise = u * interval - (interval / 2)                          # :122-123
    t = torch.clamp(lin + noise, min=near, max=far)                # :126-127
    return o, d.contiguous(), t.contiguous()


# ----------------------------------------------------------------------------------------------
# a4 / a5  positional encoding
# ----------------------------------------------------------------------------------------------
def positional_encoding(x: torch.Tensor, L: int) -> torch.Tensor:
    """keras_nerf/model/nerf/utils.py:176-186 -- [x, sin(2^0 x), cos(2^0 x), ...], no pi."""
    parts = [x]
    for i in range(L):
        s = torch.tensor(2.0 ** i, dtype=F32)
        parts.append(torch.sin(s * x))
        parts.append(torch.cos(s * x))
    return torch.cat(parts, dim=-1)


def encode_position_and_directions(o, d, t, L_xyz: int, L_dir: int):
    """keras_nerf/model/nerf/utils.py:188-210.  o,d [...,3]; t [...,S]."""
    o, d, t = as_f32(o), as_f32(d), as_f32(t)
    p = o[..., None, :] + d[..., None, :] * t[..., None]           # :193-194
    xyz = positional_encoding(p, L_xyz)
    dirs = d[..., None, :].expand(p.shape)                         # :203-205
    dir_enc = positional_encoding(dirs, L_dir)
    return xyz, dir_enc


# ----------------------------------------------------------------------------------------------
# a6  MLP
# ----------------------------------------------------------------------------------------------
def layer_shapes(cfg: NerfConfig, dx: Optional[int] = None, dd: Optional[int] = None) -> List[Tuple[str, int, int]]:
    """(name, fan_in, fan_out) in Keras variable order (mlp.py:11-27; SURVEY App. A 'Variable order')."""
    dx = cfg.dx if dx is None else dx
    dd = cfg.dd if dd is None else dd
    U = cfg.dense_units
    out = []
    width = dx
    for i in range(cfg.n_layers):
        out.append((f"layer_{i}", width, U))
        width = U
        if i % cfg.skip_layer == 0 and i > 0:                      # mlp.py:36-38
            width = U + dx
    out.append(("sigma", width, 1))
    out.append(("features", width, U))
    out.append(("rgb_features", U + dd, U // 2))
    out.append(("rgb", U // 2, 3))
    return out


def param_count(cfg: NerfConfig, dx=None, dd=None) -> int:
    return sum(i * o + o for _, i, o in layer_shapes(cfg, dx, dd))


def init_params(cfg: NerfConfig, rng: np.random.Generator, dx=None, dd=None) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Keras Dense defaults [TF-sem]: glorot_uniform kernel U(+-sqrt(6/(fan_in+fan_out))), zero bias."""
    params = []
    for _, fi, fo in layer_shapes(cfg, dx, dd):
        lim = math.sqrt(6.0 / (fi + fo))
        W = rng.uniform(-lim, lim, size=(fi, fo)).astype(np.float32)
        params.append((torch.from_numpy(W), torch.zeros(fo, dtype=F32)))
    return params


def flatten_params(params) -> torch.Tensor:
    return torch.cat([torch.cat([W.reshape(-1), b.reshape(-1)]) for W, b in params])


def unflatten_params(flat, cfg: NerfConfig, dx=None, dd=None):
    flat = as_f32(flat)
    out, off = [], 0
    for _, fi, fo in layer_shapes(cfg, dx, dd):
        W = flat[off:off + fi * fo].reshape(fi, fo).clone(); off += fi * fo
        b = flat[off:off + fo].clone(); off += fo
        out.append((W, b))
    assert off == flat.numel()
    return out


def mlp_forward(params, xyz_enc: torch.Tensor, dir_enc: torch.Tensor, cfg: NerfConfig,
                return_pre: bool = False):
    """keras_nerf/model/nerf/mlp.py:29-50.  Keras Dense: y = act(x @ W[in,out] + b)."""
    n = cfg.n_layers
    h = xyz_enc
    for i in range(n):
        W, b = params[i]
        h = torch.relu(h @ W + b)
        if i % cfg.skip_layer == 0 and i > 0:
            h = torch.cat([h, xyz_enc], dim=-1)                    # [h, x] order (mlp.py:37-38)
    Ws, bs = params[n]
    sigma_pre = h @ Ws + bs
    sigma = torch.relu(sigma_pre)                                  # mlp.py:17-18,40
    Wf, bf = params[n + 1]
    feat = h @ Wf + bf                                             # linear (mlp.py:20-21,42)
    Wg, bg = params[n + 2]
    g = torch.cat([feat, dir_enc], dim=-1) @ Wg + bg               # linear, no activation (mlp.py:23-24,43-46)
    Wc, bc = params[n + 3]
    rgb_pre = g @ Wc + bc
    rgb = torch.sigmoid(rgb_pre)                                   # mlp.py:26-27,48
    if return_pre:
        return rgb, sigma, rgb_pre, sigma_pre
    return rgb, sigma


# ----------------------------------------------------------------------------------------------
# a7  compositing
# ----------------------------------------------------------------------------------------------
def cumprod_exclusive_seq(x: torch.Tensor) -> torch.Tensor:
    """tf.math.cumprod(x, axis=-1, exclusive=True): sequential fp32 products, T_0 = 1 (differentiable)."""
    S = x.shape[-1]
    outs = [torch.ones_like(x[..., 0])]
    acc = outs[0]
    for i in range(S - 1):
        acc = acc * x[..., i]
        outs.append(acc)
    return torch.stack(outs, dim=-1)


def reduce_sum_seq(x: torch.Tensor, dim: int = -1, keepdim: bool = False) -> torch.Tensor:
    """tf.reduce_sum along `dim`, summed strictly left to right in fp32 [TF-sem: TF leaves the order
    unspecified; this is the documented choice shared with oracle/tfshim and KNERF_SCAN_SEQUENTIAL]."""
    xm = x.movedim(dim, -1)
    acc = xm[..., 0]
    for i in range(1, xm.shape[-1]):
        acc = acc + xm[..., i]
    return acc.unsqueeze(dim) if keepdim else acc


def cumsum_seq(x: torch.Tensor) -> torch.Tensor:
    """tf.cumsum(x, axis=-1) as a sequential fp32 running sum (TF CPU / numpy order)."""
    return torch.from_numpy(np.cumsum(x.detach().numpy(), axis=-1, dtype=np.float32))


def render_image_depth_chunk(rgb, sigma, t, white_background: bool, clip: bool = True,
                             epsilon: float = 1e-10):
    """keras_nerf/model/nerf/utils.py:16-58 (clip=False, white=False gives :99-134).

    rgb [...,S,3], sigma [...,S,1] or [...,S], t [...,S]."""
    if sigma.dim() == rgb.dim():
        sigma = sigma[..., 0]                                      # :32
    eps = torch.tensor(epsilon, dtype=F32)
    delta = t[..., 1:] - t[..., :-1]                               # :35
    delta = torch.cat([delta, eps.expand(delta.shape[:-1] + (1,))], dim=-1)   # :36-37
    alpha = 1.0 - torch.exp(-sigma * delta)                        # :41
    exp_alpha = 1.0 - alpha                                        # :43
    trans = cumprod_exclusive_seq(exp_alpha + eps)                 # :46-47
    weights = alpha * trans                                        # :48
    image = reduce_sum_seq(weights[..., None] * rgb, dim=-2)       # :50
    depth = reduce_sum_seq(weights * t, dim=-1)                    # :51
    if white_background:
        image = image + (1.0 - reduce_sum_seq(weights, dim=-1)[..., None])   # :53-54
    if clip:
        image = torch.clamp(image, 0.0, 1.0)                       # :56
    return image, depth, weights


def composite_backward_analytic(rgb, sigma, t, dL_dimage, white_background: bool, clip: bool = True,
                                epsilon: float = 1e-10):
    """Closed-form gradient of render_image_depth_chunk w.r.t. rgb and sigma.

    Used to cross-check the formula the fused CUDA backward implements against torch autograd."""
    sig = sigma[..., 0] if sigma.dim() == rgb.dim() else sigma
    eps = torch.tensor(epsilon, dtype=F32)
    delta = torch.cat([t[..., 1:] - t[..., :-1], eps.expand(t.shape[:-1] + (1,))], dim=-1)
    ex = torch.exp(-sig * delta)
    alpha = 1.0 - ex
    e = (1.0 - alpha) + eps
    T = cumprod_exclusive_seq(e)
    w = alpha * T
    pre = (w[..., None] * rgb).sum(dim=-2)
    bgc = 1.0 if white_background else 0.0
    if white_background:
        pre = pre + (1.0 - w.sum(dim=-1)[..., None])
    G = dL_dimage.clone()
    if clip:
        G = G * ((pre >= 0.0) & (pre <= 1.0)).to(F32)
    d_rgb = G[..., None, :] * w[..., None]
    g = (G[..., None, :] * (rgb - bgc)).sum(dim=-1)                # [.., S]
    gw = g * w
    suffix = torch.flip(torch.cumsum(torch.flip(gw, dims=[-1]), dim=-1), dims=[-1]) - gw   # sum_{k>i}
    d_alpha = g * T - suffix / e
    d_sigma = d_alpha * delta * ex
    return d_rgb, d_sigma


# ----------------------------------------------------------------------------------------------
# a8  hierarchical (fine) sampling
# ----------------------------------------------------------------------------------------------
def fine_cdf(weights: torch.Tensor) -> torch.Tensor:
    """keras_nerf/model/nerf/utils.py:63-69 -- w+=1e-5; pdf=w/sum; cdf=[0,cumsum(pdf)]."""
    w = weights + torch.tensor(1e-5, dtype=F32)
    pdf = w / reduce_sum_seq(w, dim=-1, keepdim=True)
    cdf = cumsum_seq(pdf)
    return torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)


def fine_hierarchical_sampling_chunk(mid_points, weights, u, oob_mode: str = OOB_ZERO,
                                     cdf: Optional[torch.Tensor] = None):
    """keras_nerf/model/nerf/utils.py:60-97 with the uniform draws `u` [R,Nf] made explicit.

    Returns (samples [R,Nf], indices int32 [R,Nf] (searchsorted side='right'), cdf [R,Nc+1]).
    `mid_points` has Nc-1 entries while the indices run to Nc: gathers at Nc-1 and Nc are out of
    range; oob_mode selects TF-GPU zero fill (default), clamp, or TF-CPU raise."""
    mid_points, weights, u = as_f32(mid_points), as_f32(weights), as_f32(u)
    if cdf is None:
        cdf = fine_cdf(weights)
    cdf = cdf.contiguous()
    idx = torch.searchsorted(cdf, u.contiguous(), right=True)      # :76
    ncdf = cdf.shape[-1]
    below = torch.clamp(idx - 1, min=0)                            # :78
    above = torch.clamp(idx, max=ncdf - 1)                         # :79
    c0 = torch.gather(cdf, -1, below)                              # :83-84
    c1 = torch.gather(cdf, -1, above)
    nmid = mid_points.shape[-1]
    if oob_mode == OOB_ZERO:
        pad = torch.zeros(mid_points.shape[:-1] + (ncdf - nmid,), dtype=F32)
        midp = torch.cat([mid_points, pad], dim=-1)
        m0 = torch.gather(midp, -1, below)
        m1 = torch.gather(midp, -1, above)
    elif oob_mode == OOB_CLAMP:
        m0 = torch.gather(mid_points, -1, torch.clamp(below, max=nmid - 1))
        m1 = torch.gather(mid_points, -1, torch.clamp(above, max=nmid - 1))
    elif oob_mode == OOB_RAISE:
        if int(above.max()) >= nmid:
            raise IndexError("tf.gather on CPU: index out of range for mid_points "
                             f"(max index {int(above.max())}, size {nmid})")
        m0 = torch.gather(mid_points, -1, below)
        m1 = torch.gather(mid_points, -1, above)
    else:
        raise ValueError(oob_mode)
    denom = c1 - c0                                                # :90
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)   # :91
    tt = (u - c0) / denom                                          # :92
    samples = m0 + tt * (m1 - m0)                                  # :93-94
    return samples, idx.to(torch.int32), cdf


# ----------------------------------------------------------------------------------------------
# a9 / a10  chunk + image assembly
# ----------------------------------------------------------------------------------------------
def predict_and_render_chunk_single(params, cfg: NerfConfig, o, d, t_c, white: bool,
                                    coarse_weights=None, u_fine=None, oob_mode: str = OOB_ZERO,
                                    cdf=None):
    """keras_nerf/model/nerf/nerf.py:175-216 (`_predict_and_render_chunk`)."""
    extra = {}
    if coarse_weights is not None:
        mid = 0.5 * (t_c[..., 1:] + t_c[..., :-1])                 # :182-183
        t_f, idx, cdf_used = fine_hierarchical_sampling_chunk(mid, coarse_weights.detach(), u_fine,
                                                              oob_mode, cdf)
        points, _ = torch.sort(torch.cat([t_c, t_f], dim=-1), dim=-1)   # :190-191
        extra = {"t_fine": t_f, "indices": idx, "cdf": cdf_used}
    else:
        points = t_c
    xyz, dirs = encode_position_and_directions(o, d, points, cfg.pos_emb_xyz, cfg.pos_emb_dir)
    rgb, sigma = mlp_forward(params, xyz, dirs, cfg)
    image, depth, weights = render_image_depth_chunk(rgb, sigma, points, white)
    out = {"image": image, "depth": depth, "weights": weights, "points": points,
           "rgb": rgb, "sigma": sigma}
    out.update(extra)
    return out


def render_given_points(params, cfg: NerfConfig, o, d, points, white: bool):
    """The tail of `_predict_and_render_chunk` (nerf.py:199-216) for depths that are already chosen.

    Used to check the fine network + compositing (and their backward) GIVEN the reference's sorted
    depths: end to end the fine pass is ill-conditioned (App. C-1: a 1e-7 difference in the coarse weights
    moves the ~3% of samples that fall in [0, near) by ~1e-3), so parity is pinned stage by stage."""
    xyz, dirs = encode_position_and_directions(o, d, points, cfg.pos_emb_xyz, cfg.pos_emb_dir)
    rgb, sigma = mlp_forward(params, xyz, dirs, cfg)
    image, depth, weights = render_image_depth_chunk(rgb, sigma, points, white)
    return {"image": image, "depth": depth, "weights": weights, "rgb": rgb, "sigma": sigma}


def predict_and_render_chunk(params_c, params_f, cfg, o, d, t_c, u_fine, white, oob_mode=OOB_ZERO):
    """keras_nerf/model/nerf/nerf.py:218-227."""
    coarse = predict_and_render_chunk_single(params_c, cfg, o, d, t_c, white)
    fine = predict_and_render_chunk_single(params_f, cfg, o, d, t_c, white,
                                           coarse["weights"], u_fine, oob_mode)
    return coarse, fine


def predict_and_render_images(params_c, params_f, cfg, rays, u_fine, ray_chunks: int, white: bool,
                              oob_mode: str = OOB_ZERO):
    """keras_nerf/model/nerf/nerf.py:229-304.  rays = (o[B,H,W,3], d[B,H,W,3], t[B,H,W,Nc]);
    u_fine [num_rays, Nf] in flat ray order ((b*H+y)*W+x)."""
    o, d, t = (as_f32(r) for r in rays)
    B, H, W = o.shape[:3]
    num_rays = B * H * W
    ray_chunks = min(ray_chunks, num_rays)
    assert num_rays % ray_chunks == 0                              # :100
    of, df, tf_ = o.reshape(num_rays, 3), d.reshape(num_rays, 3), t.reshape(num_rays, -1)
    u_fine = as_f32(u_fine).reshape(num_rays, -1)
    keys = ("image", "depth", "weights")
    acc_c = {k: [] for k in keys}
    acc_f = {k: [] for k in keys}
    with torch.no_grad():
        for i in range(num_rays // ray_chunks):
            s = slice(i * ray_chunks, (i + 1) * ray_chunks)
            c, f = predict_and_render_chunk(params_c, params_f, cfg, of[s], df[s], tf_[s], u_fine[s],
                                            white, oob_mode)
            for k in keys:
                acc_c[k].append(c[k])
                acc_f[k].append(f[k])

    def asm(acc):
        return {"image": torch.cat(acc["image"]).reshape(B, H, W, 3),
                "depth": torch.cat(acc["depth"]).reshape(B, H, W),
                "weights": torch.cat(acc["weights"]).reshape(B, H, W, -1)}
    return asm(acc_c), asm(acc_f)


# ----------------------------------------------------------------------------------------------
# a12 / a18  loss + PSNR
# ----------------------------------------------------------------------------------------------
def mse(target, pred):
    """tf.keras.losses.MeanSquaredError(): mean over every element (train_single.py:127)."""
    return ((pred - target) ** 2).mean()


def psnr(a, b):
    """tf.image.psnr(a, b, max_val=1.0) per image [B,H,W,3] -> [B] (nerf.py:309,311)."""
    m = ((a - b) ** 2).reshape(a.shape[0], -1).mean(dim=1)
    return -10.0 * torch.log10(m)


# ----------------------------------------------------------------------------------------------
# a13  Adam (Keras defaults) and a11 train_step
# ----------------------------------------------------------------------------------------------
@dataclass
class AdamState:
    """Keras `optimizer='adam'` defaults [TF-sem]: lr 1e-3, b1 .9, b2 .999, eps 1e-7, no amsgrad."""
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-7
    step: int = 0
    m: Optional[torch.Tensor] = None
    v: Optional[torch.Tensor] = None


def adam_apply(flat_params: torch.Tensor, flat_grads: torch.Tensor, st: AdamState) -> torch.Tensor:
    """theta -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps)   (nerf.py:163-165,455-458)."""
    if st.m is None:
        st.m = torch.zeros_like(flat_params)
        st.v = torch.zeros_like(flat_params)
    st.step += 1
    t = st.step
    lr_t = np.float32(st.lr * math.sqrt(1.0 - st.beta2 ** t) / (1.0 - st.beta1 ** t))
    st.m = st.m + (flat_grads - st.m) * np.float32(1.0 - st.beta1)
    st.v = st.v + (flat_grads * flat_grads - st.v) * np.float32(1.0 - st.beta2)
    return flat_params - lr_t * st.m / (torch.sqrt(st.v) + np.float32(st.eps))


def _req(params):
    return [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in params]


def _grads_flat(loss, params):
    leaves = [p for Wb in params for p in Wb]
    gs = torch.autograd.grad(loss, leaves, allow_unused=True)
    return torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1)
                      for g, p in zip(gs, leaves)])


def train_step(params_c, params_f, adam_c: AdamState, adam_f: AdamState, cfg: NerfConfig,
               images, rays, u_fine, ray_chunks: int, white: bool, oob_mode: str = OOB_ZERO,
               apply: bool = True) -> Dict[str, object]:
    """keras_nerf/model/nerf/nerf.py:332-473.

    Per chunk: coarse fwd+loss+grad (coarse vars only), fine fwd+loss+grad (fine vars only,
    coarse weights are constants: :361-369,390-398); accumulators += g/n_chunks (:383-384,412-413);
    two independent Adam updates (:455-458).  Returns new params, accumulated flat grads, losses and
    the reconstructed images."""
    images = as_f32(images)[..., :3]                               # :335
    o, d, t = (as_f32(r) for r in rays)
    B, H, W = o.shape[:3]
    num_rays = B * H * W
    ray_chunks = min(ray_chunks, num_rays)
    assert num_rays % ray_chunks == 0
    n_chunks = num_rays // ray_chunks
    img = images.reshape(num_rays, 3)
    of, df, tf_ = o.reshape(num_rays, 3), d.reshape(num_rays, 3), t.reshape(num_rays, -1)
    u_fine = as_f32(u_fine).reshape(num_rays, -1)
    pc, pf = _req(params_c), _req(params_f)
    gc = torch.zeros(sum(W.numel() + b.numel() for W, b in pc), dtype=F32)
    gf = torch.zeros_like(gc)
    loss_c = torch.zeros((), dtype=F32)
    loss_f = torch.zeros((), dtype=F32)
    imgs_c, imgs_f = [], []
    nch = torch.tensor(float(n_chunks), dtype=F32)
    for i in range(n_chunks):
        s = slice(i * ray_chunks, (i + 1) * ray_chunks)
        c = predict_and_render_chunk_single(pc, cfg, of[s], df[s], tf_[s], white)
        lc = mse(img[s], c["image"])
        gc += _grads_flat(lc, pc) / nch
        loss_c += lc.detach() / nch
        f = predict_and_render_chunk_single(pf, cfg, of[s], df[s], tf_[s], white,
                                            c["weights"].detach(), u_fine[s], oob_mode)
        lf = mse(img[s], f["image"])
        gf += _grads_flat(lf, pf) / nch
        loss_f += lf.detach() / nch
        imgs_c.append(c["image"].detach())
        imgs_f.append(f["image"].detach())
    assert torch.isfinite(gc).all() and torch.isfinite(gf).all()   # :381-382,410-411
    flat_c = flatten_params(params_c)
    flat_f = flatten_params(params_f)
    out = {"grad_coarse": gc, "grad_fine": gf, "coarse_loss": float(loss_c), "fine_loss": float(loss_f),
           "coarse_image": torch.cat(imgs_c).reshape(B, H, W, 3),
           "fine_image": torch.cat(imgs_f).reshape(B, H, W, 3)}
    out["coarse_psnr"] = float(psnr(images, out["coarse_image"]).mean())
    out["fine_psnr"] = float(psnr(images, out["fine_image"]).mean())
    if apply:
        out["params_coarse"] = unflatten_params(adam_apply(flat_c, gc, adam_c), cfg,
                                                params_c[0][0].shape[0], params_c[cfg.n_layers + 2][0].shape[0] - cfg.dense_units)
        out["params_fine"] = unflatten_params(adam_apply(flat_f, gf, adam_f), cfg,
                                              params_f[0][0].shape[0], params_f[cfg.n_layers + 2][0].shape[0] - cfg.dense_units)
    return out
