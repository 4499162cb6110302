"""GPU (-m gpu): the KNERF_FP32_TC mode -- the fp32 MLP on tcgen05 tensor cores through 3-way bf16 operand
splitting (csrc/mlp_fp32_tc.cu) -- against the SAME bars as the SIMT fp32 parity mode: the reference's golden
fixtures at the north star's fp32 tolerances (1e-5 composited RGB / depth / weights), the oracle's gradients, and
the SIMT mode itself on shapes the oracle cannot reach quickly."""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def maxerr(a, b):
    a = a.detach().cpu() if torch.is_tensor(a) else T(np.asarray(a))
    b = b.detach().cpu() if torch.is_tensor(b) else T(np.asarray(b))
    return float((a.double() - b.double()).abs().max())


def _mlp_f64(params, layers, o, d, t, pre_activations=False):
    """the flagship network (8 x 256, skip 4, L = 10 / 4; mlp.py:29-50 + utils.py:176-210) in float64 on the device:
    (o[R,3], d[R,3], t[R,S]) -> [R,S,4] (rgb, sigma), or the head PRE-activations (what knerf_mlp_backward's d_pre
    is the gradient of)"""
    def pe(x, L):
        out = [x]
        for i in range(L):
            out += [torch.sin(2.0 ** i * x), torch.cos(2.0 ** i * x)]
        return torch.cat(out, -1)
    def dense(i, x):
        ko, bo, fi, fo = layers[i]
        return x @ params[ko:ko + fi * fo].view(fi, fo) + params[bo:bo + fo]
    x = pe(o[:, None, :] + d[:, None, :] * t[..., None], 10)
    dr = pe(d[:, None, :].expand(-1, t.shape[1], -1), 4)
    h = x
    for i in range(8):
        h = torch.relu(dense(i, h))
        if i % 4 == 0 and i > 0:
            h = torch.cat([h, x], -1)
    sigma = dense(8, h)
    g = dense(10, torch.cat([dense(9, h), dr], -1))
    if pre_activations:
        return torch.cat([dense(11, g), sigma], -1)
    return torch.cat([torch.sigmoid(dense(11, g)), torch.relu(sigma)], -1)


def _golden_model(precision, ray_chunks=None, training=True):
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    g = load_golden("model")
    mlp_mod.set_seed(int(g["init_seed"]))
    H, W, B = int(g["H"]), int(g["W"]), int(g["B"])
    m = K.NeRF(precision=precision)
    m.compile(optimizer="adam", loss="mse", batch_size=B, image_height=H, image_width=W,
              ray_chunks=int(g["ray_chunks"]) if ray_chunks is None else ray_chunks, white_background=True,
              is_training=training)
    return g, m, (g["o"][None], g["d"][None], g["t"][None])


def test_fp32tc_coarse_render_golden():
    """whole coarse pass against the fixture produced by the reference's own code: 1e-5 (north star, fp32)"""
    g, m, rays = _golden_model("fp32_tc", training=False)
    c, f = m.predict_and_render_images(rays, u_fine=g["u_fine"])
    assert maxerr(c["image"], g["image_coarse"]) <= 1e-5
    assert maxerr(c["depth"], g["depth_coarse"]) <= 1e-5
    assert maxerr(c["weights"], g["weights_coarse"]) <= 1e-5
    mse = float(((f["image"].cpu() - T(g["image_fine"])) ** 2).mean())      # fine end to end: ill-conditioned (C-1)
    assert maxerr(f["image"], g["image_fine"]) <= 1e-2 and -10.0 * np.log10(max(mse, 1e-30)) > 55.0


@pytest.mark.parametrize("R,S", [(300, 192), (37, 192), (5, 64), (129, 320)])
def test_fp32tc_forward_and_backward_vs_simt(R, S):
    """MLP outputs and weight gradients of the tensor-core fp32 mode against the SIMT fp32 mode (itself pinned to the
    golden fixtures at 2e-6): R*S not a multiple of 128 exercises the ragged last tile of both kernels"""
    import keras_nerf_b200 as K
    from keras_nerf_b200 import _lib
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    dev = torch.device("cuda")
    gen = torch.Generator().manual_seed(R * 1000 + S)
    o = torch.zeros(R, 3)
    o[:, 2] = 4.0
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen) * 0.2 + torch.tensor([0, 0, -1.0]), dim=-1)
    t = torch.sort(torch.rand(R, S, generator=gen) * 4 + 2, dim=-1).values.contiguous()
    dpre = (torch.randn(R, S, 4, generator=gen) * 1e-3).to(dev)
    o, d, t = o.to(dev), d.to(dev), t.to(dev)
    res = {}
    for prec in ("fp32", "fp32_tc"):
        mlp_mod.set_seed(42)
        m = K.NeRF(precision=prec, n_coarse=64, n_fine=S - 64 if S > 64 else 128)
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=1, image_width=R, ray_chunks=R,
                  white_background=True)
        # trained-looking weights: scale the kernels up so that sigma is not ~0 and the heads leave their linear range
        m.fine.params.mul_(1.7)
        out = torch.full((R, S, 4), float("nan"), device=dev)
        _lib.call("knerf_mlp_forward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(o), _lib.ptr(d), _lib.ptr(t),
                  R, S, m._prec, 1, _lib.ptr(out), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
        gbuf = torch.zeros_like(m.fine.params)
        _lib.call("knerf_mlp_backward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(dpre), R, S, m._prec,
                  _lib.ptr(gbuf), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
        torch.cuda.synchronize()
        res[prec] = (out.cpu(), gbuf.cpu())
    (a, ga), (b, gb) = res["fp32"], res["fp32_tc"]
    assert not torch.isnan(b).any()
    # with the kernels scaled up the pre-activations reach ~1e2 and ANY fp32 evaluation order differs from another by
    # ~1e-5: measure both modes against a float64 evaluation of the same network -- the tensor-core mode must be as
    # accurate as the SIMT one (which is pinned to the golden fixtures at 2e-6 on the unscaled network)
    ref = _mlp_f64(m.fine.params.double(), m.fine.layers, o.double(), d.double(), t.double()).cpu()
    err_simt = float((a.double() - ref).abs().max() / ref.abs().max())
    err_tc = float((b.double() - ref).abs().max() / ref.abs().max())
    assert err_tc <= max(2.0 * err_simt, 2e-6), (err_tc, err_simt)
    assert float((a[..., :3] - b[..., :3]).abs().max()) <= 1e-4                            # rgb (sigmoid output)
    # weight gradients against float64 autograd of the same network: the hidden layers' entries carry ~1/sqrt(samples)
    # of noise in ANY fp32 evaluation (an activation within rounding of 0 flips its ReLU' bit: one sample's whole
    # contribution), so the bar is "as accurate as the SIMT mode", layer by layer
    p64 = m.fine.params.double().clone().requires_grad_(True)
    pre = _mlp_f64(p64, m.fine.layers, o.double(), d.double(), t.double(), pre_activations=True)
    (pre * dpre.double()).sum().backward()
    g64 = p64.grad.cpu()
    off = 0
    for name, fi, fo in O.layer_shapes(O.NerfConfig()):
        for n in (fi * fo, fo):
            r = g64[off:off + n]
            sc = max(float(r.abs().max()), 1e-30)
            e_simt = float((ga[off:off + n].double() - r).abs().max()) / sc
            e_tc = float((gb[off:off + n].double() - r).abs().max()) / sc
            assert e_tc <= max(3.0 * e_simt, 5e-5), (name, e_tc, e_simt)
            off += n


def test_fp32tc_train_step_vs_oracle():
    """train_step of the golden fixture in fp32_tc mode against the oracle: the coarse network at the strict
    tolerances of the SIMT mode's test (losses 2e-5, Adam update 1e-5 where the gradient is not ~0)"""
    g, m, rays = _golden_model("fp32_tc")
    cfg = O.NerfConfig()
    rng = np.random.default_rng(int(g["init_seed"]))
    pc, pf = O.init_params(cfg, rng), O.init_params(cfg, rng)
    logs = m.train_step((g["images"], rays), u_fine=g["u_fine"])
    orays = tuple(T(np.asarray(r)) for r in rays)
    ref = O.train_step(pc, pf, O.AdamState(), O.AdamState(), cfg, g["images"], orays, g["u_fine"], int(g["ray_chunks"]), True)
    assert logs["coarse_loss"] == pytest.approx(ref["coarse_loss"], rel=2e-5)
    assert logs["fine_loss"] == pytest.approx(ref["fine_loss"], rel=5e-3)
    gmask = ref["grad_coarse"].abs() > 1e-3 * ref["grad_coarse"].abs().max()
    new = O.flatten_params(ref["params_coarse"])
    assert maxerr(m.coarse.params.cpu()[gmask], new[gmask]) <= 1e-5


def test_fp32tc_other_widths_fall_back_per_layer():
    """dense_units 128 runs on the tensor cores too (multiples of 64); 100 stays on SIMT -- both equal the oracle"""
    import keras_nerf_b200 as K
    from keras_nerf_b200 import _lib
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    dev = torch.device("cuda")
    for units, layers, skip in ((128, 5, 2), (100, 3, 4)):
        mlp_mod.set_seed(3)
        m = K.NeRF(precision="fp32_tc", n_layers=layers, dense_units=units, skip_layer=skip, pos_emb_xyz=6, pos_emb_dir=2)
        R, S = 70, 64
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=1, image_width=R, ray_chunks=R,
                  white_background=True, is_training=False)
        gen = torch.Generator().manual_seed(units)
        o = torch.zeros(R, 3)
        o[:, 2] = 4.0
        d = torch.nn.functional.normalize(torch.randn(R, 3, generator=gen), dim=-1)
        t = torch.sort(torch.rand(R, S, generator=gen) * 4 + 2, dim=-1).values.contiguous()
        out = torch.empty(R, S, 4, device=dev)
        od, dd, td = o.to(dev), d.to(dev), t.to(dev)
        _lib.call("knerf_mlp_forward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(od), _lib.ptr(dd),
                  _lib.ptr(td), R, S, m._prec, 0, _lib.ptr(out), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
        cfg = O.NerfConfig(n_layers=layers, dense_units=units, skip_layer=skip, pos_emb_xyz=6, pos_emb_dir=2)
        params = O.unflatten_params(m.fine.params.cpu(), cfg)
        xyz, dirs = O.encode_position_and_directions(o, d, t, 6, 2)
        rgb, sigma = O.mlp_forward(params, xyz, dirs, cfg)
        assert maxerr(out[..., :3], rgb) <= 2e-6 and maxerr(out[..., 3:], sigma) <= 5e-6
