"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes shim), against
 (1) the golden fixtures produced by the reference's own Python (tests/golden/make_golden.py),
 (2) the CPU oracle on the same seeded inputs,
 (3) size-independent properties at BASELINE.json's full sizes.
Tolerances are the north star's: bit-exact bins given the reference cdf; 1e-5 max-abs on fp32 depths and
composited RGB/depth/acc."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu

T = torch.from_numpy


def gpu():
    import keras_nerf_b200 as K
    return K


def maxerr(a, b):
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0


# ---- a1/a2 host helpers ---------------------------------------------------------------------------
def test_camera_helpers():
    K = gpu()
    g = load_golden("camera")
    assert K.get_focal_from_fov(0.6911112070083618, 100) == pytest.approx(138.88887889922103)   # reference KAT
    for th, pose in zip(g["thetas"], g["poses"]):
        assert maxerr(K.pose_spherical(float(th), float(g["phi"]), float(g["radius"])), pose) <= 1e-6


# ---- a3 rays ----------------------------------------------------------------------------------------
def test_rays_reference_fixture():
    K = gpu()
    g = load_golden("rays")
    H, W, N = int(g["H"]), int(g["W"]), int(g["N"])
    u = O.uniform24(np.random.default_rng(1234), (H, W, N))
    gen = K.RaysGenerator(focal_length=float(g["focal"]), image_width=W, image_height=H, near=2.0, far=6.0, n_sample=N)
    o, d, t = gen(g["pose"], u=u)
    rows = g["rows"]
    assert o.shape == (128, 128, 3) and d.shape == (128, 128, 3) and t.shape == (128, 128, 32)
    assert o.dtype == torch.float32 and t.dtype == torch.float32
    assert maxerr(o.cpu()[rows], g["o_rows"]) == 0.0
    assert maxerr(d, g["d"]) <= 1e-6
    assert maxerr(t.cpu()[rows], g["t_rows"]) <= 1e-6
    oo, od, ot = O.generate_rays(g["pose"], H, W, float(g["focal"]), 2.0, 6.0, N, u)
    assert maxerr(t, ot) <= 1e-6 and maxerr(d, od) <= 1e-6


def test_rays_reference_test_asserts():
    """tests/data/test_rays.py:50-87 with the library's own RNG."""
    K = gpu()
    g = load_golden("rays")
    gen = K.RaysGenerator(138.88887889922103, 128, 128, 2.0, 6.0, 32)
    last = None
    for _ in range(4):
        o, d, t = gen(g["pose"])
        assert not torch.isnan(o).any() and not torch.isnan(d).any() and not torch.isnan(t).any()
        assert o.shape == (128, 128, 3) and t.shape == (128, 128, 32)
        if last is not None:
            assert torch.allclose(last[0], o) and torch.allclose(last[1], d)
            assert torch.allclose(last[2], t, atol=4.0 / 32.0)
            assert not torch.equal(last[2], t)                # fresh jitter on every call
        assert float(t.min()) >= 2.0 - 4.0 / 32 and float(t.max()) <= 6.0 + 4.0 / 32
        assert bool((t[..., 1:] > t[..., :-1]).all())
        assert (o[..., None, :] + d[..., None, :] * t[..., None]).shape == (128, 128, 32, 3)
        last = (o.clone(), d.clone(), t.clone())
    # jitter is uniform: mean offset ~ 0, spread ~ interval/sqrt(12)
    lin = O.linspace_tf(2.0, 6.0, 32).cuda()
    off = (t - lin)[..., 1:-1]
    assert abs(float(off.mean())) < 2e-3 and abs(float(off.std()) - (4 / 32) / 12 ** 0.5) < 2e-3


def test_rays_odd_sizes_scalar_path():
    K = gpu()
    pose = K.pose_spherical(10.0, -30.0, 4.0)
    for (H, W, N) in ((5, 7, 3), (1, 1, 1), (16, 16, 33)):
        u = O.uniform24(np.random.default_rng(H * W + N), (H, W, N))
        gen = K.RaysGenerator(40.0, W, H, 2.0, 6.0, N)
        o, d, t = gen(pose, u=u)
        oo, od, ot = O.generate_rays(pose, H, W, 40.0, 2.0, 6.0, N, u)
        assert maxerr(o, oo) == 0 and maxerr(d, od) <= 1e-6 and maxerr(t, ot) <= 1e-6


# ---- a4/a5 positional encoding -------------------------------------------------------------------------
def test_positional_encoding_golden():
    K = gpu()
    g = load_golden("posenc")
    ut = K.NeRFUtils(1, 8, 8, 64, 10, 4, True)
    xyz, dirs = ut.encode_position_and_directions(g["o"], g["d"], g["t"])
    assert xyz.shape == g["xyz"].shape and dirs.shape == g["dirs"].shape
    assert maxerr(xyz, g["xyz"]) <= 2e-6      # |arg| up to ~5e3: 1 ulp of the argument reduction
    assert maxerr(dirs, g["dirs"]) <= 1e-6
    pe = ut.positional_encoding(g["o"], 10)
    assert pe.shape[-1] == 3 * 2 * 10 + 3
    assert maxerr(pe, g["pe_o"]) <= 2e-6


def test_encode_shapes_like_reference_tests():
    """tests/model/nerf/test_nerf_utils.py:54-110"""
    K = gpu()
    ut = K.NeRFUtils(2, 128, 128, 1024, 10, 4, True)
    rays = torch.rand(2, 32, 32, 32, 3)
    assert ut.positional_encoding(rays, 10).shape == (2, 32, 32, 32, 63)
    o, d, t = torch.rand(2, 128, 128, 3), torch.rand(2, 128, 128, 3), torch.rand(2, 128, 128, 32)
    a, b = ut.encode_position_and_directions(o, d, t)
    assert a.shape == (2, 128, 128, 32, 63) and b.shape == (2, 128, 128, 32, 27)
    a2, b2 = ut.encode_position_and_directions(o.reshape(-1, 3), d.reshape(-1, 3), t.reshape(-1, 32))
    assert a2.shape == (2 * 128 * 128, 32, 63) and b2.shape == (2 * 128 * 128, 32, 27)
    assert torch.equal(a.reshape(a2.shape), a2)


# ---- a7 compositing ----------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [32, 64, 192])
def test_composite_forward_golden(S):
    K = gpu()
    g = load_golden(f"composite_S{S}")
    R = g["t"].shape[0]
    for white, tag in ((True, "white"), (False, "black")):
        ut = K.NeRFUtils(1, 1, R, R, 10, 4, white)
        img, dep, w = ut.render_image_depth_chunk(g["rgb"], g["sigma"], g["t"])
        assert img.shape == (R, 3) and dep.shape == (R,) and w.shape == (R, S)
        assert maxerr(img, g[f"image_{tag}"]) <= 1e-5
        assert maxerr(dep, g[f"depth_{tag}"]) <= 1e-5
        assert maxerr(w, g[f"weights_{tag}"]) <= 1e-6
    hw = g["image_full"].shape[1]
    ut = K.NeRFUtils(1, hw, hw, R, 10, 4)
    img, dep, w = ut.render_image_depth(g["rgb"].reshape(1, hw, hw, S, 3), g["sigma"].reshape(1, hw, hw, S, 1),
                                        g["t"].reshape(1, hw, hw, S))
    assert img.shape == (1, hw, hw, 3) and dep.shape == (1, hw, hw) and w.shape == (1, hw, hw, S)
    assert maxerr(img, g["image_full"]) <= 1e-5 and maxerr(w, g["weights_full"]) <= 1e-6


@pytest.mark.parametrize("S", [1, 7, 33, 320, 512])
def test_composite_forward_ragged_sizes(S):
    K = gpu()
    from keras_nerf_b200 import _lib
    rng = np.random.default_rng(S)
    R = 37
    rgb = rng.uniform(0, 1, (R, S, 3)).astype(np.float32)
    sigma = (rng.uniform(0, 30, (R, S, 1)) * (rng.uniform(size=(R, S, 1)) < 0.5)).astype(np.float32)
    t = np.sort(rng.uniform(2, 6, (R, S)).astype(np.float32), axis=-1)
    ut = K.NeRFUtils(1, 1, R, R, 10, 4, True)
    img, dep, w = ut.render_image_depth_chunk(rgb, sigma, t)
    oi, od, ow = O.render_image_depth_chunk(T(rgb), T(sigma), T(t), True)
    assert maxerr(img, oi) <= 1e-5 and maxerr(dep, od) <= 1e-5 and maxerr(w, ow) <= 1e-6
    # packed float4 input path == separate-input path
    dev = torch.device("cuda")
    packed = torch.cat([T(rgb), T(sigma)], dim=-1).to(dev).contiguous()
    tt = T(t).to(dev)
    img2 = torch.empty(R, 3, device=dev); dep2 = torch.empty(R, device=dev); w2 = torch.empty(R, S, device=dev)
    acc = torch.empty(R, device=dev)
    _lib.call("knerf_composite_forward", _lib.ptr(packed), None, None, _lib.ptr(tt), R, S, 1, 1, 1e-10,
              _lib.ptr(img2), _lib.ptr(dep2), _lib.ptr(w2), _lib.ptr(acc), _lib.stream())
    assert torch.equal(img, img2) and torch.equal(dep, dep2) and torch.equal(w, w2)
    assert maxerr(acc, w2.sum(-1)) <= 1e-5


@pytest.mark.parametrize("S,white", [(64, True), (192, True), (64, False), (40, True)])
def test_composite_backward_vs_autograd(S, white):
    gpu()
    from keras_nerf_b200 import _lib
    rng = np.random.default_rng(10 * S + white)
    R = 96
    rgb = torch.from_numpy(rng.uniform(0.02, 0.98, (R, S, 3)).astype(np.float32)).requires_grad_(True)
    sig_np = (rng.uniform(0, 1, (R, S)) ** 2 * 25).astype(np.float32)
    sig_np[rng.uniform(size=sig_np.shape) < 0.3] = 0.0
    sig_np[:8] *= 40.0                                   # nearly opaque rays: the 1e-10 epsilon matters
    sigma = torch.from_numpy(sig_np).requires_grad_(True)
    t = torch.from_numpy(np.sort(rng.uniform(2, 6, (R, S)).astype(np.float32), axis=-1))
    target = torch.from_numpy(rng.uniform(0, 1, (R, 3)).astype(np.float32))
    target[:4] = 2.0                                     # drives C above the clip -> masked gradient... only if C>1
    img, _, _ = O.render_image_depth_chunk(rgb, sigma, t, white)
    loss = ((img - target) ** 2).mean()
    g_rgb, g_sig = torch.autograd.grad(loss, [rgb, sigma])
    dev = torch.device("cuda")
    packed = torch.cat([rgb.detach(), sigma.detach()[..., None]], dim=-1).to(dev).contiguous()
    t_d, target_d = t.to(dev), target.to(dev)       # keep alive: the library only sees raw pointers
    d_out = torch.empty(R, S, 4, device=dev)
    sq = torch.empty(R, device=dev)
    scale = 2.0 / (3.0 * R)
    _lib.call("knerf_composite_backward", _lib.ptr(packed), _lib.ptr(t_d), R, S, int(white), 1, 1e-10, None,
              _lib.ptr(target_d), scale, 0, _lib.ptr(d_out), _lib.ptr(sq), _lib.stream())
    d = d_out.cpu()
    s_r, s_s = float(g_rgb.abs().max()), float(g_sig.abs().max())
    assert maxerr(d[..., :3] / s_r, g_rgb / s_r) <= 2e-5
    assert maxerr(d[..., 3] / s_s, g_sig / s_s) <= 5e-5
    assert float(sq.sum().cpu()) / (3 * R) == pytest.approx(float(loss), rel=1e-5)
    # through_activations folds sigmoid' and relu' in
    d2 = torch.empty_like(d_out)
    _lib.call("knerf_composite_backward", _lib.ptr(packed), _lib.ptr(t_d), R, S, int(white), 1, 1e-10, None,
              _lib.ptr(target_d), scale, 1, _lib.ptr(d2), None, _lib.stream())
    exp_rgb = d[..., :3] * (rgb.detach() * (1 - rgb.detach()))
    exp_sig = d[..., 3] * (sigma.detach() > 0)
    assert maxerr(d2.cpu()[..., :3], exp_rgb) <= 1e-9 + 1e-6 * s_r
    assert maxerr(d2.cpu()[..., 3], exp_sig) <= 1e-9 + 1e-6 * s_s
    # explicit dL/dimage input form
    dimg = (scale * (img.detach() - target)).to(dev).contiguous()
    d3 = torch.empty_like(d_out)
    _lib.call("knerf_composite_backward", _lib.ptr(packed), _lib.ptr(t_d), R, S, int(white), 1, 1e-10,
              _lib.ptr(dimg), None, 0.0, 0, _lib.ptr(d3), None, _lib.stream())
    assert maxerr(d3[..., :3].cpu() / s_r, g_rgb / s_r) <= 2e-5


# ---- a8 hierarchical sampling ----------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["flagship", "reftest"])
def test_sampler_golden(tag):
    K = gpu()
    g = load_golden(f"sampler_{tag}")
    R, Nf = g["u"].shape
    ut = K.NeRFUtils(1, 1, R, R, 10, 4, True)
    # (1) given the reference's cdf: bins bit-exact, samples to the last bit (same IEEE ops)
    s, idx, cdf = ut.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], Nf, u=g["u"], cdf=g["cdf"],
                                                      return_aux=True)
    assert s.shape == (R, Nf)
    np.testing.assert_array_equal(idx.cpu().numpy(), g["idx"])
    assert maxerr(cdf, g["cdf"]) == 0.0
    assert maxerr(s, g["samples"]) <= 1e-6
    # (2a) own cdf, sequential (TF-CPU order, the fp32 parity default): the whole chain is bit-exact
    s1, idx1, cdf1 = ut.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], Nf, u=g["u"], return_aux=True)
    assert maxerr(cdf1, g["cdf"]) == 0.0
    np.testing.assert_array_equal(idx1.cpu().numpy(), g["idx"])
    assert maxerr(s1, g["samples"]) <= 1e-6
    # (2b) own cdf, warp-shuffle scan: cdf within 1.2e-6, bins flip only where u sits within that of an edge
    ut = K.NeRFUtils(1, 1, R, R, 10, 4, True, scan_mode="warp")
    s2, idx2, cdf2 = ut.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], Nf, u=g["u"], return_aux=True)
    assert maxerr(cdf2, g["cdf"]) <= 3e-6
    flips = idx2.cpu().numpy() != g["idx"]
    assert flips.mean() <= 2e-4
    # depths are 1e-5-exact GIVEN the cdf (checked above and in test_sampler_sorted_merge); with its own cdf
    # (1e-6 away) a sample moves by |dcdf|/(c1-c0)*(m1-m0), i.e. more in nearly empty bins (peaked rows)
    so2, _, _ = O.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], g["u"], cdf=cdf2.cpu())
    assert maxerr(s2, so2) <= 1e-6
    assert maxerr(s2, g["samples"]) <= 1e-2   # out-of-range bins span [0, ~6): d(depth)/d(cdf) ~ 400 there
    # (3) TF-CPU semantics: out-of-range mid-point gather raises
    ut_raise = K.NeRFUtils(1, 1, R, R, 10, 4, True, oob_mode="raise")
    with pytest.raises(IndexError):
        ut_raise.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], Nf, u=g["u"])
    ut_clamp = K.NeRFUtils(1, 1, R, R, 10, 4, True, oob_mode="clamp")
    sc = ut_clamp.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], Nf, u=g["u"])
    so, _, _ = O.fine_hierarchical_sampling_chunk(g["mid"], g["weights"], g["u"], oob_mode=O.OOB_CLAMP)
    assert maxerr(sc, so) <= 1e-4


def test_sampler_full_image_shape_like_reference_test():
    """tests/model/nerf/test_nerf_utils.py:65-76 (random, unsorted coarse points; Nc=32, Nf=64)"""
    K = gpu()
    ut = K.NeRFUtils(2, 128, 128, 1024, 10, 4, True)
    cp = torch.rand(2, 128, 128, 32)
    mid = 0.5 * (cp[..., 1:] + cp[..., :-1])
    w = torch.rand(2, 128, 128, 32)
    fine = ut.fine_hierarchical_sampling(mid, w, 64)
    assert fine.shape == (2, 128, 128, 64) and not torch.isnan(fine).any()


@pytest.mark.parametrize("Nc,Nf", [(64, 128), (64, 256), (32, 64), (16, 40), (100, 77)])
def test_sampler_sorted_merge(Nc, Nf):
    gpu()
    from keras_nerf_b200 import _lib
    rng = np.random.default_rng(Nc * 1000 + Nf)
    R = 203
    t_c = np.sort(rng.uniform(2, 6, (R, Nc)).astype(np.float32), axis=-1)
    w = (rng.uniform(0, 1, (R, Nc)) ** 4).astype(np.float32)
    u = O.uniform24(rng, (R, Nf))
    dev = torch.device("cuda")
    ts = torch.empty(R, Nc + Nf, device=dev)
    samples = torch.empty(R, Nf, device=dev)
    cdf = torch.empty(R, Nc + 1, device=dev)
    tc_d, w_d, u_d = T(t_c).to(dev), T(w).to(dev), T(u).to(dev)   # keep alive across the raw-pointer call
    _lib.call("knerf_sample_fine", _lib.ptr(tc_d), None, _lib.ptr(w_d), _lib.ptr(u_d), 0,
              None, R, Nc, Nf, 0, _lib.ptr(ts), _lib.ptr(samples), None, _lib.ptr(cdf), None, _lib.stream())
    # the merged row is exactly sort(concat(t_coarse, the kernel's own samples))  (nerf.py:190-191)
    exp, _ = torch.sort(torch.cat([tc_d, samples], dim=-1), dim=-1)
    assert torch.equal(ts, exp)
    mid = 0.5 * (t_c[:, 1:] + t_c[:, :-1])
    so, _, _ = O.fine_hierarchical_sampling_chunk(mid, w, u, cdf=cdf.cpu())
    assert maxerr(samples, so) <= 1e-6


@pytest.mark.parametrize("Nc,Nf", [(64, 128), (64, 256), (32, 64), (16, 40)])
def test_sampler_builtin_rng_is_the_knerf_uniform_stream(Nc, Nf):
    """u=NULL draws sample e of the ray batch from Philox index e: same rows as passing knerf_uniform()'s output
    (covers the 4-draws-per-lane path and the scalar one)."""
    gpu()
    from keras_nerf_b200 import _lib
    dev = torch.device("cuda")
    R = 77
    g = torch.Generator().manual_seed(Nc + Nf)
    tc_d = torch.sort(torch.rand(R, Nc, generator=g) * 4 + 2, dim=-1).values.to(dev).contiguous()
    w_d = (torch.rand(R, Nc, generator=g) ** 4).to(dev)
    u_d = torch.empty(R, Nf, device=dev)
    _lib.call("knerf_uniform", _lib.ptr(u_d), R * Nf, 99, 0, _lib.stream())
    outs = []
    for u_ptr in (None, _lib.ptr(u_d)):
        ts, sm = torch.empty(R, Nc + Nf, device=dev), torch.empty(R, Nf, device=dev)
        _lib.call("knerf_sample_fine", _lib.ptr(tc_d), None, _lib.ptr(w_d), u_ptr, 99, None, R, Nc, Nf, 0,
                  _lib.ptr(ts), _lib.ptr(sm), None, None, None, _lib.stream())
        outs.append((ts, sm))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_sampler_builtin_rng_statistics():
    gpu()
    from keras_nerf_b200 import _lib
    dev = torch.device("cuda")
    n = 1 << 20
    a, b = torch.empty(n, device=dev), torch.empty(n, device=dev)
    _lib.call("knerf_uniform", _lib.ptr(a), n, 7, 0, _lib.stream())
    _lib.call("knerf_uniform", _lib.ptr(b), n, 8, 0, _lib.stream())
    assert 0.0 <= float(a.min()) and float(a.max()) < 1.0
    assert abs(float(a.mean()) - 0.5) < 2e-3 and abs(float(a.var()) - 1 / 12) < 2e-3
    assert not torch.equal(a, b)
    c = torch.empty(n, device=dev)
    _lib.call("knerf_uniform", _lib.ptr(c), n, 7, 0, _lib.stream())
    assert torch.equal(a, c)
    assert float(((a * 2 ** 24) % 1).abs().max()) == 0.0   # multiples of 2^-24


# ---- a6 MLP -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag,dx,dd", [("flagship", 63, 27), ("reftest", 99, 99)])
def test_mlp_forward_golden(tag, dx, dd):
    K = gpu()
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    g = load_golden(f"mlp_{tag}")
    mlp_mod.set_seed(int(g["init_seed"]))
    net = K.NeRFMLP(n_layers=8, dense_units=256, skip_layer=4)
    rgb, sigma = net((g["x"], g["dirs"]))
    assert rgb.shape == g["rgb"].shape and sigma.shape == g["sigma"].shape
    assert net.count_params() == int(g["n_params"])
    assert maxerr(rgb, g["rgb"]) <= 2e-6 and maxerr(sigma, g["sigma"]) <= 5e-6
    cfgd = net.get_config()
    assert cfgd['n_layers'] == 8 and cfgd['dense_units'] == 256 and cfgd['skip_layer'] == 4
    assert len(net.trainable_variables) == 24


@pytest.mark.parametrize("n_layers,units,skip", [(5, 64, 4), (3, 32, 1), (2, 100, 4)])
def test_mlp_forward_other_shapes(n_layers, units, skip):
    K = gpu()
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(5)
    net = K.NeRFMLP(n_layers=n_layers, dense_units=units, skip_layer=skip)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (70, 3, 15)).astype(np.float32)
    dr = rng.uniform(-1, 1, (70, 3, 9)).astype(np.float32)
    rgb, sigma = net((x, dr))
    cfg = O.NerfConfig(n_layers=n_layers, dense_units=units, skip_layer=skip)
    params = O.unflatten_params(net.params.cpu(), cfg, 15, 9)
    orgb, osig = O.mlp_forward(params, T(x), T(dr), cfg)
    assert maxerr(rgb, orgb) <= 2e-6 and maxerr(sigma, osig) <= 5e-6


# ---- a9/a10/a11/a13 whole model ---------------------------------------------------------------------------
def _model(precision="fp32", ray_chunks=None):
    K = gpu()
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    g = load_golden("model")
    mlp_mod.set_seed(int(g["init_seed"]))
    H, W, B = int(g["H"]), int(g["W"]), int(g["B"])
    model = K.NeRF(precision=precision)
    model.compile(optimizer="adam", loss="mse", batch_size=B, image_height=H, image_width=W,
                  ray_chunks=int(g["ray_chunks"]) if ray_chunks is None else ray_chunks, white_background=True)
    rays = (g["o"][None], g["d"][None], g["t"][None])
    return g, model, rays


def test_model_render_golden():
    g, model, rays = _model()
    c, f = model.predict_and_render_images(rays, u_fine=g["u_fine"])
    for name, res in (("coarse", c), ("fine", f)):
        assert res["image"].shape == g[f"image_{name}"].shape
        assert res["depth"].shape == g[f"depth_{name}"].shape and res["weights"].shape == g[f"weights_{name}"].shape
    # coarse pass: strict north-star tolerance
    assert maxerr(c["image"], g["image_coarse"]) <= 1e-5
    assert maxerr(c["depth"], g["depth_coarse"]) <= 1e-5
    assert maxerr(c["weights"], g["weights_coarse"]) <= 1e-5
    # fine pass END TO END is ill-conditioned in the reference itself: the out-of-range mid-point gather
    # (App. C-1) maps ~3% of the draws into [0, near) with d(depth)/d(cdf) ~ 6/pdf ~ 400, so the 1e-7 rounding
    # difference between two correct coarse networks (here 8e-8 on the weights) moves those depths by ~1e-3.
    # Parity of the fine pass is therefore pinned stage by stage (test_fine_pass_given_reference_depths,
    # test_sampler_golden); end to end only a loose bound is meaningful.
    assert maxerr(f["image"], g["image_fine"]) <= 1e-2
    mse = float(((f["image"].cpu() - T(g["image_fine"])) ** 2).mean())
    assert -10.0 * np.log10(max(mse, 1e-30)) > 55.0
    # chunking does not change the picture (same device arithmetic, different chunk boundaries)
    _, model2, _ = _model(ray_chunks=256)
    c2, f2 = model2.predict_and_render_images(rays, u_fine=g["u_fine"])
    assert maxerr(f2["image"], f["image"]) <= 1e-6 and maxerr(c2["weights"], c["weights"]) == 0.0


def _fine_inputs(g):
    """reference-side inputs of the fine pass: oracle coarse pass + sampler on the fixture (== golden)."""
    cfg = O.NerfConfig()
    rng = np.random.default_rng(int(g["init_seed"]))
    pc, pf = O.init_params(cfg, rng), O.init_params(cfg, rng)
    o, d, t = T(g["o"]).reshape(-1, 3), T(g["d"]).reshape(-1, 3), T(g["t"]).reshape(-1, cfg.n_coarse)
    with torch.no_grad():
        c = O.predict_and_render_chunk_single(pc, cfg, o, d, t, True)
        f = O.predict_and_render_chunk_single(pf, cfg, o, d, t, True, c["weights"], T(g["u_fine"]))
    assert maxerr(f["image"].reshape(g["image_fine"].shape), g["image_fine"]) <= 2e-6
    return cfg, pf, o, d, t, c, f


def test_fine_pass_given_reference_depths():
    """fine network + compositing through the C ABI on the reference's sorted depths: 1e-5 (north star)."""
    g, model, _ = _model(ray_chunks=256)
    from keras_nerf_b200 import _lib
    import ctypes as C
    cfg, pf, o, d, t, c, f = _fine_inputs(g)
    dev = model.device
    R, S = o.shape[0], cfg.n_coarse + cfg.n_fine
    # sampler given the reference's coarse weights: bit-exact cdf and bins, depths to 1e-6, sorted row exact
    ts = torch.empty(R, S, device=dev)
    idx = torch.empty(R, cfg.n_fine, dtype=torch.int32, device=dev)
    t_d, w_d, u_d = t.to(dev), c["weights"].to(dev), T(g["u_fine"]).to(dev)
    _lib.call("knerf_sample_fine", _lib.ptr(t_d), None, _lib.ptr(w_d), _lib.ptr(u_d), 0, None, R, cfg.n_coarse,
              cfg.n_fine, _lib.OOB_ZERO | _lib.SCAN_SEQUENTIAL, _lib.ptr(ts), None, _lib.ptr(idx, torch.int32), None,
              None, _lib.stream())
    np.testing.assert_array_equal(idx.cpu().numpy(), f["indices"].numpy())
    assert maxerr(ts, f["points"]) <= 1e-6
    # fine MLP + compositing on those depths
    o_d, d_d, pts = o.to(dev), d.to(dev), f["points"].to(dev).contiguous()
    rgbs = torch.empty(R, S, 4, device=dev)
    _lib.call("knerf_mlp_forward", C.byref(model.cfg), _lib.ptr(model.fine.params), None, _lib.ptr(o_d), _lib.ptr(d_d),
              _lib.ptr(pts), R, S, _lib.FP32, 0, _lib.ptr(rgbs), model._ws.data_ptr(), model._ws.numel(), _lib.stream())
    assert maxerr(rgbs[..., :3], f["rgb"]) <= 1e-5 and maxerr(rgbs[..., 3:], f["sigma"]) <= 1e-5
    img, dep, w = (torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, S, device=dev))
    _lib.call("knerf_composite_forward", _lib.ptr(rgbs), None, None, _lib.ptr(pts), R, S, 1, 1, 1e-10, _lib.ptr(img),
              _lib.ptr(dep), _lib.ptr(w), None, _lib.stream())
    assert maxerr(img.reshape(g["image_fine"].shape), g["image_fine"]) <= 1e-5
    assert maxerr(dep.reshape(g["depth_fine"].shape), g["depth_fine"]) <= 1e-5
    assert maxerr(w.reshape(g["weights_fine"].shape), g["weights_fine"]) <= 1e-5


def test_fine_backward_given_reference_depths():
    """fused compositing backward + MLP backward through the C ABI vs torch autograd of the oracle."""
    g, model, _ = _model(ray_chunks=256)
    from keras_nerf_b200 import _lib
    import ctypes as C
    cfg, pf, o, d, t, c, f = _fine_inputs(g)
    dev = model.device
    R, S = o.shape[0], cfg.n_coarse + cfg.n_fine
    target = T(g["images"][..., :3].reshape(-1, 3).copy())
    leaves = [(W.clone().requires_grad_(True), b.clone().requires_grad_(True)) for W, b in pf]
    out = O.render_given_points(leaves, cfg, o, d, f["points"], True)
    loss = O.mse(target, out["image"])
    gref = torch.cat([x.reshape(-1) for x in torch.autograd.grad(loss, [p for wb in leaves for p in wb])])
    o_d, d_d, pts, tgt = o.to(dev), d.to(dev), f["points"].to(dev).contiguous(), target.to(dev)
    rgbs, dpre = torch.empty(R, S, 4, device=dev), torch.empty(R, S, 4, device=dev)
    sq = torch.empty(R, device=dev)
    grads = torch.zeros_like(model.fine.params)
    ws, wsn = model._ws.data_ptr(), model._ws.numel()
    _lib.call("knerf_mlp_forward", C.byref(model.cfg), _lib.ptr(model.fine.params), None, _lib.ptr(o_d), _lib.ptr(d_d),
              _lib.ptr(pts), R, S, _lib.FP32, 1, _lib.ptr(rgbs), ws, wsn, _lib.stream())
    _lib.call("knerf_composite_backward", _lib.ptr(rgbs), _lib.ptr(pts), R, S, 1, 1, 1e-10, None, _lib.ptr(tgt),
              2.0 / (3.0 * R), 1, _lib.ptr(dpre), _lib.ptr(sq), _lib.stream())
    _lib.call("knerf_mlp_backward", C.byref(model.cfg), _lib.ptr(model.fine.params), None, _lib.ptr(dpre), R, S,
              _lib.FP32, _lib.ptr(grads), ws, wsn, _lib.stream())
    assert float(sq.sum().cpu()) / (3 * R) == pytest.approx(float(loss), rel=1e-5)
    scale = float(gref.abs().max())
    assert maxerr(grads / scale, gref / scale) <= 2e-5
    # per-tensor check so that a small tensor (biases, the sigma / rgb heads) cannot hide behind a large one
    off = 0
    for _, fi_, fo in O.layer_shapes(cfg):
        for n in (fi_ * fo, fo):
            seg, ref = grads[off:off + n].cpu(), gref[off:off + n]
            s = float(ref.abs().max())
            assert s > 0 and maxerr(seg / s, ref / s) <= 1e-4
            off += n


def test_model_train_two_steps_golden():
    g, model, rays = _model()
    cfg = O.NerfConfig()
    shapes = O.layer_shapes(cfg)
    for step in range(2):
        for m in model.metrics:
            m.reset_state()
        ci, fi = model.accumulate_gradients(g["images"], rays, u_fine=g["u_fine"])
        torch.cuda.synchronize()
        lc, lf = model._losses.tolist()
        model._losses.zero_()
        # coarse network: strict.  fine network end to end: loose, see test_model_render_golden (its strict
        # check is test_fine_backward_given_reference_depths); from step 1 on the coarse weights themselves
        # carry Adam's sign(g) noise on ~zero gradients, so only step 0 is strict.
        strict = step == 0
        assert lc == pytest.approx(float(g[f"s{step}_coarse_loss"]), rel=2e-5 if strict else 2e-3)
        assert lf == pytest.approx(float(g[f"s{step}_fine_loss"]), rel=5e-3)
        for name, gr in (("coarse", model.coarse_gradients_accumulator), ("fine", model.fine_gradients_accumulator)):
            gr = gr.cpu()
            gmax = float(gr.abs().max())
            tight = strict and name == "coarse"
            off = k = 0
            for _, fi_, fo in shapes:
                for n in (fi_ * fo, fo):
                    seg = gr[off:off + n].numpy()
                    m = min(n, 32)
                    np.testing.assert_allclose(seg[:m], g[f"s{step}_grad_{name}_head"][k][:m],
                                               atol=(5e-5 if tight else 5e-2) * gmax)
                    assert float(np.abs(seg).sum(dtype=np.float64)) == pytest.approx(
                        float(g[f"s{step}_grad_{name}_abssum"][k]), rel=5e-4 if tight else 5e-2, abs=1e-9)
                    off += n
                    k += 1
        gmax_c = float(np.abs(g[f"s{step}_grad_coarse_head"]).max())
        model.apply_gradients()
        assert float(model.coarse_gradients_accumulator.abs().max()) == 0.0     # nerf.py:465-471
        checked = 0
        for name, net in (("coarse", model.coarse), ("fine", model.fine)):
            for k, v in enumerate(net.trainable_variables):
                m = min(v.numel(), 32)
                got, want = v.reshape(-1)[:m].cpu().numpy(), g[f"s{step}_param_{name}_head"][k][:m]
                # Adam's first steps move every weight by ~lr * sign(g): where the gradient is ~0 the sign is noise,
                # so 2 * lr per step taken is all that can be said there ...
                np.testing.assert_allclose(got, want, atol=2.1e-3 * (step + 1))
                # ... but where the reference's gradient is clearly non-zero (coarse network, first step: the
                # gradients themselves agree to 5e-5 * max) the update must be the reference's, to 1e-5
                if name == "coarse" and step == 0:
                    sel = np.abs(g[f"s{step}_grad_{name}_head"][k][:m]) > 1e-2 * gmax_c
                    checked += int(sel.sum())
                    np.testing.assert_allclose(got[sel], want[sel], atol=1e-5)
        assert step > 0 or checked >= 50


def test_train_step_metrics_and_oracle():
    """train_step end to end (metrics dict of nerf.py:323-330) against the oracle's step on the same inputs."""
    g, model, rays = _model()
    cfg = O.NerfConfig()
    rng = np.random.default_rng(int(g["init_seed"]))
    pc, pf = O.init_params(cfg, rng), O.init_params(cfg, rng)
    logs = model.train_step((g["images"], rays), u_fine=g["u_fine"])
    assert set(logs) == {"coarse_loss", "coarse_psnr", "coarse_ssim", "fine_loss", "fine_psnr", "fine_ssim"}
    orays = tuple(T(np.asarray(r)) for r in rays)
    ref = O.train_step(pc, pf, O.AdamState(), O.AdamState(), cfg, g["images"], orays, g["u_fine"],
                       int(g["ray_chunks"]), True)
    assert logs["coarse_loss"] == pytest.approx(ref["coarse_loss"], rel=2e-5)
    assert logs["coarse_psnr"] == pytest.approx(ref["coarse_psnr"], abs=1e-3)
    assert logs["coarse_ssim"] == pytest.approx(float(g["s0_coarse_ssim"]), abs=1e-4)
    assert logs["fine_loss"] == pytest.approx(ref["fine_loss"], rel=5e-3)       # end-to-end fine: loose (C-1)
    assert logs["fine_psnr"] == pytest.approx(float(g["s0_fine_psnr"]), abs=0.05)
    assert logs["fine_ssim"] == pytest.approx(float(g["s0_fine_ssim"]), abs=5e-3)
    # Adam moved the weights exactly as the oracle's Keras-Adam restatement does where gradients are not ~0
    gmask = ref["grad_coarse"].abs() > 1e-3 * ref["grad_coarse"].abs().max()
    new = O.flatten_params(ref["params_coarse"])
    assert maxerr(model.coarse.params.cpu()[gmask], new[gmask]) <= 1e-5


def test_adam_kernel_vs_oracle():
    gpu()
    from keras_nerf_b200.model.nerf.nerf import Adam
    rng = np.random.default_rng(0)
    n = 100003
    p0 = rng.normal(size=n).astype(np.float32)
    st = O.AdamState()
    opt = Adam()
    p = T(p0.copy()).cuda()
    pref = T(p0.copy())
    for step in range(3):
        gnp = (rng.normal(size=n) * 10.0 ** rng.integers(-6, 1, size=n)).astype(np.float32)
        gbuf = T(gnp.copy()).cuda()
        opt.apply_flat(p, gbuf, zero_grads=True)
        pref = O.adam_apply(pref, T(gnp), st)
        assert float(gbuf.abs().max()) == 0.0
        assert maxerr(p, pref) <= 2e-6


# ---- BASELINE config 2: 4096-ray fp32 training step; properties that need no oracle at that size ---------
def test_config2_4096_ray_step_properties():
    K = gpu()
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    H = W = 64
    rng = np.random.default_rng(7)
    pose = K.pose_spherical(30.0, -30.0, 4.0)
    focal = K.get_focal_from_fov(0.6911112070083618, W)
    u_c = O.uniform24(rng, (H, W, 64))
    u_f = O.uniform24(rng, (H * W, 128))
    o, d, t = K.RaysGenerator(focal, W, H, 2.0, 6.0, 64)(pose, u=u_c)
    rays = (o[None], d[None], t[None])
    images = rng.uniform(0, 1, (1, H, W, 4)).astype(np.float32)
    grads = {}
    for chunks in (4096, 1024):
        mlp_mod.set_seed(42)
        model = K.NeRF(precision="fp32")
        model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=H, image_width=W, ray_chunks=chunks,
                      white_background=True)
        ci, fi = model.accumulate_gradients(images, rays, u_fine=u_f)
        torch.cuda.synchronize()
        grads[chunks] = (model._grad_flat.clone(), model._losses.clone(), fi.clone())
    g1, l1, i1 = grads[4096]
    g4, l4, i4 = grads[1024]
    # accumulating 4 chunks of g/4 == the un-chunked gradient
    scale = float(g1.abs().max())
    assert torch.isfinite(g1).all() and scale > 0
    assert maxerr(g4 / scale, g1 / scale) <= 2e-5
    assert maxerr(l4, l1) <= 1e-6
    assert maxerr(i4, i1) <= 1e-6
    assert float(i1.min()) >= 0.0 and float(i1.max()) <= 1.0
    # finite-difference check of the fused backward on one weight direction (fine net)
    mlp_mod.set_seed(42)
    model = K.NeRF(precision="fp32")
    model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=H, image_width=W, ray_chunks=4096,
                  white_background=True)
    n = model.fine.params.numel()
    direction = torch.from_numpy(np.random.default_rng(3).normal(size=n).astype(np.float32)).cuda()
    direction /= direction.norm()
    base = model.fine.params.clone()
    vals = []
    for eps in (+2e-3, -2e-3):
        model.fine.params.copy_(base + eps * direction)
        _, f = model.predict_and_render_images(rays, u_fine=u_f)
        tgt = torch.from_numpy(images[..., :3]).cuda()
        vals.append(float(((f["image"].double() - tgt.double()) ** 2).mean()))
    fd = (vals[0] - vals[1]) / 4e-3
    analytic = float((g1[n:].double() * direction.double()).sum())
    assert fd == pytest.approx(analytic, rel=5e-2, abs=1e-6)


# ---- empty inputs and argument errors through the C ABI ------------------------------------------------------
def test_empty_inputs_and_argument_errors():
    """R = 0 is a no-op for every per-ray entry point (a tf op on a [0, ...] tensor); bad arguments come back as a
    negative status + message (KnerfError), never a crash."""
    gpu()
    from keras_nerf_b200 import _lib
    dev = torch.device("cuda")
    z = torch.zeros(16, device=dev)
    st = _lib.stream()
    P = _lib.ptr(z)
    _lib.call("knerf_composite_forward", P, None, None, P, 0, 64, 1, 1, 1e-10, P, P, P, P, st)
    _lib.call("knerf_composite_backward", P, P, 0, 64, 1, 1, 1e-10, None, P, 1.0, 1, P, P, st)
    _lib.call("knerf_sample_fine", P, None, P, None, 1, None, 0, 64, 128, 0, P, None, None, None, None, st)
    _lib.call("knerf_positional_encoding", P, 0, 3, 10, P, 63, st)
    _lib.call("knerf_encode_position_and_directions", P, P, P, 0, 64, 10, 4, P, 63, P, 27, st)
    _lib.call("knerf_uniform", P, 0, 1, 0, st)
    _lib.call("knerf_adam_step", P, P, P, P, 0, 1e-3, 0.9, 0.999, 1e-7, 1, 1, st)
    torch.cuda.synchronize()
    assert float(z.abs().sum()) == 0.0                        # nothing was written
    bad = [("knerf_composite_forward", (None, None, None, P, 4, 64, 1, 1, 1e-10, P, P, P, P, st)),      # no inputs
           ("knerf_sample_fine", (P, P, P, None, 1, None, 1, 64, 128, 0, P, None, None, None, None, st)),  # both forms
           ("knerf_sample_fine", (P, None, P, None, 1, None, 1, 1, 128, 0, P, None, None, None, None, st)),  # Nc < 2
           ("knerf_sample_fine", (P, None, P, None, 1, None, 1, 64, 4096, 0, P, None, None, None, None, st)),  # Nf max
           ("knerf_sample_fine", (P, None, P, None, 1, None, 1, 64, 128, 7, P, None, None, None, None, st)),  # oob_mode
           ("knerf_generate_rays", (None, 4, 4, 100.0, 2.0, 6.0, 8, None, 1, P, P, P, st)),
           ("knerf_generate_rays", (P, 0, 4, 100.0, 2.0, 6.0, 8, None, 1, P, P, P, st)),
           ("knerf_image_metrics", (P, P, 1, 4, 4, 3, 1.0, P, P, P, 16, st))]                        # below the window
    for name, args in bad:
        with pytest.raises(_lib.KnerfError):
            _lib.call(name, *args)
        assert len(_lib.load().knerf_last_error()) > 0
    with pytest.raises(_lib.KnerfError):                      # host tensors are refused before the call
        _lib.ptr(torch.zeros(4))
