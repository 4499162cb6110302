"""GPU (-m gpu): parity against the oracle AT THE SIZES BASELINE.json NAMES.

  configs[1]  4096-ray training step in fp32 on one B200, "checked against the reference (bit-exact fine-sample
              bins, 1e-5 composited RGB)"                                  -> test_config1_*
  configs[0]  batch 2 x 128 x 128, ray_chunks 2048 (train_single.py:16-17): one train step + one full-image render
              -> test_config0_* (the oracle runs a subsample of the 16 chunks; the chunk-accumulation rule is
              checked over all of them)
  and the bf16 tolerance of the north star (2e-3 per pixel, 0.05 dB) on TRAINED weights (peaked sigma, saturated
  heads), not only at glorot initialisation                                -> test_bf16_tolerance_on_trained_weights
"""
import ctypes as C

import numpy as np
import pytest
import torch

import oracle as O
from oracle import nerf_oracle as ON

pytestmark = pytest.mark.gpu


def T(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def maxerr(a, b):
    a = a.detach().cpu() if torch.is_tensor(a) else T(np.asarray(a))
    b = b.detach().cpu() if torch.is_tensor(b) else T(np.asarray(b))
    return float((a.double() - b.double()).abs().max())


def _scene(H, W, B, seed):
    """rays of B orbit views (lego-shaped synthetic scene: fov / near / far of the reference's defaults) with
    explicit uniform draws, from the ORACLE's ray generator so that both sides start from identical inputs"""
    rng = np.random.default_rng(seed)
    focal = O.get_focal_from_fov(0.6911112070083618, W)
    os_, ds_, ts_ = [], [], []
    for b in range(B):
        pose = O.pose_spherical(30.0 + 97.0 * b, -30.0, 4.0)
        u_c = O.uniform24(rng, (H, W, 64))
        o, d, t = O.generate_rays(pose, H, W, focal, 2.0, 6.0, 64, u_c)
        os_.append(o), ds_.append(d), ts_.append(t)
    rays = tuple(torch.stack(x) for x in (os_, ds_, ts_))
    images = rng.uniform(0, 1, (B, H, W, 4)).astype(np.float32)
    u_f = O.uniform24(rng, (B * H * W, 128))
    return rays, images, u_f


def _gpu_model(B, H, W, ray_chunks, precision="fp32", seed=42, training=True):
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(seed)
    m = K.NeRF(precision=precision, scan_mode="sequential")
    m.compile(optimizer="adam", loss="mse", batch_size=B, image_height=H, image_width=W, ray_chunks=ray_chunks,
              white_background=True, is_training=training)
    return m


def _oracle_params(seed=42):
    cfg = O.NerfConfig()
    rng = np.random.default_rng(seed)
    return cfg, O.init_params(cfg, rng), O.init_params(cfg, rng)


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).norm() / b.norm())


# ---- configs[1]: 4096-ray fp32 training step -------------------------------------------------------------------
def test_config1_4096_ray_train_step_vs_oracle():
    from keras_nerf_b200 import _lib
    H = W = 64
    rays, images, u_f = _scene(H, W, 1, seed=11)
    cfg, pc, pf = _oracle_params()
    m = _gpu_model(1, H, W, 4096)
    assert torch.equal(m.coarse.params.cpu(), O.flatten_params(pc))

    # --- forward: coarse image / weights at 1e-5; bins bit-exact given the reference's cdf; fine pass at 1e-5
    #     given the reference's sorted depths (its end-to-end form is ill-conditioned: test_illconditioning_cpu.py)
    o, d, t = (r.reshape(4096, -1) for r in rays)
    with torch.no_grad():
        oc = O.predict_and_render_chunk_single(pc, cfg, o, d, t, True)
        of = O.predict_and_render_chunk_single(pf, cfg, o, d, t, True, oc["weights"], T(u_f))
    c, f = m.predict_and_render_images(rays, u_fine=u_f)
    assert maxerr(c["image"].reshape(-1, 3), oc["image"]) <= 1e-5
    assert maxerr(c["depth"].reshape(-1), oc["depth"]) <= 1e-5
    assert maxerr(c["weights"].reshape(4096, 64), oc["weights"]) <= 1e-5
    dev = torch.device("cuda")
    idx = torch.empty(4096, 128, dtype=torch.int32, device=dev)
    smp = torch.empty(4096, 128, device=dev)
    mid = (0.5 * (t[:, 1:] + t[:, :-1])).contiguous().to(dev)
    # (named tensors: a temporary's memory would be handed to the next allocation while the kernel still reads it)
    w_d, u_d, cdf_d = oc["weights"].contiguous().to(dev), T(u_f).to(dev), of["cdf"].contiguous().to(dev)
    _lib.call("knerf_sample_fine", None, _lib.ptr(mid), _lib.ptr(w_d), _lib.ptr(u_d), 0, _lib.ptr(cdf_d), 4096, 64,
              128, _lib.OOB_ZERO, None, _lib.ptr(smp), idx.data_ptr(), None, None, _lib.stream())
    assert torch.equal(idx.cpu(), of["indices"])                            # bit-exact bins (north star)
    assert maxerr(smp, of["t_fine"]) <= 1e-5                                # fp32 sample depths
    rgbs = torch.empty(4096, 192, 4, device=dev)
    pts, o_d, d_d = of["points"].contiguous().to(dev), o.contiguous().to(dev), d.contiguous().to(dev)
    _lib.call("knerf_mlp_forward", C.byref(m.cfg), _lib.ptr(m.fine.params), None, _lib.ptr(o_d), _lib.ptr(d_d),
              _lib.ptr(pts), 4096, 192, _lib.FP32, 0, _lib.ptr(rgbs), m._ws.data_ptr(), m._ws.numel(), _lib.stream())
    img = torch.empty(4096, 3, device=dev)
    dep = torch.empty(4096, device=dev)
    _lib.call("knerf_composite_forward", _lib.ptr(rgbs), None, None, _lib.ptr(pts), 4096, 192, 1, 1, 1e-10,
              _lib.ptr(img), _lib.ptr(dep), None, None, _lib.stream())
    assert maxerr(img, of["image"]) <= 1e-5 and maxerr(dep, of["depth"]) <= 1e-5

    # --- the training step: losses, accumulated gradients, Adam
    ref = O.train_step(pc, pf, O.AdamState(), O.AdamState(), cfg, images, rays, u_f, 4096, True)
    m.accumulate_gradients(images, rays, u_fine=u_f)
    torch.cuda.synchronize()
    lc, lf = m._losses.tolist()
    assert lc == pytest.approx(ref["coarse_loss"], rel=2e-5)
    assert lf == pytest.approx(ref["fine_loss"], rel=5e-3)                  # end-to-end fine: loose (App. C-1)
    gc = m.coarse_gradients_accumulator
    assert _rel(gc, ref["grad_coarse"]) <= 1e-4
    assert maxerr(gc, ref["grad_coarse"]) <= 2e-5 * float(ref["grad_coarse"].abs().max())
    assert _rel(m.fine_gradients_accumulator, ref["grad_fine"]) <= 5e-2     # through the ill-conditioned depths
    m._losses.zero_()
    m.apply_gradients()
    new_c = O.flatten_params(ref["params_coarse"])
    mask = ref["grad_coarse"].abs() > 1e-3 * ref["grad_coarse"].abs().max()  # Adam's sign(g) on ~0 gradients is noise
    assert int(mask.sum()) > 1000
    assert maxerr(m.coarse.params.cpu()[mask], new_c[mask]) <= 1e-5


# ---- configs[0]: batch 2 x 128 x 128, ray_chunks 2048 ------------------------------------------------------------
def test_config0_render_and_train_step_vs_oracle_on_a_subsample_of_chunks():
    B, H, W, RC = 2, 128, 128, 2048
    rays, images, u_f = _scene(H, W, B, seed=5)
    cfg, pc, pf = _oracle_params()
    m = _gpu_model(B, H, W, RC)
    n = B * H * W
    assert m.sequential_chunks == 16
    c, f = m.predict_and_render_images(rays, u_fine=u_f)
    o, d, t = (r.reshape(n, -1) for r in rays)
    tgt = T(images[..., :3]).reshape(n, 3)
    sub = (0, 7, 15)                       # first chunk, a middle one of image 0, the last one of image 1
    grads_ref = {}
    for i in sub:
        s = slice(i * RC, (i + 1) * RC)
        pcr, pfr = ON._req(pc), ON._req(pf)
        oc = O.predict_and_render_chunk_single(pcr, cfg, o[s], d[s], t[s], True)
        of = O.predict_and_render_chunk_single(pfr, cfg, o[s], d[s], t[s], True, oc["weights"].detach(), T(u_f)[s])
        assert maxerr(c["image"].reshape(n, 3)[s], oc["image"]) <= 1e-5
        assert maxerr(c["depth"].reshape(n)[s], oc["depth"]) <= 1e-5
        assert maxerr(c["weights"].reshape(n, 64)[s], oc["weights"]) <= 1e-5
        fi = f["image"].reshape(n, 3)[s].cpu()
        assert maxerr(fi, of["image"]) <= 1e-2                                # end to end: loose (App. C-1)
        assert -10 * np.log10(float(((fi - of["image"].detach()) ** 2).mean()) + 1e-30) > 55.0
        grads_ref[i] = (ON._grads_flat(O.mse(tgt[s], oc["image"]), pcr), ON._grads_flat(O.mse(tgt[s], of["image"]), pfr))

    # per-chunk gradients of the same three chunks through the C ABI (a one-chunk model over the chunk's rays)
    m1 = _gpu_model(1, 16, 128, RC)
    per_chunk = []
    for i in range(16):
        s = slice(i * RC, (i + 1) * RC)
        r1 = tuple(x.reshape(n, -1)[s].reshape(1, 16, 128, -1) for x in rays)
        m1.accumulate_gradients(images.reshape(n, 4)[s].reshape(1, 16, 128, 4), r1, u_fine=u_f[s], want_images=False)
        torch.cuda.synchronize()
        per_chunk.append(m1._grad_flat.clone())
        m1._grad_flat.zero_()
        m1._losses.zero_()
        if i in grads_ref:
            gc, gf = grads_ref[i]
            np_ = gc.numel()
            assert maxerr(per_chunk[-1][:np_], gc) <= 2e-5 * float(gc.abs().max())
            assert _rel(per_chunk[-1][:np_], gc) <= 1e-4
            assert _rel(per_chunk[-1][np_:], gf) <= 5e-2
    # the step over all 16 chunks accumulates g_i / 16 (nerf.py:383-384,412-413)
    m.accumulate_gradients(images, rays, u_fine=u_f, want_images=False)
    torch.cuda.synchronize()
    want = torch.stack(per_chunk).double().mean(0)
    got = m._grad_flat.double()
    # (fp32 atomics inside every chunk's weight-gradient kernels: 2e-5 of the largest entry, as in
    # test_gpu_parity.py::test_config2_4096_ray_step_properties)
    assert float((got - want.to(got.device)).abs().max()) <= 2e-5 * float(want.abs().max())


# ---- bf16 tolerance on trained weights ----------------------------------------------------------------------------
def test_bf16_tolerance_on_trained_weights():
    """The north star states the bf16 tolerance (max-abs 2e-3 per pixel, 0.05 dB PSNR) for RANDOM-INIT weights, where
    tests/test_gpu_tc.py checks it.  This test measures the same on TRAINED weights (600 optimizer steps on the
    synthetic scene: sigma peaked at the surface, rgb heads saturated), the regime a user renders in.  What holds
    there, and is asserted: the PSNR criterion (<= 0.05 dB) on every view; per pixel the median error is 0 and
    >= 97 % of the pixels are inside 2e-3 -- but NOT all of them: silhouette pixels (accumulated opacity 0.03-0.4,
    where d(pixel)/d(sigma) is largest) reach 1e-2 .. 3e-2 (measured: 99th percentile 2.5e-3 .. 3.0e-3, maximum
    1.0e-2 .. 2.6e-2 over three views), because eight layers of bf16 operands leave ~0.5 % on the sigma
    pre-activation.  A user who needs 1e-5 on trained weights has the fp32 modes; DESIGN.md §2 says so."""
    import keras_nerf_b200 as K
    from keras_nerf_b200 import _lib
    from keras_nerf_b200.data.synthetic import SyntheticScene
    dev = torch.device("cuda")
    R = 128 * 128                      # whole 128 x 128 views per step, as train.py feeds them (batch_size 1)
    scene = SyntheticScene(128, 64, n_views=40, device=dev)
    m16 = _gpu_model(1, 64, 256, R, precision="bf16")
    first = last = None
    for step in range(600):
        img, rays = scene.ray_batch((step * 7) % 40, R, seed=step)
        m16.accumulate_gradients(img, rays, seed=10_000 + step, want_images=False)
        m16.apply_gradients()
        if step in (0, 599):
            torch.cuda.synchronize()
            lf = float(m16._losses[1])
            first, last = (lf, last) if step == 0 else (first, lf)
        m16._losses.zero_()
    assert last < 0.5 * first, (first, last)                       # it did train
    m32 = _gpu_model(1, 64, 256, R, precision="fp32", training=False)
    m32.coarse.params.copy_(m16.coarse.params)
    m32.fine.params.copy_(m16.fine.params)
    m16._repack()
    worst_c = worst_f = 0.0
    for view in (3, 17, 31):
        img, rays = scene.ray_batch(view, R, seed=777)
        o, d, t = (r.reshape(R, -1).contiguous() for r in rays)
        u = torch.rand(R, 128, generator=torch.Generator().manual_seed(view)).to(dev)
        c32, f32 = m32.predict_and_render_images(rays, u_fine=u)
        c16, _ = m16.predict_and_render_images(rays, u_fine=u)
        tgt = img[..., :3].reshape(R, 3)
        psnr = lambda x: float(-10 * torch.log10(((x.reshape(R, 3) - tgt) ** 2).mean()))  # noqa: E731
        # coarse pass end to end
        e = (c16["image"] - c32["image"]).abs().amax(-1).reshape(-1)
        print(f"view {view} coarse: max {float(e.max()):.2e} q99 {float(torch.quantile(e, 0.99)):.2e} "
              f">2e-3 {float((e > 2e-3).float().mean()):.4f}")
        worst_c = max(worst_c, float(e.max()))
        assert float((e > 2e-3).float().mean()) <= 0.03 and float(torch.quantile(e, 0.99)) <= 5e-3
        assert abs(psnr(c16["image"]) - psnr(c32["image"])) <= 0.05
        # fine network on the SAME sorted depths (the fp32 run's): the MLP mode's own error, without the reference's
        # out-of-range-gather amplification of last-bit differences in the coarse weights
        ts = torch.empty(R, 192, device=dev)
        outs = [torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, 64, device=dev)]
        outf = [torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, 192, device=dev)]
        m32._render_rays(o, d, t, u, 0, outs, outf, t_sorted=ts)
        imgs = {}
        for name, mm in (("fp32", m32), ("bf16", m16)):
            rgbs = torch.empty(R, 192, 4, device=dev)
            _lib.call("knerf_mlp_forward", C.byref(mm.cfg), _lib.ptr(mm.fine.params), mm._packed_ptr("fine"),
                      _lib.ptr(o), _lib.ptr(d), _lib.ptr(ts), R, 192, mm._prec, 0, _lib.ptr(rgbs), mm._ws.data_ptr(),
                      mm._ws.numel(), _lib.stream())
            im = torch.empty(R, 3, device=dev)
            _lib.call("knerf_composite_forward", _lib.ptr(rgbs), None, None, _lib.ptr(ts), R, 192, 1, 1, 1e-10,
                      _lib.ptr(im), None, None, None, _lib.stream())
            imgs[name] = im
        e = (imgs["bf16"] - imgs["fp32"]).abs().amax(-1).reshape(-1)
        print(f"view {view} fine (same depths): max {float(e.max()):.2e} q99 {float(torch.quantile(e, 0.99)):.2e} "
              f">2e-3 {float((e > 2e-3).float().mean()):.4f}")
        worst_f = max(worst_f, float(e.max()))
        assert float((e > 2e-3).float().mean()) <= 0.05 and float(torch.quantile(e, 0.99)) <= 8e-3
        assert abs(psnr(imgs["bf16"]) - psnr(imgs["fp32"])) <= 0.05
    assert worst_c <= 5e-2, worst_c
    assert worst_f <= 5e-2, worst_f
