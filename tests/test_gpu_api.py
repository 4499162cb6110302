"""GPU (-m gpu): the reference-facing class API around the hot path -- compile-time checks, test_step, fit loop,
checkpoint round trip (keras_nerf/model/nerf/nerf.py:45-136,475-497; train_single.py:137-148)."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def _setup(precision="fp32", **kw):
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    g = load_golden("model")
    mlp_mod.set_seed(42)
    m = K.NeRF(precision=precision, **kw)
    m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=16, image_width=16, ray_chunks=128,
              white_background=True)
    rays = (g["o"][None], g["d"][None], g["t"][None])
    return K, g, m, rays


def test_compile_checks_like_reference():
    import keras_nerf_b200 as K
    m = K.NeRF()
    with pytest.raises(AssertionError):        # nerf.py:100: ray_chunks must divide the number of rays
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=10, image_width=10, ray_chunks=33)
    m2 = K.NeRF()
    m2.compile(optimizer="adam", loss="mse", batch_size=1, image_height=4, image_width=4, ray_chunks=4096)
    assert m2.ray_chunks == 16 and m2.sequential_chunks == 1      # nerf.py:95-98 clamps to num_rays
    assert m2.coarse.count_params() == 595844 and len(m2.fine.trainable_variables) == 24
    with pytest.raises(NotImplementedError):
        K.NeRF().compile(optimizer="sgd", loss="mse", batch_size=1, image_height=4, image_width=4, ray_chunks=16)


def test_test_step_matches_oracle_losses():
    K, g, m, rays = _setup()
    logs = m.test_step((g["images"], rays), u_fine=g["u_fine"])
    cfg = O.NerfConfig()
    rng = np.random.default_rng(42)
    pc, pf = O.init_params(cfg, rng), O.init_params(cfg, rng)
    orays = tuple(torch.from_numpy(np.asarray(r)) for r in rays)
    c, f = O.predict_and_render_images(pc, pf, cfg, orays, g["u_fine"], 128, True)
    tgt = torch.from_numpy(g["images"][..., :3])
    assert logs["coarse_loss"] == pytest.approx(float(O.mse(tgt, c["image"])), rel=1e-5)
    assert logs["fine_loss"] == pytest.approx(float(O.mse(tgt, f["image"])), rel=5e-3)
    assert logs["coarse_psnr"] == pytest.approx(float(O.psnr(tgt, c["image"]).mean()), abs=1e-3)
    assert set(logs) == {"coarse_loss", "coarse_psnr", "coarse_ssim", "fine_loss", "fine_psnr", "fine_ssim"}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fit_reduces_loss_and_checkpoint_round_trip(precision, tmp_path):
    K, g, m, rays = _setup(precision)
    # a learnable target: the analytic sphere of the synthetic scene instead of noise
    from keras_nerf_b200.data.synthetic import analytic_rgba
    img = analytic_rgba(torch.from_numpy(g["o"]).cuda(), torch.from_numpy(g["d"]).cuda(), True, albedo=1.0)[None]
    seen = []

    class Monitor:   # the two hooks NeRFTrainMonitor uses (callback.py:62,113)
        def on_train_batch_end(self, step, logs):
            seen.append(("batch", step))

        def on_epoch_end(self, epoch, logs):
            seen.append(("epoch", epoch, logs["fine_loss"]))

    hist = m.fit([(img, rays)] * 2, epochs=4, validation_data=[(img, rays)], callbacks=[Monitor()], verbose=0)
    assert len(hist["coarse_loss"]) == 4 and "val_fine_psnr" in hist
    assert hist["coarse_loss"][-1] < hist["coarse_loss"][0]
    assert [s for s in seen if s[0] == "epoch"][-1][1] == 3 and ("batch", 1) in seen
    # save_model / load_model (nerf.py:45-76,132-136): config json + per-net weights, Keras [in,out] layout
    path = str(tmp_path / "model")
    m.save_model(path)
    cfg = json.load(open(os.path.join(path, "model_config.json")))
    assert cfg == {"n_coarse": 64, "n_fine": 128, "pos_emb_xyz": 10, "pos_emb_dir": 4, "n_layers": 8,
                   "dense_units": 256, "skip_layer": 4}
    m2 = K.NeRF(model_path=path, precision=precision)
    m2.compile(optimizer="adam", loss="mse", batch_size=1, image_height=16, image_width=16, ray_chunks=256,
               white_background=True, is_training=False)
    assert torch.equal(m2.coarse.params, m.coarse.params) and torch.equal(m2.fine.params, m.fine.params)
    a = m.predict_and_render_images(rays, u_fine=g["u_fine"])[1]["image"]
    b = m2.predict_and_render_images(rays, u_fine=g["u_fine"])[1]["image"]
    assert float((a - b).abs().max()) <= 1e-6
    w = m2.coarse.get_weights()
    assert w[0].shape == (63, 256) and w[1].shape == (256,) and w[10].shape == (319, 256) and w[-2].shape == (128, 3)


def test_rays_generator_feeds_model_like_inference_script():
    """inference.py:61-114: orbit poses -> RaysGenerator -> predict_and_render_images -> frames"""
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(42)
    m = K.NeRF(precision="bf16")
    wh = 32
    m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=wh, image_width=wh, ray_chunks=512,
              white_background=True, is_training=False)
    gen = K.RaysGenerator(K.get_focal_from_fov(0.6911112070083618, wh), wh, wh, 2.0, 6.0, m.n_coarse)
    frames = []
    for theta in range(0, 360, 120):
        o, d, t = gen(K.pose_spherical(float(theta), -30.0, 4.0))
        _, fine = m.predict_and_render_images((o[None], d[None], t[None]))
        frames.append(fine["image"][0])
        assert fine["image"].shape == (1, wh, wh, 3) and fine["depth"].shape == (1, wh, wh)
        assert fine["weights"].shape == (1, wh, wh, 192)
        assert float(fine["image"].min()) >= 0.0 and float(fine["image"].max()) <= 1.0
    assert not torch.equal(frames[0], frames[1])


def test_class_api_accepts_dlpack_producers():
    """The shim consumes any DLPack producer zero-copy (how a tf.Tensor reaches libknerf): the same
    render from torch tensors and from objects that expose only __dlpack__."""
    K, g, m, rays = _setup()

    class Foreign:
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, stream=None, **kw):
            return self._t.__dlpack__()

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    dev = m.device
    t_rays = tuple(torch.as_tensor(r).to(dev) for r in rays)
    u = torch.as_tensor(g["u_fine"]).to(dev)
    a = m.predict_and_render_images(t_rays, u_fine=u)
    b = m.predict_and_render_images(tuple(Foreign(r) for r in t_rays), u_fine=Foreign(u))
    assert torch.equal(a[1]["image"], b[1]["image"]) and torch.equal(a[0]["depth"], b[0]["depth"])
    ut = K.NeRFUtils(1, 1, 8, 8, 10, 4, True)
    rgb, sig = torch.rand(8, 16, 3, device=dev), torch.rand(8, 16, 1, device=dev)
    t = torch.sort(torch.rand(8, 16, device=dev) * 4 + 2, dim=-1).values
    o1 = ut.render_image_depth_chunk(rgb, sig, t)
    o2 = ut.render_image_depth_chunk(Foreign(rgb), Foreign(sig), Foreign(t))
    assert all(torch.equal(x, y) for x, y in zip(o1, o2))


def test_bf16_request_on_other_shape_falls_back_loudly(caplog):
    """`--precision bf16` is the scripts' default, and the reference's CLI accepts any --num_units / --num_layers /
    --skip_layer / --pos_emb_*: a shape the fused bf16 kernels do not implement must train (fp32_tc mode), not raise,
    and must say so."""
    import logging
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(1)
    m = K.NeRF(precision="bf16", n_layers=8, dense_units=128, skip_layer=2, pos_emb_xyz=6, pos_emb_dir=2)   # three concats
    with caplog.at_level(logging.WARNING):
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=8, image_width=16, ray_chunks=64,
                  white_background=True)
    assert m.precision == "fp32_tc" and any("fp32_tc" in r.message for r in caplog.records)
    rng = np.random.default_rng(0)
    pose = K.pose_spherical(10.0, -30.0, 4.0)
    o, d, t = K.RaysGenerator(K.get_focal_from_fov(0.69, 16), 16, 8, 2.0, 6.0, 64)(pose, seed=3)
    images = rng.uniform(0, 1, (1, 8, 16, 4)).astype(np.float32)
    logs = [m.train_step((images, (o[None], d[None], t[None])), seed=5 + i) for i in range(3)]
    assert all(np.isfinite(l["fine_loss"]) for l in logs) and logs[2]["coarse_loss"] != logs[0]["coarse_loss"]


def test_bf16_train_step_with_fewer_encoding_frequencies():
    """--pos_emb_xyz 6 --pos_emb_dir 2 stays on the fused bf16 kernels (no fall-back) through the whole train step and
    tracks the fp32 mode's losses"""
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    rng = np.random.default_rng(0)
    pose = K.pose_spherical(10.0, -30.0, 4.0)
    o, d, t = K.RaysGenerator(K.get_focal_from_fov(0.69, 32), 32, 16, 2.0, 6.0, 64)(pose, seed=3)
    images = rng.uniform(0, 1, (1, 16, 32, 4)).astype(np.float32)
    logs = {}
    for prec in ("fp32", "bf16"):
        mlp_mod.set_seed(1)
        m = K.NeRF(precision=prec, pos_emb_xyz=6, pos_emb_dir=2)
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=16, image_width=32, ray_chunks=256,
                  white_background=True)
        assert m.precision == prec
        logs[prec] = [m.train_step((images, (o[None], d[None], t[None])), seed=5 + i) for i in range(3)]
    for a, b in zip(logs["fp32"], logs["bf16"]):
        for k in ("coarse_loss", "fine_loss"):
            assert abs(a[k] - b[k]) <= 2e-2 * abs(a[k]) + 1e-4, (k, a[k], b[k])
    assert logs["bf16"][2]["coarse_loss"] != logs["bf16"][0]["coarse_loss"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fused_chunks_take_the_same_step(precision):
    """compile(fuse_chunks=...): the ray chunks of a training step are independent and their gradients summed
    (nerf.py:351-421), so executing several per library call must give the same losses and the same updated weights
    (fp32 summation order apart) when the fine-sample draws are given explicitly."""
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    H = W = 16
    rng = np.random.default_rng(0)
    pose = K.pose_spherical(20.0, -30.0, 4.0)
    o, d, t = K.RaysGenerator(K.get_focal_from_fov(0.69, W), W, H, 2.0, 6.0, 64)(pose, seed=3)
    images = rng.uniform(0, 1, (1, H, W, 4)).astype(np.float32)
    u_f = rng.uniform(0, 1, (H * W, 128)).astype(np.float32)
    res = []
    for fuse in (None, "auto", 4):
        mlp_mod.set_seed(1)
        m = K.NeRF(precision=precision, scan_mode="sequential")
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=H, image_width=W, ray_chunks=32,
                  white_background=True, fuse_chunks=fuse)
        assert (m._train_ray_chunks, m._train_chunks) == {None: (32, 8), "auto": (256, 1), 4: (128, 2)}[fuse]
        logs = m.train_step((images, (o[None], d[None], t[None])), u_fine=u_f)
        res.append((logs, m.coarse.params.cpu().clone(), m.fine.params.cpu().clone()))
    tol = 1e-6 if precision == "fp32" else 2e-3        # bf16: tile boundaries move, so do single roundings
    for logs, pc, pf in res[1:]:
        for k in ("coarse_loss", "fine_loss"):
            assert logs[k] == pytest.approx(res[0][0][k], rel=10 * tol)
        # one Adam step moves every weight by about lr = 1e-3: compare the steps, not the weights
        assert float((pc - res[0][1]).abs().max()) <= (2e-5 if precision == "fp32" else 2e-3)
        assert float((pf - res[0][2]).abs().max()) <= (2e-5 if precision == "fp32" else 2e-3)


def test_custom_loss_is_refused_and_initializers_accepted():
    import keras_nerf_b200 as K
    m = K.NeRF(precision="fp32")
    with pytest.raises(NotImplementedError):
        m.compile(optimizer="adam", loss=lambda a, b: abs(a - b), batch_size=1, image_height=4, image_width=4,
                  ray_chunks=16)

    def compute_distributed_loss(y_true, y_pred):      # train.py:130-136 wraps MeanSquaredError under this name
        return ((y_true - y_pred) ** 2).mean()
    m.compile(optimizer="adam", loss=compute_distributed_loss, batch_size=1, image_height=4, image_width=4, ray_chunks=16)
    for init in ("he_normal", "glorot_normal", "lecun_uniform", "zeros"):
        net = K.NeRFMLP(n_layers=2, dense_units=16, skip_layer=4, initializer=init).build(15, 9)
        w = net.trainable_variables[0]
        assert bool(torch.isfinite(w).all()) and (init != "zeros" or float(w.abs().max()) == 0.0)
    with pytest.raises(NotImplementedError):
        K.NeRFMLP(initializer="orthogonal_typo")
