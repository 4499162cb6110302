"""GPU (-m gpu): the rows either side of the hot path -- f3 ImageLoader kernel vs its oracle and
DatasetLoader, f1 fit + NeRFTrainMonitor, f4 orbit GIF, and the train.py / inference.py command lines
(keras_nerf/data/image.py:17-35, loader.py:55-113, model/nerf/callback.py, train_single.py, inference.py)."""
import csv
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as IO

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("in_hw,out_wh,white", [((800, 800), (400, 400), True), ((800, 800), (128, 128), False),
                                                ((100, 100), (128, 128), True), ((64, 64), (64, 64), False),
                                                ((97, 61), (40, 23), True), ((16, 300), (5, 7), False)])
def test_image_prepare_kernel_matches_oracle(in_hw, out_wh, white):
    """uint8 RGBA -> [image_width, image_height, 4] float32: same fp32 operations in the same order as the oracle."""
    from keras_nerf_b200 import ImageLoader
    rng = np.random.default_rng(in_hw[0] * 7 + out_wh[0])
    rgba = rng.integers(0, 256, in_hw + (4,), dtype=np.uint8)
    rgba[: in_hw[0] // 3, :, 3] = 0                         # transparent and opaque bands like a rendered object
    rgba[in_hw[0] // 3: in_hw[0] // 2, :, 3] = 255
    out = ImageLoader(out_wh[0], out_wh[1], white)(rgba)
    exp = IO.image_loader(rgba, out_wh[0], out_wh[1], white)
    assert out.shape == exp.shape == (out_wh[0], out_wh[1], 4) and out.dtype == torch.float32
    assert np.abs(out.cpu().numpy() - exp).max() <= 1e-6    # tolerance of the f3 row (observed: 0)
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0


def test_image_loader_paths_rgb_png_and_errors(tmp_path):
    from PIL import Image
    from keras_nerf_b200 import ImageLoader
    from keras_nerf_b200._lib import KnerfError
    rng = np.random.default_rng(3)
    rgb = rng.integers(0, 256, (50, 50, 3), dtype=np.uint8)
    p = str(tmp_path / "rgb.png")
    Image.fromarray(rgb, mode="RGB").save(p)
    out = ImageLoader(25, 25, white_background=True)(p)     # decode_image(channels=4): opaque alpha
    exp = IO.image_loader(np.concatenate([rgb, np.full((50, 50, 1), 255, np.uint8)], -1), 25, 25, True)
    assert np.abs(out.cpu().numpy() - exp).max() <= 1e-6 and float(out[..., 3].min()) == 1.0
    with pytest.raises(FileNotFoundError):
        ImageLoader(8, 8)(str(tmp_path / "nope.png"))
    with pytest.raises(ValueError):
        ImageLoader(8, 8)(np.zeros((4, 4), np.uint8))
    from keras_nerf_b200 import _lib
    with pytest.raises(KnerfError):                          # C-ABI argument check
        _lib.call("knerf_image_prepare", None, 4, 4, 2, 2, 0, None, _lib.stream())


def _scene(tmp_path, wh=32, **kw):
    from keras_nerf_b200.data.synthetic import write_nerf_synthetic_like
    return write_nerf_synthetic_like(str(tmp_path / "scene"), image_wh=wh, **kw)


def test_dataset_loader_shapes_like_reference_test(tmp_path):
    """tests/data/test_loader.py:13-49 on a synthetic directory: three datasets, batch shapes [B,H,W,.]"""
    from keras_nerf_b200 import DatasetLoader
    d = _scene(tmp_path, wh=40, n_train=5, n_val=2, n_test=2)
    train, val, test = DatasetLoader(d, white_background=True).load_dataset(2, 20, 20, 2.0, 6.0, 32)
    assert (len(train), len(val), len(test)) == (2, 1, 1)
    for images, (o, dr, t) in train:
        assert images.shape == (2, 20, 20, 4) and images.dtype == torch.float32 and images.is_cuda
        assert o.shape == (2, 20, 20, 3) and dr.shape == (2, 20, 20, 3) and t.shape == (2, 20, 20, 32)
        assert float(images.min()) >= 0 and float(images.max()) <= 1
        assert float(t.min()) >= 2.0 and float(t.max()) <= 6.0
        corner = images[:, 0, 0]                              # background corner: white, alpha 0
        assert torch.allclose(corner[:, :3], torch.ones_like(corner[:, :3])) and float(corner[:, 3].max()) == 0.0
    t1 = next(iter(train))[1][2]
    t2 = next(iter(train))[1][2]
    assert not torch.equal(t1, t2)                            # fresh stratified jitter on every pass


def test_train_script_monitor_resume_and_inference_gif(tmp_path):
    """train_single.py command line end to end (2 epochs), resume from log.csv, then inference.py -> GIF."""
    from PIL import Image
    import inference
    import train
    d = _scene(tmp_path, wh=32, n_train=4, n_val=2, n_test=3)
    logs, models = str(tmp_path / "logs"), str(tmp_path / "model")
    argv = ["--name", "ball", "--data_dir", d, "--img_wh", "16", "--white_bg", "--batch_size", "1", "--ray_chunks",
            "128", "--num_coarse_samples", "16", "--num_fine_samples", "32", "--log_dir", logs, "--model_dirs", models,
            "--log_freq", "1", "--precision", "fp32"]
    nerf = train.main(argv + ["--num_epochs", "2"], multi_gpu=False)
    run = os.path.join(logs, "ball")
    rows = list(csv.DictReader(open(os.path.join(run, "log.csv"))))
    assert [int(r["epoch"]) for r in rows] == [0, 1]
    assert {"coarse_loss", "fine_loss", "val_coarse_loss", "val_fine_loss", "fine_psnr", "val_fine_ssim"} <= set(rows[0])
    assert all(np.isfinite(float(r["fine_loss"])) for r in rows)
    for f in ("test_0_0.png", "test_sample_0_1.png", "model/model_config.json"):
        assert os.path.exists(os.path.join(run, f)), f
    assert nerf.has_checkpoint(os.path.join(run, "model")) and nerf.has_checkpoint(os.path.join(models, "ball"))
    # resume: the monitor reports epoch 2, the logged model is loaded, one more epoch is appended
    nerf2 = train.main(argv + ["--num_epochs", "3"], multi_gpu=False)
    rows = list(csv.DictReader(open(os.path.join(run, "log.csv"))))
    assert [int(r["epoch"]) for r in rows] == [0, 1, 2]
    assert nerf2.model_path == os.path.join(run, "model")    # train_single.py:89-97
    out = inference.main(["--model_dirs", os.path.join(models, "ball"), "--img_wh", "16", "--ray_chunks", "256",
                          "--white_bg", "--output_freq", "90", "--output_dir", str(tmp_path / "out"),
                          "--precision", "fp32"])
    assert out.endswith("ball.gif")
    with Image.open(out) as im:
        assert 1 <= im.n_frames <= 4 and im.size == (16, 16)   # Pillow merges identical consecutive frames
    with pytest.raises(FileNotFoundError):
        inference.main(["--model_dirs", str(tmp_path / "empty")])


@pytest.mark.parametrize("shape", [(1, 128, 256, 3), (2, 16, 16, 3), (3, 37, 53, 3), (1, 11, 11, 1), (1, 400, 400, 3)])
def test_image_metrics_kernel_matches_oracle(shape):
    """per-image PSNR / SSIM of update_and_return_metrics (nerf.py:306-330) from knerf_image_metrics"""
    from keras_nerf_b200.model.nerf.nerf import image_metrics
    rng = np.random.default_rng(shape[1])
    a = rng.random(shape, dtype=np.float32)
    b = np.clip(a + 0.2 * rng.standard_normal(shape).astype(np.float32), 0, 1).astype(np.float32)
    b[0, : shape[1] // 2] = a[0, : shape[1] // 2]                          # a flat-error region and a noisy one
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = image_metrics(ta, tb).cpu().numpy()
    assert out.shape == (2, shape[0])
    assert np.abs(out[0] - IO.psnr(a, b)).max() <= 1e-4                    # dB
    assert np.abs(out[1] - IO.ssim(a, b)).max() <= 5e-6
    again = image_metrics(ta, tb).cpu().numpy()
    assert np.array_equal(out, again)                                      # fixed-order sums: bit-reproducible
    same = image_metrics(ta, ta).cpu().numpy()
    assert np.allclose(same[1], 1.0, atol=1e-6) and np.isinf(same[0]).all()
    small = image_metrics(ta[:, -8:, -8:], tb[:, -8:, -8:]).cpu().numpy()  # below the window: PSNR only
    assert np.isnan(small[1]).all() and np.abs(small[0] - IO.psnr(a[:, -8:, -8:], b[:, -8:, -8:])).max() <= 1e-4
