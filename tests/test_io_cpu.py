"""CPU tests of the rows either side of the hot path: the image-loader oracle, the dataset / callback /
GIF host logic.  No libknerf compute here."""
import csv
import json
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as IO


# ---- f3: oracle of tf.image.resize(antialias=True) ------------------------------------------------------------
@pytest.mark.parametrize("n_in,n_out", [(800, 400), (800, 128), (100, 128), (128, 128), (37, 11), (5, 1)])
def test_spans_are_normalised_and_in_range(n_in, n_out):
    starts, w = IO.compute_spans(n_out, n_in)
    assert starts.min() >= 0 and (starts + (w > 0).sum(1)).max() <= n_in
    np.testing.assert_allclose(w.sum(1), 1.0, atol=2e-7)
    assert (w >= 0).all()
    if n_in == n_out:                                      # identity: one tap of weight 1 on the same pixel
        assert np.array_equal(starts + np.argmax(w, axis=1), np.arange(n_out)) and np.all(w.max(1) == 1.0)


def test_resize_matches_pillows_antialiased_bilinear():
    """Independent implementation of the same filter (Pillow's BILINEAR scales its support when reducing): pins
    the span / weight structure of the restatement.  Pillow works per channel in float ('F' mode)."""
    from PIL import Image
    rng = np.random.default_rng(0)
    for (h, w), (oh, ow) in (((80, 80), (40, 40)), ((96, 64), (32, 48)), ((50, 50), (64, 64)), ((64, 64), (13, 13))):
        img = rng.random((h, w, 1), dtype=np.float32)
        ours = IO.resize_bilinear_antialias(img, oh, ow)[..., 0]
        pil = np.asarray(Image.fromarray(img[..., 0], mode="F").resize((ow, oh), Image.BILINEAR))
        assert np.abs(ours - pil).max() <= 2e-6


def test_image_loader_oracle_semantics():
    rng = np.random.default_rng(1)
    rgba = rng.integers(0, 256, (32, 32, 4), dtype=np.uint8)
    for white in (False, True):
        out = IO.image_loader(rgba, 32, 32, white)         # same size: resize is the identity
        x = rgba.astype(np.float32) * np.float32(1 / 255)
        a = x[..., 3:4]
        exp = a * x[..., :3] + (1 - a) * (1.0 if white else 0.0)
        np.testing.assert_allclose(out[..., :3], np.clip(exp, 0, 1), atol=1e-7)
        np.testing.assert_array_equal(out[..., 3:4], a)
    out = IO.image_loader(rgba, 16, 8, True)               # (sic) [image_width, image_height, 4]
    assert out.shape == (16, 8, 4) and out.min() >= 0 and out.max() <= 1
    const = np.full((40, 40, 4), 200, np.uint8)            # constants survive any resize
    np.testing.assert_allclose(IO.resize_bilinear_antialias(const.astype(np.float32), 7, 9), 200.0, rtol=1e-6)


# ---- f3: Dataset host logic (no GPU: stand-in element functions) ------------------------------------------------
def _fake_dataset(n, batch, **kw):
    from keras_nerf_b200.data.loader import Dataset
    calls = {"img": 0, "rays": 0}

    def img(path):
        calls["img"] += 1
        return torch.full((4, 4, 4), float(path))

    def rays(c2w):
        calls["rays"] += 1
        return (torch.full((4, 4, 3), float(c2w)), torch.zeros(4, 4, 3), torch.rand(4, 4, 8))

    return Dataset(list(range(n)), list(range(n)), img, rays, batch, seed=0, **kw), calls


def test_dataset_batches_shuffle_and_drop_remainder():
    ds, calls = _fake_dataset(11, 2)
    assert len(ds) == 5
    seen = []
    for images, (o, d, t) in ds:
        assert images.shape == (2, 4, 4, 4) and o.shape == (2, 4, 4, 3) and t.shape == (2, 4, 4, 8)
        assert torch.equal(images[:, 0, 0, 0], o[:, 0, 0, 0])           # image i stays zipped with pose i
        seen += images[:, 0, 0, 0].tolist()
    assert len(seen) == 10 and len(set(seen)) == 10                      # one of 11 dropped, none repeated
    # shuffle(buffer=batch_size): element k cannot be emitted before position k - (buffer - 1)
    order = ds._shuffled_indices()
    assert sorted(order) == list(range(11)) and all(pos >= i - 1 for pos, i in enumerate(order))
    e1 = [b[0][:, 0, 0, 0].tolist() for b in ds]
    e2 = [b[0][:, 0, 0, 0].tolist() for b in ds]
    assert e1 != e2                                                      # reshuffled each pass
    assert 10 <= calls["img"] <= 11 and calls["rays"] == 30                    # images cached, rays re-drawn every pass


def test_dataset_take_and_get_next():
    ds, _ = _fake_dataset(6, 2)
    assert len(list(ds.take(1))) == 1 and len(ds.take(1)) == 1
    it = iter(ds)
    for _ in range(3):
        it.get_next()
    with pytest.raises(IndexError):
        it.get_next()


def test_dataset_loader_reads_transforms(tmp_path):
    from keras_nerf_b200.data.loader import DatasetLoader
    from keras_nerf_b200.data.synthetic import write_nerf_synthetic_like
    d = write_nerf_synthetic_like(str(tmp_path / "scene"), image_wh=16, n_train=3, n_val=1, n_test=1)
    cfg = json.load(open(os.path.join(d, "transforms_train.json")))
    dl = DatasetLoader(d, white_background=True)
    paths, cams = dl._load_image_path_and_camera_param(dl._load_json(os.path.join(d, "transforms_train.json")))
    assert len(paths) == 3 and all(os.path.exists(p) and p.endswith(".png") for p in paths)
    assert np.asarray(cams).shape == (3, 4, 4) and abs(cfg["camera_angle_x"] - 0.6911112070083618) < 1e-12
    from keras_nerf_b200.data.image import decode_image_rgba
    rgba = decode_image_rgba(paths[0])
    assert rgba.shape == (16, 16, 4) and rgba.dtype == np.uint8
    assert decode_image_rgba(open(paths[0], "rb").read()).tobytes() == rgba.tobytes()
    with pytest.raises(FileNotFoundError):
        decode_image_rgba(str(tmp_path / "missing.png"))


# ---- f1: NeRFTrainMonitor host logic with a stand-in model ----------------------------------------------------
class _FakeModel:
    def __init__(self):
        self.saved = []

    def predict_and_render_images(self, rays):
        o = rays[0]
        B, H, W = o.shape[:3]
        res = {"image": torch.rand(B, H, W, 3), "depth": torch.rand(B, H, W) * 4 + 2}
        return res, res

    def save_model(self, path, weights_only=False):
        self.saved.append((path, weights_only))


def test_train_monitor_files_csv_and_resume(tmp_path):
    from keras_nerf_b200.model.nerf.callback import NeRFTrainMonitor, depth_to_color
    ds, _ = _fake_dataset(8, 1)
    log_dir = str(tmp_path / "logs" / "lego")
    mon = NeRFTrainMonitor(ds, log_dir, batch_size=1, update_freq=2)
    assert mon.last_epoch == 0 and os.path.isdir(os.path.join(log_dir, "model"))
    model = _FakeModel()
    mon.set_model(model)
    for epoch in range(4):
        mon.on_epoch_end(epoch, {"coarse_loss": 0.1 / (epoch + 1), "fine_loss": 0.05 / (epoch + 1),
                                 "val_coarse_loss": 0.2, "val_fine_loss": 0.1})
    files = set(os.listdir(log_dir))
    assert {"log.csv", "model", "test_0_0.png", "test_sample_0_0.png", "test_0_2.png", "test_sample_0_2.png"} <= files
    assert "test_0_1.png" not in files                                   # update_freq
    rows = list(csv.DictReader(open(os.path.join(log_dir, "log.csv"))))
    assert [int(r["epoch"]) for r in rows] == [0, 2] and float(rows[1]["coarse_loss"]) == pytest.approx(0.1 / 3)
    assert model.saved == [(os.path.join(log_dir, "model"), False), (os.path.join(log_dir, "model"), True)]
    mon2 = NeRFTrainMonitor(ds, log_dir, batch_size=1, update_freq=2)    # resume = last logged epoch + 1
    assert mon2.last_epoch == 3 and mon2.coarse_log_list == [pytest.approx(0.1 / 3)]
    rgb = depth_to_color(torch.linspace(2, 6, 64).reshape(8, 8))
    assert rgb.shape == (8, 8, 3) and rgb[0, 0].sum() < 30 and rgb[-1, -1].sum() > 500   # inferno: black -> pale yellow
    import cv2
    fig = cv2.imread(os.path.join(log_dir, "test_0_0.png"))
    assert fig.shape[1] == 5 * 264 and fig.shape[0] == 284 + 300


# ---- f4: GIF writer ---------------------------------------------------------------------------------------------
def test_mimwrite_gif_round_trip(tmp_path):
    from PIL import Image
    from keras_nerf_b200.utils.video import frames_to_uint8, mimwrite
    frames = [np.full((16, 16, 3), v, np.float32) for v in (0.0, 0.25, 0.5, 1.0)]
    assert [int(f[0, 0, 0]) for f in frames_to_uint8(frames)] == [0, 64, 127, 255]
    path = str(tmp_path / "orbit.gif")
    mimwrite(path, frames, fps=20)
    with Image.open(path) as im:
        assert im.n_frames == 4 and im.size == (16, 16) and im.info["duration"] == 50
        im.seek(3)
        assert np.asarray(im.convert("RGB"))[0, 0].tolist() == [255, 255, 255]


# ---- f1: oracle of tf.image.ssim ----------------------------------------------------------------------------------
def test_ssim_oracle_properties_and_golden():
    rng = np.random.default_rng(5)
    a = rng.random((2, 24, 20, 3), dtype=np.float32)
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    assert np.allclose(IO.ssim(a, a), 1.0, atol=1e-6)
    s_ab = IO.ssim(a, b)
    assert np.allclose(s_ab, IO.ssim(b, a), atol=1e-6) and (s_ab < 0.99).all() and (s_ab > 0).all()
    assert IO.ssim(a[:, :11, :11], b[:, :11, :11]).shape == (2,)           # exactly one window position
    assert np.allclose(IO.psnr(a, b), -10 * np.log10(((a - b) ** 2).reshape(2, -1).mean(1)), atol=1e-4)
    # the reference's own metric code (nerf.py:306-330) executed over the TF stand-in: golden s0_*_ssim / psnr
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "tfshim"))
    try:
        import tensorflow as tf                                             # the torch stand-in, not TensorFlow
        ta, tb = torch.from_numpy(a), torch.from_numpy(b)
        assert np.allclose(np.asarray(tf.image.ssim(ta, tb, max_val=1.0)), s_ab, atol=2e-6)
        assert np.allclose(np.asarray(tf.image.psnr(ta, tb, max_val=1.0)), IO.psnr(a, b), atol=1e-4)
    finally:
        sys.path.pop(0)


# ---- train.py: per-replica slices of the global batch (MirroredStrategy semantics, train.py:75-87) ------------
def test_replica_batches_split_the_global_batch():
    import train

    class FakeStrategy:
        def __init__(self, rank):
            self.rank, self.num_replicas_in_sync = rank, 2

    per_rank = []
    for rank in range(2):
        ds, _ = _fake_dataset(9, 4)                        # global batch 4 = 2 replicas x batch_size 2
        ds._rng.seed(42)                                   # what train.py does so that all ranks agree on the order
        rb = train.ReplicaBatches(ds, FakeStrategy(rank), per_replica=2)
        assert len(rb) == 2 and len(rb.take(1)) == 1 and iter(rb).get_next()[0].shape[0] == 2
        per_rank.append([(im[:, 0, 0, 0].tolist(), o[:, 0, 0, 0].tolist()) for im, (o, d, t) in rb])
    for b0, b1 in zip(*per_rank):
        assert len(b0[0]) == len(b1[0]) == 2 and not set(b0[0]) & set(b1[0])   # disjoint halves of one batch
        assert b0[0] == b0[1] and b1[0] == b1[1]
    seen = [x for rank in per_rank for b in rank for x in b[0]]
    assert len(seen) == 8 and len(set(seen)) == 8


# ---- the DLPack side of the shim ------------------------------------------------------------------------------
class _Foreign:
    """a tensor of 'another framework': exposes only the DLPack protocol (what tf.Tensor / cupy / jax arrays do)"""

    def __init__(self, t):
        self._t = t

    def __dlpack__(self, stream=None, **kw):
        return self._t.__dlpack__()

    def __dlpack_device__(self):
        return self._t.__dlpack_device__()


def test_dev_accepts_dlpack_producers():
    from keras_nerf_b200 import _lib
    src = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    out = _lib.dev(_Foreign(src), torch.device("cpu"))
    assert torch.equal(out, src) and out.data_ptr() == src.data_ptr()      # zero-copy on the same device
    assert torch.equal(_lib.dev(np.arange(4.0), torch.device("cpu")), torch.arange(4, dtype=torch.float32))
