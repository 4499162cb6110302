"""NeRFTrainMonitor (reference: keras_nerf/model/nerf/callback.py) -- same constructor, hooks, files and CSV
layout; the figures are composed with OpenCV (matplotlib is not a dependency here): the same five panels
(coarse image / coarse depth / fine image / fine depth / ground truth, depth through the 'inferno' colour map with
imshow's min-max normalisation) and the same log-scale loss plot underneath."""
from __future__ import annotations

import logging
import math
import os
from csv import DictReader, DictWriter

import numpy as np
import torch

_TILE = 256


def _to_np(x):
    return x.detach().float().cpu().numpy() if torch.is_tensor(x) else np.asarray(x, dtype=np.float32)


def depth_to_color(depth) -> np.ndarray:
    """plt.imshow(depth, cmap='inferno') (callback.py:84-94): min-max normalise, 256-entry inferno LUT -> RGB uint8."""
    import cv2
    d = _to_np(depth)
    lo, hi = float(np.nanmin(d)), float(np.nanmax(d))
    n = np.zeros_like(d) if hi <= lo else (d - lo) / (hi - lo)
    idx = np.clip(np.nan_to_num(n) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    return cv2.cvtColor(cv2.applyColorMap(idx, cv2.COLORMAP_INFERNO), cv2.COLOR_BGR2RGB)


def image_to_uint8(image) -> np.ndarray:
    """plt.imshow of a float RGB image: clip to [0,1], scale to 8 bits."""
    return (np.clip(_to_np(image)[..., :3], 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def _tile(rgb_u8, title):
    import cv2
    t = cv2.resize(rgb_u8, (_TILE, _TILE), interpolation=cv2.INTER_NEAREST)
    canvas = np.full((_TILE + 28, _TILE + 8, 3), 255, np.uint8)
    canvas[24:24 + _TILE, 4:4 + _TILE] = t
    cv2.putText(canvas, title, (6, 17), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    return canvas


def loss_plot(curves, title, width, height=300) -> np.ndarray:
    """ax.plot(...); ax.set_yscale('log') (callback.py:96-106,150-163).  curves: [(values, rgb, dashed, label)]."""
    import cv2
    img = np.full((height, width, 3), 255, np.uint8)
    vals = [v for c in curves for v in c[0] if v is not None and v > 0 and math.isfinite(v)]
    cv2.putText(img, title, (width // 2 - 4 * len(title), 16), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (0, 0, 0), 1, cv2.LINE_AA)
    x0, x1, y0, y1 = 60, width - 10, 24, height - 20
    cv2.rectangle(img, (x0, y0), (x1, y1), (0, 0, 0), 1)
    if vals:
        lo, hi = math.log10(min(vals)), math.log10(max(vals))
        if hi - lo < 1e-6:
            lo, hi = lo - 0.5, hi + 0.5
        n = max(max(len(c[0]) for c in curves), 2)
        for e in range(math.floor(lo), math.ceil(hi) + 1):   # decade grid lines
            if lo <= e <= hi:
                y = int(y1 - (e - lo) / (hi - lo) * (y1 - y0))
                cv2.line(img, (x0, y), (x1, y), (220, 220, 220), 1)
                cv2.putText(img, f"1e{e}", (8, y + 4), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (0, 0, 0), 1, cv2.LINE_AA)
        for k, (values, rgb, dashed, label) in enumerate(curves):
            pts = [(int(x0 + i / (n - 1) * (x1 - x0)), int(y1 - (math.log10(v) - lo) / (hi - lo) * (y1 - y0)))
                   for i, v in enumerate(values) if v is not None and v > 0 and math.isfinite(v)]
            for i in range(len(pts) - 1):
                if not dashed or i % 2 == 0:
                    cv2.line(img, pts[i], pts[i + 1], rgb, 2, cv2.LINE_AA)
            if len(pts) == 1:
                cv2.circle(img, pts[0], 2, rgb, -1)
            cv2.line(img, (x1 - 190, y0 + 14 + 16 * k), (x1 - 165, y0 + 14 + 16 * k), rgb, 2)
            cv2.putText(img, label, (x1 - 160, y0 + 18 + 16 * k), cv2.FONT_HERSHEY_SIMPLEX, 0.4, (0, 0, 0), 1, cv2.LINE_AA)
    return img


def save_figure(path, coarse_image, coarse_depth, fine_image, fine_depth, ground_truth, curves=None, title=''):
    import cv2
    row = np.concatenate([_tile(image_to_uint8(coarse_image), 'Coarse Image'),
                          _tile(depth_to_color(coarse_depth), 'Coarse Depth'),
                          _tile(image_to_uint8(fine_image), 'Fine Image'),
                          _tile(depth_to_color(fine_depth), 'Fine Depth'),
                          _tile(image_to_uint8(ground_truth), 'Ground Truth')], axis=1)
    if curves is not None:
        row = np.concatenate([row, loss_plot(curves, title, row.shape[1])], axis=0)
    cv2.imwrite(path, cv2.cvtColor(row, cv2.COLOR_RGB2BGR))


_BLUE, _ORANGE = (31, 119, 180), (255, 127, 14)


class NeRFTrainMonitor:
    """callback.py:8-226.  Duck-typed Keras callback: NeRF.fit calls set_model / on_train_batch_end / on_epoch_end."""

    def __init__(self, dataset, log_dir: str, batch_size: int, update_freq: int = 1, verbose: bool = False, **kwargs):
        logging.info('Initializing NeRFTrainMonitor')
        logging.info(f'Log Directory: {log_dir}, Batch Size: {batch_size}, Update Frequency: {update_freq}')
        self.model = None
        self.dataset, self.log_dir, self.batch_size = dataset, log_dir, batch_size
        self.update_freq, self.verbose = update_freq, verbose
        self.log_model_dir = os.path.join(log_dir, 'model')
        os.makedirs(self.log_model_dir, exist_ok=True)
        self.coarse_log_list, self.val_coarse_log_list = [], []
        self.fine_log_list, self.val_fine_log_list = [], []
        if self.verbose:
            self.coarse_log_list_batch, self.fine_log_list_batch = [], []
        # resume: the last `epoch` of log.csv + 1 (callback.py:33-47; its `i > 0` skips the first data row too)
        self.last_epoch = 0
        self.log_csv = os.path.join(log_dir, 'log.csv')
        if os.path.exists(self.log_csv):
            with open(self.log_csv, 'r') as f:
                for i, row in enumerate(DictReader(f)):
                    if i > 0:
                        self.coarse_log_list.append(float(row['coarse_loss']))
                        self.val_coarse_log_list.append(float(row['val_coarse_loss']))
                        self.fine_log_list.append(float(row['fine_loss']))
                        self.val_fine_log_list.append(float(row['val_fine_loss']))
                        self.last_epoch = int(row['epoch'])
            self.last_epoch += 1
        os.makedirs(self.log_dir, exist_ok=True)
        for inputs in self.dataset.take(1):                  # callback.py:51-55
            self.images, self.rays = inputs
            ray_origin, ray_direction, coarse_points = self.rays
            self.ray_origin, self.ray_direction, self.coarse_points = (
                ray_origin[:self.batch_size], ray_direction[:self.batch_size], coarse_points[:self.batch_size])
        self.dataset_iterator = iter(self.dataset)
        self.dataset_iterator.get_next()

    def set_model(self, model):
        self.model = model

    def _render_fixed_view(self):
        coarse, fine = self.model.predict_and_render_images((self.ray_origin, self.ray_direction, self.coarse_points))
        return coarse['image'], coarse['depth'], fine['image'], fine['depth']

    def on_train_batch_end(self, batch, logs=None):
        if not self.verbose:
            return
        logging.debug(f'Batch {batch}: {logs}')
        self.coarse_log_list_batch.append(logs['coarse_loss'])
        self.fine_log_list_batch.append(logs['fine_loss'])
        ci, cd, fi, fd = self._render_fixed_view()
        curves = [(self.coarse_log_list_batch, _BLUE, False, 'Coarse Train Loss'),
                  (self.fine_log_list_batch, _ORANGE, False, 'Fine Train Loss')]
        for i in range(self.batch_size):
            save_figure(os.path.join(self.log_dir, f'debug_{i}_{batch}.png'), ci[i], cd[i], fi[i], fd[i],
                        self.images[i, ..., :3], curves, f'Loss Batch Plot: {batch}')

    def on_epoch_end(self, epoch, logs):
        self.coarse_log_list.append(logs['coarse_loss'])
        self.val_coarse_log_list.append(logs['val_coarse_loss'])
        self.fine_log_list.append(logs['fine_loss'])
        self.val_fine_log_list.append(logs['val_fine_loss'])
        if epoch % self.update_freq == 0:
            ci, cd, fi, fd = self._render_fixed_view()
            curves = [(self.coarse_log_list, _BLUE, False, 'Coarse Train Loss'),
                      (self.val_coarse_log_list, _BLUE, True, 'Coarse Val Loss'),
                      (self.fine_log_list, _ORANGE, False, 'Fine Train Loss'),
                      (self.val_fine_log_list, _ORANGE, True, 'Fine Val Loss')]
            for i in range(self.batch_size):
                save_figure(os.path.join(self.log_dir, f'test_{i}_{epoch}.png'), ci[i], cd[i], fi[i], fd[i],
                            self.images[i, ..., :3], curves, f'Loss Plot: {epoch}')
            # another view from the iterator (callback.py:176-214)
            try:
                images, rays = self.dataset_iterator.get_next()
            except IndexError:
                # the reference's iterator raises OutOfRangeError once the test split is used up (more logging epochs
                # than test batches); here it starts over
                self.dataset_iterator = iter(self.dataset)
                images, rays = self.dataset_iterator.get_next()
            images = images[..., :3]
            o, d, t = (r[:self.batch_size] for r in rays)
            coarse, fine = self.model.predict_and_render_images((o, d, t))
            for i in range(self.batch_size):
                save_figure(os.path.join(self.log_dir, f'test_sample_{i}_{epoch}.png'), coarse['image'][i],
                            coarse['depth'][i], fine['image'][i], fine['depth'][i], images[i])
            with open(self.log_csv, 'a') as f:               # callback.py:216-222
                new_logs = {'epoch': epoch}
                new_logs.update({k: float(v) for k, v in logs.items()})
                writer = DictWriter(f, new_logs.keys())
                if epoch == 0:
                    writer.writeheader()
                writer.writerow(new_logs)
            self.model.save_model(self.log_model_dir, weights_only=(epoch != 0))
        if self.verbose:
            self.coarse_log_list_batch, self.fine_log_list_batch = [], []
