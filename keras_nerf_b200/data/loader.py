"""DatasetLoader (reference: keras_nerf/data/loader.py) -- nerf_synthetic-style directories
(`transforms_{train,val,test}.json` + PNG frames) -> three batched datasets of (images, rays)."""
from __future__ import annotations

import json
import logging
import os
import random

import torch

from .image import ImageLoader
from .rays import RaysGenerator
from .utils import get_focal_from_fov


class _Iterator:
    """What `iter(tf.data.Dataset)` gives the callback (callback.py:56-57, 177): next() and get_next()."""

    def __init__(self, gen):
        self._gen = gen

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._gen)

    def get_next(self):
        try:
            return next(self._gen)
        except StopIteration:
            raise IndexError("End of sequence") from None   # tf.errors.OutOfRangeError


class Dataset:
    """The slice of tf.data.Dataset the reference uses on `zip((images, rays)).shuffle(B).batch(B,
    drop_remainder=True).prefetch(AUTOTUNE)` (loader.py:99-107): re-iterable, reshuffled on every pass with a
    buffer of `shuffle_buffer` elements, `take(n)`, `len()`.

    Every pass re-runs the rays map, so the stratified jitter is fresh each epoch as in the reference (its
    `map(rays_generator)` is not cached).  Prepared images ARE cached on the GPU after the first pass (decode +
    resize are deterministic), which is what `prefetch` hides in the reference."""

    def __init__(self, image_paths, camera_params, image_loader, rays_generator, batch_size, shuffle_buffer=None,
                 limit=None, seed=None):
        self.image_paths, self.camera_params = list(image_paths), list(camera_params)
        self.image_loader, self.rays_generator = image_loader, rays_generator
        self.batch_size = int(batch_size)
        self.shuffle_buffer = self.batch_size if shuffle_buffer is None else int(shuffle_buffer)
        self.limit = limit
        self._rng = random.Random(seed)
        self._cache = {}

    def _element(self, i):
        if i not in self._cache:
            self._cache[i] = self.image_loader(self.image_paths[i])
        return self._cache[i], self.rays_generator(self.camera_params[i])

    def _shuffled_indices(self):
        # tf.data shuffle(buffer): fill a buffer, emit a uniformly chosen slot, refill it from the stream
        buf, out = [], []
        for i in range(len(self.image_paths)):
            buf.append(i)
            if len(buf) > max(self.shuffle_buffer, 1) - 1:
                out.append(buf.pop(self._rng.randrange(len(buf))))
        while buf:
            out.append(buf.pop(self._rng.randrange(len(buf))))
        return out

    def _batches(self):
        order, B = self._shuffled_indices(), self.batch_size
        n = len(order) // B                                # drop_remainder=True
        if self.limit is not None:
            n = min(n, self.limit)
        for b in range(n):
            elems = [self._element(i) for i in order[b * B:(b + 1) * B]]
            images = torch.stack([e[0] for e in elems])
            rays = tuple(torch.stack([e[1][k] for e in elems]) for k in range(3))
            yield images, rays

    def __iter__(self):
        return _Iterator(self._batches())

    def __len__(self):
        n = len(self.image_paths) // self.batch_size
        return n if self.limit is None else min(n, self.limit)

    def take(self, count):
        return Dataset(self.image_paths, self.camera_params, self.image_loader, self.rays_generator,
                       self.batch_size, self.shuffle_buffer, limit=count, seed=self._rng.random())


class DatasetLoader:
    """Same constructor and methods as keras_nerf/data/loader.py:12-113."""

    def __init__(self, data_dir: str, white_background: bool = False, device=None, **kwargs):
        self.data_dir = data_dir
        self.white_background = white_background
        self.device = device

    def _load_json(self, filename: str) -> dict:
        with open(filename, 'r') as f:
            return json.load(f)

    def _load_image_path_and_camera_param(self, json_config: dict) -> tuple:
        image_paths, camera_params = [], []
        for frame in json_config['frames']:                 # loader.py:47-51
            image_paths.append(os.path.join(self.data_dir, f"{frame['file_path']}.png"))
            camera_params.append(frame['transform_matrix'])
        return image_paths, camera_params

    def load_dataset(self, batch_size: int, image_width: int, image_height: int, near: float, far: float,
                     n_sample: int) -> list:
        image_loader = ImageLoader(image_width, image_height, self.white_background, device=self.device)
        datasets = []
        for subset in ['train', 'val', 'test']:
            json_config = self._load_json(os.path.join(self.data_dir, f"transforms_{subset}.json"))
            focal_length = get_focal_from_fov(json_config['camera_angle_x'], image_width)
            rays_generator = RaysGenerator(focal_length=focal_length, image_width=image_width,
                                           image_height=image_height, near=near, far=far, n_sample=n_sample,
                                           device=self.device)
            image_paths, camera_params = self._load_image_path_and_camera_param(json_config)
            datasets.append(Dataset(image_paths, camera_params, image_loader, rays_generator, batch_size))
            logging.info(f"Loaded {subset} dataset. {len(image_paths)} images.")
        return datasets
