"""NeRFUtils (reference: keras_nerf/model/nerf/utils.py) -- same method names and argument meaning, every
method a call into libknerf (CUDA).  Inputs may be numpy or torch (any device); outputs are CUDA tensors."""
from __future__ import annotations

import itertools

import torch

from ... import _lib

_seed_counter = _lib.seed_stream(0xF17E0000)


class NeRFUtils:
    def __init__(self, batch_size, image_height, image_width, ray_chunks, pos_emb_xyz, pos_emb_dir,
                 white_background=False, device=None, oob_mode="zero", scan_mode="sequential"):
        # keras_nerf/model/nerf/utils.py:5-14
        self.batch_size = batch_size
        self.image_height = image_height
        self.image_width = image_width
        self.ray_chunks = ray_chunks
        self.pos_emb_xyz = pos_emb_xyz
        self.pos_emb_dir = pos_emb_dir
        self.num_rays = self.batch_size * self.image_height * self.image_width
        self.sequential_chunks = self.num_rays // ray_chunks
        self.white_background = white_background
        self.device = torch.device(device) if device is not None else None
        self.oob_mode = oob_mode
        # "sequential": pdf normaliser / cdf summed left to right (TF-CPU order, bit-identical to the oracle);
        # "warp": warp-shuffle scan (the throughput default of the bf16 mode)
        self.scan_mode = scan_mode

    def _dev(self):
        return self.device or _lib.default_device()

    # ---- compositing --------------------------------------------------------------------------
    def _render(self, rgb, sigma, sample_points, epsilon, white, clip):
        device = self._dev()
        rgb, sigma, t = _lib.dev(rgb, device), _lib.dev(sigma, device), _lib.dev(sample_points, device)
        lead, S = t.shape[:-1], t.shape[-1]
        R = t.numel() // S
        assert rgb.shape == lead + (S, 3) and sigma.numel() == R * S, (rgb.shape, sigma.shape, t.shape)
        image = torch.empty(lead + (3,), dtype=torch.float32, device=device)
        depth = torch.empty(lead, dtype=torch.float32, device=device)
        weights = torch.empty(lead + (S,), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.call("knerf_composite_forward", None, _lib.ptr(rgb), _lib.ptr(sigma), _lib.ptr(t), R, S,
                      int(bool(white)), int(bool(clip)), float(epsilon), _lib.ptr(image), _lib.ptr(depth),
                      _lib.ptr(weights), None, _lib.stream())
        return image, depth, weights

    def render_image_depth_chunk(self, rgb, sigma, sample_points, epsilon=1e-10):
        """keras_nerf/model/nerf/utils.py:16-58: rgb[R,S,3], sigma[R,S,1], sample_points[R,S]."""
        return self._render(rgb, sigma, sample_points, epsilon, self.white_background, True)

    def render_image_depth(self, rgb, sigma, sample_points, epsilon=1e-10):
        """keras_nerf/model/nerf/utils.py:99-134 (test-only full-image form: no white term, no clip)."""
        return self._render(rgb, sigma, sample_points, epsilon, False, False)

    # ---- hierarchical sampling ----------------------------------------------------------------
    def _sample(self, mid_points, weights, n_samples, u, seed, cdf=None, return_aux=False):
        device = self._dev()
        mid, w = _lib.dev(mid_points, device), _lib.dev(weights, device)
        lead, Nc = w.shape[:-1], w.shape[-1]
        assert mid.shape == lead + (Nc - 1,), (mid.shape, w.shape)
        R = w.numel() // Nc
        n_samples = int(n_samples)
        samples = torch.empty(lead + (n_samples,), dtype=torch.float32, device=device)
        if u is not None:
            u = _lib.dev(u, device).reshape(lead + (n_samples,))
        if cdf is not None:
            cdf = _lib.dev(cdf, device).reshape(lead + (Nc + 1,))
        idx = cdf_out = None
        if return_aux:
            idx = torch.empty(lead + (n_samples,), dtype=torch.int32, device=device)
            cdf_out = torch.empty(lead + (Nc + 1,), dtype=torch.float32, device=device)
        mode = _lib.OOB_MODES[self.oob_mode]
        flags = mode | (_lib.SCAN_SEQUENTIAL if self.scan_mode == "sequential" else 0)
        count = torch.zeros(1, dtype=torch.int32, device=device) if mode == _lib.OOB_COUNT else None
        if seed is None:
            seed = next(_seed_counter)
        with torch.cuda.device(device):
            _lib.call("knerf_sample_fine", None, _lib.ptr(mid), _lib.ptr(w), _lib.ptr(u), int(seed), _lib.ptr(cdf),
                      R, Nc, n_samples, flags, None, _lib.ptr(samples), _lib.ptr(idx, torch.int32),
                      _lib.ptr(cdf_out), _lib.ptr(count, torch.int32), _lib.stream())
        if count is not None and int(count.item()) > 0:
            # TF's CPU gather kernel raises InvalidArgumentError here
            raise IndexError(f"fine sampling: {int(count.item())} rays gather mid_points out of range "
                             f"(size {Nc - 1}); this is what tf.gather does on CPU")
        if return_aux:
            return samples, idx, cdf_out
        return samples

    def fine_hierarchical_sampling_chunk(self, mid_points, weights, n_samples, u=None, seed=None, **kw):
        """keras_nerf/model/nerf/utils.py:60-97.  `u` [R,n_samples] replaces the internal tf.random.uniform."""
        return self._sample(mid_points, weights, n_samples, u, seed, **kw)

    def fine_hierarchical_sampling(self, mid_points, weights, n_samples, u=None, seed=None, **kw):
        """keras_nerf/model/nerf/utils.py:136-174 (full-image form, same arithmetic)."""
        return self._sample(mid_points, weights, n_samples, u, seed, **kw)

    # ---- positional encoding ------------------------------------------------------------------
    def positional_encoding(self, inputs, pos_embedding_dim):
        """keras_nerf/model/nerf/utils.py:176-186"""
        device = self._dev()
        x = _lib.dev(inputs, device)
        dim = x.shape[-1]
        width = dim * (1 + 2 * int(pos_embedding_dim))
        out = torch.empty(x.shape[:-1] + (width,), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.call("knerf_positional_encoding", _lib.ptr(x), x.numel() // dim, dim, int(pos_embedding_dim),
                      _lib.ptr(out), width, _lib.stream())
        return out

    def encode_position_and_directions(self, ray_origin, ray_direction, coarse_points):
        """keras_nerf/model/nerf/utils.py:188-210"""
        device = self._dev()
        o, d, t = _lib.dev(ray_origin, device), _lib.dev(ray_direction, device), _lib.dev(coarse_points, device)
        lead, S = t.shape[:-1], t.shape[-1]
        R = t.numel() // S
        wx, wd = 3 + 6 * self.pos_emb_xyz, 3 + 6 * self.pos_emb_dir
        xyz = torch.empty(lead + (S, wx), dtype=torch.float32, device=device)
        dirs = torch.empty(lead + (S, wd), dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            _lib.call("knerf_encode_position_and_directions", _lib.ptr(o), _lib.ptr(d), _lib.ptr(t), R, S,
                      self.pos_emb_xyz, self.pos_emb_dir, _lib.ptr(xyz), wx, _lib.ptr(dirs), wd, _lib.stream())
        return xyz, dirs
