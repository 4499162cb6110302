"""360-degree orbit render -> GIF.  Same command line as the reference's inference.py (its flags at :14-40);
`--precision bf16` (default) selects the tcgen05 path, `--precision fp32` the parity mode.  Under torchrun the
frames are sharded across the GPUs and gathered on rank 0."""
import argparse
import logging
import os

import numpy as np
import torch

from keras_nerf_b200 import NeRF, RaysGenerator, get_focal_from_fov, pose_spherical
from keras_nerf_b200.utils.video import mimwrite


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--name', type=str, default='', help='Name of the nerf model')
    parser.add_argument('--model_dirs', type=str, required=True)
    parser.add_argument('--ray_chunks', type=int, default=4096)
    parser.add_argument('--img_wh', type=int, default=128)
    parser.add_argument('--near', type=float, default=2.0)
    parser.add_argument('--far', type=float, default=6.0)
    parser.add_argument('--fov', type=float, default=0.6911112070083618)
    parser.add_argument('--eagerly', action='store_true')
    parser.add_argument('--white_bg', action='store_true')
    parser.add_argument('--phi', type=float, default=-30.0)
    parser.add_argument('--z_translate', type=float, default=4.0)
    parser.add_argument('--output_dir', type=str, default='output')
    parser.add_argument('--output_freq', type=int, default=10)
    parser.add_argument('--verbose', action='store_true')
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'fp32', 'fp32_tc'])
    args = parser.parse_args(argv)
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO,
                        format='%(asctime)s | %(name)s | %(levelname)s | %(message)s')
    logging.info(args)
    if args.name == '':
        args.name = args.model_dirs.rstrip('/').split('/')[-1]
    if not NeRF.has_checkpoint(args.model_dirs):
        raise FileNotFoundError(f"Model not found for {args.model_dirs}")

    strategy = None
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        from keras_nerf_b200.distributed import RayShardedStrategy
        strategy = RayShardedStrategy()
    nerf = NeRF(model_path=args.model_dirs, precision=args.precision)
    poses = [pose_spherical(float(theta), args.phi, args.z_translate) for theta in range(0, 360, args.output_freq)]
    logging.info(f'Camera Matrix Shape: {np.stack(poses).shape}')
    rays_generator = RaysGenerator(focal_length=get_focal_from_fov(args.fov, args.img_wh), image_width=args.img_wh,
                                   image_height=args.img_wh, near=args.near, far=args.far, n_sample=nerf.n_coarse)
    nerf.compile(optimizer='adam', loss='mean_squared_error', batch_size=1, image_width=args.img_wh,
                 image_height=args.img_wh, ray_chunks=args.ray_chunks, white_background=args.white_bg,
                 is_training=False)
    nerf.coarse.summary()
    nerf.fine.summary()

    lo, hi = (0, len(poses)) if strategy is None else strategy.shard_bounds(len(poses))
    images = []
    for k in range(lo, hi):                                  # whole frames per rank
        rays = tuple(r[None] for r in rays_generator(poses[k]))
        _, fine = nerf.predict_and_render_images(rays)
        images.append(fine['image'][0])
    frames = torch.stack(images) if images else torch.empty((0, args.img_wh, args.img_wh, 3), device=nerf.device)
    if strategy is not None:
        frames = strategy.gather_rows(frames, len(poses))
        rank = strategy.rank
        torch.distributed.destroy_process_group()
        if rank != 0:
            return None
    os.makedirs(args.output_dir, exist_ok=True)
    logging.info("creating the video from the frames...")
    out = os.path.join(args.output_dir, f"{args.name}.gif")
    mimwrite(out, list(frames.cpu().numpy()), fps=20)
    return out


if __name__ == "__main__":
    main()
