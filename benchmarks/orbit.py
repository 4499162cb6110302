#!/usr/bin/env python
"""BASELINE config[4], the MLP-bound stress: bf16 tensor-core MLP, 64 coarse + 256 fine samples per ray, a
360-degree orbit of 1024x1024 frames (the inference.py flow: pose_spherical -> RaysGenerator ->
predict_and_render_images), whole frames sharded over the GPUs of one box and gathered on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/orbit.py [frames] [wh]

One JSON line from rank 0.  frames defaults to 40 (output_freq 9); pass fewer on one GPU (16.1 G samples in all)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    wh = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    from keras_nerf_b200 import NeRF, RaysGenerator, get_focal_from_fov, pose_spherical
    from keras_nerf_b200.distributed import RayShardedStrategy
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    st = RayShardedStrategy(backend="nccl", device=dev) if world > 1 else None
    mlp_mod.set_seed(42)
    model = NeRF(n_coarse=64, n_fine=256, precision="bf16", strategy=st, device=dev)
    model.compile(optimizer="adam", loss="mse", batch_size=1, image_height=wh, image_width=wh, ray_chunks=wh * 32,
                  white_background=True, is_training=False)
    if st is not None:
        st.broadcast_parameters(model)
        model._repack()
    gen = RaysGenerator(get_focal_from_fov(0.6911112070083618, wh), wh, wh, 2.0, 6.0, model.n_coarse, device=dev)
    poses = [pose_spherical(360.0 * k / frames, -30.0, 4.0) for k in range(frames)]
    lo, hi = (0, frames) if st is None else st.shard_bounds(frames)

    def render(k):
        rays = tuple(r[None] for r in gen(poses[k], seed=k))
        return model.predict_and_render_images(rays, seed=1000 + k)[1]["image"][0]

    def sync():
        torch.cuda.synchronize()
        if st is not None:
            st.barrier()
            torch.cuda.synchronize()

    render(lo if hi > lo else 0)                              # warm-up frame
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mine = [render(k) for k in range(lo, hi)]
    stack = torch.stack(mine) if mine else torch.empty((0, wh, wh, 3), device=dev)
    if st is not None:
        stack = st.gather_rows(stack, frames)                 # every rank ends up with the whole orbit
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if st is not None:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    assert stack.shape == (frames, wh, wh, 3) and bool(torch.isfinite(stack).all())
    if rank == 0:
        samples = frames * wh * wh * (2 * model.n_coarse + model.n_fine)
        print(json.dumps({"metric": f"orbit_seconds_{frames}x{wh}x{wh}_nf256", "value": float(ms) / 1e3, "unit": "s",
                          "n_gpus": world, "frames": frames, "ms_per_frame_per_gpu": float(ms) / max(hi - lo, 1),
                          "samples": samples, "tflops": 1_186_816 * samples / (float(ms) * 1e-3) / 1e12,
                          "note": "tflops counts the unfolded 593,408 MAC/sample; includes ray generation, both "
                                  "samplers, compositing and the all-gather of the frames",
                          "frames_gathered_bytes": frames * wh * wh * 12, "precision_mode": "bf16"}))
    if st is not None:
        st.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
