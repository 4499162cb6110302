#!/usr/bin/env python
"""Does training with fp8 records (KNERF_REC_FP8) converge like training with bf16 records?  Two identically seeded
bf16-mode models, one per record format, take the same steps on the synthetic scene at the bench's training shape
(32,768 rays per step, random windows of 100 views of 400 x 400); every `every` steps both are evaluated on the same
held-out rays.  With `spread` a third run repeats the bf16-records training with other fine-sample draws: the spread
between two runs of the SAME format is the yardstick for the difference between the formats (trajectories at a constant
learning rate of 1e-3 are chaotic: any perturbation, even the order of the fp32 atomics, separates them).
usage: python benchmarks/records_convergence.py [steps=600] [every=100] [spread]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    every = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    from keras_nerf_b200 import NeRF
    from keras_nerf_b200.data.synthetic import SyntheticScene
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    dev = torch.device("cuda", 0)
    R = 32768
    scene = SyntheticScene(400, 64, n_views=100, device=dev)
    rng = np.random.default_rng(0)
    plan = [(int(rng.integers(0, 90)), int(rng.integers(0, 400 * 400 - R))) for _ in range(steps)]
    held = [scene.ray_batch(90 + k, R, offset=40000 + 9000 * k, seed=500 + k) for k in range(4)]   # views 90..93: never trained on
    out = {"steps": steps, "rays_per_step": R, "eval_rays": 4 * R}
    variants = [("bf16", "bf16", 0), ("fp8", "fp8", 0)]
    if len(sys.argv) > 3 and sys.argv[3] == "spread":
        variants.append(("bf16_other_draws", "bf16", 1000003))
    for tag, records, seed_off in variants:
        mlp_mod.set_seed(42)
        torch.manual_seed(0)
        m = NeRF(precision="bf16", device=dev, records=records)
        m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=R // 256, image_width=256, ray_chunks=R,
                  white_background=True)
        curve = []
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for i, (k, off) in enumerate(plan):
            img, rays = scene.ray_batch(k, R, offset=off, seed=10000 + i)
            m.train_step((img, rays), seed=7000 + i + seed_off)
            if (i + 1) % every == 0 or i + 1 == steps:
                mse = np.zeros(2)
                for hi, hr in held:
                    res = m.predict_and_render_images(hr, seed=1)
                    for j in range(2):
                        mse[j] += float(((res[j]["image"] - hi[..., :3].to(dev)) ** 2).mean()) / len(held)
                curve.append({"step": i + 1, "val_coarse_psnr": round(float(-10 * np.log10(mse[0])), 3),
                              "val_fine_psnr": round(float(-10 * np.log10(mse[1])), 3)})
        t1.record(); torch.cuda.synchronize()
        out[tag] = {"curve": curve, "seconds_incl_eval": round(t0.elapsed_time(t1) / 1e3, 2)}
        del m
    print(json.dumps(out))


if __name__ == "__main__":
    main()
