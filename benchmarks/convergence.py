#!/usr/bin/env python
"""End-to-end check of the whole user path on a synthetic nerf_synthetic-shaped directory: PNG frames ->
DatasetLoader/ImageLoader -> NeRF.fit (train.py command line, NeRFTrainMonitor) -> log.csv -> inference.py GIF.
usage: python benchmarks/convergence.py [--precision bf16] [--epochs 30] [--img_wh 100] [--white_bg]"""
import argparse
import csv
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--records", default="fp8", choices=["fp8", "bf16"])
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--img_wh", type=int, default=100)
    ap.add_argument("--views", type=int, default=24)
    ap.add_argument("--white_bg", action="store_true")
    ap.add_argument("--ray_chunks", type=int, default=0, help="default: the whole view in one chunk")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import inference
    import train
    from keras_nerf_b200.data.synthetic import write_nerf_synthetic_like
    tmp = a.out or tempfile.mkdtemp(prefix="knerf_conv_")
    data = write_nerf_synthetic_like(os.path.join(tmp, "scene"), image_wh=2 * a.img_wh, n_train=a.views, n_val=2, n_test=4)
    argv = ["--name", "ball", "--data_dir", data, "--img_wh", str(a.img_wh), "--batch_size", "1", "--ray_chunks",
            str(a.ray_chunks or a.img_wh * a.img_wh), "--log_dir", os.path.join(tmp, "logs"), "--model_dirs", os.path.join(tmp, "model"),
            "--log_freq", str(max(a.epochs // 3, 1)), "--precision", a.precision, "--records", a.records, "--num_epochs", str(a.epochs)]
    if a.white_bg:
        argv.append("--white_bg")
    t0 = time.time()
    nerf = train.main(argv, multi_gpu=False)
    dt = time.time() - t0
    hist = nerf.history if hasattr(nerf, "history") else None
    rows = list(csv.DictReader(open(os.path.join(tmp, "logs", "ball", "log.csv"))))
    gif = inference.main(["--model_dirs", os.path.join(tmp, "model", "ball"), "--img_wh", str(a.img_wh), "--ray_chunks",
                          str(a.ray_chunks or a.img_wh * a.img_wh), "--output_freq", "30", "--output_dir", os.path.join(tmp, "out"),
                          "--precision", a.precision] + (["--white_bg"] if a.white_bg else []))
    print(json.dumps({"precision": a.precision, "img_wh": a.img_wh, "views": a.views, "epochs": a.epochs,
                      "white_bg": a.white_bg, "train_wall_s": round(dt, 1), "steps": a.views * a.epochs,
                      "history": hist, "logged": [{k: round(float(v), 5) for k, v in r.items()} for r in rows],
                      "gif": os.path.basename(gif), "gif_bytes": os.path.getsize(gif)}))


if __name__ == "__main__":
    main()
