"""GPU (-m gpu), needs >= 2 devices (skipped otherwise): ray-sharded data parallelism has MirroredStrategy's
semantics (train.py:75,110,130-136; keras_nerf/model/nerf/nerf.py:455-458) --

  * the gradient every replica applies is the SUM over the replicas of the gradients of their own shards,
  * all replicas hold identical parameters after every step,
  * which is the step a single process takes when it accumulates the same shards itself.

Two processes (one per GPU, NCCL) through the product path: NeRF.train_step -> knerf_train_chunk_dp -> the
library's own communicator (knerf_comm_create) with the coarse all-reduce overlapped on a side stream."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs(rank, R, step):
    """shard `rank` of step `step`: R rays of a synthetic view, with explicit fine-sample draws"""
    from keras_nerf_b200.data.synthetic import SyntheticScene
    dev = torch.device("cuda", torch.cuda.current_device())
    scene = SyntheticScene(64, 64, n_views=8, device=dev)
    img, rays = scene.ray_batch(2 * step + rank, R, offset=700 * rank + 100 * step, seed=50 + 2 * step + rank)
    u = torch.rand(R, 128, generator=torch.Generator().manual_seed(1000 + 2 * step + rank)).to(dev)
    return img, rays, u


def _model(precision, R, strategy=None):
    import keras_nerf_b200 as K
    from keras_nerf_b200.model.nerf import mlp as mlp_mod
    mlp_mod.set_seed(42)
    m = K.NeRF(precision=precision, strategy=strategy, scan_mode="sequential")
    m.compile(optimizer="adam", loss="mse", batch_size=1, image_height=R // 256, image_width=256, ray_chunks=R // 2,
              white_background=True)           # two accumulation chunks per step: the reduction rides in the last
    return m


def _worker(rank, world, port, precision, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    from keras_nerf_b200.distributed import RayShardedStrategy
    st = RayShardedStrategy(backend="nccl", device=torch.device("cuda", rank))
    R = 1024
    m = _model(precision, R, st)
    st.broadcast_parameters(m)
    res = {}
    for step in range(2):
        img, rays, u = _inputs(rank, R, step)
        # (1) this rank's own shard gradient, NOT reduced
        m.accumulate_gradients(img, rays, u_fine=u, want_images=False, reduce=False)
        torch.cuda.synchronize()
        local = m._grad_flat.clone()
        m._grad_flat.zero_()
        m._losses.zero_()
        # (2) the product path: accumulate + overlapped all-reduce inside libknerf
        m.accumulate_gradients(img, rays, u_fine=u, want_images=False)
        torch.cuda.synchronize()
        reduced = m._grad_flat.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        torch.distributed.all_gather(gathered, local)
        res[f"reduced{step}"] = reduced.cpu().numpy()
        res[f"sum_of_shards{step}"] = torch.stack(gathered).double().sum(0).float().cpu().numpy()
        m._losses.zero_()
        m.apply_gradients()
        torch.cuda.synchronize()
        res[f"params{step}"] = torch.cat([m.coarse.params, m.fine.params]).cpu().numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    st.barrier()
    st.close()
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_two_replicas_equal_one_process_accumulating_both_shards(precision, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), precision, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # single process, same weights, both shards accumulated by itself (grad = g_0 + g_1: SUM, not mean)
    torch.cuda.set_device(0)
    R = 1024
    m = _model(precision, R)
    for step in range(2):
        for k in range(world):
            # every rank ends with the same reduced gradient, bit for bit (one all-reduce result)
            assert np.array_equal(r[0][f"reduced{step}"], r[k][f"reduced{step}"])
            assert np.array_equal(r[0][f"params{step}"], r[k][f"params{step}"])       # replicas stay identical
        red, want = r[0][f"reduced{step}"], r[0][f"sum_of_shards{step}"]
        scale = float(np.abs(want).max())
        assert scale > 0
        # == SUM of the shard gradients (fp32 sum order of the ring / bf16 atomics: 1e-5 relative to the largest)
        tol = 1e-6 if precision == "fp32" else 2e-5
        assert float(np.abs(red - want).max()) <= tol * scale
        for k in range(world):
            img, rays, u = _inputs(k, R, step)
            m.accumulate_gradients(img, rays, u_fine=u, want_images=False)
            torch.cuda.synchronize()
        total = m._grad_flat.clone()       # the accumulators SUM over calls until apply_gradients clears them
        assert float((total.cpu() - torch.from_numpy(want)).abs().max()) <= (2e-6 if precision == "fp32" else 5e-3) * scale
        m._losses.zero_()
        m.apply_gradients()
        torch.cuda.synchronize()
        if precision == "fp32":
            p = torch.cat([m.coarse.params, m.fine.params]).cpu().numpy()
            # Adam where the gradient is clearly non-zero (sign(g) noise elsewhere)
            sel = np.abs(want) > 1e-3 * scale
            assert float(np.abs(p[sel] - r[0][f"params{step}"][sel]).max()) <= 1e-5
        # continue from the replicas' weights: Adam's sign(g) on ~zero gradients is not comparable across summation
        # orders, and would otherwise leak into the next step's gradients
        n = m.coarse.params.numel()
        both = torch.from_numpy(r[0][f"params{step}"]).cuda()
        m.coarse.params.copy_(both[:n])
        m.fine.params.copy_(both[n:])
        m._repack()
