"""Ray-sharded data parallelism: the B200 replacement for tf.distribute.MirroredStrategy (train.py:75-79,110).

One process per GPU (torchrun).  Rays are independent units, so the data path has NO collective; the only
exchange is the SUM all-reduce of the flat accumulated MLP gradient (2 x 595,844 fp32 = 4.77 MB) that
MirroredStrategy performs inside `apply_gradients` (nerf.py:455-458; SUM, not mean: train.py:134-136), and a
gather of rendered pixels for inference.  torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo
for the CPU tests of the host logic)."""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist


class RayShardedStrategy:
    def __init__(self, backend: str | None = None, device=None):
        if not dist.is_initialized():
            if "RANK" not in os.environ:
                os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
                os.environ.setdefault("MASTER_PORT", "29511")
                os.environ.setdefault("RANK", "0")
                os.environ.setdefault("WORLD_SIZE", "1")
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
                torch.cuda.set_device(local)
                device = device or torch.device("cuda", local)
                dist.init_process_group(backend, device_id=device)
            else:
                dist.init_process_group(backend)
        self.rank = dist.get_rank()
        self.num_replicas_in_sync = dist.get_world_size()
        self.device = device

    @contextlib.contextmanager
    def scope(self):  # API parity with `with strategy.scope():` (train.py:110)
        yield self

    # ---- sharding ------------------------------------------------------------------------------
    def shard_bounds(self, n: int):
        """contiguous [lo, hi) slice of n units (rays / image rows / frames) owned by this rank"""
        w, r = self.num_replicas_in_sync, self.rank
        base, rem = divmod(n, w)
        lo = r * base + min(r, rem)
        return lo, lo + base + (1 if r < rem else 0)

    def shard_bounds_of(self, rank: int, n: int):
        """shard_bounds of another rank"""
        w = self.num_replicas_in_sync
        base, rem = divmod(n, w)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)

    def shard(self, x, dim=0):
        lo, hi = self.shard_bounds(x.shape[dim])
        return x.narrow(dim, lo, hi - lo)

    # ---- collectives ---------------------------------------------------------------------------
    def all_reduce_sum(self, *tensors):
        if self.num_replicas_in_sync == 1:
            return
        # the two accumulators are views of one flat buffer when allocated by NeRF: one collective
        if len(tensors) == 2 and tensors[0].untyped_storage().data_ptr() == tensors[1].untyped_storage().data_ptr():
            base = tensors[0]._base if tensors[0]._base is not None else tensors[0]
            dist.all_reduce(base, op=dist.ReduceOp.SUM)
            return
        for t in tensors:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def broadcast_parameters(self, model, src=0):
        """replicas start from identical weights (MirroredStrategy mirrors variables at creation)"""
        if self.num_replicas_in_sync == 1:
            return
        for net in (model.coarse, model.fine):
            dist.broadcast(net.params, src=src)

    def gather_rows(self, x: torch.Tensor, n_total: int, dim=0, sizes=None):
        """all ranks' shards -> full tensor on every rank.  Shards are those of shard()/shard_bounds over n_total, or
        `sizes[r]` rows from rank r (e.g. whole ray chunks) when given."""
        w = self.num_replicas_in_sync
        if w == 1:
            return x
        x = x.movedim(dim, 0).contiguous()
        if sizes is None:
            base, rem = divmod(n_total, w)
            sizes = [base + (1 if r < rem else 0) for r in range(w)]
        assert sum(sizes) == n_total and x.shape[0] == sizes[self.rank]
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        pad[:x.shape[0]] = x
        out = torch.empty((w * mx,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, pad)
        parts = [out[r * mx:r * mx + sizes[r]] for r in range(w)]
        return torch.cat(parts, dim=0).movedim(0, dim)

    def mean_scalar(self, v: float) -> float:
        if self.num_replicas_in_sync == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item()) / self.num_replicas_in_sync

    def mean_dict(self, logs: dict) -> dict:
        """cross-replica mean of every value of a metrics dict in ONE all-reduce (Keras aggregates the `Mean` metric
        variables of a MirroredStrategy over the replicas; every rank sees the same number of batches)"""
        if self.num_replicas_in_sync == 1 or not logs:
            return dict(logs)
        keys = sorted(logs)
        t = torch.tensor([float(logs[k]) for k in keys], dtype=torch.float64, device=self.device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return {k: float(v) / self.num_replicas_in_sync for k, v in zip(keys, t.tolist())}

    def barrier(self):
        if self.num_replicas_in_sync > 1:
            dist.barrier()
