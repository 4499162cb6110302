"""Ray-sharded data parallelism: the B200 replacement for tf.distribute.MirroredStrategy (train.py:75-79,110).

One process per GPU (torchrun).  Rays are independent units, so the data path has NO collective; the only
exchange is the SUM all-reduce of the flat accumulated MLP gradient (2 x 595,844 fp32 = 4.77 MB) that
MirroredStrategy performs inside `apply_gradients` (nerf.py:455-458; SUM, not mean: train.py:134-136), and a
gather of rendered pixels for inference.

The gradient all-reduce of the training step lives behind the C ABI: `knerf_comm()` builds a libknerf
communicator (its own NCCL communicator, bootstrapped by broadcasting the 128-byte unique id over
torch.distributed) and `NeRF.accumulate_gradients` hands it to `knerf_train_chunk_dp`, which issues the coarse
network's all-reduce on a side stream while the fine network is still in its backward.  torch.distributed is the
rendezvous / bootstrap plumbing and carries the rest (parameter broadcast, pixel gather, metric means; gloo for
the CPU tests of the host logic, where the gradient SUM falls back to `dist.all_reduce`)."""
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
        if device is None and dist.get_backend() == "nccl":
            device = torch.device("cuda", torch.cuda.current_device())   # NCCL rejects CPU tensors
        self.device = torch.device(device) if device is not None else None
        self._comm = None            # libknerf communicator (knerf_comm*), created on first use
        self._comm_stream = None

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
    def knerf_comm(self):
        """(knerf_comm*, cudaStream_t) for knerf_train_chunk_dp / knerf_allreduce_grads, or (None, None) when there
        is nothing to reduce or no GPU (gloo tests).  Collective on first use: every rank must call it."""
        if self.num_replicas_in_sync == 1 or self.device is None or self.device.type != "cuda":
            return None, None
        if self._comm is None:
            import ctypes as C
            from . import _lib
            lib = _lib.load()
            with torch.cuda.device(self.device):
                ident = (C.c_ubyte * _lib.COMM_ID_BYTES)()
                if self.rank == 0:
                    _lib.check(lib.knerf_comm_unique_id(ident))
                dev = self.device if dist.get_backend() == "nccl" else torch.device("cpu")
                t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
                dist.broadcast(t, src=0)
                ident = (C.c_ubyte * _lib.COMM_ID_BYTES)(*t.cpu().tolist())
                comm = C.c_void_p()
                _lib.check(lib.knerf_comm_create(ident, self.rank, self.num_replicas_in_sync, C.byref(comm)))
                self._comm = comm
                self._comm_stream = torch.cuda.Stream(device=self.device)
        return self._comm, self._comm_stream.cuda_stream

    def close(self):
        if self._comm is not None:
            from . import _lib
            _lib.load().knerf_comm_destroy(self._comm)
            self._comm = None

    def all_reduce_sum(self, *tensors):
        if self.num_replicas_in_sync == 1:
            return
        comm, _ = self.knerf_comm()
        if comm is not None and all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in tensors):
            from . import _lib
            with torch.cuda.device(self.device):
                for t in tensors:        # behind the C ABI, on the caller's current stream
                    _lib.call("knerf_allreduce_grads", comm, t.data_ptr(), t.numel(), _lib.stream())
            return
        # CPU / gloo: the two accumulators are views of one flat buffer when allocated by NeRF: one collective
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
