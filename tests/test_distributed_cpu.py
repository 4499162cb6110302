"""CPU, world_size 2 over gloo: the host-side logic of the ray-sharded data-parallel path
(keras_nerf_b200/distributed.py replaces tf.distribute.MirroredStrategy, train.py:75-79,110)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from keras_nerf_b200.distributed import RayShardedStrategy
    st = RayShardedStrategy(backend="gloo")
    assert st.num_replicas_in_sync == world and st.rank == rank
    with st.scope():
        pass
    # shards tile [0, n) exactly, also when n is not divisible
    for n in (10, 11, 640000, 1):
        lo, hi = st.shard_bounds(n)
        allb = [None] * world
        torch.distributed.all_gather_object(allb, (lo, hi))
        assert allb[0][0] == 0 and allb[-1][1] == n
        assert all(allb[i][1] == allb[i + 1][0] for i in range(world - 1))
    # gradient all-reduce is a SUM over replicas (train.py:134-136), one collective on the flat buffer
    n = 1000
    flat = torch.full((2 * n,), float(rank + 1))
    st.all_reduce_sum(flat[:n], flat[n:])
    assert torch.equal(flat, torch.full((2 * n,), float(sum(range(1, world + 1)))))
    a, b = torch.full((5,), float(rank)), torch.full((7,), 2.0 * rank)
    st.all_reduce_sum(a, b)
    assert float(a[0]) == sum(range(world)) and float(b[0]) == 2.0 * sum(range(world))
    # render: rows sharded, pixels gathered
    H, Wd = 7, 3
    full = torch.arange(H * Wd * 3, dtype=torch.float32).reshape(H, Wd, 3)
    mine = st.shard(full, 0)
    got = st.gather_rows(mine.clone(), H, 0)
    assert torch.equal(got, full)
    assert st.mean_scalar(float(rank)) == pytest.approx(sum(range(world)) / world)
    md = st.mean_dict({"fine_loss": float(rank), "coarse_loss": 2.0 * rank + 1.0})
    assert md == {"coarse_loss": pytest.approx(float(world)), "fine_loss": pytest.approx((world - 1) / 2.0)}
    # sharded rendering gathers whole ray chunks: uneven chunk counts per rank (5 chunks of 4 rays over 2 ranks)
    chunks, rc = 5, 4
    bounds = [st.shard_bounds_of(r, chunks) for r in range(world)]
    assert bounds[rank] == st.shard_bounds(chunks)
    sizes = [(b - a) * rc for a, b in bounds]
    px = torch.arange(chunks * rc * 8, dtype=torch.float32).reshape(chunks * rc, 8)
    lo, hi = st.shard_bounds(chunks)
    got = st.gather_rows(px[lo * rc:hi * rc].clone(), chunks * rc, 0, sizes=sizes)
    assert torch.equal(got, px)

    # orbit frames sharded over ranks (inference.py, benchmarks/orbit.py): with fewer frames than ranks a rank
    # holds an empty shard
    for nf in (1, 3):
        lo, hi = st.shard_bounds(nf)
        frames = torch.arange(nf * 4, dtype=torch.float32).reshape(nf, 2, 2)
        got = st.gather_rows(frames[lo:hi].clone(), nf)
        assert torch.equal(got, frames)

    class Net:  # broadcast of replicated weights
        def __init__(self, v):
            self.params = torch.full((4,), float(v))

    class M:
        coarse, fine = Net(rank + 10), Net(rank + 20)
    st.broadcast_parameters(M)
    assert float(M.coarse.params[0]) == 10.0 and float(M.fine.params[0]) == 20.0
    st.barrier()
    q.put((rank, "ok"))
    torch.distributed.destroy_process_group()


def test_ray_sharded_strategy_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
