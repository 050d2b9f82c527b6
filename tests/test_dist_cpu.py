"""world_size-2 gloo tests (CPU) of the data-parallel host logic: ray sharding + flat-bucket
gradient allreduce reproduce the single-process gradients; (sum,count) means; tile gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _model():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Softplus(beta=100), torch.nn.Linear(16, 3))
    # a channels-last 4-D parameter like the VM planes
    plane = torch.empty_strided((1, 4, 5, 5), (100, 1, 20, 4)).copy_(torch.randn(1, 4, 5, 5))
    m.register_parameter("plane", torch.nn.Parameter(plane))
    return m


def _loss_terms(m, x, y):
    pred = m(x) + m.plane.mean()
    per_ray = ((pred - y) ** 2).sum(-1)
    keep = x[:, 0] > 0                      # rank-dependent sample count, like culled samples
    return per_ray, (pred[keep] ** 2).sum(), keep.sum()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensoflow_b200.dist import FlatGradBucket, shard_slice, global_mean, gather_tiles, interleaved_ids, gather_interleaved
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(37, 6, generator=g), torch.randn(37, 3, generator=g)
    m = _model()
    sl = shard_slice(37, rank, world)
    per_ray, s, c = _loss_terms(m, X[sl], Y[sl])
    mean_kept = global_mean(s.detach(), c)
    # loss = mean over ALL rays + mean over ALL kept samples: scale local sums by the global counts
    c_all = c.clone().float()
    dist.all_reduce(c_all)
    loss = per_ray.sum() / 37 + s / c_all
    loss.backward()
    bucket = FlatGradBucket(m.parameters())
    bucket.allreduce()
    tiles = gather_tiles(per_ray.detach()[:, None])
    # inference split (config 5): strided pixel ids, results gathered back into pixel order
    ids = interleaved_ids(37, rank, world)
    inter = gather_interleaved(X[ids] * 2.0, 37)
    assert torch.equal(inter, X * 2.0)
    # numpy payloads are pickled by value (torch tensors travel as file descriptors the exiting worker may close first)
    q.put((rank, [p.grad.clone().numpy() for p in m.parameters()], float(mean_kept), tiles.numpy()))
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    res = [(r, [torch.from_numpy(g) for g in gs], mk, torch.from_numpy(t)) for r, gs, mk, t in res]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = torch.Generator().manual_seed(1)
    X, Y = torch.randn(37, 6, generator=g), torch.randn(37, 3, generator=g)
    m = _model()
    per_ray, s, c = _loss_terms(m, X, Y)
    (per_ray.mean() + s / c).backward()
    for r in range(world):
        for got, p in zip(res[r][1], m.parameters()):
            assert got.stride() == p.grad.stride() or got.shape == p.grad.shape
            assert torch.allclose(got, p.grad, rtol=1e-5, atol=1e-6)
        assert abs(res[r][2] - float(s / c)) < 1e-5
        assert torch.allclose(res[r][3][:, 0], per_ray.detach(), rtol=1e-5, atol=1e-6)


def test_shard_slice_covers_everything():
    from tensoflow_b200.dist import shard_slice
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            idx = []
            for r in range(w):
                s = shard_slice(n, r, w)
                idx += list(range(s.start, s.stop))
            assert idx == list(range(n))


def test_interleaved_ids_cover_everything():
    from tensoflow_b200.dist import interleaved_ids
    for n in (0, 1, 7, 64, 640000):
        for w in (1, 2, 3, 8):
            ids = torch.cat([interleaved_ids(n, r, w) for r in range(w)])
            assert ids.numel() == n and torch.equal(torch.sort(ids).values, torch.arange(n))
            assert max(interleaved_ids(n, r, w).numel() for r in range(w)) - min(interleaved_ids(n, r, w).numel() for r in range(w)) <= 1


def _worker_untouched(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tensoflow_b200.dist import FlatGradBucket
    torch.manual_seed(0)
    a = torch.nn.Parameter(torch.randn(5))
    b = torch.nn.Parameter(torch.randn(3))           # only rank 0's shard reaches it (e.g. inner_light without occluded directions)
    bucket = FlatGradBucket([a, b])
    # 5- and 3-element parameters: every view still starts on a 256-byte boundary of the flat buffer (vector reductions of the kernels)
    assert all((v.data_ptr() - bucket.flat.data_ptr()) % 256 == 0 for v in bucket.views)
    outs = []
    for step in range(2):                            # twice: the views of step 1 must not leak into step 2
        bucket.begin_step()
        loss = (a * (rank + 1.0)).sum() * (step + 1)
        if rank == 0:
            loss = loss + (b * 2.0).sum()
        loss.backward()
        assert (b.grad is None) == (rank != 0)
        work = bucket.allreduce(average=True, async_op=True)
        bucket.finish()
        assert work is not None and a.grad.data_ptr() == bucket.views[0].data_ptr()
        outs.append((a.grad.clone().numpy(), b.grad.clone().numpy()))
    q.put((rank, outs))
    dist.destroy_process_group()


def test_bucket_delivers_gradients_to_untouched_parameters_and_averages():
    """a rank whose shard never touched a parameter must still receive the gradient summed from the other ranks
    (replicas would diverge otherwise), and average=True must divide also on the asynchronous path."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_untouched, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in range(world):
        for step, (ga, gb) in enumerate(res[rank]):
            assert torch.allclose(torch.from_numpy(ga), torch.full((5,), (1.0 + 2.0) / 2 * (step + 1)))
            assert torch.allclose(torch.from_numpy(gb), torch.full((3,), 2.0 / 2))
