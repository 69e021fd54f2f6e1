"""CPU, world_size 2 over gloo: the data-parallel exchange of the train step (parallel.GradAllReduce on
flat gradient arenas, batch sharding) — all-reduced gradients scaled by 1/world must equal the mean of
the per-shard gradients, which is what the fused Adam kernel consumes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from denoise_gan_b200.parallel import GradAllReduce, shard_batch


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        full = torch.randn(8, 4, 4, 3)                       # a "global batch"
        mine = shard_batch(full, rank, world)
        assert mine.shape[0] == 4 and torch.equal(mine, full[rank * 4:(rank + 1) * 4])
        # toy per-shard "gradient arenas" (generator, discriminator): functions of the shard only
        g_grad = torch.stack([mine.mean(), mine.std(), mine.abs().sum()]).repeat(5)
        d_grad = torch.stack([mine.min(), mine.max()]).repeat(3)

        class M:  # the attributes GradAllReduce touches
            pass
        m = M(); m.gen_params = M(); m.disc_params = M()
        m.gen_params.grad, m.disc_params.grad = g_grad.clone(), d_grad.clone()
        comm = GradAllReduce("cpu")
        assert comm.world_size == world
        comm.start(m.disc_params.grad)      # the step overlaps this bucket with the generator backward pass
        comm.start(m.gen_params.grad)
        comm.wait()
        torch.save({"g": m.gen_params.grad / world, "d": m.disc_params.grad / world, "g_local": g_grad, "d_local": d_grad},
                   os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_allreduce_is_mean_of_shard_gradients(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    for key, loc in (("g", "g_local"), ("d", "d_local")):
        mean = (r[0][loc] + r[1][loc]) / 2
        assert torch.allclose(r[0][key], mean) and torch.allclose(r[1][key], mean)
        assert torch.equal(r[0][key], r[1][key])             # every replica applies the same update


def _bucket_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from collections import OrderedDict
        from denoise_gan_b200.params import ParamSet
        g = torch.Generator().manual_seed(5)
        shapes = [(3, 3, 16, 64), (64,), (64,), (3, 3, 64, 64), (64,), (64,), (3, 3, 64, 64), (1, 1, 64, 3), (3,)]
        tensors = OrderedDict((f"g/p{i}", torch.randn(*s, generator=g)) for i, s in enumerate(shapes))
        ps = ParamSet("g", tensors, "cpu")
        ps.grad.copy_(torch.arange(ps.grad.numel(), dtype=torch.float32) * (rank + 1))     # rank-dependent "gradients"
        local = ps.grad.clone()
        comm = GradAllReduce("cpu", bucket_mb=0.1)                                           # ~100 KB buckets -> several of them
        b = comm.buckets(ps)
        assert len(b) >= 2 and sum(c for _, c, _ in b) == ps.numel and b[0][0] + b[0][1] == ps.numel and b[-1][0] == 0
        assert set().union(*[n for _, _, n in b]) == set(tensors)                              # every variable in exactly one bucket
        assert sum(len(n) for _, _, n in b) == len(tensors)
        comm.begin(ps)
        done, launched_at = set(), []
        for name in reversed(list(tensors)):                                                  # the backward pass completes the LAST variables first
            done.add(name)
            comm.poll(ps, done)
            launched_at.append(comm.launched)
        assert launched_at[-1] == len(b) and launched_at[0] <= 1 and sorted(launched_at) == launched_at
        comm.poll(ps, done)                                                                    # idempotent
        comm.finish(ps)
        assert comm.launched == len(b)
        # a second step where one variable never completes: finish() must still exchange its bucket
        ps.grad.copy_(local)
        comm.begin(ps)
        comm.poll(ps, set(list(tensors)[1:]))
        assert comm.launched < 2 * len(b)
        comm.finish(ps)
        assert comm.launched == 2 * len(b)
        torch.save({"sum": ps.grad.clone(), "local": local}, os.path.join(out_dir, f"b{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_bucketed_exchange_reverse_layer_order(tmp_path):
    """parallel.GradAllReduce buckets: contiguous ranges of the flat arena in reverse layer order, launched by poll() as soon
    as all their variables are complete, remaining ones by finish(); the union of the bucket all-reduces equals ONE all-reduce of
    the whole arena."""
    world, port = 2, _free_port()
    mp.spawn(_bucket_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"b{i}.pt")) for i in range(world)]
    total = r[0]["local"] + r[1]["local"]
    assert torch.equal(r[0]["sum"], total) and torch.equal(r[1]["sum"], total)
