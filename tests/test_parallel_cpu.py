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
