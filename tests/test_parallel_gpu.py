"""Two GPUs, NCCL through the C ABI (dg_comm_*): one data-parallel SRGAN train step on two ranks with different batch shards.
The all-reduced flat gradient arenas divided by the world size must equal the MEAN of the two per-shard oracle gradients
(SURVEY.md 8e), every rank must end the step with identical parameters, and those must equal the oracle's Adam update with
the averaged gradients.  Skipped when fewer than two GPUs are visible."""
import os
import socket
from types import SimpleNamespace

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir, fp16):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from denoise_gan_b200.dataloader import synthetic_pair
        from denoise_gan_b200.parallel import GradAllReduce
        from denoise_gan_b200.srgan import SRGAN
        from denoise_gan_b200.train_srgan import train_step
        model = SRGAN(SimpleNamespace(crop_size=64, scale=4, lr=1e-3, fp16=fp16, vgg=0, seed=0))
        model.comm = GradAllReduce(model.device, bucket_mb=1.0)
        assert model.comm._comm is not None, "the CUDA data plane must be dg_comm_allreduce (NCCL through the C ABI)"
        model.world_size = world
        assert len(model.comm.buckets(model.gen_params)) >= 3
        x, y = synthetic_pair(2, 64, 4, step=100 + rank)          # a different shard per rank
        # the same shard through the same kernels WITHOUT the exchange: the local gradients the all-reduce has to average
        solo = SRGAN(SimpleNamespace(crop_size=64, scale=4, lr=1e-3, fp16=fp16, vgg=0, seed=0))
        train_step(solo, x.cuda(), y.cuda())
        torch.cuda.synchronize()
        local = {"g_grad": solo.gen_params.grads(), "d_grad": solo.disc_params.grads()}
        losses = [float(v) for v in train_step(model, x.cuda(), y.cuda())]
        torch.cuda.synchronize()
        assert model.comm.launched == len(model.comm.buckets(model.gen_params)) + len(model.comm.buckets(model.disc_params))
        torch.save({"g_grad": model.gen_params.grads(), "d_grad": model.disc_params.grads(), "g": model.gen_params.export(),
                    "d": model.disc_params.export(), "losses": losses, "local": local}, os.path.join(out_dir, f"r{rank}.pt"))
        model.comm.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("fp16", [0, 1])
def test_two_rank_step_matches_mean_of_shard_oracle_gradients(tmp_path, fp16):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    from denoise_gan_b200 import params as P
    from denoise_gan_b200.dataloader import synthetic_pair
    from oracle import ops_torch as OT
    from oracle import steps as OS
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), fp16), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    # per-shard oracle gradients (float64) from the same initial weights
    shard = []
    for rank in range(world):
        g = {k: v.double() for k, v in P.init_srgan_generator(0, 4).items()}
        d = {k: v.double() for k, v in P.init_patch_discriminator(1).items()}
        x, y = synthetic_pair(2, 64, 4, step=100 + rank)
        out = {}
        OS.srgan_train_step(g, d, None, OT.KerasAdam(1e-3, decay_steps=100000), OT.KerasAdam(5e-3, decay_steps=100000), x.double(), y.double(), out=out)
        shard.append(out)
    # fp32: against the float64 oracle.  bf16: a 2-image batch at 64 px through a discriminator at initialisation carries O(0.5)
    # relative bf16 noise in the gradients it sends back (test_srgan_gpu.py measures it on the bf16-emulating ORACLE), so the oracle
    # cannot tell a correct exchange from a broken one there; the exchange itself is exact arithmetic, so the all-reduced arena is
    # compared with the mean of the two ranks' LOCAL gradients (same kernels, same shards, no exchange) to fp32 summation accuracy.
    tol = 1e-5 if fp16 else 2e-4
    for key, ok in (("g_grad", "gen_grads"), ("d_grad", "disc_grads")):
        worst = 0.0
        for name, t0 in r[0][key].items():
            assert torch.equal(t0, r[1][key][name]), f"{name}: the ranks hold different all-reduced gradients"
            if fp16:
                mean = (r[0]["local"][key][name].double() + r[1]["local"][key][name].double()) / world
            else:
                mean = (shard[0][ok][name] + shard[1][ok][name]) / world
            if name.endswith("/bias") and mean.abs().max() < 1e-9:
                continue            # conv bias in front of a BatchNorm: mathematically zero gradient, rounding noise only
            got = t0.double() / world
            err = ((got - mean).norm() / mean.norm().clamp_min(1e-30)).item()
            worst = max(worst, err)
            assert err < tol, (name, err)
        print(f"fp16={fp16} {key}: worst relative L2 error of the averaged gradient vs the mean of the shard oracles {worst:.3e}")
    for key in ("g", "d"):
        for name, t0 in r[0][key].items():
            if name.endswith(("moving_mean", "moving_variance")):
                continue            # BatchNorm moving statistics are per replica by design (rank 0's are the exported ones)
            assert torch.equal(t0, r[1][key][name]), f"{name}: replicas diverged after one step"
