"""world_size-2 gloo test of the data-parallel exchange the stage-2 step performs: every rank flattens its
per-parameter gradients into one fp32 arena (FlatAdamW.set_grads), one SUM all-reduce runs over the arena, and the
1/world factor is applied by the optimizer kernel.  Checked against single-process gradients of the global batch,
with the oracle's discriminator as the (cheap) network."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from easevoice_trainer_b200.train.s2_step import FlatAdamW
    from oracle import s2_oracle
    spec = {k: v for k, v in s2_oracle.discriminator_param_spec().items() if k.startswith("discriminators.0.")}
    P = s2_oracle.init_params(spec, 4321)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in P.items()}
    opt = FlatAdamW(params.items(), [(1.0, list(params))], (0.8, 0.99), 1e-9)
    g = torch.Generator().manual_seed(11)
    y = torch.rand(4, 1, 4096, generator=g) - 0.5
    yh = torch.rand(4, 1, 4096, generator=g) - 0.5

    def loss_of(pp, ys, yhs):
        r, _ = s2_oracle.disc_s(pp, "discriminators.0", ys)
        f, _ = s2_oracle.disc_s(pp, "discriminators.0", yhs)
        return torch.mean((1 - r) ** 2) + torch.mean(f ** 2)

    sl = slice(rank * 2, rank * 2 + 2)                       # per-rank shard of the global batch
    grads = torch.autograd.grad(loss_of(params, y[sl], yh[sl]), list(params.values()))
    opt.set_grads(grads)
    dist.all_reduce(opt.flat_g)                              # SUM, as S2Step._allreduce does
    flat = opt.flat_g / world                                # the 1/world factor lives in the optimizer kernel
    if rank == 0:
        full = torch.autograd.grad(loss_of(params, y, yh), list(params.values()))
        ref = torch.cat([t.reshape(-1) for t in full])
        ret["err"] = float((flat - ref).norm() / ref.norm())
        ret["n"] = flat.numel()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_equals_global_batch():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, nprocs=2, args=(2, 29512, ret), join=True)
    assert ret["n"] == 5_641_362
    assert ret["err"] < 1e-5, ret["err"]
