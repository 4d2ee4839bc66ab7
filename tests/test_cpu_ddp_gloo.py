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


def _gpt_worker(rank, world, port, ret):
    """stage-1: micro-batch gradients accumulate in the flat arena, ONE all-reduce per optimizer step (the reference's
    Lightning DDP reduces every micro-batch), the 1/world mean is applied by the ScaledAdam kernels."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from easevoice_trainer_b200.train.gpt_step import FlatScaledAdam
    from oracle import gpt_oracle
    m = dict(gpt_oracle.GPT_MODEL, n_layer=1)
    P = gpt_oracle.init_params(gpt_oracle.gpt_param_spec(m), 5)
    params = {k: torch.nn.Parameter(v.clone()) for k, v in P.items()}
    opt = FlatScaledAdam(params.items())
    total = [torch.zeros_like(v) for v in params.values()]
    for mb in range(3):                                       # three accumulated micro-batches, global batch 4 = 2 ranks x 2
        x, xl, y, yl, bert = gpt_oracle.synthetic_gpt_batch(4, 6, 10, 100 + mb, ragged=True)
        sl = slice(rank * 2, rank * 2 + 2)
        loss = gpt_oracle.forward_old(params, x[sl], xl[sl], y[sl], yl[sl], bert[sl], m)[0]
        opt.accumulate(torch.autograd.grad(loss, list(params.values())))
        if rank == 0:                                         # DDP semantics: mean over ranks of the per-rank (summed) loss gradients
            for r in range(world):
                s2 = slice(r * 2, r * 2 + 2)
                lf = gpt_oracle.forward_old(params, x[s2], xl[s2], y[s2], yl[s2], bert[s2], m)[0]
                for t, g in zip(total, torch.autograd.grad(lf, list(params.values()))):
                    t += g / world
    dist.all_reduce(opt.flat_g)
    flat = opt.flat_g / world
    if rank == 0:
        ref = torch.cat([t.reshape(-1) for t in total])
        ret["gpt_err"] = float((flat - ref).norm() / ref.norm())
        ret["gpt_n"] = flat.numel()
        ret["chunks_cover"] = int(opt.chunks[:, 2].sum()) == flat.numel() and int(opt.numel.sum()) == flat.numel()
    dist.destroy_process_group()


def test_two_rank_gpt_accumulate_then_single_allreduce():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gpt_worker, nprocs=2, args=(2, 29514, ret), join=True)
    assert ret["chunks_cover"]
    assert ret["gpt_err"] < 1e-5, ret["gpt_err"]


def _shape_worker(rank, world, port, ret):
    """ranks collate batches of different lengths; every rank must end up with the SAME padded shape each
    iteration (otherwise they would capture CUDA graphs at different iterations and mis-pair their all-reduces)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from easevoice_trainer_b200.train import s2_step
    from easevoice_trainer_b200.train.data import TextAudioSpeakerCollate

    class Host:                                               # the two attributes S2Step.agree_shape uses
        world, _shape_group = 2, None
    host = Host()
    g = torch.Generator().manual_seed(100 + rank)             # different data per rank
    shapes = []
    for it in range(6):
        items = []
        for _ in range(3):
            T = int(torch.randint(40, 200, (1,), generator=g))
            X = int(torch.randint(5, 90, (1,), generator=g))
            items.append((torch.randn(1, 768, T), torch.rand(1, T * 640 + 17) - 0.5, torch.randint(0, 732, (X,))))
        b = TextAudioSpeakerCollate(640)(items)
        Tq, Xq = s2_step.S2Step.agree_shape(host, b["ssl"].shape[2], b["text"].shape[1])
        p = s2_step.pad_host_batch(b, Tq, Xq, 640)
        assert p["ssl"].shape[2] == Tq and p["wav"].shape[2] == Tq * 640 and p["text"].shape[1] == Xq
        assert Tq % 32 == 0 and Xq % 32 == 0 and Tq >= b["ssl"].shape[2]
        assert torch.equal(p["ssl"][:, :, :b["ssl"].shape[2]], b["ssl"]) and float(p["ssl"][:, :, b["ssl"].shape[2]:].abs().sum()) == 0.0
        shapes.append((Tq, Xq))
    ret[rank] = shapes
    dist.destroy_process_group()


def test_two_ranks_agree_on_padded_shapes():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_shape_worker, nprocs=2, args=(2, 29516, ret), join=True)
    assert ret[0] == ret[1] and len(set(ret[0])) > 1, (ret[0], ret[1])
