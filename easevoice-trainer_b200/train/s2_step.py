"""One stage-2 (SoVITS + HiFi-GAN) optimisation step on the sm_100a kernels.

Mirrors /root/reference/src/train/sovits.py:459-525 (G forward, mel/slice features, D step, G step, two AdamW
updates) with these deliberate differences, none of which changes the math of the update:
  * fp32 storage / TF32 tensor-core math instead of fp16 autocast + GradScaler (no loss scaling needed);
  * the discriminators see real and generated audio as one 2B batch (models.py:606-612 runs them sequentially);
  * the G step does not compute (and all-reduce) discriminator weight gradients that the reference computes and
    then discards at the next ``optim_d.zero_grad()``;
  * the grad-norm probe is one fused reduction inside the optimizer kernel instead of 883 ``.item()`` syncs
    (commons.py:140-155); nothing in the step synchronises with the host;
  * slice ids / posterior noise / dropout masks come from device-side Philox streams.
Data-parallel: gradients are summed across ranks with NCCL over one flat fp32 buffer per network, then the
1/world scaling is folded into the optimizer kernel.
"""
import math

import torch
import torch.distributed as dist

from .. import ops
from ..mel_processing import get_bank


class FlatAdamW:
    """AdamW (torch.optim.AdamW semantics, sovits.py:294-319) over per-group flat fp32 arenas.

    Parameters are re-pointed into one contiguous buffer per lr group, so the update is one kernel per group and the
    data-parallel all-reduce is one collective per network."""

    def __init__(self, named_params, groups, betas, eps, weight_decay=0.01):
        # groups: list of (lr_scale, [names]) ; every param must appear exactly once
        self.betas, self.eps, self.wd = betas, eps, weight_decay
        named = dict(named_params)
        self.groups = []
        dev = next(iter(named.values())).device
        total = sum(named[n].numel() for _, names in groups for n in names)
        self.flat_p = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        off = 0
        self.slots = {}
        for lr_scale, names in groups:
            beg = off
            for n in names:
                p = named[n]
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.data.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)
                self.slots[n] = (off, k)
                off += k
            self.groups.append(dict(lr_scale=lr_scale, beg=beg, end=off, names=list(names)))
        self.params = [named[n] for _, names in groups for n in names]
        self.names = [n for _, names in groups for n in names]
        self.hyper = torch.zeros(2, device=dev, dtype=torch.float32)        # [lr, step] lives on the device
        self.gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)

    @property
    def step_count(self):
        return int(self.hyper[1].item())

    def set_lr(self, lr):
        """host-side (outside any captured graph): the lr only changes at epoch boundaries (ExponentialLR)."""
        self.hyper[0] = float(lr)

    def set_grads(self, grads):
        """copy autograd's per-parameter gradients into the flat buffer (unused parameters get zeros)."""
        views, srcs = [], []
        for n, g in zip(self.names, grads):
            off, k = self.slots[n]
            if g is None:
                self.flat_g[off:off + k].zero_()
            else:
                views.append(self.flat_g[off:off + k].view_as(g))
                srcs.append(g)
        torch._foreach_copy_(views, srcs)

    def step(self, grad_scale=1.0):
        """graph-capturable: step counter, bias corrections and lr are read from device memory."""
        ops.scalar_add(self.hyper[1:], 1.0)
        self.gnorm_sq.zero_()
        for g in self.groups:
            s, e = g["beg"], g["end"]
            ops.adamw_flat(self.flat_p[s:e], self.flat_g[s:e], self.flat_m[s:e], self.flat_v[s:e], self.hyper, g["lr_scale"],
                           self.betas, self.eps, self.wd, grad_scale, self.gnorm_sq)

    # torch.optim-compatible state for checkpoints (ckpt.py:78-93 stores optimizer.state_dict())
    def state_dict(self):
        state = {}
        sc = self.step_count
        for i, n in enumerate(self.names):
            off, k = self.slots[n]
            shape = self.params[i].shape
            state[i] = dict(step=torch.tensor(float(sc)), exp_avg=self.flat_m[off:off + k].view(shape).clone(),
                            exp_avg_sq=self.flat_v[off:off + k].view(shape).clone())
        pg, idx = [], 0
        for g in self.groups:
            n = len(g["names"])
            pg.append(dict(lr_scale=g["lr_scale"], betas=self.betas, eps=self.eps, weight_decay=self.wd,
                           params=list(range(idx, idx + n))))
            idx += n
        return dict(state=state, param_groups=pg)

    def load_state_dict(self, sd):
        for i, n in enumerate(self.names):
            if i in sd["state"]:
                off, k = self.slots[n]
                st = sd["state"][i]
                self.flat_m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.flat_v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                self.hyper[1] = float(st["step"])


def g_param_groups(net_g, text_low_lr_rate):
    """sovits.py:286-312: text_embedding / encoder_text / mrte train at lr * text_low_lr_rate."""
    base, te, et, mr = [], [], [], []
    for n, _ in net_g.named_parameters():
        if n.startswith("enc_p.text_embedding."):
            te.append(n)
        elif n.startswith("enc_p.encoder_text."):
            et.append(n)
        elif n.startswith("enc_p.mrte."):
            mr.append(n)
        else:
            base.append(n)
    return [(1.0, base), (text_low_lr_rate, te), (text_low_lr_rate, et), (text_low_lr_rate, mr)]


class S2Step:
    """Holds the two networks, both optimisers and runs `sovits.py:459-525` for one batch."""

    def __init__(self, net_g, net_d, hps_train, hps_data, world_size=1):
        self.net_g, self.net_d = net_g, net_d
        self.t, self.d = hps_train, hps_data
        self.world = world_size
        self.seg_frames = hps_train["segment_size"] // hps_data["hop_length"]
        betas = tuple(hps_train["betas"])
        self.opt_g = FlatAdamW(net_g.named_parameters(), g_param_groups(net_g, hps_train["text_low_lr_rate"]), betas,
                               hps_train["eps"])
        self.opt_d = FlatAdamW(net_d.named_parameters(), [(1.0, [n for n, _ in net_d.named_parameters()])], betas,
                               hps_train["eps"])
        self.lr = hps_train["learning_rate"]
        self.opt_g.set_lr(self.lr)
        self.opt_d.set_lr(self.lr)
        dev = next(net_g.parameters()).device
        self.bank = get_bank(hps_data["sampling_rate"], hps_data["filter_length"], hps_data["n_mel_channels"],
                             hps_data["mel_fmin"], hps_data["mel_fmax"], dev)
        self._graphs = {}

    def losses(self, batch, noise=None, ids_slice=None):
        """Forward + both losses (no optimiser).  batch: dict of channels-last device tensors:
        ssl [B,T,768], spec [B,T,1025], lengths int32 [B], wav [B,L,1], text int64 [B,X], text_lengths int32 [B]."""
        hop, seg = self.d["hop_length"], self.t["segment_size"]
        r = self.net_g.forward_cl(batch["ssl"], batch["spec"], batch["lengths"], batch["text"], batch["text_lengths"],
                                  noise, ids_slice)
        ids = r["ids_slice"]
        y_hat = r["y_hat"]                                                        # [B, seg, 1]
        B = y_hat.shape[0]
        mel = ops.spec_to_mel(batch["spec"], self.bank)                            # sovits.py:470-477
        y_mel = ops.slice_rows(mel, ids, self.seg_frames)
        _, y_hat_mel = ops.mel_frontend(y_hat.reshape(B, seg), self.bank, hop)     # sovits.py:481-490
        y = ops.slice_rows(batch["wav"], ids, seg, mul=hop)                        # sovits.py:492-494
        r.update(y=y, y_mel=y_mel, y_hat_mel=y_hat_mel)
        return r

    def d_loss(self, r):
        B = r["y"].shape[0]
        outs = self.net_d.forward_cl(r["y"], r["y_hat"].detach())                  # sovits.py:497
        loss = 0
        for logit, _ in outs:
            loss = loss + ops.mean_sq_one_minus(logit[:B]) + ops.mean_sq(logit[B:])
        return loss

    def g_loss(self, r):
        B = r["y"].shape[0]
        outs = self.net_d.forward_cl(r["y"], r["y_hat"], weights_need_grad=False)  # sovits.py:511
        loss_mel = ops.mean_abs_diff(r["y_hat_mel"], r["y_mel"]) * self.t["c_mel"]
        loss_kl = ops.kl_loss(r["z_p"], r["logs_q"], r["m_p"], r["logs_p"], r["lengths"]) * self.t["c_kl"]
        loss_fm, loss_gen = 0, 0
        for logit, fmap in outs:
            for f in fmap:
                loss_fm = loss_fm + ops.mean_abs_diff(f[B:], f[:B])
            loss_gen = loss_gen + ops.mean_sq_one_minus(logit[B:])
        loss_fm = loss_fm * 2
        total = loss_gen + loss_fm + loss_mel + loss_kl                            # + kl_ssl == 0 (frozen quantizer)
        return total, dict(loss_gen=loss_gen, loss_fm=loss_fm, loss_mel=loss_mel, loss_kl=loss_kl)

    def _allreduce(self, opt):
        if self.world > 1:
            dist.all_reduce(opt.flat_g)

    # The step in three segments, split where the data-parallel gradient exchanges happen:
    #   A: G forward, features, D forward, D backward -> flat D grads          | all-reduce(D grads)
    #   B: D AdamW, D forward (updated weights), G backward -> flat G grads    | all-reduce(G grads)
    #   C: G AdamW, RNG advance
    def _seg_a(self, batch, noise=None, ids_slice=None):
        r = self.losses(batch, noise, ids_slice)
        loss_d = self.d_loss(r)
        self.opt_d.set_grads(torch.autograd.grad(loss_d, self.opt_d.params, allow_unused=True))
        return r, loss_d

    def _seg_b(self, r):
        self.opt_d.step(1.0 / self.world)
        loss_g, parts = self.g_loss(r)
        self.opt_g.set_grads(torch.autograd.grad(loss_g, self.opt_g.params, allow_unused=True))
        return loss_g, parts

    def _seg_c(self):
        self.opt_g.step(1.0 / self.world)
        ops.advance_rng()

    @staticmethod
    def _outputs(loss_d, loss_g, parts, opt_d, opt_g):
        out = dict(loss_disc=loss_d.detach(), loss_gen_all=loss_g.detach(), grad_norm_d=opt_d.gnorm_sq, grad_norm_g=opt_g.gnorm_sq)
        out.update({k: v.detach() for k, v in parts.items()})
        return out

    def step(self, batch, noise=None, ids_slice=None):
        """Eager step (every kernel launched from Python).  No host synchronisation anywhere."""
        r, loss_d = self._seg_a(batch, noise, ids_slice)
        self._allreduce(self.opt_d)
        loss_g, parts = self._seg_b(r)
        self._allreduce(self.opt_g)
        self._seg_c()
        return self._outputs(loss_d, loss_g, parts, self.opt_d, self.opt_g)

    # ---- CUDA-graph path: the ~3000 launches of one step are captured once per batch shape and replayed --------
    def graph_step(self, batch):
        """Copy `batch` into the static input buffers of the graph captured for its shape, replay, return the static
        loss tensors.  Shapes are bucketed by the sampler (sovits.py:233-252), so a handful of graphs covers a run."""
        key = (tuple(batch["ssl"].shape), tuple(batch["text"].shape))
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(batch)
            self._graphs[key] = g
        graphs, static, out = g
        for k in static:
            if static[k] is not batch[k]:
                static[k].copy_(batch[k], non_blocking=True)
        if len(graphs) == 1:
            graphs[0].replay()
        else:                               # data parallel: NCCL all-reduces run between the captured segments
            graphs[0].replay()
            self._allreduce(self.opt_d)
            graphs[1].replay()
            self._allreduce(self.opt_g)
            graphs[2].replay()
        return out

    def _capture(self, batch):
        static = {k: (v.clone() if k != "spec" else v) for k, v in batch.items()}
        # `spec` keeps its padded row pitch (a view of a wider buffer): clone the parent storage explicitly
        sp = batch["spec"]
        wide = torch.empty((sp.shape[0], sp.shape[1], sp.stride(1)), device=sp.device, dtype=sp.dtype)
        static["spec"] = wide[:, :, :sp.shape[2]]
        static["spec"].copy_(sp)
        # warm-up (allocator, smem attributes, NCCL) must not count as training: snapshot and restore all mutable state
        snap = [t.clone() for t in (self.opt_g.flat_p, self.opt_g.flat_m, self.opt_g.flat_v, self.opt_g.hyper,
                                    self.opt_d.flat_p, self.opt_d.flat_m, self.opt_d.flat_v, self.opt_d.hyper,
                                    ops.rng_state(sp.device))]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.step(static)
        torch.cuda.current_stream().wait_stream(side)
        for dst, src in zip((self.opt_g.flat_p, self.opt_g.flat_m, self.opt_g.flat_v, self.opt_g.hyper, self.opt_d.flat_p,
                             self.opt_d.flat_m, self.opt_d.flat_v, self.opt_d.hyper, ops.rng_state(sp.device)), snap):
            dst.copy_(src)
        torch.cuda.synchronize()
        if self.world == 1:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.step(static)
            return [graph], static, out
        # world > 1: three graphs sharing one memory pool (replayed in capture order), collectives in between
        ga, gb, gc = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            r, loss_d = self._seg_a(static)
        with torch.cuda.graph(gb, pool=ga.pool()):
            loss_g, parts = self._seg_b(r)
        with torch.cuda.graph(gc, pool=ga.pool()):
            self._seg_c()
        out = self._outputs(loss_d, loss_g, parts, self.opt_d, self.opt_g)
        return [ga, gb, gc], static, out

    def set_lr(self, lr):
        self.lr = lr
        self.opt_g.set_lr(lr)
        self.opt_d.set_lr(lr)

    def decay_lr(self):
        self.set_lr(self.lr * self.t["lr_decay"])


def synthetic_batch(B, T, X, device, seed=1234, hop=640, bank=None, ragged=False):
    """BASELINE config-3 shaped synthetic batch, generated on the HOST (pinned) like a DataLoader would deliver it,
    in the reference's collate layout (data_utils.py:167-226): ssl [B,768,T], wav [B,1,L], text [B,X]."""
    g = torch.Generator().manual_seed(seed)
    L = T * hop
    wav = (torch.rand(B, 1, L, generator=g) - 0.5)
    ssl = torch.randn(B, 768, T, generator=g)
    text = torch.randint(0, 732, (B, X), generator=g)
    if ragged:
        lengths = torch.randint(max(T // 2, 34), T + 1, (B,), generator=g)
        lengths[0] = T
        lengths, _ = torch.sort(lengths, descending=True)
        text_lengths = torch.randint(max(X // 2, 1), X + 1, (B,), generator=g)
        text_lengths[0] = X
    else:
        lengths = torch.full((B,), T, dtype=torch.long)
        text_lengths = torch.full((B,), X, dtype=torch.long)
    return dict(ssl=ssl, wav=wav, text=text, lengths=lengths, text_lengths=text_lengths)


def to_device_batch(host, device, bank, hop=640):
    """H2D + layout change + feature extraction that the reference does on CPU workers (data_utils.py:119-128):
    wav -> |X| on the GPU with the fused mel kernel.  Returns the channels-last dict S2Step consumes."""
    wav = host["wav"].to(device, non_blocking=True)
    ssl = host["ssl"].to(device, non_blocking=True)
    text = host["text"].to(device, non_blocking=True)
    lengths = host["lengths"].to(device, non_blocking=True).to(torch.int32)
    text_lengths = host["text_lengths"].to(device, non_blocking=True).to(torch.int32)
    B, _, L = wav.shape
    spec, _ = ops.mel_frontend(wav.reshape(B, L), bank, hop, want_spec=True, want_mel=False)   # [B,T,1025], pitch 1028
    return dict(ssl=ops.to_channels_last(ssl), spec=spec, lengths=lengths, wav=wav.reshape(B, L, 1), text=text,
                text_lengths=text_lengths)
