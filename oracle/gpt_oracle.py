"""Oracle (test infrastructure): stage-1 AR semantic-token GPT train step, functional torch fp32.

Restates, on a flat {state_dict_key: tensor} dict (keys as in Text2SemanticDecoder.state_dict(), without the
Lightning "model." prefix):
  src/easevoice/soundstorm/auto_reg/models/t2s_model.py   forward_old :431-490, pad_y_eos :557-561
  .../models/utils.py                                     make_pad_mask :16-41
  .../modules/embedding.py                                TokenEmbedding :8-33, SinePositionalEmbedding :36-81
  .../modules/transformer.py                              TransformerEncoderLayer.forward :266-315 (post-LN, ReLU)
  .../modules/patched_mha_with_cache.py                   multi_head_attention_forward_patched :14-465 (SDPA, additive mask)
  .../modules/optim.py                                    ScaledAdam :123-622 (per-tensor restatement; the reference stacks
                                                          tensors of equal shape, which does not change the per-tensor math)
  .../models/t2s_lightning_module.py                      training_step :41-89 (manual optimisation, step when batch_idx>0 and %4==0)
  .../modules/lr_schedulers.py                            :36-65 (lr is 0.01 for the first optimizer step, then locked to 0.002)
Dropout is off (parity configuration); the reference hard-codes 0.1 in the layers even though gpt.yaml says 0.
"""
import math

import torch
import torch.nn.functional as F

GPT_MODEL = dict(vocab_size=1025, phoneme_vocab_size=732, embedding_dim=512, hidden_dim=512, head=16, linear_units=2048,
                 n_layer=24, dropout=0, EOS=1024, random_bert=0)


def gpt_param_spec(m=GPT_MODEL):
    D, H = m["hidden_dim"], m["linear_units"]
    s = {"bert_proj.weight": (D, 1024), "bert_proj.bias": (D,),
         "ar_text_embedding.word_embeddings.weight": (m["phoneme_vocab_size"], D), "ar_text_position.alpha": (1,),
         "ar_audio_embedding.word_embeddings.weight": (m["vocab_size"], D), "ar_audio_position.alpha": (1,)}
    for i in range(m["n_layer"]):
        p = f"h.layers.{i}."
        s[p + "self_attn.in_proj_weight"] = (3 * D, D); s[p + "self_attn.in_proj_bias"] = (3 * D,)
        s[p + "self_attn.out_proj.weight"] = (D, D); s[p + "self_attn.out_proj.bias"] = (D,)
        s[p + "linear1.weight"] = (H, D); s[p + "linear1.bias"] = (H,)
        s[p + "linear2.weight"] = (D, H); s[p + "linear2.bias"] = (D,)
        s[p + "norm1.weight"] = (D,); s[p + "norm1.bias"] = (D,)
        s[p + "norm2.weight"] = (D,); s[p + "norm2.bias"] = (D,)
    s["ar_predict_layer.weight"] = (m["vocab_size"], D)
    return s


def init_params(spec, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shape in spec.items():
        if k.endswith("alpha"):
            t = torch.ones(shape)
        elif k.endswith(("norm1.weight", "norm2.weight")):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif k.endswith(".bias") or k.endswith("in_proj_bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif "word_embeddings" in k:
            t = torch.randn(shape, generator=g)
        else:
            t = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        out[k] = t.float()
    return out


def sine_pe(L, D, device=None):
    """embedding.py:53-69: interleaved sin/cos, base 10000."""
    pos = torch.arange(0, L, dtype=torch.float32, device=device).unsqueeze(1)
    div = torch.exp(torch.arange(0, D, 2, dtype=torch.float32, device=device) * -(math.log(10000.0) / D))
    pe = torch.zeros(L, D, device=device)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def prefix_lm_mask(x_lens, y_lens, X, Y):
    """t2s_model.py:456-479: True = masked.  [B, X+Y, X+Y]"""
    dev = x_lens.device
    xm = torch.arange(X, device=dev)[None, :] >= x_lens[:, None]
    ym = torch.arange(Y, device=dev)[None, :] >= y_lens[:, None]
    pad = torch.cat([xm, ym], 1)                                                # key padding
    xa = F.pad(torch.zeros(X, X, dtype=torch.bool, device=dev), (0, Y), value=True)
    ya = F.pad(torch.triu(torch.ones(Y, Y, dtype=torch.bool, device=dev), diagonal=1), (X, 0), value=False)
    return torch.cat([xa, ya], 0)[None] | pad[:, None, :]


def forward_old(P, x, x_lens, y, y_lens, bert, m=GPT_MODEL, taps=None):
    """-> (loss (sum CE), top-3 accuracy ignoring EOS targets, logits [B, Y, V], targets [B, Y])."""
    D, H = m["hidden_dim"], m["head"]
    B, X = x.shape
    Y = y.shape[1]
    xe = F.embedding(x, P["ar_text_embedding.word_embeddings.weight"]) + F.linear(bert.transpose(1, 2), P["bert_proj.weight"], P["bert_proj.bias"])
    pe = sine_pe(max(X, Y), D, xe.device)
    xe = xe + P["ar_text_position.alpha"] * pe[:X]
    ymask = (torch.arange(Y, device=y.device)[None, :] >= y_lens[:, None]).long()
    codes = y.long() * (1 - ymask)
    tg = F.pad(codes, (0, 1), value=0) + m["EOS"] * F.pad(ymask, (0, 1), value=1)      # pad_y_eos
    y_in, targets = tg[:, :-1], tg[:, 1:]
    ye = F.embedding(y_in, P["ar_audio_embedding.word_embeddings.weight"]) + P["ar_audio_position.alpha"] * pe[:Y]
    h = torch.cat([xe, ye], 1)
    if taps is not None:
        h.retain_grad()
        taps["h0"] = h
    mask = prefix_lm_mask(x_lens, y_lens, X, Y)
    add = torch.zeros(mask.shape, dtype=h.dtype, device=h.device).masked_fill(mask, float("-inf")).unsqueeze(1)   # [B,1,L,L]
    L = X + Y
    dk = D // H
    for i in range(m["n_layer"]):
        p = f"h.layers.{i}."
        qkv = F.linear(h, P[p + "self_attn.in_proj_weight"], P[p + "self_attn.in_proj_bias"])
        q, k, v = [t.view(B, L, H, dk).transpose(1, 2) for t in qkv.split(D, dim=-1)]
        att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dk) + add, dim=-1) @ v
        att = F.linear(att.transpose(1, 2).reshape(B, L, D), P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"])
        h = F.layer_norm(h + att, (D,), P[p + "norm1.weight"], P[p + "norm1.bias"], 1e-5)
        ff = F.linear(torch.relu(F.linear(h, P[p + "linear1.weight"], P[p + "linear1.bias"])), P[p + "linear2.weight"], P[p + "linear2.bias"])
        h = F.layer_norm(h + ff, (D,), P[p + "norm2.weight"], P[p + "norm2.bias"], 1e-5)
    logits = F.linear(h[:, X:], P["ar_predict_layer.weight"])
    loss = F.cross_entropy(logits.permute(0, 2, 1), targets, reduction="sum")
    top3 = logits.detach().topk(3, dim=-1).indices
    valid = targets != m["EOS"]
    acc = ((top3 == targets.unsqueeze(-1)).any(-1) & valid).sum().float() / valid.sum().clamp(min=1).float()
    return loss, acc, logits, targets


def logits_to_probs(logits, previous_tokens=None, temperature=1.0, top_k=None, top_p=None, repetition_penalty=1.0):
    """utils.py:109-145.  The repetition penalty is written into `logits` IN PLACE (scatter_), as in the reference."""
    if previous_tokens is not None and repetition_penalty != 1.0:
        previous_tokens = previous_tokens.long()
        score = torch.gather(logits, dim=1, index=previous_tokens)
        score = torch.where(score < 0, score * repetition_penalty, score / repetition_penalty)
        logits.scatter_(dim=1, index=previous_tokens, src=score)
    if top_p is not None and top_p < 1.0:
        sl, si = torch.sort(logits, descending=True)
        remove = torch.cumsum(torch.softmax(sl, dim=-1), dim=-1) > top_p
        remove[:, 0] = False
        logits = logits.masked_fill(remove.scatter(dim=1, index=si, src=remove), -float("inf"))
    logits = logits / max(temperature, 1e-5)
    if top_k is not None:
        v, _ = torch.topk(logits, min(top_k, logits.size(-1)))
        logits = torch.where(logits < v[:, -1].unsqueeze(-1), -float("inf"), logits)
    return torch.softmax(logits, dim=-1)


def infer_panel(P, x, bert, prompts, top_k=-100, top_p=100, early_stop_num=-1, temperature=1.0, repetition_penalty=1.35,
                m=GPT_MODEL, max_steps=1500, trace=None):
    """Text2SemanticDecoder.infer_panel_naive, t2s_model.py:762-867, B = 1, prompts given.  No KV cache: every step re-runs
    the stack on the whole sequence under the prefix-LM mask (text rows see text only, audio rows see text + earlier audio),
    which is what process_prompt + decode_next_token compute incrementally.  -> (y[:, :-1], idx - 1); `trace` receives the raw
    [1, V] logits of every step."""
    D, H = m["hidden_dim"], m["head"]
    dk = D // H
    X = x.shape[1]
    y = prompts.long()
    Yp = y.shape[1]
    xe = F.embedding(x, P["ar_text_embedding.word_embeddings.weight"]) + F.linear(bert.transpose(1, 2), P["bert_proj.weight"], P["bert_proj.bias"])
    pe = sine_pe(max(X, Yp + max_steps + 2), D)
    xe = xe + P["ar_text_position.alpha"] * pe[:X]
    stop, idx = False, 0
    for idx in range(max_steps):
        Y = y.shape[1]
        ye = F.embedding(y, P["ar_audio_embedding.word_embeddings.weight"]) + P["ar_audio_position.alpha"] * pe[:Y]
        h = torch.cat([xe, ye], 1)
        mask = prefix_lm_mask(torch.tensor([X]), torch.tensor([Y]), X, Y)
        add = torch.zeros(mask.shape).masked_fill(mask, float("-inf")).unsqueeze(1)
        L = X + Y
        for i in range(m["n_layer"]):
            p = f"h.layers.{i}."
            qkv = F.linear(h, P[p + "self_attn.in_proj_weight"], P[p + "self_attn.in_proj_bias"])
            q, k, v = [t.view(1, L, H, dk).transpose(1, 2) for t in qkv.split(D, dim=-1)]
            att = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dk) + add, dim=-1) @ v
            att = F.linear(att.transpose(1, 2).reshape(1, L, D), P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"])
            h = F.layer_norm(h + att, (D,), P[p + "norm1.weight"], P[p + "norm1.bias"], 1e-5)
            ff = F.linear(torch.relu(F.linear(h, P[p + "linear1.weight"], P[p + "linear1.bias"])), P[p + "linear2.weight"], P[p + "linear2.bias"])
            h = F.layer_norm(h + ff, (D,), P[p + "norm2.weight"], P[p + "norm2.bias"], 1e-5)
        logits = F.linear(h[:, -1], P["ar_predict_layer.weight"])
        if trace is not None:
            trace.append(logits.clone())
        if idx < 11:
            logits = logits[:, :-1]
        probs = logits_to_probs(logits, y, temperature=temperature, top_k=top_k, top_p=top_p, repetition_penalty=repetition_penalty)
        q = torch.empty_like(probs).exponential_(1)
        samples = torch.argmax(probs / q, dim=-1, keepdim=True).long()
        y = torch.cat([y, samples], dim=1)
        if early_stop_num != -1 and (y.shape[1] - Yp) > early_stop_num:
            stop = True
        if int(torch.argmax(logits, dim=-1)[0]) == m["EOS"] or int(samples[0, 0]) == m["EOS"]:
            stop = True
        if stop:
            break
    return y[:, :-1], idx - 1


def forward_dpo(P, x, x_lens, y, y_lens, bert, reject_y, reject_lens, m=GPT_MODEL, beta=0.2):
    """t2s_model.py:393-429 with the rejected batch given (make_reject_y draws it at random, utils.py:195-232);
    dpo_loss / get_batch_logps: utils.py:160-192 (reference_free=True)."""
    loss1, acc, logits, targets = forward_old(P, x, x_lens, y, y_lens, bert, m)
    _, _, rlogits, rtargets = forward_old(P, x, x_lens, reject_y, reject_lens, bert, m)
    A = torch.gather(logits.log_softmax(-1), 2, targets.unsqueeze(2)).squeeze(2).sum(-1)
    R = torch.gather(rlogits.log_softmax(-1), 2, rtargets.unsqueeze(2)).squeeze(2).sum(-1)
    loss2 = (-F.logsigmoid(beta * (A - R))).mean()
    return loss1 + loss2, acc, loss1, loss2


def make_reject_given(y, spans):
    """the repeat_P corruption of utils.py:196-202 with the random span endpoints supplied."""
    rows = [torch.cat([y[b][:i0], y[b][i0:i1], y[b][i0:i1], y[b][i1:]]) for b, (i0, i1) in enumerate(spans)]
    lens = [len(r) for r in rows]
    out = torch.zeros((len(rows), max(lens)), dtype=y.dtype)
    for b, r in enumerate(rows):
        out[b, :len(r)] = r
    return out, torch.tensor(lens)


class ScaledAdamOracle:
    """Per-tensor restatement of ScaledAdam as configured at t2s_lightning_module.py:100-108."""

    def __init__(self, params, lr=0.01, betas=(0.9, 0.95), clipping_scale=2.0, clipping_update_period=1000, scalar_lr_scale=0.1,
                 eps=1e-8, param_min_rms=1e-5, param_max_rms=3.0, scalar_max=10.0, size_update_period=4):
        self.params = list(params)
        self.lr, self.betas, self.clip, self.period = lr, betas, clipping_scale, clipping_update_period
        self.slr, self.eps, self.rmin, self.rmax, self.smax, self.sup = scalar_lr_scale, eps, param_min_rms, param_max_rms, scalar_max, size_update_period
        self.state = [None] * len(self.params)
        self.model_norms = torch.zeros(clipping_update_period)
        self.threshold = None

    @torch.no_grad()
    def step(self, grads):
        b1, b2 = self.betas
        first = self.state[0] is None
        # ---- clipping scale (optim.py:300-390)
        cs = 1.0
        if not first:
            step0 = self.state[0]["step"]
            if step0 > 0:
                tot = 0.0
                for p, g, st in zip(self.params, grads, self.state):
                    tot = tot + ((g ** 2).sum() if p.numel() == 1 else ((g * st["rms"]) ** 2).sum())
                tot_norm = tot.sqrt()
                self.model_norms[step0 % self.period] = tot_norm
                if step0 % self.period == 0:
                    srt = self.model_norms.sort()[0]
                    self.threshold = self.clip * float(srt[min(self.period - 1, (self.period // 4) * 2)])
                if step0 >= self.period and self.threshold is not None:
                    cs = min(1.0, self.threshold / (float(tot_norm) + 1e-20))
        for i, (p, g) in enumerate(zip(self.params, grads)):
            if self.state[i] is None:
                st = dict(step=0, delta=torch.zeros_like(p), v=torch.zeros_like(p))
                if p.numel() > 1:
                    st.update(rms=(p ** 2).mean().sqrt(), sv=torch.zeros(()), sg=torch.zeros(self.sup))
                self.state[i] = st
            st = self.state[i]
            # optim.py:466-468 scales a LOCAL copy of the gradient; _step/_step_scalar re-read p.grad (:573,:610), so in the
            # reference the clipping scale only reaches the size-update statistics.  Restated as is.
            gc = g * cs if cs != 1.0 else g
            step = st["step"]
            st["delta"].mul_(b1)
            if p.numel() > 1:
                st["sg"][step % self.sup] = (p * gc).sum()
                if step % self.sup == self.sup - 1:
                    st["rms"] = (p ** 2).mean().sqrt()
                    if step > 0:                                            # _size_update, optim.py:499-558
                        b2c = b2 ** self.sup
                        st["sv"] = st["sv"] * b2c + (st["sg"] ** 2).mean() * (1 - b2c)
                        size_step = (step + 1) // self.sup
                        bc2 = 1 - b2c ** size_step
                        sstep = -(self.lr * self.slr) * (bc2 ** 0.5) * st["sg"].sum() / (st["sv"].sqrt() + self.eps)
                        if st["rms"] < self.rmin:
                            sstep = torch.zeros(())
                        if st["rms"] > self.rmax:
                            sstep = torch.tensor(-(self.lr * self.slr) * self.sup)
                        st["delta"].add_(p * sstep, alpha=1 - b1)
                st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)                 # _step, optim.py:560-598
                bc2 = 1 - b2 ** (step + 1)
                v = st["v"] * (1.0 / bc2) if bc2 < 0.99 else st["v"]
                alpha = -self.lr * (1 - b1) * st["rms"].clamp(min=self.rmin)
                st["delta"].add_(g / (v.sqrt() + self.eps) * alpha)
                p.add_(st["delta"])
            else:                                                           # _step_scalar, optim.py:600-622
                st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
                bc2 = 1 - b2 ** (step + 1)
                st["delta"].add_(g / ((st["v"] / bc2).sqrt() + self.eps), alpha=-(self.lr * self.slr) * (1 - b1))
                p.clamp_(min=-self.smax, max=self.smax)
                p.add_(st["delta"])
            st["step"] = step + 1
        return cs


def synthetic_gpt_batch(B, X, Y, seed, ragged=False):
    """BASELINE config-2 shaped batch (dataset.py:226-271 collate layout)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 732, (B, X), generator=g)
    y = torch.randint(0, 1024, (B, Y), generator=g)
    bert = torch.randn(B, 1024, X, generator=g)
    if ragged:
        xl = torch.randint(max(X // 2, 1), X + 1, (B,), generator=g); xl[0] = X
        yl = torch.randint(max(Y // 2, 1), Y + 1, (B,), generator=g); yl[0] = Y
        for b in range(B):
            y[b, yl[b]:] = 1024
            bert[b, :, xl[b]:] = 0
    else:
        xl, yl = torch.full((B,), X), torch.full((B,), Y)
    return x, xl, y, yl, bert
