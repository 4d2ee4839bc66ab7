"""Stage-1 AR semantic-token GPT on the sm_100a kernels.

Mirror of /root/reference/src/easevoice/soundstorm/auto_reg/models/t2s_model.py `Text2SemanticDecoder` (training
path: forward_old :431-490) with the reference's parameter names, shapes and dtypes, so `state_dict()` is
interchangeable (Lightning checkpoints carry these keys under a "model." prefix, t2s_lightning_module.py:26).

Execution is channels-last [B, L, D] fp32 throughout:
  bert_proj / in_proj / out_proj / linear1(+ReLU) / linear2 / ar_predict_layer  -> ops.linear (tcgen05 TF32 GEMM tiles)
  prefix-LM masked SDPA with probability dropout                               -> ops.flash_attention (fused, O(L) memory)
  residual + post-LayerNorm (transformer.py:300-315, norm_first=False)          -> ops.layernorm(res=...)
  token embeddings, alpha * sinusoid + concat                                   -> ops.embedding / ops.gpt_embed
  CrossEntropyLoss(sum) + top-3 accuracy ignoring EOS                           -> ops.ce_sum_topk
The reference hard-codes dropout 0.1 in the positional embeddings, attention probabilities and the three layer
dropouts even though configs/gpt.yaml says `dropout: 0` (t2s_model.py:276-293); `layer_dropout` reproduces that
(default 0.1) and can be set to 0 for parity runs.
"""
import math

import torch

from . import ops
from .models import ParamTree


def sine_table(length, dim):
    """embedding.py:53-69 computed the same way (fp32 torch ops on the host), uploaded once as a constant."""
    pos = torch.arange(0, length, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * -(math.log(10000.0) / dim))
    pe = torch.zeros(length, dim)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


class Text2SemanticDecoder(ParamTree):
    """t2s_model.py:255-300 (constructor contract: config["model"] keys)."""

    def __init__(self, config, norm_first=False, top_k=3, layer_dropout=0.1, seed=None):
        super().__init__()
        m = config["model"]
        assert not norm_first, "the reference trains the post-LN variant only"
        self.model_dim, self.embedding_dim = m["hidden_dim"], m["embedding_dim"]
        self.num_head, self.num_layers = m["head"], m["n_layer"]
        self.vocab_size, self.phoneme_vocab_size = m["vocab_size"], m["phoneme_vocab_size"]
        self.p_dropout, self.EOS, self.top_k = float(m["dropout"]), m["EOS"], top_k
        self.layer_dropout = float(layer_dropout)
        assert self.EOS == self.vocab_size - 1
        assert self.model_dim == self.embedding_dim and self.model_dim // self.num_head == 32, "head dim 32 kernels"
        D, F = self.model_dim, self.model_dim * 4
        gen = torch.Generator().manual_seed(0 if seed is None else seed)

        def uni(shape, bound):
            return (torch.rand(shape, generator=gen) * 2 - 1) * bound

        def xavier(shape):
            return uni(shape, math.sqrt(6.0 / (shape[0] + shape[1])))

        self._register("bert_proj.weight", uni((D, 1024), 1 / math.sqrt(1024)))
        self._register("bert_proj.bias", uni((D,), 1 / math.sqrt(1024)))
        self._register("ar_text_embedding.word_embeddings.weight", torch.randn((self.phoneme_vocab_size, D), generator=gen))
        self._register("ar_text_position.alpha", torch.ones(1))
        self._register("ar_audio_embedding.word_embeddings.weight", torch.randn((self.vocab_size, D), generator=gen))
        self._register("ar_audio_position.alpha", torch.ones(1))
        for i in range(self.num_layers):
            p = f"h.layers.{i}."
            self._register(p + "self_attn.in_proj_weight", xavier((3 * D, D)))
            self._register(p + "self_attn.in_proj_bias", torch.zeros(3 * D))
            self._register(p + "self_attn.out_proj.weight", uni((D, D), 1 / math.sqrt(D)))
            self._register(p + "self_attn.out_proj.bias", torch.zeros(D))
            self._register(p + "linear1.weight", uni((F, D), 1 / math.sqrt(D)))
            self._register(p + "linear1.bias", uni((F,), 1 / math.sqrt(D)))
            self._register(p + "linear2.weight", uni((D, F), 1 / math.sqrt(F)))
            self._register(p + "linear2.bias", uni((D,), 1 / math.sqrt(F)))
            self._register(p + "norm1.weight", torch.ones(D))
            self._register(p + "norm1.bias", torch.zeros(D))
            self._register(p + "norm2.weight", torch.ones(D))
            self._register(p + "norm2.bias", torch.zeros(D))
        self._register("ar_predict_layer.weight", uni((self.vocab_size, D), 1 / math.sqrt(D)))
        self._pe = None

    def pe(self, length, device):
        if self._pe is None or self._pe.shape[0] < length or self._pe.device != device:
            self._pe = sine_table(max(length, 2048), self.model_dim).to(device)
        return self._pe

    def _drop(self, x, tag):
        return ops.dropout(x, self.layer_dropout, tag) if (self.training and self.layer_dropout > 0) else x

    def make_targets(self, y, y_lens):
        """pad_y_eos (t2s_model.py:557-561) on the host-visible int tensors (exact integer work, tiny)."""
        Y = y.shape[1]
        ymask = (torch.arange(Y, device=y.device)[None, :] >= y_lens[:, None]).to(torch.int64)
        codes = y.to(torch.int64) * (1 - ymask)
        tg = torch.nn.functional.pad(codes, (0, 1), value=0) + self.EOS * torch.nn.functional.pad(ymask, (0, 1), value=1)
        return tg[:, :-1].contiguous(), tg[:, 1:].contiguous()

    def _decode(self, xe, y_in, x_lens, y_lens, X, tagp=""):
        """embedded text prefix + shifted semantic tokens -> padded logits [B, Y, Vp] (t2s_model.py:462-487)."""
        B, Y = y_in.shape
        D, H, dev = self.model_dim, self.num_head, xe.device
        ye = ops.embedding(self.P("ar_audio_embedding.word_embeddings.weight"), y_in)
        h = ops.gpt_embed(xe, ye, self.P("ar_text_position.alpha"), self.P("ar_audio_position.alpha"), self.pe(max(X, Y), dev))
        h = self._drop(h, tagp + "gpt.pos")
        p_attn = self.layer_dropout if self.training else 0.0
        for i in range(self.num_layers):
            p = f"h.layers.{i}."
            qkv = ops.linear(h, self.w(p + "self_attn.in_proj", suffix="_weight"), self.P(p + "self_attn.in_proj_bias"))
            a = ops.flash_attention(qkv, heads=H, prefix=X, xlen=x_lens, ylen=y_lens, p_drop=p_attn, tag=f"{tagp}gpt.attn{i}")
            a = ops.linear(a, self.w(p + "self_attn.out_proj"), self.b(p + "self_attn.out_proj"))
            pd = self.layer_dropout if self.training else 0.0
            # the three layer dropouts are fused: into the two LayerNorm kernels (residual branch) and into linear1's epilogue
            h = ops.layernorm(h, self.P(p + "norm1.weight"), self.P(p + "norm1.bias"), res=a, res_drop=(pd, f"{tagp}gpt.d1.{i}"))
            f = ops.linear(h, self.w(p + "linear1"), self.b(p + "linear1"), act=ops.ACT_RELU, drop=(pd, f"{tagp}gpt.df.{i}"))
            f = ops.linear(f, self.w(p + "linear2"), self.b(p + "linear2"))
            h = ops.layernorm(h, self.P(p + "norm2.weight"), self.P(p + "norm2.bias"), res=f, res_drop=(pd, f"{tagp}gpt.d2.{i}"))
        key = (B, X, str(dev))
        if getattr(self, "_xoff_key", None) != key:
            self._xoff, self._xoff_key = torch.full((B,), X, device=dev, dtype=torch.int64), key
        hy = ops.slice_rows(h, self._xoff, Y)
        Vp = (self.vocab_size + 3) // 4 * 4
        return ops.linear(hy, self.w("ar_predict_layer", pad0=Vp))

    def _embed_text(self, x, bert_feature, bert_channels_last):
        bert_cl = bert_feature if bert_channels_last else ops.to_channels_last(bert_feature)
        xe = ops.embedding(self.P("ar_text_embedding.word_embeddings.weight"), x.to(torch.int64))
        return ops.linear(bert_cl, self.w("bert_proj", need_pb=False), self.b("bert_proj"), res=xe)

    def forward(self, x, x_lens, y, y_lens, bert_feature, reject=None, bert_channels_last=False):
        """DPO variant (t2s_model.py:393-429): CE(sum) on the given y plus a reference-free DPO term (beta 0.2) against
        a synthetically corrupted `reject` = (reject_y, reject_y_lens); built by `make_reject_y` when not given.
        -> (loss, acc)."""
        if reject is None:
            reject = make_reject_y(y, y_lens)
        ry, ryl = reject
        B, X = x.shape
        x_lens, y_lens, ryl = [t.to(torch.int64).contiguous() for t in (x_lens, y_lens, ryl)]
        y_in, tg = self.make_targets(y, y_lens)
        ry_in, rtg = self.make_targets(ry, ryl)
        self.begin_pack()
        try:
            return self._forward_dpo(x, bert_feature, bert_channels_last, y_in, tg, ry_in, rtg, x_lens, y_lens, ryl, X)
        finally:
            self.end_pack()

    def _forward_dpo(self, x, bert_feature, bert_channels_last, y_in, tg, ry_in, rtg, x_lens, y_lens, ryl, X):
        # the text prefix is embedded once per branch in the reference (two make_input_data calls); the branches only share
        # weights, so the second pass re-runs it to keep the dropout streams independent as well
        lc = self._decode(self._embed_text(x, bert_feature, bert_channels_last), y_in, x_lens, y_lens, X)
        lr_ = self._decode(self._embed_text(x, bert_feature, bert_channels_last), ry_in, x_lens, ryl, X, tagp="rej.")
        loss, metrics = ops.dpo_ce(lc, tg, lr_, rtg, self.top_k, self.EOS, V=self.vocab_size, beta=0.2)
        self.last_logits, self.last_dpo = lc.detach(), metrics
        return loss, metrics[2]

    def forward_old(self, x, x_lens, y, y_lens, bert_feature, targets=None, bert_channels_last=False):
        """-> (loss, acc) like t2s_model.py:431-490; `loss` is differentiable, `acc` a device scalar.
        bert_feature: [B, 1024, X] (reference layout), or [B, X, 1024] with bert_channels_last=True."""
        B, X = x.shape
        x_lens, y_lens = x_lens.to(torch.int64).contiguous(), y_lens.to(torch.int64).contiguous()
        y_in, tg = self.make_targets(y, y_lens) if targets is None else targets
        self.begin_pack()                      # all 147 weight packs of the step in one launch (ops.PackPlan)
        try:
            logits = self._decode(self._embed_text(x, bert_feature, bert_channels_last), y_in, x_lens, y_lens, X)
        finally:
            self.end_pack()
        loss, out2 = ops.ce_sum_topk(logits, tg.reshape(-1), self.top_k, self.EOS, V=self.vocab_size)
        self.last_logits = logits.detach()        # detached: a retained graph would pin AccumulateGrad nodes to this stream
        return loss, out2[1]


def make_reject_y(y_o, y_lens, generator=None):
    """utils.py:195-232: per item, duplicate a random span of the PADDED row (the reference's `randint(0, 1)` always picks
    the repeat branch); rows are re-padded with 0 to the longest result.  Host-side integer work, as in the reference."""
    rows, lens = [], []
    for b in range(len(y_lens)):
        y = y_o[b]
        i0, i1 = sorted(torch.randint(0, len(y), size=(2,), generator=generator).tolist())
        rows.append(torch.cat([y[:i0], y[i0:i1], y[i0:i1], y[i1:]]))
        lens.append(len(rows[-1]))
    Ym = max(lens)
    out = torch.zeros((len(rows), Ym), dtype=y_o.dtype, device=y_o.device)
    for b, r in enumerate(rows):
        out[b, :len(r)] = r
    return out, torch.tensor(lens, device=y_lens.device)
