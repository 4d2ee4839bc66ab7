"""Oracle (test infrastructure): stage-2 SoVITS + HiFi-GAN train step, functional torch fp32.

A compact functional restatement of the reference's stage-2 networks and losses
operating on a flat ``{state_dict_key: tensor}`` dict, in the reference's
[B, C, T] layout.  All randomness is injected (posterior noise, slice ids);
dropout is off (the parity configuration, SURVEY.md section 7 "Hard parts").

Reference (all under /root/reference/src/easevoice/module/):
  models.py    SynthesizerTrn.forward :904-946, TextEncoder :228-251, PosteriorEncoder :348-359,
               ResidualCouplingBlock :308-315, Generator :452-471, DiscriminatorP :538-557,
               DiscriminatorS :576-587, MultiPeriodDiscriminator :601-614
  modules.py   LayerNorm :28-31, WN :187-212, ResBlock1 :298-311, ResidualCouplingLayer :439-458,
               Flip :376-383, MelStyleEncoder :739-763 (+ MultiHeadAttention :627-657,
               ScaledDotProductAttention :669-682, Conv1dGLU :553-559, Mish :542-543)
  attentions.py Encoder.forward :67-90, MultiHeadAttention :233-292 (+ rel-pos helpers :312-365), FFN :408-416
  mrte_model.py MRTE.forward :25-61
  core_vq.py   EuclideanCodebook.quantize :172-180, VectorQuantization.forward :300-320 (eval branch)
  commons.py   slice_segments :42-48, sequence_mask :115-119, fused_add_tanh_sigmoid_multiply :94-101
  losses.py    feature_loss :7-15, discriminator_loss :18-31, generator_loss :34-43, kl_loss :46-61
and the step in /root/reference/src/train/sovits.py:459-525.
"""
import math
import torch
import torch.nn.functional as F

from . import mel_oracle

LRELU = 0.1

S2_MODEL = dict(
    inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2, n_layers=6,
    kernel_size=3, p_dropout=0.1, resblock="1", resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], upsample_rates=[10, 8, 2, 2, 2],
    upsample_initial_channel=512, upsample_kernel_sizes=[16, 16, 8, 2, 2], gin_channels=512,
    semantic_frame_rate="25hz", freeze_quantizer=True)
S2_DATA = dict(sampling_rate=32000, filter_length=2048, hop_length=640, win_length=2048,
               n_mel_channels=128, mel_fmin=0.0, mel_fmax=None)
S2_TRAIN = dict(segment_size=20480, c_mel=45, c_kl=1.0, learning_rate=1e-4, betas=(0.8, 0.99),
                eps=1e-9, lr_decay=0.999875, text_low_lr_rate=0.4)
N_SYMBOLS = 732           # len(SYMBOLS), text/symbols.py:410-412
PERIODS = [2, 3, 5, 7, 11]


# ----------------------------------------------------------------------------
# parameter specs (names/shapes must equal the reference's state_dict)
# ----------------------------------------------------------------------------
def _conv(spec, name, cout, cin_g, k, bias=True, wn=False, wn_dim0=None):
    if wn:
        spec[name + ".weight_g"] = (wn_dim0 if wn_dim0 is not None else cout, 1, 1)
        spec[name + ".weight_v"] = (cout, cin_g, k) if wn_dim0 is None else (wn_dim0, cin_g, k)
    else:
        spec[name + ".weight"] = (cout, cin_g, k)
    if bias:
        spec[name + ".bias"] = (cout,)


def _attn_encoder_spec(spec, pfx, n_layers, hidden=192, filt=768, k=3, window=4, heads=2):
    for i in range(n_layers):
        a = f"{pfx}.attn_layers.{i}"
        spec[a + ".emb_rel_k"] = (1, 2 * window + 1, hidden // heads)
        spec[a + ".emb_rel_v"] = (1, 2 * window + 1, hidden // heads)
        for c in "qkvo":
            _conv(spec, f"{a}.conv_{c}", hidden, hidden, 1)
        spec[f"{pfx}.norm_layers_1.{i}.gamma"] = (hidden,)
        spec[f"{pfx}.norm_layers_1.{i}.beta"] = (hidden,)
        _conv(spec, f"{pfx}.ffn_layers.{i}.conv_1", filt, hidden, k)
        _conv(spec, f"{pfx}.ffn_layers.{i}.conv_2", hidden, filt, k)
        spec[f"{pfx}.norm_layers_2.{i}.gamma"] = (hidden,)
        spec[f"{pfx}.norm_layers_2.{i}.beta"] = (hidden,)


def _wn_spec(spec, pfx, hidden, k, n_layers, gin):
    spec[pfx + ".cond_layer.bias"] = (2 * hidden * n_layers,)
    spec[pfx + ".cond_layer.weight_g"] = (2 * hidden * n_layers, 1, 1)
    spec[pfx + ".cond_layer.weight_v"] = (2 * hidden * n_layers, gin, 1)
    for i in range(n_layers):
        spec[f"{pfx}.in_layers.{i}.bias"] = (2 * hidden,)
        spec[f"{pfx}.in_layers.{i}.weight_g"] = (2 * hidden, 1, 1)
        spec[f"{pfx}.in_layers.{i}.weight_v"] = (2 * hidden, hidden, k)
    for i in range(n_layers):
        rs = 2 * hidden if i < n_layers - 1 else hidden
        spec[f"{pfx}.res_skip_layers.{i}.bias"] = (rs,)
        spec[f"{pfx}.res_skip_layers.{i}.weight_g"] = (rs, 1, 1)
        spec[f"{pfx}.res_skip_layers.{i}.weight_v"] = (rs, hidden, 1)


def generator_param_spec(m=S2_MODEL, spec_channels=1025):
    """name -> shape for SynthesizerTrn (parameters AND buffers), models.py:854-902."""
    s = {}
    H, I, gin = m["hidden_channels"], m["inter_channels"], m["gin_channels"]
    # enc_p (TextEncoder, models.py:199-226)
    _conv(s, "enc_p.ssl_proj", H, 768, 1)
    _attn_encoder_spec(s, "enc_p.encoder_ssl", m["n_layers"] // 2, H, m["filter_channels"], m["kernel_size"])
    _attn_encoder_spec(s, "enc_p.encoder_text", m["n_layers"], H, m["filter_channels"], m["kernel_size"])
    s["enc_p.text_embedding.weight"] = (N_SYMBOLS, H)
    for c in "qkvo":
        _conv(s, f"enc_p.mrte.cross_attention.conv_{c}", 512, 512, 1)
    _conv(s, "enc_p.mrte.c_pre", 512, H, 1)
    _conv(s, "enc_p.mrte.text_pre", 512, H, 1)
    _conv(s, "enc_p.mrte.c_post", H, 512, 1)
    _attn_encoder_spec(s, "enc_p.encoder2", m["n_layers"] // 2, H, m["filter_channels"], m["kernel_size"])
    _conv(s, "enc_p.proj", 2 * I, H, 1)
    # dec (Generator, models.py:416-450)
    C0 = m["upsample_initial_channel"]
    _conv(s, "dec.conv_pre", C0, I, 7)
    for i, (u, k) in enumerate(zip(m["upsample_rates"], m["upsample_kernel_sizes"])):
        cin, cout = C0 // 2 ** i, C0 // 2 ** (i + 1)
        s[f"dec.ups.{i}.bias"] = (cout,)
        s[f"dec.ups.{i}.weight_g"] = (cin, 1, 1)
        s[f"dec.ups.{i}.weight_v"] = (cin, cout, k)
    for i in range(len(m["upsample_rates"])):
        ch = C0 // 2 ** (i + 1)
        for j, k in enumerate(m["resblock_kernel_sizes"]):
            r = f"dec.resblocks.{i * len(m['resblock_kernel_sizes']) + j}"
            for grp in ("convs1", "convs2"):
                for l in range(3):
                    s[f"{r}.{grp}.{l}.bias"] = (ch,)
                    s[f"{r}.{grp}.{l}.weight_g"] = (ch, 1, 1)
                    s[f"{r}.{grp}.{l}.weight_v"] = (ch, ch, k)
    _conv(s, "dec.conv_post", 1, C0 // 2 ** len(m["upsample_rates"]), 7, bias=False)
    _conv(s, "dec.cond", C0, gin, 1)
    # enc_q (PosteriorEncoder, models.py:338-346)
    _conv(s, "enc_q.pre", H, spec_channels, 1)
    _wn_spec(s, "enc_q.enc", H, 5, 16, gin)
    _conv(s, "enc_q.proj", 2 * I, H, 1)
    # flow (ResidualCouplingBlock, models.py:293-306)
    for f in range(4):
        p = f"flow.flows.{2 * f}"
        _conv(s, p + ".pre", H, I // 2, 1)
        _wn_spec(s, p + ".enc", H, 5, 4, gin)
        _conv(s, p + ".post", I // 2, H, 1)
    # ref_enc (MelStyleEncoder(704, style_vector_dim=512), modules.py:700-737)
    s["ref_enc.spectral.0.fc.weight"] = (128, 704); s["ref_enc.spectral.0.fc.bias"] = (128,)
    s["ref_enc.spectral.3.fc.weight"] = (128, 128); s["ref_enc.spectral.3.fc.bias"] = (128,)
    for i in range(2):
        s[f"ref_enc.temporal.{i}.conv1.conv.weight"] = (256, 128, 5)
        s[f"ref_enc.temporal.{i}.conv1.conv.bias"] = (256,)
    for n in ("w_qs", "w_ks", "w_vs", "fc"):
        s[f"ref_enc.slf_attn.{n}.weight"] = (128, 128); s[f"ref_enc.slf_attn.{n}.bias"] = (128,)
    s["ref_enc.fc.fc.weight"] = (gin, 128); s["ref_enc.fc.fc.bias"] = (gin,)
    # top level
    _conv(s, "ssl_proj", 768, 768, 2)
    cb = "quantizer.vq.layers.0._codebook."
    s[cb + "inited"] = (1,); s[cb + "cluster_size"] = (1024,)
    s[cb + "embed"] = (1024, 768); s[cb + "embed_avg"] = (1024, 768)
    return s


GEN_BUFFERS = tuple("quantizer.vq.layers.0._codebook." + n for n in ("inited", "cluster_size", "embed", "embed_avg"))


def discriminator_param_spec():
    """name -> shape for MultiPeriodDiscriminator, models.py:481-614 (all weight-normed)."""
    s = {}
    ds = [(16, 1, 15), (64, 4, 41), (256, 4, 41), (1024, 4, 41), (1024, 4, 41), (1024, 1024, 5)]
    for i, (co, cig, k) in enumerate(ds):
        p = f"discriminators.0.convs.{i}"
        s[p + ".bias"] = (co,); s[p + ".weight_g"] = (co, 1, 1); s[p + ".weight_v"] = (co, cig, k)
    p = "discriminators.0.conv_post"
    s[p + ".bias"] = (1,); s[p + ".weight_g"] = (1, 1, 1); s[p + ".weight_v"] = (1, 1024, 3)
    chans = [(32, 1), (128, 32), (512, 128), (1024, 512), (1024, 1024)]
    for d in range(1, 6):
        for i, (co, ci) in enumerate(chans):
            p = f"discriminators.{d}.convs.{i}"
            s[p + ".bias"] = (co,); s[p + ".weight_g"] = (co, 1, 1, 1); s[p + ".weight_v"] = (co, ci, 5, 1)
        p = f"discriminators.{d}.conv_post"
        s[p + ".bias"] = (1,); s[p + ".weight_g"] = (1, 1, 1, 1); s[p + ".weight_v"] = (1, 1024, 3, 1)
    return s


def init_params(spec, seed, requires_grad=False):
    """Deterministic synthetic weights (CPU generator => identical on every host).

    Not the reference's init distributions -- parity does not depend on them; scales are
    chosen so activations stay O(1) through the deep stacks.
    """
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in spec.items():
        if name.endswith("inited"):
            t = torch.ones(shape)
        elif name.endswith("cluster_size"):
            t = torch.ones(shape)
        elif name.endswith("weight_g"):
            t = 0.5 + torch.rand(shape, generator=g)
            if ".resblocks." in name or "res_skip_layers" in name:
                t = 0.3 * t        # keep the residual stacks (and tanh / exp(logs)) out of saturation
        elif name.endswith(".gamma"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith((".bias", ".beta")):
            t = 0.05 * torch.randn(shape, generator=g)
        elif "emb_rel" in name:
            t = torch.randn(shape, generator=g) * shape[-1] ** -0.5
        elif "_codebook.embed" in name:
            t = torch.randn(shape, generator=g)
        elif name.endswith("weight_v"):
            t = torch.randn(shape, generator=g)
        elif name.endswith("text_embedding.weight"):
            t = torch.randn(shape, generator=g)
        else:  # plain conv / linear weight: fan-in scaled
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(max(fan_in, 1)))
            if name in ("enc_q.proj.weight", "enc_p.proj.weight", "dec.conv_post.weight"):
                t = 0.25 * t
            if name == "ref_enc.fc.fc.weight":
                t = 0.1 * t
        out[name] = t.float()
    # flow post convs are zero-initialised in the reference (modules.py:436-437); use small
    # non-zero values instead so the flow's backward path is exercised by parity tests.
    if requires_grad:
        for name, t in out.items():
            if name not in GEN_BUFFERS:
                t.requires_grad_(True)
    return out


# ----------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------
def wn_weight(P, pfx):
    """torch.nn.utils.weight_norm(dim=0): w = g * v / ||v|| (norm over all dims but 0)."""
    v, g = P[pfx + ".weight_v"], P[pfx + ".weight_g"]
    n = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def _w(P, pfx):
    return P[pfx + ".weight"] if (pfx + ".weight") in P else wn_weight(P, pfx)


def conv1d(P, pfx, x, stride=1, padding=0, dilation=1, groups=1):
    return F.conv1d(x, _w(P, pfx), P.get(pfx + ".bias"), stride, padding, dilation, groups)


def sequence_mask(lengths, T):
    return (torch.arange(T, device=lengths.device)[None, :] < lengths[:, None])


def chan_layernorm(P, pfx, x, eps=1e-5):
    """modules.py:28-31: LayerNorm over the channel dim of [B, C, T]."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), P[pfx + ".gamma"], P[pfx + ".beta"], eps).transpose(1, -1)


def relpos_attention(P, pfx, x, c, attn_mask, n_heads, window):
    """attentions.py:233-292.  x,c [B, C, T]; attn_mask [B, 1, T_t, T_s] (0 => fill -1e4)."""
    q, k, v = conv1d(P, pfx + ".conv_q", x), conv1d(P, pfx + ".conv_k", c), conv1d(P, pfx + ".conv_v", c)
    B, C, Tt = q.shape
    Ts = k.shape[2]
    dk = C // n_heads
    q = q.view(B, n_heads, dk, Tt).transpose(2, 3)
    k = k.view(B, n_heads, dk, Ts).transpose(2, 3)
    v = v.view(B, n_heads, dk, Ts).transpose(2, 3)
    qs = q / math.sqrt(dk)
    scores = qs @ k.transpose(-2, -1)
    if window is not None:
        Ek = P[pfx + ".emb_rel_k"][0]                     # [2w+1, dk], shared across heads
        rel = qs @ Ek.t()                                 # [B, H, T, 2w+1]
        add = torch.zeros_like(scores)
        for r in range(-window, window + 1):              # scores[i, i+r] += rel[i, r+w]
            n = Tt - abs(r)
            if n <= 0:
                continue
            i0 = max(0, -r)
            idx = torch.arange(i0, i0 + n, device=scores.device)
            add[:, :, idx, idx + r] = rel[:, :, idx, r + window]
        scores = scores + add
    scores = scores.masked_fill(attn_mask == 0, -1e4)
    p = F.softmax(scores, dim=-1)
    out = p @ v
    if window is not None:
        Ev = P[pfx + ".emb_rel_v"][0]
        for r in range(-window, window + 1):              # out[i] += p[i, i+r] * Ev[r+w]
            n = Tt - abs(r)
            if n <= 0:
                continue
            i0 = max(0, -r)
            idx = torch.arange(i0, i0 + n, device=p.device)
            contrib = p[:, :, idx, idx + r].unsqueeze(-1) * Ev[r + window]
            out = out.index_add(2, idx, contrib.to(out.dtype))      # .to(): no-op in fp32; keeps the autocast comparator of bench.py running
    out = out.transpose(2, 3).contiguous().view(B, C, Tt)
    return conv1d(P, pfx + ".conv_o", out)


def ffn(P, pfx, x, x_mask, k):
    """attentions.py:408-416 (same-padding, ReLU)."""
    pl, pr = (k - 1) // 2, k // 2
    h = conv1d(P, pfx + ".conv_1", F.pad(x * x_mask, (pl, pr)))
    h = torch.relu(h)
    h = conv1d(P, pfx + ".conv_2", F.pad(h * x_mask, (pl, pr)))
    return h * x_mask


def attn_encoder(P, pfx, x, x_mask, n_layers, n_heads=2, window=4, k=3):
    """attentions.py:67-90 (dropout off)."""
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    for i in range(n_layers):
        y = relpos_attention(P, f"{pfx}.attn_layers.{i}", x, x, attn_mask, n_heads, window)
        x = chan_layernorm(P, f"{pfx}.norm_layers_1.{i}", x + y)
        y = ffn(P, f"{pfx}.ffn_layers.{i}", x, x_mask, k)
        x = chan_layernorm(P, f"{pfx}.norm_layers_2.{i}", x + y)
    return x * x_mask


def mrte(P, pfx, ssl_enc, ssl_mask, text, text_mask, ge):
    """mrte_model.py:25-61 (test=None branch)."""
    attn_mask = text_mask.unsqueeze(2) * ssl_mask.unsqueeze(-1)
    ssl_enc = conv1d(P, pfx + ".c_pre", ssl_enc * ssl_mask)
    text_enc = conv1d(P, pfx + ".text_pre", text * text_mask)
    a = relpos_attention(P, pfx + ".cross_attention", ssl_enc * ssl_mask, text_enc * text_mask, attn_mask, 4, None)
    x = a + ssl_enc + ge
    return conv1d(P, pfx + ".c_post", x * ssl_mask)


def mish(x):
    return x * torch.tanh(F.softplus(x))


def mel_style_encoder(P, pfx, x, mask):
    """modules.py:739-763.  x [B, 704, T] (already masked), mask [B, 1, T] float -> [B, 512, 1]."""
    x = x.transpose(1, 2)
    pad = (mask == 0).squeeze(1)                                    # True where padded
    x = mish(F.linear(x, P[pfx + ".spectral.0.fc.weight"], P[pfx + ".spectral.0.fc.bias"]))
    x = mish(F.linear(x, P[pfx + ".spectral.3.fc.weight"], P[pfx + ".spectral.3.fc.bias"]))
    x = x.transpose(1, 2)
    for i in range(2):                                              # Conv1dGLU, modules.py:553-559
        h = F.conv1d(x, P[f"{pfx}.temporal.{i}.conv1.conv.weight"], P[f"{pfx}.temporal.{i}.conv1.conv.bias"], padding=2)
        h1, h2 = h.split(128, dim=1)
        x = x + h1 * torch.sigmoid(h2)
    x = x.transpose(1, 2).masked_fill(pad.unsqueeze(-1), 0)
    B, T, _ = x.shape
    res = x                                                          # MultiHeadAttention, modules.py:627-657
    def heads(n):
        y = F.linear(x, P[f"{pfx}.slf_attn.{n}.weight"], P[f"{pfx}.slf_attn.{n}.bias"])
        return y.view(B, T, 2, 64).permute(2, 0, 1, 3).reshape(2 * B, T, 64)
    q, k, v = heads("w_qs"), heads("w_ks"), heads("w_vs")
    attn = torch.bmm(q, k.transpose(1, 2)) / (128 ** 0.5)           # temperature sqrt(d_model)
    attn = attn.masked_fill(pad.unsqueeze(1).expand(-1, T, -1).repeat(2, 1, 1), float("-inf"))
    attn = F.softmax(attn, dim=2)
    o = torch.bmm(attn, v).view(2, B, T, 64).permute(1, 2, 0, 3).reshape(B, T, 128)
    x = F.linear(o, P[pfx + ".slf_attn.fc.weight"], P[pfx + ".slf_attn.fc.bias"]) + res
    x = F.linear(x, P[pfx + ".fc.fc.weight"], P[pfx + ".fc.fc.bias"])
    lens = (~pad).sum(dim=1).unsqueeze(1)
    w = x.masked_fill(pad.unsqueeze(-1), 0).sum(dim=1) / lens
    return w.unsqueeze(-1)


def wn_stack(P, pfx, x, x_mask, g, hidden, k, n_layers):
    """modules.py:187-212 (dilation_rate 1, dropout 0)."""
    out = torch.zeros_like(x)
    g = conv1d(P, pfx + ".cond_layer", g)
    for i in range(n_layers):
        a = conv1d(P, f"{pfx}.in_layers.{i}", x, padding=(k - 1) // 2)
        a = a + g[:, i * 2 * hidden:(i + 1) * 2 * hidden, :]
        acts = torch.tanh(a[:, :hidden]) * torch.sigmoid(a[:, hidden:])
        rs = conv1d(P, f"{pfx}.res_skip_layers.{i}", acts)
        if i < n_layers - 1:
            x = (x + rs[:, :hidden]) * x_mask
            out = out + rs[:, hidden:]
        else:
            out = out + rs
    return out * x_mask


def posterior_encoder(P, pfx, y, y_mask, g, noise, hidden=192, inter=192):
    """models.py:348-359; g is detached (:349-350); noise replaces randn_like (:358)."""
    x = conv1d(P, pfx + ".pre", y) * y_mask
    x = wn_stack(P, pfx + ".enc", x, y_mask, g.detach(), hidden, 5, 16)
    stats = conv1d(P, pfx + ".proj", x) * y_mask
    m, logs = stats.split(inter, dim=1)
    z = (m + noise * torch.exp(logs)) * y_mask
    return z, m, logs


def flow_forward(P, pfx, x, x_mask, g, hidden=192):
    """models.py:308-315 + modules.py:439-458 (mean_only) + Flip :376-383."""
    half = x.shape[1] // 2
    for f in range(4):
        p = f"{pfx}.flows.{2 * f}"
        x0, x1 = x[:, :half], x[:, half:]
        h = conv1d(P, p + ".pre", x0) * x_mask
        h = wn_stack(P, p + ".enc", h, x_mask, g, hidden, 5, 4)
        m = conv1d(P, p + ".post", h) * x_mask
        x1 = m + x1 * x_mask
        x = torch.cat([x0, x1], 1).flip(1)
    return x


def flow_reverse(P, pfx, x, x_mask, g, hidden=192):
    """ResidualCouplingBlock(reverse=True), models.py:316-319: the flows in reverse order, i.e. Flip first, then the
    mean-only coupling inverted, modules.py:459-462: x1 = (x1 - m) * exp(-0) * mask, x0 passes through."""
    half = x.shape[1] // 2
    for f in reversed(range(4)):
        p = f"{pfx}.flows.{2 * f}"
        x = x.flip(1)
        x0, x1 = x[:, :half], x[:, half:]
        h = conv1d(P, p + ".pre", x0) * x_mask
        h = wn_stack(P, p + ".enc", h, x_mask, g, hidden, 5, 4)
        m = conv1d(P, p + ".post", h) * x_mask
        x1 = (x1 - m) * x_mask
        x = torch.cat([x0, x1], 1)
    return x


def text_encoder(P, quantized, y_mask, text, t_mask, ge, m=S2_MODEL, speed=1):
    """TextEncoder.forward, models.py:228-251 -> (m_p, logs_p, y_mask); speed != 1 resamples the encoder output, :246-248."""
    q = conv1d(P, "enc_p.ssl_proj", quantized * y_mask) * y_mask
    q = attn_encoder(P, "enc_p.encoder_ssl", q * y_mask, y_mask, m["n_layers"] // 2)
    t = F.embedding(text, P["enc_p.text_embedding.weight"]).transpose(1, 2)
    t = attn_encoder(P, "enc_p.encoder_text", t * t_mask, t_mask, m["n_layers"])
    q = mrte(P, "enc_p.mrte", q, y_mask, t, t_mask, ge)
    q = attn_encoder(P, "enc_p.encoder2", q * y_mask, y_mask, m["n_layers"] // 2)
    if speed != 1:
        q = F.interpolate(q, size=int(q.shape[-1] / speed) + 1, mode="linear")
        y_mask = F.interpolate(y_mask, size=q.shape[-1], mode="nearest")
    stats = conv1d(P, "enc_p.proj", q) * y_mask
    m_p, logs_p = stats.split(m["inter_channels"], dim=1)
    return m_p, logs_p, y_mask


def decode(P, codes, text, refers, noise, noise_scale=0.5, m=S2_MODEL, speed=1):
    """SynthesizerTrn.decode, models.py:973-1013 (the vocoder call of TTS, inference/tts.py), dropout off:
    codes [1, 1, T] int64, text [1, X] int64, refers = list of reference spectrograms [1, 1025, Tr], noise [1, 192, F]
    (stands for torch.randn_like; F = 2T, or int(2T / speed) + 1) -> waveform [1, 1, F * 640]."""
    ges = []
    for r in refers:
        rm = torch.ones(1, 1, r.shape[2], dtype=r.dtype)
        ges.append(mel_style_encoder(P, "ref_enc", r[:, :704] * rm, rm))
    ge = torch.stack(ges, 0).mean(0)
    T2 = codes.shape[2] * 2
    y_mask = torch.ones(1, 1, T2)
    t_mask = torch.ones(1, 1, text.shape[1])
    embed = P["quantizer.vq.layers.0._codebook.embed"]
    quantized = F.embedding(codes[0], embed).transpose(1, 2).repeat_interleave(2, dim=2)   # quantizer.decode + nearest x2
    m_p, logs_p, y_mask = text_encoder(P, quantized, y_mask, text, t_mask, ge, m, speed)
    z_p = m_p + noise * torch.exp(logs_p) * noise_scale
    z = flow_reverse(P, "flow", z_p, y_mask, ge)
    return generator(P, "dec", z * y_mask, ge, m)


def slice_segments(x, ids, size):
    """commons.py:42-48."""
    return torch.stack([x[i, :, int(ids[i]):int(ids[i]) + size] for i in range(x.shape[0])])


def generator(P, pfx, x, g, m=S2_MODEL):
    """models.py:452-471 + ResBlock1 modules.py:298-311."""
    x = conv1d(P, pfx + ".conv_pre", x, padding=3) + conv1d(P, pfx + ".cond", g)
    nk = len(m["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(m["upsample_rates"], m["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU)
        x = F.conv_transpose1d(x, wn_weight(P, f"{pfx}.ups.{i}"), P[f"{pfx}.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, rd) in enumerate(zip(m["resblock_kernel_sizes"], m["resblock_dilation_sizes"])):
            r = f"{pfx}.resblocks.{i * nk + j}"
            h = x
            for l, d in enumerate(rd):
                t = conv1d(P, f"{r}.convs1.{l}", F.leaky_relu(h, LRELU), padding=(rk * d - d) // 2, dilation=d)
                t = conv1d(P, f"{r}.convs2.{l}", F.leaky_relu(t, LRELU), padding=(rk - 1) // 2)
                h = t + h
            xs = h if xs is None else xs + h
        x = xs / nk
    x = F.leaky_relu(x)                                   # default slope 0.01 (models.py:467)
    x = conv1d(P, pfx + ".conv_post", x, padding=3)
    return torch.tanh(x)


def vq_nearest(x, embed):
    """core_vq.py:172-180: x [N, D], embed [K, D] -> argmax of -(|x|^2 - 2 x.e + |e|^2)."""
    e = embed.t()
    dist = -(x.pow(2).sum(1, keepdim=True) - 2 * x @ e + e.pow(2).sum(0, keepdim=True))
    return dist.max(dim=-1).indices


def extract_latent(P, ssl):
    """SynthesizerTrn.extract_latent, models.py:1015-1018 (Normalize.token's only model call, normalize.py:203): ssl [B, 768, T]
    -> ssl_proj (k = 2, stride 2) -> nearest codebook entry -> codes.transpose(0, 1) = [B, 1, T // 2] int64."""
    s = F.conv1d(ssl, P["ssl_proj.weight"], P["ssl_proj.bias"], stride=2)
    B, D, N = s.shape
    embed = P["quantizer.vq.layers.0._codebook.embed"]
    return vq_nearest(s.transpose(1, 2).reshape(B * N, D), embed).view(B, 1, N)


def synthesizer_forward(P, ssl, y, y_lengths, text, text_lengths, noise, ids_slice, m=S2_MODEL, segment=32):
    """SynthesizerTrn.forward, models.py:904-946, with frozen quantizer (eval), dropout off,
    `noise` [B,192,T] for enc_q and `ids_slice` [B] injected."""
    T = y.shape[2]
    y_mask = sequence_mask(y_lengths, T).unsqueeze(1).to(y.dtype)
    ge = mel_style_encoder(P, "ref_enc", y[:, :704] * y_mask, y_mask)
    s = F.conv1d(ssl, P["ssl_proj.weight"], P["ssl_proj.bias"], stride=2)
    B, D, N = s.shape
    embed = P["quantizer.vq.layers.0._codebook.embed"]
    codes = vq_nearest(s.transpose(1, 2).reshape(B * N, D), embed).view(B, N)
    quantized = F.embedding(codes, embed).transpose(1, 2)
    quantized = quantized.repeat_interleave(2, dim=2)     # F.interpolate(nearest, x2), models.py:924-927
    # enc_p, models.py:228-251
    q = conv1d(P, "enc_p.ssl_proj", quantized * y_mask) * y_mask
    q = attn_encoder(P, "enc_p.encoder_ssl", q * y_mask, y_mask, m["n_layers"] // 2)
    t_mask = sequence_mask(text_lengths, text.shape[1]).unsqueeze(1).to(y.dtype)
    t = F.embedding(text, P["enc_p.text_embedding.weight"]).transpose(1, 2)
    t = attn_encoder(P, "enc_p.encoder_text", t * t_mask, t_mask, m["n_layers"])
    q = mrte(P, "enc_p.mrte", q, y_mask, t, t_mask, ge)
    q = attn_encoder(P, "enc_p.encoder2", q * y_mask, y_mask, m["n_layers"] // 2)
    stats = conv1d(P, "enc_p.proj", q) * y_mask
    m_p, logs_p = stats.split(m["inter_channels"], dim=1)
    z, m_q, logs_q = posterior_encoder(P, "enc_q", y, y_mask, ge, noise)
    z_p = flow_forward(P, "flow", z, y_mask, ge)
    z_slice = slice_segments(z, ids_slice, segment)
    o = generator(P, "dec", z_slice, ge, m)
    return dict(y_hat=o, ids_slice=ids_slice, y_mask=y_mask, z=z, z_p=z_p, m_p=m_p, logs_p=logs_p,
                m_q=m_q, logs_q=logs_q, quantized=quantized, codes=codes, ge=ge)


def disc_s(P, pfx, x):
    """models.py:576-587."""
    cfg = [(1, 7, 1), (4, 20, 4), (4, 20, 16), (4, 20, 64), (4, 20, 256), (1, 2, 1)]
    fmap = []
    for i, (s, p, g) in enumerate(cfg):
        x = F.leaky_relu(conv1d(P, f"{pfx}.convs.{i}", x, stride=s, padding=p, groups=g), LRELU)
        fmap.append(x)
    x = conv1d(P, pfx + ".conv_post", x, padding=1)
    fmap.append(x)
    return x.flatten(1), fmap


def disc_p(P, pfx, x, period):
    """models.py:538-557."""
    B, C, T = x.shape
    if T % period:
        x = F.pad(x, (0, period - T % period), "reflect")
        T = x.shape[2]
    x = x.view(B, C, T // period, period)
    fmap = []
    for i, s in enumerate([3, 3, 3, 3, 1]):
        x = F.conv2d(x, wn_weight(P, f"{pfx}.convs.{i}"), P[f"{pfx}.convs.{i}.bias"], stride=(s, 1), padding=(2, 0))
        x = F.leaky_relu(x, LRELU)
        fmap.append(x)
    x = F.conv2d(x, wn_weight(P, pfx + ".conv_post"), P[pfx + ".conv_post.bias"], padding=(1, 0))
    fmap.append(x)
    return x.flatten(1), fmap


def mpd(P, y, y_hat):
    """MultiPeriodDiscriminator.forward, models.py:601-614."""
    rs, gs, frs, fgs = [], [], [], []
    for d in range(6):
        pfx = f"discriminators.{d}"
        fn = (lambda t: disc_s(P, pfx, t)) if d == 0 else (lambda t: disc_p(P, pfx, t, PERIODS[d - 1]))
        r, fr = fn(y)
        g, fg = fn(y_hat)
        rs.append(r); gs.append(g); frs.append(fr); fgs.append(fg)
    return rs, gs, frs, fgs


# ----------------------------------------------------------------------------
# losses (losses.py) and the step (sovits.py:459-525)
# ----------------------------------------------------------------------------
def discriminator_loss(rs, gs):
    return sum(torch.mean((1 - r) ** 2) + torch.mean(g ** 2) for r, g in zip(rs, gs))


def generator_loss(gs):
    return sum(torch.mean((1 - g) ** 2) for g in gs)


def feature_loss(frs, fgs):
    return 2 * sum(torch.mean(torch.abs(r.detach() - g)) for fr, fg in zip(frs, fgs) for r, g in zip(fr, fg))


def kl_loss(z_p, logs_q, m_p, logs_p, z_mask):
    kl = logs_p - logs_q - 0.5 + 0.5 * ((z_p - m_p) ** 2) * torch.exp(-2.0 * logs_p)
    return torch.sum(kl * z_mask) / torch.sum(z_mask)


def s2_losses(PG, PD, batch, noise, ids_slice, data=S2_DATA, train=S2_TRAIN, m=S2_MODEL):
    """One stage-2 step's two losses (no optimizer), sovits.py:459-518, fp32, dropout off.

    batch = (ssl [B,768,T], spec [B,1025,T], spec_lengths, wav [B,1,L], text [B,X], text_lengths).
    Returns dict with loss_disc (graph through PD only) and loss_gen_all (graph through PG; the
    discriminator weights are treated as constants of the *same* values -- the reference steps D
    in between, so callers comparing a full step must update PD first).
    """
    ssl, spec, spec_len, wav, text, text_len = batch
    seg = train["segment_size"] // data["hop_length"]
    out = synthesizer_forward(PG, ssl, spec, spec_len, text, text_len, noise, ids_slice, m, seg)
    margs = (data["filter_length"], data["n_mel_channels"], data["sampling_rate"], data["mel_fmin"], data["mel_fmax"])
    mel = mel_oracle.spec_to_mel(spec, *margs)
    y_mel = slice_segments(mel, ids_slice, seg)
    y_hat = out["y_hat"]
    y_hat_mel = mel_oracle.mel_spectrogram(y_hat.squeeze(1), data["filter_length"], data["n_mel_channels"],
                                           data["sampling_rate"], data["hop_length"], data["win_length"],
                                           data["mel_fmin"], data["mel_fmax"])
    y = slice_segments(wav, ids_slice * data["hop_length"], train["segment_size"])
    rs, gs, _, _ = mpd(PD, y, y_hat.detach())
    loss_disc = discriminator_loss(rs, gs)
    rs, gs, frs, fgs = mpd(PD, y, y_hat)
    loss_mel = F.l1_loss(y_mel, y_hat_mel) * train["c_mel"]
    loss_kl = kl_loss(out["z_p"], out["logs_q"], out["m_p"], out["logs_p"], out["y_mask"]) * train["c_kl"]
    loss_fm = feature_loss(frs, fgs)
    loss_gen = generator_loss(gs)
    total = loss_gen + loss_fm + loss_mel + loss_kl     # + kl_ssl == 0 (frozen quantizer in eval)
    out.update(loss_disc=loss_disc, loss_gen_all=total, loss_gen=loss_gen, loss_fm=loss_fm,
               loss_mel=loss_mel, loss_kl=loss_kl, y=y, y_mel=y_mel, y_hat_mel=y_hat_mel)
    return out


def adamw_step(p, g, m, v, step, lr, betas=(0.8, 0.99), eps=1e-9, wd=0.01):
    """torch.optim.AdamW single-tensor math (sovits.py:294-319 uses torch defaults wd=0.01)."""
    b1, b2 = betas
    p = p * (1 - lr * wd)
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / ((v.sqrt() / math.sqrt(bc2)) + eps)
    return p, m, v


def synthetic_batch(B, T, X, seed, ragged=False, hop=640):
    """BASELINE config-3 shaped synthetic batch on the CPU generator (identical on every host)."""
    g = torch.Generator().manual_seed(seed)
    L = T * hop
    wav = (torch.rand(B, 1, L, generator=g) - 0.5)
    ssl = torch.randn(B, 768, T, generator=g)
    text = torch.randint(0, N_SYMBOLS, (B, X), generator=g)
    if ragged:
        spec_len = torch.randint(max(T // 2, 34), T + 1, (B,), generator=g)
        spec_len[0] = T
        spec_len, _ = torch.sort(spec_len, descending=True)
        text_len = torch.randint(max(X // 2, 1), X + 1, (B,), generator=g)
        text_len[0] = X
    else:
        spec_len = torch.full((B,), T, dtype=torch.long)
        text_len = torch.full((B,), X, dtype=torch.long)
    return wav, ssl, text, spec_len, text_len
