"""CPU restatement of the HuBERT forward `Normalize.ssl` runs (reference src/normalization/normalize.py:166-168:
`CNHubert(...).model(wav16k.unsqueeze(0))["last_hidden_state"]`, src/easevoice/feature_extractor/cnhubert.py:14-33).

TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench CPU arms).  The algorithm lives in a third-party dependency of
the reference, `transformers` (this image: 5.5.0; `HubertModel`, models/hubert/modeling_hubert.py); the architecture is the
published HuBERT-base / wav2vec 2.0 one and is restated here from that file's forward passes:

  feature extractor  7 x Conv1d(no bias): (1->512, k10, s5) + GroupNorm(512 groups) + GELU, then (512->512, k3, s2) x 4 and
                     (k2, s2) x 2, each + GELU                                   HubertFeatureEncoder / *ConvLayer
  feature projection LayerNorm(512) -> Linear(512 -> 768)                        HubertFeatureProjection
  positional conv    Conv1d(768, 768, k128, pad 64, groups 16, weight_norm over dim 2), last frame dropped, GELU; added to input
  encoder            LayerNorm, then 12 post-LN blocks: MHA(12 x 64, q scaled by 64^-0.5) + residual + LN, FFN 3072 GELU + LN

Pinned against `transformers.HubertModel` itself by oracle/pin_against_reference.py --hubert (golden tests/golden/hubert.pt)."""
import math

import torch
import torch.nn.functional as F

HUBERT_BASE = dict(hidden=768, heads=12, ffn=3072, layers=12, conv_dim=512, conv_kernel=(10, 3, 3, 3, 3, 2, 2),
                   conv_stride=(5, 2, 2, 2, 2, 2, 2), pos_k=128, pos_groups=16, eps=1e-5)


def param_spec(m=HUBERT_BASE):
    """state_dict contract of transformers.HubertModel (5.5.0) minus `masked_spec_embed` (unused in inference)."""
    s = {}
    cd = m["conv_dim"]
    for i, k in enumerate(m["conv_kernel"]):
        s[f"feature_extractor.conv_layers.{i}.conv.weight"] = (cd, 1 if i == 0 else cd, k)
        if i == 0:
            s["feature_extractor.conv_layers.0.layer_norm.weight"] = (cd,)
            s["feature_extractor.conv_layers.0.layer_norm.bias"] = (cd,)
    H = m["hidden"]
    s["feature_projection.layer_norm.weight"] = (cd,)
    s["feature_projection.layer_norm.bias"] = (cd,)
    s["feature_projection.projection.weight"] = (H, cd)
    s["feature_projection.projection.bias"] = (H,)
    s["encoder.pos_conv_embed.conv.bias"] = (H,)
    s["encoder.pos_conv_embed.conv.parametrizations.weight.original0"] = (1, 1, m["pos_k"])
    s["encoder.pos_conv_embed.conv.parametrizations.weight.original1"] = (H, H // m["pos_groups"], m["pos_k"])
    s["encoder.layer_norm.weight"] = (H,)
    s["encoder.layer_norm.bias"] = (H,)
    for i in range(m["layers"]):
        p = f"encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[p + f"attention.{n}.weight"] = (H, H)
            s[p + f"attention.{n}.bias"] = (H,)
        s[p + "layer_norm.weight"] = (H,)
        s[p + "layer_norm.bias"] = (H,)
        s[p + "feed_forward.intermediate_dense.weight"] = (m["ffn"], H)
        s[p + "feed_forward.intermediate_dense.bias"] = (m["ffn"],)
        s[p + "feed_forward.output_dense.weight"] = (H, m["ffn"])
        s[p + "feed_forward.output_dense.bias"] = (H,)
        s[p + "final_layer_norm.weight"] = (H,)
        s[p + "final_layer_norm.bias"] = (H,)
    return s


def init_params(spec, seed):
    """Seeded synthetic weights with trained-model-like scales (there is no checkpoint in the image): fan-in scaled normals,
    LayerNorm / GroupNorm gains near 1, small biases."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for k, shp in spec.items():
        if k.endswith("original0"):
            P[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("norm.weight"):
            P[k] = 1.0 + 0.05 * torch.randn(shp, generator=g)
        elif k.endswith("bias"):
            P[k] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            gain = 1.6 if "conv_layers" in k else 1.0              # keeps activations O(1) through the GELU conv stack
            P[k] = torch.randn(shp, generator=g) * (gain / math.sqrt(fan_in))
    return P


def pos_conv_weight(P):
    """torch weight_norm(dim=2): w = g * v / ||v|| with the norm over dims (0, 1) for every kernel position."""
    g, v = P["encoder.pos_conv_embed.conv.parametrizations.weight.original0"], P["encoder.pos_conv_embed.conv.parametrizations.weight.original1"]
    return v * (g / v.norm(p=2, dim=(0, 1), keepdim=True))


def forward(P, wav, m=HUBERT_BASE, taps=None):
    """wav [B, L] float32 (16 kHz, un-normalised, as normalize.py feeds it) -> last_hidden_state [B, T, 768]."""
    x = wav[:, None, :]
    for i, (k, s) in enumerate(zip(m["conv_kernel"], m["conv_stride"])):
        x = F.conv1d(x, P[f"feature_extractor.conv_layers.{i}.conv.weight"], stride=s)
        if i == 0:
            x = F.group_norm(x, m["conv_dim"], P["feature_extractor.conv_layers.0.layer_norm.weight"],
                             P["feature_extractor.conv_layers.0.layer_norm.bias"], m["eps"])
        x = F.gelu(x)
    x = x.transpose(1, 2)                                                        # [B, T, 512]
    if taps is not None:
        taps["features"] = x
    x = F.layer_norm(x, (m["conv_dim"],), P["feature_projection.layer_norm.weight"], P["feature_projection.layer_norm.bias"], m["eps"])
    h = F.linear(x, P["feature_projection.projection.weight"], P["feature_projection.projection.bias"])
    pos = F.conv1d(h.transpose(1, 2), pos_conv_weight(P), P["encoder.pos_conv_embed.conv.bias"], padding=m["pos_k"] // 2,
                   groups=m["pos_groups"])
    pos = F.gelu(pos[:, :, :-1] if m["pos_k"] % 2 == 0 else pos).transpose(1, 2)
    h = F.layer_norm(h + pos, (m["hidden"],), P["encoder.layer_norm.weight"], P["encoder.layer_norm.bias"], m["eps"])
    B, T, D = h.shape
    H = m["heads"]
    dk = D // H
    for i in range(m["layers"]):
        p = f"encoder.layers.{i}."
        q = F.linear(h, P[p + "attention.q_proj.weight"], P[p + "attention.q_proj.bias"]) * dk ** -0.5
        k = F.linear(h, P[p + "attention.k_proj.weight"], P[p + "attention.k_proj.bias"])
        v = F.linear(h, P[p + "attention.v_proj.weight"], P[p + "attention.v_proj.bias"])
        q, k, v = [t.view(B, T, H, dk).transpose(1, 2) for t in (q, k, v)]
        a = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).transpose(1, 2).reshape(B, T, D)
        a = F.linear(a, P[p + "attention.out_proj.weight"], P[p + "attention.out_proj.bias"])
        h = F.layer_norm(h + a, (D,), P[p + "layer_norm.weight"], P[p + "layer_norm.bias"], m["eps"])
        f = F.linear(F.gelu(F.linear(h, P[p + "feed_forward.intermediate_dense.weight"], P[p + "feed_forward.intermediate_dense.bias"])),
                     P[p + "feed_forward.output_dense.weight"], P[p + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h + f, (D,), P[p + "final_layer_norm.weight"], P[p + "final_layer_norm.bias"], m["eps"])
    return h
