"""Semantic-token extraction of the dataset preparation step (token half).

Mirror of `Normalize.token` (reference src/normalization/normalize.py:185-211): for every line of the refinements list
(`wav_name|language|text`), load `4-cnhubert/<name>.pt` (HuBERT features [1, 768, T], written by `Normalize.ssl`), run
`SynthesizerTrn.extract_latent` and write `6-name2semantic.tsv`:

    item_name\\tsemantic_audio
    <name>\\t<space separated ints>

Same files in, same file out; the model call runs on the GPU through the library's exact-fp32 projection and codeword search
(`models.SynthesizerTrn.extract_latent`), several utterances per launch (zero-padded to a common length: the stride-2 projection
and the per-frame codeword search are local, padding cannot change a valid token).  The reference re-runs the model per file on
the CPU.  There is no CPU fallback: without the CUDA library the call fails.
"""
import os

import torch

from . import configs, models


def format_path(p):
    """Behaviour of the reference's `utils/path/path.py:7-9`: trailing separators dropped, both separator kinds mapped to the
    platform's, then blanks / quotes / newlines / the U+202A mark the file dialogs leave around a path trimmed."""
    p = p.rstrip("/\\")
    for sep in ("/", "\\"):
        p = p.replace(sep, os.sep)
    return p.strip(" '\"\n\u202a")


def load_vq_model(weights_path=None, device="cuda", hps=None, state_dict=None):
    """SynthesizerTrn as `Normalize.token` builds it (normalize.py:186-194): s2 config, eval, `strict=False` load of the
    `weight` entry of a pretrained s2G checkpoint."""
    hps = hps or configs.load_s2_config()
    net = models.SynthesizerTrn(hps["data"]["filter_length"] // 2 + 1, hps["train"]["segment_size"] // hps["data"]["hop_length"],
                                n_speakers=hps["data"]["n_speakers"], **hps["model"])
    if state_dict is None and weights_path:
        state_dict = torch.load(str(weights_path), map_location="cpu", weights_only=False)["weight"]
    if state_dict is not None:
        net.load_state_dict({k: v.float() for k, v in state_dict.items()}, strict=False)
    return net.to(device).eval()


def extract_tokens(vq_model, feats, max_batch=32, max_frames=1 << 16):
    """feats: list of HuBERT feature tensors [1, 768, T_i] (CPU or GPU) -> list of python int lists (codes[0, 0, :]).
    Utterances are grouped longest-first into zero-padded batches of at most `max_batch` rows / `max_frames` padded frames."""
    dev = next(iter(vq_model.parameters())).device if hasattr(vq_model, "parameters") else torch.device("cuda")
    order = sorted(range(len(feats)), key=lambda i: -feats[i].shape[-1])
    out = [None] * len(feats)
    i = 0
    while i < len(order):
        Tm = feats[order[i]].shape[-1]
        nb = max(1, min(max_batch, max_frames // max(Tm, 1)))
        grp = order[i:i + nb]
        i += nb
        x = torch.zeros((len(grp), 768, Tm), dtype=torch.float32, device=dev)
        lens = torch.empty(len(grp), dtype=torch.int64)
        for r, j in enumerate(grp):
            f = feats[j]
            assert f.dim() == 3 and f.shape[0] == 1 and f.shape[1] == 768, f"expected [1, 768, T], got {tuple(f.shape)}"
            x[r, :, :f.shape[-1]] = f[0].to(dev, torch.float32)
            lens[r] = f.shape[-1]
        codes = vq_model.extract_latent(x, lens).cpu()
        for r, j in enumerate(grp):
            out[j] = codes[r, 0, :int(lens[r]) // 2].tolist()
    return out


def write_semantic_tsv(refinements_path, hubert_dir, semantic_output_path, vq_model, max_batch=32):
    """`Normalize.token` body (normalize.py:195-211).  Lines whose feature file is missing are skipped, like the reference.
    Returns the number of utterances written."""
    with open(refinements_path, "r", encoding="utf8") as f:
        lines = f.read().strip("\n").split("\n")
    names, feats = [], []
    for line in lines:
        if not line:
            continue
        wav_name, _language, _text = line.split("|")
        wav_name = os.path.basename(format_path(wav_name))
        hubert_path = os.path.join(hubert_dir, wav_name + ".pt")
        if not os.path.exists(hubert_path):
            continue
        names.append(wav_name)
        feats.append(torch.load(hubert_path, map_location="cpu", weights_only=False))
    toks = extract_tokens(vq_model, feats, max_batch=max_batch)
    opt = ["item_name\tsemantic_audio"]
    for n, t in zip(names, toks):
        opt.append("%s\t%s" % (n, " ".join(str(i) for i in t)))
    with open(semantic_output_path, "w", encoding="utf8") as f:
        f.write("\n".join(opt) + "\n")
    return len(names)
