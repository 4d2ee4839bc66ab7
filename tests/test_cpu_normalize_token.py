"""Normalize.token (token half): the oracle restatement of extract_latent reproduces the golden produced by the
reference (oracle/pin_against_reference.py --extract-latent), and the TSV writer keeps the reference's file format
(normalize.py:195-211).  The model call of the writer is GPU-only; here it is replaced by the oracle to test the host logic."""
import json
import os

import torch

from oracle import s2_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _golden():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "extract_latent.json")))


def test_oracle_extract_latent_matches_reference_golden():
    gold = _golden()
    P = s2_oracle.init_params(s2_oracle.generator_param_spec(), gold["g_seed"])
    g = torch.Generator().manual_seed(gold["ssl_seed"])
    for case in gold["cases"]:
        ssl = torch.randn(1, 768, case["T"], generator=g)
        codes = s2_oracle.extract_latent(P, ssl)
        assert codes.shape == (1, 1, case["T"] // 2) and codes.dtype == torch.int64
        assert codes[0, 0].tolist() == case["codes"]


class _OracleVQ:
    """stands in for models.SynthesizerTrn in the host-logic test: same extract_latent contract, computed by the oracle"""

    def __init__(self, P):
        self.P = P

    def parameters(self):
        return iter([torch.zeros(1)])

    def extract_latent(self, x, lengths=None):
        codes = s2_oracle.extract_latent(self.P, x[:, :, :x.shape[2] // 2 * 2])
        if lengths is not None:
            keep = torch.arange(codes.shape[2])[None, :] < (lengths // 2)[:, None]
            codes = codes * keep[:, None, :]
        return codes


def test_semantic_tsv_writer_format_and_padding_invariance(tmp_path):
    from easevoice_trainer_b200 import normalize_token as nt
    gold = _golden()
    P = s2_oracle.init_params(s2_oracle.generator_param_spec(), gold["g_seed"])
    g = torch.Generator().manual_seed(gold["ssl_seed"])
    hub = tmp_path / "4-cnhubert"
    hub.mkdir()
    names = ["a.wav", "b.wav", "c.wav", "d.wav"]
    for n, case in zip(names, gold["cases"]):
        torch.save(torch.randn(1, 768, case["T"], generator=g), str(hub / (n + ".pt")))
    ref_list = tmp_path / "refinements.list"
    # the second entry has no feature file (skipped like the reference does), quotes and a directory part are stripped
    ref_list.write_text('"/data/denoise/a.wav"|zh|x\nmissing.wav|zh|y\nb.wav|en|z\n c.wav |zh|w\nd.wav|zh|v\n', encoding="utf8")
    out = tmp_path / "6-name2semantic.tsv"
    n = nt.write_semantic_tsv(str(ref_list), str(hub), str(out), _OracleVQ(P), max_batch=3)
    assert n == 4
    lines = out.read_text(encoding="utf8").split("\n")
    assert lines[0] == "item_name\tsemantic_audio" and lines[-1] == ""
    got = {l.split("\t")[0]: [int(v) for v in l.split("\t")[1].split(" ")] for l in lines[1:-1]}
    assert list(got) == names                                  # file order = list order
    for nme, case in zip(names, gold["cases"]):
        assert got[nme] == case["codes"]                       # batched + zero-padded == one file at a time (the golden)


def test_oracle_decode_matches_reference_golden():
    """SURVEY 8 row f4 (vocoder half): the oracle restatement of SynthesizerTrn.decode (flow reversed) reproduces the waveform
    of the reference (oracle/pin_against_reference.py --decode) from the same seeds."""
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "decode.pt"), weights_only=False)
    c = gold["cfg"]
    P = s2_oracle.init_params(s2_oracle.generator_param_spec(), c["g_seed"])
    g = torch.Generator().manual_seed(c["seed"])
    codes = torch.randint(0, 1024, (1, 1, c["T"]), generator=g)
    text = torch.randint(0, 300, (1, c["X"]), generator=g)
    refers = [torch.rand(1, 1025, tr, generator=g) * 2.0 for tr in c["Tr"]]
    noise = torch.randn(1, 192, 2 * c["T"], generator=g)
    with torch.no_grad():
        w = s2_oracle.decode(P, codes, text, refers, noise, c["noise_scale"])
    assert w.shape == gold["wave"].shape
    assert float((w - gold["wave"]).abs().max()) < 2e-5
    Fs = int(2 * c["T"] / c["speed"]) + 1
    noise_s = torch.randn(1, 192, Fs, generator=g)
    with torch.no_grad():
        ws = s2_oracle.decode(P, codes, text, refers, noise_s, c["noise_scale"], speed=c["speed"])
    assert ws.shape == gold["wave_speed"].shape == (1, 1, Fs * 640)
    assert float((ws - gold["wave_speed"]).abs().max()) < 2e-5


def test_oracle_infer_panel_matches_reference_golden():
    """SURVEY 8 row f4 (AR half): the oracle restatement of infer_panel_naive (no cache: full recompute under the prefix-LM mask)
    reproduces the greedy token sequence the reference decoded with its KV cache (oracle/pin_against_reference.py --infer-panel)."""
    from oracle import gpt_oracle
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "infer_panel.json")))
    c = gold["cfg"]
    m = dict(gpt_oracle.GPT_MODEL, n_layer=c["n_layer"])
    P = gpt_oracle.init_params(gpt_oracle.gpt_param_spec(m), c["param_seed"])
    P["ar_text_position.alpha"].fill_(0.8); P["ar_audio_position.alpha"].fill_(1.3)
    g = torch.Generator().manual_seed(c["seed"])
    x = torch.randint(0, m["phoneme_vocab_size"], (1, c["X"]), generator=g)
    bert = torch.randn(1, 1024, c["X"], generator=g)
    prompts = torch.randint(0, 1024, (1, c["Yp"]), generator=g)
    tr = []
    with torch.no_grad():
        y, idx = gpt_oracle.infer_panel(P, x, bert, prompts, top_k=c["top_k"], top_p=100, early_stop_num=c["early_stop_num"],
                                        temperature=c["temperature"], repetition_penalty=c["repetition_penalty"], m=m, trace=tr)
    assert y[0].tolist() == gold["tokens"] and int(idx) == gold["idx"]
    for s_, ref in gold["logits_step"].items():
        assert float((tr[int(s_)][0] - torch.tensor(ref)).abs().max()) < 2e-4


def test_oracle_hubert_matches_transformers_golden():
    """SURVEY 8 row f3 (ssl half): the oracle restatement of the HuBERT forward reproduces what transformers.HubertModel produced
    from the same seeded weights (oracle/pin_against_reference.py --hubert); when transformers is importable the model itself is
    re-run as well."""
    from oracle import hubert_oracle as ho
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "hubert.pt"), weights_only=False)
    c = gold["cases"][0]                                        # 2 layers, 1 s of audio
    m = dict(ho.HUBERT_BASE, layers=c["layers"])
    P = ho.init_params(ho.param_spec(m), c["param_seed"])
    wav = torch.randn(1, c["L"], generator=torch.Generator().manual_seed(c["wav_seed"])) * c["wav_scale"]
    with torch.no_grad():
        o = ho.forward(P, wav, m)
    assert o.shape == c["out"].shape and float((o - c["out"]).abs().max()) < 5e-5
