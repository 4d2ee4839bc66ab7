"""CPU tests: the oracle against the committed reference goldens (tests/golden/, written by
oracle/pin_against_reference.py from the reference executed in the authoring container)."""
import json
import os

import numpy as np
import torch

from oracle import mel_oracle, s2_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_mel_kat_matches_reference_golden():
    gold = torch.load(os.path.join(GOLD, "mel_kat_22050.pt"))
    y = mel_oracle.kat_sines()
    mel = mel_oracle.mel_spectrogram(y, 2048, 128, 22050, 640, 2048, 0.0, None)
    assert mel.shape == (8, 128, 34)
    assert torch.equal(mel, gold["mel"])
    assert torch.equal(mel_oracle.spectrogram(y, 2048, 640, 2048)[0], gold["spec_b0"])
    # SURVEY.md 8(c) known answers from the reference
    assert abs(float(mel.sum()) + 286936.5209) < 0.5
    assert np.allclose(mel[0, :4, 0].numpy(), [0.13125055, 0.20751116, 0.28897765, 0.36968514], atol=1e-5)
    assert mel[:, :, 17].argmax(1).tolist() == [8, 16, 25, 33, 41, 48, 54, 59]


def test_mel_random_audio_matches_reference_golden():
    g = torch.Generator().manual_seed(7)
    for sr, L in ((32000, 32000), (48000, 24000)):
        yy = torch.rand(3, L, generator=g) - 0.5
        gold = torch.load(os.path.join(GOLD, f"mel_rand_{sr}.pt"))
        assert torch.equal(mel_oracle.mel_spectrogram(yy, 2048, 128, sr, 640, 2048, 0.0, None), gold["mel"])


def test_filterbank_is_two_sparse_and_product_restatement_agrees():
    from easevoice_trainer_b200.mel_processing import slaney_filterbank
    for sr in (22050, 32000, 48000):
        fb = mel_oracle.mel_filterbank(sr, 2048, 128, 0.0, None)
        assert fb.shape == (128, 1025) and (fb != 0).sum(0).max() <= 2
        assert np.array_equal(fb, slaney_filterbank(sr, 2048, 128, 0.0, None))


def test_f64_and_f32_oracles_agree_above_the_floor():
    y = mel_oracle.kat_sines()
    f32 = mel_oracle.mel_spectrogram(y, 2048, 128, 22050, 640, 2048, 0.0, None).numpy()
    f64 = mel_oracle.mel_spectrogram_f64(y.numpy(), 2048, 128, 22050, 640, 2048, 0.0, None)
    strong = f64 > np.log(1e-2)
    assert np.abs(f32 - f64)[strong].max() < 2e-4


def test_param_specs_match_reference_counts():
    g, d = s2_oracle.generator_param_spec(), s2_oracle.discriminator_param_spec()
    assert len(g) == 776 and len(d) == 111                       # state_dict entries
    ng = sum(int(np.prod(s)) for k, s in g.items() if k not in s2_oracle.GEN_BUFFERS)
    nd = sum(int(np.prod(s)) for s in d.values())
    assert ng == 51_310_080 and nd == 46_747_132                 # SURVEY.md 8(c) parameter counts


def test_s2_small_losses_match_reference_golden():
    """oracle stage-2 forward (B=2, T=48) reproduces the losses the reference produced on the same inputs."""
    gold = json.load(open(os.path.join(GOLD, "s2_small.json")))
    c = gold["cfg"]
    PG = s2_oracle.init_params(s2_oracle.generator_param_spec(), c["g_seed"])
    PD = s2_oracle.init_params(s2_oracle.discriminator_param_spec(), c["d_seed"])
    wav, ssl, text, spec_len, text_len = s2_oracle.synthetic_batch(c["B"], c["T"], c["X"], c["batch_seed"], c["ragged"])
    spec = mel_oracle.spectrogram(wav.squeeze(1), 2048, 640, 2048)
    g = torch.Generator().manual_seed(c["noise_seed"])
    noise = torch.randn(c["B"], 192, c["T"], generator=g)
    ids = (torch.rand(c["B"], generator=g) * (spec_len - 32 + 1)).long()
    assert ids.tolist() == gold["ids_slice"]
    with torch.no_grad():
        o = s2_oracle.s2_losses(PG, PD, (ssl, spec, spec_len, wav, text, text_len), noise, ids)
    for k in ("loss_disc", "loss_gen", "loss_fm", "loss_mel", "loss_kl", "loss_gen_all"):
        assert abs(float(o[k]) - gold[k]) <= 2e-4 * abs(gold[k]), (k, float(o[k]), gold[k])
    assert int(o["codes"].sum()) == gold["codes_sum"]
    assert np.allclose(o["y_hat"][0, 0, 100:108].numpy(), gold["y_hat_0_0_100_108"], rtol=1e-3, atol=1e-6)
