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


# ---- stage-1 AR GPT -------------------------------------------------------------------------------------------------
def _gpt_case(tag):
    import json
    from oracle import gpt_oracle
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", f"gpt_{tag}.json")))
    m = gold["model"]
    P = gpt_oracle.init_params(gpt_oracle.gpt_param_spec(m), gold["param_seed"])
    P["ar_text_position.alpha"].fill_(gold["alpha_text"]); P["ar_audio_position.alpha"].fill_(gold["alpha_audio"])
    batch = gpt_oracle.synthetic_gpt_batch(gold["B"], gold["X"], gold["Y"], gold["batch_seed"], gold["ragged"])
    return gold, m, P, batch


def test_gpt_param_spec_matches_reference_count():
    from oracle import gpt_oracle
    spec = gpt_oracle.gpt_param_spec()
    n = sum(int(torch.tensor(s).prod()) for s in spec.values())
    assert n == 77_606_402                      # SURVEY.md section 8 / reference Text2SemanticDecoder at configs/gpt.yaml
    assert sum(int(torch.tensor(s).prod()) for k, s in spec.items() if k.startswith("h.")) == 75_657_216


def test_gpt_forward_old_matches_reference_golden():
    from oracle import gpt_oracle
    for tag in ("small", "ragged"):
        gold, m, P, (x, xl, y, yl, bert) = _gpt_case(tag)
        Pq = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        loss, acc, logits, targets = gpt_oracle.forward_old(Pq, x, xl, y, yl, bert, m)
        loss.backward()
        assert abs(float(loss) - gold["loss"]) / gold["loss"] < 1e-5
        assert abs(float(acc) - gold["acc"]) < 1e-6
        assert int(targets.sum()) == gold["targets_sum"]
        for k, n in gold["grad_norms"].items():
            assert abs(float(Pq[k].grad.norm()) - n) / n < 2e-4, k


def test_gpt_dpo_matches_reference_golden():
    from oracle import gpt_oracle
    gold, m, P, (x, xl, y, yl, bert) = _gpt_case("ragged")
    ry, ryl = gpt_oracle.make_reject_given(y, [tuple(s) for s in gold["dpo"]["spans"]])
    loss, acc, l1, l2 = gpt_oracle.forward_dpo(P, x, xl, y, yl, bert, ry, ryl, m)
    assert abs(float(loss) - gold["dpo"]["loss"]) / gold["dpo"]["loss"] < 1e-5
    assert abs(float(l1) + float(l2) - float(loss)) < 1e-3


def test_prefix_lm_mask_properties():
    """t2s_model.py:456-479: text rows never see audio; audio row i sees text + audio <= i; padded keys are never seen."""
    from oracle import gpt_oracle
    X, Y = 5, 7
    xl, yl = torch.tensor([5, 3]), torch.tensor([7, 4])
    mk = gpt_oracle.prefix_lm_mask(xl, yl, X, Y)
    assert mk[:, :X, X:].all()
    assert not mk[0, X:, :X].any() and mk[1, :, 3:X].all()
    for i in range(Y):
        assert not mk[0, X + i, X:X + i + 1].any() and mk[0, X + i, X + i + 1:].all()
    assert mk[1, :, X + 4:].all()
    assert (~mk).any(-1).all()                  # no fully-masked row (would be NaN in the reference's softmax)


def test_scaled_adam_oracle_matches_reference_golden_trajectory():
    import json
    from oracle import gpt_oracle
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "scaled_adam.json")))
    g = torch.Generator().manual_seed(gold["seed"])
    shapes = [tuple(s) for s in gold["shapes"]]
    params = [torch.randn(s, generator=g) * sc for s, sc in zip(shapes, gold["scales"])]
    opt = gpt_oracle.ScaledAdamOracle(params, lr=gold["lr_first"], clipping_update_period=gold["clipping_update_period"])
    clipped = False
    for it in range(gold["steps"]):
        grads = [torch.randn(s, generator=g) * (5.0 if it in gold["big_grad_steps"] else 1.0) for s in shapes]
        clipped |= opt.step(grads) < 1.0
        if it == 0:
            opt.lr = gold["lr_rest"]
        for p, n in zip(params, gold["param_norms"][it]):
            assert abs(float(p.double().norm()) - n) / (n + 1e-12) < 1e-6
    assert clipped
