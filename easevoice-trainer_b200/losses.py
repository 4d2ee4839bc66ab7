"""Drop-in for /root/reference/src/easevoice/module/losses.py on libevk reductions.

The functions accept the reference's argument structure (lists of per-discriminator outputs / feature maps);
tensors may be in any layout since every loss is a full reduction.
"""
from . import ops


def feature_loss(fmap_r, fmap_g):
    """losses.py:7-15: 2 * sum mean|rl.detach() - gl|."""
    loss = 0
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            loss = loss + ops.mean_abs_diff(gl, rl)
    return loss * 2


def discriminator_loss(disc_real_outputs, disc_generated_outputs):
    """losses.py:18-31 without its per-head .item() host syncs; r_losses/g_losses are device scalars."""
    loss, r_losses, g_losses = 0, [], []
    for dr, dg in zip(disc_real_outputs, disc_generated_outputs):
        r, g = ops.mean_sq_one_minus(dr), ops.mean_sq(dg)
        loss = loss + r + g
        r_losses.append(r.detach()); g_losses.append(g.detach())
    return loss, r_losses, g_losses


def generator_loss(disc_outputs):
    """losses.py:34-43."""
    loss, gen_losses = 0, []
    for dg in disc_outputs:
        l = ops.mean_sq_one_minus(dg)
        gen_losses.append(l)
        loss = loss + l
    return loss, gen_losses


def kl_loss(z_p, logs_q, m_p, logs_p, lengths):
    """losses.py:46-61 on channels-last tensors with an int32 length vector instead of a dense mask."""
    return ops.kl_loss(z_p, logs_q, m_p, logs_p, lengths)


def multi_resolution_stft_loss(y_hat, y, window_sizes=(4096, 2048, 1024, 512, 256), hop_size=147, n_fft=2048):
    """The only multi-resolution STFT loss defined in the reference tree (audiokit/uvr5/.../bs_roformer.py:565-581): sum over
    window sizes of F.l1_loss between the complex torch.stft(n_fft=max(win, 2048), hop=147, win_length=win, Hann, center=True)
    of the generated and the target waveform, computed by the general STFT kernel and its adjoint.  y_hat, y: [B, L] or
    [B, 1, L].  BASELINE.json lists it for config 5; SovitsTrain's own loss has no such term (sovits.py:513-518), so
    S2Step only adds it when hps.train["c_mrstft"] > 0 (off by default)."""
    B = y_hat.shape[0]
    return ops.mrstft_loss(y_hat.reshape(B, -1), y.reshape(B, -1), tuple(window_sizes), hop_size, n_fft)
