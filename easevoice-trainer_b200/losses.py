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
