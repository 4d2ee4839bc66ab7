#!/usr/bin/env python
"""Small driver for ncu captures: a few launches of each kernel the roofline numbers are about.
  ncu --set full --clock-control none --import-source on -k regex:'gemm_tma|mel_fwd_warp|stft_|flash' -o gpurun_out/prof python tests/prof_kernels.py
Order of launches (ncu ids): [gemm_tma slab=0] k11 128ch x2, [slab=1] k11 128ch x2, [slab=1] discP 1024 x2, GPT linear1 x2, mel x2 (|X|+mel), flash fwd/bwd x1.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops                    # noqa: E402
from easevoice_trainer_b200.mel_processing import get_bank      # noqa: E402

L = lib.init()
dev = torch.device("cuda", 0)
x = torch.randn(16, 2560, 128, device=dev)
w = ops.pack_weight(torch.randn(128, 128, 11, device=dev) * 0.02, None)
bias = torch.zeros(128, device=dev)
with torch.no_grad():
    for slab in (0, 1):
        L.evk_set_tma_options(slab, 0, 3.52e-4)
        for _ in range(2):
            ops.conv(x, w, bias, pad=5)
    L.evk_set_tma_options(1, 1, 3.52e-4)
    xd = torch.randn(32, 254, 1024, device=dev)
    wd = ops.pack_weight(torch.randn(1024, 1024, 5, device=dev) * 0.01, None)
    bd = torch.zeros(1024, device=dev)
    for _ in range(2):
        ops.conv(xd, wd, bd, pad=2, P=2)
# GPT linear1 (20 480 x 512 -> 2048, ReLU): launches 6, 7 of gemm_tma
xg = torch.randn(1, 20480, 512, device=dev)
wg = ops.pack_weight(torch.randn(2048, 512, 1, device=dev) * 0.02, None)
bg = torch.zeros(2048, device=dev)
with torch.no_grad():
    for _ in range(2):
        ops.linear(xg, wg, bg, act=ops.ACT_RELU)
bank = get_bank(32000, 2048, 128, 0.0, None, dev)
wav = torch.rand(64, 221440, device=dev) - 0.5
for _ in range(2):
    ops.mel_frontend(wav, bank, 640, want_spec=True, want_mel=True)
if os.environ.get("PROF_FLASH", "1") == "1":
    B, H, X, Y = 16, 16, 256, 1024
    qkv = torch.randn(B, X + Y, 3 * 512, device=dev, requires_grad=True)
    xl = torch.full((B,), X, dtype=torch.int64, device=dev)
    yl = torch.full((B,), Y, dtype=torch.int64, device=dev)
    o = ops.flash_attention(qkv, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=0.0)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
print("done")
