#!/usr/bin/env python
"""Small driver for ncu captures: launches the fused mel kernel and one heavy conv layer a few times.
  ncu --set full --clock-control none --import-source on -k regex:'mel_fwd_warp|gconv_tc' -c 4 -o gpurun_out/prof python tests/prof_kernels.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops                    # noqa: E402
from easevoice_trainer_b200.mel_processing import get_bank      # noqa: E402

lib.init()
dev = torch.device("cuda", 0)
bank = get_bank(32000, 2048, 128, 0.0, None, dev)
wav = torch.rand(64, 221440, device=dev) - 0.5
for _ in range(3):
    ops.mel_frontend(wav, bank, 640, want_spec=True, want_mel=True)
x = torch.randn(16, 2560, 128, device=dev)
w = ops.pack_weight(torch.randn(128, 128, 11, device=dev) * 0.02, None)
bias = torch.zeros(128, device=dev)
with torch.no_grad():
    for _ in range(3):
        ops.conv(x, w, bias, pad=5)
    xd = torch.randn(32, 254, 1024, device=dev)
    wd = ops.pack_weight(torch.randn(1024, 1024, 5, device=dev) * 0.01, None)
    bd = torch.zeros(1024, device=dev)
    for _ in range(2):
        ops.conv(xd, wd, bd, pad=2, P=2)
torch.cuda.synchronize()
print("done")
