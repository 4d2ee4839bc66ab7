#!/usr/bin/env python
"""ncu driver for the prefix-LM attention kernels at the benchmarked shape (B = 16, 16 heads, X = 256, Y = 1024, p_drop = 0.1):
one forward + backward per kernel family inside cudaProfilerStart / Stop (mma.sync family, then tcgen05 family with the
two-threads-per-row forward).
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:flash -o gpurun_out/prof_flash python tests/prof_flash.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops                    # noqa: E402

L = lib.init()
dev = torch.device("cuda", 0)
B, H, X, Y = 16, 16, 256, 1024
qkv = torch.randn(B, X + Y, 3 * 512, device=dev, requires_grad=True)
xl = torch.full((B,), X, dtype=torch.int64, device=dev)
yl = torch.full((B,), Y, dtype=torch.int64, device=dev)
go = torch.randn(B, X + Y, 512, device=dev)


def run(tc):
    L.evk_set_flash_tc(tc, -1.0)
    qkv.grad = None
    o = ops.flash_attention(qkv, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=0.1, tag="prof")
    o.backward(go)
    torch.cuda.synchronize()


for tc in (0, 3):
    run(tc)                                  # warm-up outside the capture
torch.cuda.profiler.start()
for tc in (0, 3):
    run(tc)
torch.cuda.profiler.stop()
L.evk_set_flash_tc(0, -1.0)
print("done")
