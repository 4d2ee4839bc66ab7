#!/usr/bin/env python
"""Where does a tile of the tcgen05 attention forward go?  Runs the benchmarked shape on an instrumented build
(tools/exp/build_ftprof.sh) and prints thread 0's cycle counters per key tile.   EVK_LIB_PATH=tools/exp/libevk_ftprof.so python tools/exp/ft_profile.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, ops  # noqa: E402

L_ = lib.init()
raw = ctypes.CDLL(os.environ["EVK_LIB_PATH"])
dev = torch.device("cuda", 0)
NAMES = ["stage+sync", "issue S", "wait S", "prefetch issue", "tcgen05.ld S", "softmax+P stores", "fence+sync", "issue PV", "wait PV",
         "ld O + acc", "tiles", "kernel"]
B, H, X, Y = 16, 16, 256, 1024
L = X + Y
g = torch.Generator().manual_seed(1)
qkv = torch.randn(B, L, 3 * H * 32, generator=g).to(dev)
xl = torch.full((B,), X, device=dev, dtype=torch.int64)
yl = torch.full((B,), Y, device=dev, dtype=torch.int64)
TC = int(os.environ.get("AB_TC", "1"))
NAMES2 = ["softmax: wait S", "softmax: ld S+max+exp(+drop)", "softmax: wait O_tile", "softmax: ld O + acc", "softmax: P stores+fence+arrive",
          "mma: wait kv_full", "mma: issue S", "mma: wait p_full", "mma: issue PV", "prod: wait kv_empty", "prod: issue loads",
          "prod: rotate + V^T stores", "prod: cp.async wait+fence+arrive"]
L_.evk_set_flash_tc(TC, -1.0)
for p in (0.0, 0.1):
    with torch.no_grad():
        for _ in range(2):
            ops.flash_attention(qkv, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=p, tag="ft")
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        raw.evk_ft_prof_read(buf, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.flash_attention(qkv, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=p, tag="ft")
        e1.record()
        torch.cuda.synchronize()
        raw.evk_ft_prof_read(buf, 0)
    v = list(buf)
    if TC >= 2:
        tiles, ptiles = max(v[13], 1), max(v[15], 1)
        print(f"== v{TC} p_drop={p}: {e0.elapsed_time(e1):.3f} ms, {tiles} tiles, {v[14] / tiles:.0f} cycles per tile per CTA")
        for i in range(13):
            print(f"   {NAMES2[i]:36s} {v[i] / (ptiles if i >= 9 else tiles):8.0f} cycles/tile")
        continue
    tiles = max(v[10], 1)
    print(f"== {os.path.basename(os.environ['EVK_LIB_PATH'])} p_drop={p}: {e0.elapsed_time(e1):.3f} ms, {tiles} tiles, {v[11] / tiles:.0f} cycles per tile per CTA")
    for i in range(10):
        print(f"   {NAMES[i]:18s} {v[i] / tiles:8.0f} cycles/tile")
