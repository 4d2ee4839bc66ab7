#!/usr/bin/env python
"""A/B of the gemm_tma_kernel options on the layer shapes of the step (GPU): slab staging, 256-row tiles.
   python tools/ab_tma.py > gpurun_out/ab_tma.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from easevoice_trainer_b200 import lib, ops  # noqa: E402

L = lib.init()
dev = torch.device("cuda", 0)


def gemm_table():
    rows = []
    for name, M, K, N in (("gpt in_proj 512->1536", 20480, 512, 1536), ("gpt linear1 512->2048", 20480, 512, 2048),
                          ("gpt linear2 2048->512", 20480, 2048, 512), ("gpt out_proj 512->512", 20480, 512, 512),
                          ("s2 enc_q.pre 1028->192", 5536, 1028, 192), ("s2 res_skip 192->384", 5536, 192, 384)):
        xs = [torch.randn(1, M, K, device=dev) for _ in range(4)]
        w = ops.pack_weight(torch.randn(N, K, 1, device=dev) * 0.02, None)
        b = torch.zeros(N, device=dev)

        def burst():
            for _ in range(4):
                for x in xs:
                    y = ops.linear(x, w, b)
            return y
        with torch.no_grad():
            burst()
            ms = bench.graph_time(burst) / 16
        rows.append((name, ms, 2.0 * M * K * N / (ms * 1e-3) / 1e12))
    return rows


for slab, mt2 in ((0, 0), (1, 0), (0, 1), (1, 1)):
    L.evk_set_tma_options(slab, mt2, 3.52e-4)
    print(f"==== slab={slab} mt2={mt2}")
    for r in bench.kernel_table(ops, dev):
        print(f"  {r['layer']:38s} {r['ms'] * 1e3:8.1f} us  {r['tflops']:7.1f} TFLOP/s")
    for name, ms, tf in gemm_table():
        print(f"  {name:38s} {ms * 1e3:8.1f} us  {tf:7.1f} TFLOP/s")
    sys.stdout.flush()
L.evk_set_tma_options(1, 1, 3.52e-4)
