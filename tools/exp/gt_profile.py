#!/usr/bin/env python
"""Where does a gemm_tma tile's time go?  Runs layer shapes of the step on the instrumented build (tools/exp/build_prof.sh)
and prints the per-CTA cycle counters of the three roles.
   python tools/exp/gt_profile.py > gpurun_out/gt_profile.txt
slots: 0 kernel, 1 producer waits emptyA, 2 producer waits emptyB, 3 MMA waits fullA, 4 MMA waits fullB, 5 MMA waits acc_empty,
       6 epilogue waits acc_full, 7 epilogue body, 8 tiles, 9 prologue"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PLAIN = "--plain" in sys.argv          # product library, timings only (what the instrumentation costs)
if not PLAIN:
    os.environ["EVK_LIB_PATH"] = os.path.join(ROOT, "tools", "exp", "libevk_prof.so")
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, ops  # noqa: E402

L = lib.init()
raw = None if PLAIN else ctypes.CDLL(os.environ["EVK_LIB_PATH"])
dev = torch.device("cuda", 0)
NAMES = ["kernel", "prod:emptyA", "prod:emptyB", "mma:fullA", "mma:fullB", "mma:acc_empty", "epi:acc_full", "epi:body", "tiles", "prologue"]


def read(reset):
    if PLAIN:
        return torch.zeros(148, 16, dtype=torch.float64)
    buf = (ctypes.c_ulonglong * (160 * 16))()
    raw.evk_gt_prof_read(buf, int(reset))
    return torch.tensor(list(buf), dtype=torch.float64).view(160, 16)[:148]


def run(name, fn, flops):
    with torch.no_grad():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        read(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
    c = read(True) / 10.0
    us = e0.elapsed_time(e1) * 100.0
    tiles = c[:, 8]
    busy = c[tiles > 0]
    print(f"== {name}: {us:.1f} us/launch, {flops / us / 1e6:.0f} TFLOP/s; tiles/CTA min {tiles.min():.0f} max {tiles.max():.0f}")
    if PLAIN:
        return
    worst = busy[busy[:, 0].argmax()]
    for i, n in enumerate(NAMES):
        if i == 8:
            continue
        print(f"   {n:14s} mean {busy[:, i].mean():10.0f}   slowest-CTA {worst[i]:10.0f}   per tile (slowest CTA) {worst[i] / max(worst[8], 1):9.0f}")
    sys.stdout.flush()


def conv_case(name, b, T, C, N, k, P=1, stride=1):
    x = torch.randn(b, T * P, C, device=dev)
    w = ops.pack_weight(torch.randn(N, C, k, device=dev) * 0.02, None)
    bias = torch.zeros(N, device=dev)
    y = ops.conv(x, w, bias, stride=stride, pad=(k - 1) // 2, P=P)
    run(name, lambda: ops.conv(x, w, bias, stride=stride, pad=(k - 1) // 2, P=P), 2.0 * b * y.shape[1] * N * C * k)


def lin_case(name, M, K, N):
    x = torch.randn(1, M, K, device=dev)
    w = ops.pack_weight(torch.randn(N, K, 1, device=dev) * 0.02, None)
    bias = torch.zeros(N, device=dev)
    run(name, lambda: ops.linear(x, w, bias), 2.0 * M * K * N)


for slab in (1, 0):
    L.evk_set_tma_options(slab, 1, 3.52e-4)
    print(f"######## slab={slab}")
    conv_case("k11 128ch T=2560", 16, 2560, 128, 128, 11)
    conv_case("k7 128ch T=2560", 16, 2560, 128, 128, 7)
    conv_case("k11 256ch T=320", 16, 320, 256, 256, 11)
    conv_case("k11 64ch T=5120", 16, 5120, 64, 64, 11)
    if slab:
        conv_case("discP 1024 k5 p2", 32, 127, 1024, 1024, 5, P=2)
        lin_case("gpt linear1 512->2048", 20480, 512, 2048)
        lin_case("gpt linear2 2048->512", 20480, 2048, 512)
        lin_case("gpt out_proj 512->512", 20480, 512, 512)
