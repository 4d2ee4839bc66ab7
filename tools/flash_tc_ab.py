#!/usr/bin/env python
"""A/B of the two prefix-LM attention kernel families behind evk_flash_attn_fwd / _bwd: mma.sync (csrc/flash.cu) vs tcgen05 /
TMEM (csrc/flash_tc.cu).  Parity of O, dq, dk, dv between the families and against a float64 torch reference (small case), then
CUDA-event timing of forward and backward at the benchmarked shape (B = 16, 16 heads, X = 256, Y = 1024).
    python tools/flash_tc_ab.py [--no-time] > gpurun_out/flash_tc_ab.json"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, ops  # noqa: E402

L_ = lib.init()
dev = torch.device("cuda", 0)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def run(qkv, dout, H, X, xl, yl, p, tc):
    L_.evk_set_flash_tc(int(tc), -1.0)
    q = qkv.clone().requires_grad_(True)
    o = ops.flash_attention(q, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=p, tag="ab")
    o.backward(dout)
    torch.cuda.synchronize()
    return o.detach(), q.grad.detach()


def ref64(qkv, dout, H, X, xl, yl):
    B, L, D3 = qkv.shape
    D = D3 // 3
    x = qkv.double().cpu().requires_grad_(True)
    q, k, v = (x[..., i * D:(i + 1) * D].view(B, L, H, 32).transpose(1, 2) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(32.0)
    i = torch.arange(L)[:, None]
    j = torch.arange(L)[None, :]
    masks = []
    for b in range(B):
        text = (j < X) & (j < int(xl[b]))
        audio = (j >= X) & ((j - X) < int(yl[b])) & (j <= i)
        masks.append(text | audio)
    m = torch.stack(masks)[:, None]
    s = s.masked_fill(~m, float("-inf"))
    pr = torch.softmax(s, -1)
    pr = torch.nan_to_num(pr, nan=0.0)
    o = (pr @ v).transpose(1, 2).reshape(B, L, D)
    o.backward(dout.double().cpu())
    return o.detach(), x.grad.detach()


TC = int(os.environ.get("AB_TC", "3"))          # kernel generation under test (1: single-stage, 2: warp-specialised forward)
out = dict(parity=[], timing={}, tc=TC)
g = torch.Generator().manual_seed(11)
cases = [  # B, H, X, Y, xlens, ylens, p_drop
    (2, 4, 12, 20, [12, 7], [20, 13], 0.0),
    (2, 16, 256, 300, [256, 190], [300, 211], 0.0),
    (3, 16, 100, 413, [100, 64, 1], [413, 129, 300], 0.0),
    (2, 16, 256, 300, [256, 190], [300, 211], 0.1),
    (1, 16, 0, 130, [0], [130], 0.1),
]
bad = 0
for (B, H, X, Y, xls, yls, p) in cases:
    L = X + Y
    qkv = (torch.randn(B, L, 3 * H * 32, generator=g) * 1.5).to(dev)
    dout = torch.randn(B, L, H * 32, generator=g).to(dev)
    xl = torch.tensor(xls, device=dev, dtype=torch.int64)
    yl = torch.tensor(yls, device=dev, dtype=torch.int64)
    o0, g0 = run(qkv, dout, H, X, xl, yl, p, tc=0)
    o1, g1 = run(qkv, dout, H, X, xl, yl, p, tc=TC)
    D = H * 32
    row = dict(case=f"B{B} H{H} X{X} Y{Y} p{p}", o=rel(o1, o0), dq=rel(g1[..., :D], g0[..., :D]), dk=rel(g1[..., D:2 * D], g0[..., D:2 * D]),
               dv=rel(g1[..., 2 * D:], g0[..., 2 * D:]), finite=bool(torch.isfinite(o1).all() and torch.isfinite(g1).all()))
    if p == 0.0 and B * L * L * H <= 3 * 16 * 520 * 520:
        orf, grf = ref64(qkv, dout, H, X, xl.cpu(), yl.cpu())
        # rows of padded queries (i >= X + ylen) carry whatever the mask gives both implementations alike; compare everything
        row.update(o_vs_f64_tc=rel(o1, orf), o_vs_f64_mma=rel(o0, orf), dq_vs_f64_tc=rel(g1[..., :D], grf[..., :D]),
                   dq_vs_f64_mma=rel(g0[..., :D], grf[..., :D]), dk_vs_f64_tc=rel(g1[..., D:2 * D], grf[..., D:2 * D]),
                   dk_vs_f64_mma=rel(g0[..., D:2 * D], grf[..., D:2 * D]), dv_vs_f64_tc=rel(g1[..., 2 * D:], grf[..., 2 * D:]),
                   dv_vs_f64_mma=rel(g0[..., 2 * D:], grf[..., 2 * D:]))
    ok = row["finite"] and max(row["o"], row["dq"], row["dk"], row["dv"]) < 4e-3
    row["ok"] = ok
    bad += 0 if ok else 1
    out["parity"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)

if "--no-time" not in sys.argv:
    B, H, X, Y = 16, 16, 256, 1024
    L = X + Y
    qkv = torch.randn(B, L, 3 * H * 32, generator=g).to(dev)
    dout = torch.randn(B, L, H * 32, generator=g).to(dev)
    xl = torch.full((B,), X, device=dev, dtype=torch.int64)
    yl = torch.full((B,), Y, device=dev, dtype=torch.int64)
    pairs = B * H * (X * L + Y * (Y + 1) // 2)
    for tc in (0, 2, 3):
        for p in (0.0, 0.1):
            L_.evk_set_flash_tc(tc, -1.0)
            q = qkv.clone().requires_grad_(True)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            tf = tb = 0.0
            for it in range(6):
                q.grad = None
                ev[0].record()
                o = ops.flash_attention(q, heads=H, prefix=X, xlen=xl, ylen=yl, p_drop=p, tag="ab")
                ev[1].record()
                o.backward(dout)
                ev[2].record()
                torch.cuda.synchronize()
                if it >= 2:
                    tf += ev[0].elapsed_time(ev[1]) / 4
                    tb += ev[1].elapsed_time(ev[2]) / 4
            out["timing"][f"tc{tc}_p{p}"] = dict(fwd_ms=tf, bwd_ms=tb, fwd_tflops=pairs * 128 / tf / 1e9, bwd_tflops=pairs * 320 / tb / 1e9)
            print(json.dumps({f"tc{tc}_p{p}": out["timing"][f"tc{tc}_p{p}"]}), file=sys.stderr, flush=True)
L_.evk_set_flash_tc(1, -1.0)
out["bad"] = bad
print(json.dumps(out))
sys.exit(1 if bad else 0)
