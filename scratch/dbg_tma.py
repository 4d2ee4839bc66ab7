import os, sys, torch, math, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops
L = lib.init()
dev = torch.device("cuda", 0)
def sync(tag):
    try:
        torch.cuda.synchronize(); print("OK  ", tag, flush=True)
    except Exception as e:
        print("FAIL", tag, str(e).split("\n")[0], flush=True); sys.exit(0)
g = torch.Generator().manual_seed(0)
# forward variants by N
for N in (256, 128, 64, 32):
    x = torch.randn(4, 1024, 64, generator=g).to(dev)
    w = ops.pack_weight((torch.randn(N, 64, 3, generator=g) / 14).to(dev))
    with torch.no_grad():
        y = ops.conv(x, w, None, pad=1)
    sync(f"fwd conv N={N}")
    ref = torch.nn.functional.conv1d(x.transpose(1, 2).cpu(), w.pa.cpu().permute(1, 2, 0)[:, :64, :].contiguous(), padding=1).transpose(1, 2)
    print("   rel", float((y.cpu() - ref).norm() / ref.norm()))
# wgrad
for (B, T, C, N, Q, pad, dil, P) in ((4, 1024, 64, 64, 3, 1, 1, 1), (3, 1500, 128, 128, 11, 5, 1, 1), (3, 310, 512, 1024, 5, 2, 1, 3), (2, 2100, 64, 64, 7, 15, 5, 1)):
    x = torch.randn(B, T * P, C, generator=g).to(dev).requires_grad_(True)
    v = (torch.randn(N, C, Q, generator=g) / math.sqrt(C * Q)).to(dev).requires_grad_(True)
    ops.USE_TMA_WGRAD = False
    y = ops.conv(x, ops.pack_weight(v), None, pad=pad, dil=dil, P=P)
    gy = torch.randn(y.shape, generator=g).to(dev)
    (g0,) = torch.autograd.grad(y, v, gy)
    sync("old wgrad")
    ops.USE_TMA_WGRAD = True
    y = ops.conv(x, ops.pack_weight(v), None, pad=pad, dil=dil, P=P)
    (g1,) = torch.autograd.grad(y, v, gy)
    sync(f"tma wgrad {B,T,C,N,Q,P}")
    print("   rel", float((g1 - g0).norm() / g0.norm()))
