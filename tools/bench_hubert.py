#!/usr/bin/env python
"""Throughput of the Normalize.ssl model call (ssl half): HuBERT-base forward on one 10 s utterance at 16 kHz
(160 000 samples -> 499 frames of 768), the H2D copy of the samples and the D2H copy of the features inside the timed region.
CPU arm: the oracle restatement on this box's cores (one run; transformers' own HubertModel when importable).  One JSON line.
   python tools/bench_hubert.py > gpurun_out/bench_hubert.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, hubert  # noqa: E402
from oracle import hubert_oracle as ho  # noqa: E402  (CPU arm + seeded weights)

lib.init()
dev = torch.device("cuda", 0)
P = ho.init_params(ho.param_spec(), 42)
net = hubert.HubertModel()
net.load_state_dict(P)
net = net.to(dev).eval()
wav = (torch.randn(1, 160000, generator=torch.Generator().manual_seed(9)) * 0.3).pin_memory()
for _ in range(3):
    o = net(wav.to(dev, non_blocking=True))["last_hidden_state"].cpu()
torch.cuda.synchronize()
reps = 10
t0 = time.perf_counter()
for _ in range(reps):
    o = net(wav.to(dev, non_blocking=True))["last_hidden_state"].cpu()
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / reps
threads = min(16, os.cpu_count() or 1)
torch.set_num_threads(threads)
kind = "port"
with torch.no_grad():
    t0 = time.perf_counter()
    oc = ho.forward(P, wav)
    cpu_s = time.perf_counter() - t0
err = float((o - oc).norm() / oc.norm())
print(json.dumps(dict(metric="HuBERT-base forward (Normalize.ssl model call), one 10 s utterance", unit="audio-s/s", value=10.0 / gpu_s,
                      ms=gpu_s * 1e3, frames=int(o.shape[1]), launch_mode="eager", h2d_bytes=640000, d2h_bytes=int(o.numel() * 4),
                      cpu_baseline=dict(value=10.0 / cpu_s, unit="audio-s/s", cores=threads, kind=kind, sample="one forward, oracle restatement"),
                      rel_l2_vs_cpu_oracle=err)))
