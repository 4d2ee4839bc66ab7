#!/usr/bin/env python
"""Throughput of the TTS vocoder call (vocoder half): SynthesizerTrn.decode of one 10 s utterance
(250 semantic tokens -> 500 frames -> 320 000 samples at 32 kHz), eager launches, inputs on the device.  CPU arm: the oracle
restatement of the reference's decode on this box's cores (one run).  One JSON line.
   python tools/bench_decode.py > gpurun_out/bench_decode.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, normalize_token as nt  # noqa: E402
from oracle import s2_oracle  # noqa: E402  (CPU arm only)

lib.init()
dev = torch.device("cuda", 0)
T, X, Tr = 250, 120, 300
g = torch.Generator().manual_seed(5)
codes = torch.randint(0, 1024, (1, 1, T), generator=g)
text = torch.randint(0, 300, (1, X), generator=g)
refer = torch.rand(1, 1025, Tr, generator=g) * 2.0
noise = torch.randn(1, 192, 2 * T, generator=g)
P = s2_oracle.init_params(s2_oracle.generator_param_spec(), 1234)
net = nt.load_vq_model(device=dev, state_dict=P)
cd, td, rd, nd = codes.to(dev), text.to(dev), refer.to(dev), noise.to(dev)
for _ in range(3):
    o = net.decode(cd, td, rd, noise=nd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    o = net.decode(cd, td, rd, noise=nd)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / reps
gpu_ms = e0.elapsed_time(e1) / reps
secs = o.shape[-1] / 32000.0
threads = min(16, os.cpu_count() or 1)
torch.set_num_threads(threads)
with torch.no_grad():
    t0 = time.perf_counter()
    oc = s2_oracle.decode(P, codes, text, [refer], noise)
    cpu_s = time.perf_counter() - t0
err = float((o.cpu() - oc).norm() / oc.norm())
print(json.dumps(dict(metric="SynthesizerTrn.decode (TTS vocoder call), one 10 s utterance", unit="audio-s/s", value=secs / wall,
                      ms_wall=wall * 1e3, ms_gpu_events=gpu_ms, samples=int(o.shape[-1]), launch_mode="eager (host-bound: ~1 100 launches)",
                      cpu_baseline=dict(value=secs / cpu_s, unit="audio-s/s", cores=threads, kind="port", sample="one decode call"),
                      rel_l2_vs_cpu_oracle=err)))
