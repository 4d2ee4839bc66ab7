#!/usr/bin/env python
"""Throughput of the Normalize.token model call: HuBERT features [1, 768, T] -> semantic tokens.
GPU: normalize_token.extract_tokens (zero-padded batches, exact-fp32 projection + codeword search), features resident in host
memory (the H2D copy is inside the timed region, as when they come from 4-cnhubert/*.pt).  CPU arm: the oracle restatement of
the reference's per-file call (normalize.py:203) on this box's cores, bounded sample.  One JSON line.
   python tools/bench_normalize_token.py > gpurun_out/bench_normalize_token.json"""
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
N, T = 256, 500                                     # 256 utterances of 10 s (50 Hz features)
g = torch.Generator().manual_seed(3)
feats = [torch.randn(1, 768, T - (i % 7) * 10, generator=g).pin_memory() for i in range(N)]
P = s2_oracle.init_params(s2_oracle.generator_param_spec(), 1234)
net = nt.load_vq_model(device=dev, state_dict=P)
nt.extract_tokens(net, feats[:32])
torch.cuda.synchronize()
reps = 5
t0 = time.perf_counter()
for _ in range(reps):
    toks = nt.extract_tokens(net, feats)
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / reps
frames = sum(f.shape[-1] for f in feats)
threads = min(16, os.cpu_count() or 1)
torch.set_num_threads(threads)
n_cpu = 24
t0 = time.perf_counter()
with torch.no_grad():
    cpu = [s2_oracle.extract_latent(P, f)[0, 0].tolist() for f in feats[:n_cpu]]
cpu_s = time.perf_counter() - t0
mism = sum(int(a != b) for x, y in zip(toks[:n_cpu], cpu) for a, b in zip(x, y))
print(json.dumps(dict(metric="Normalize.token semantic-token extraction", unit="feature-frames/s", value=frames / gpu_s,
                      utterances=N, frames=frames, ms=gpu_s * 1e3, h2d_bytes=frames * 768 * 4,
                      cpu_baseline=dict(value=sum(f.shape[-1] for f in feats[:n_cpu]) / cpu_s, unit="feature-frames/s", cores=threads, kind="port",
                                        sample=f"{n_cpu} utterances, one model call per file as normalize.py:203"),
                      token_mismatches_vs_cpu_oracle=mism)))
