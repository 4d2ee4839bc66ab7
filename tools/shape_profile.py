#!/usr/bin/env python
"""Which launch shapes carry the stage-2 step?  One eager step (single stream, host run-ahead behind a spin kernel) with a
CUDA-event pair around every library call, keyed by kernel family AND shape.
   python tools/shape_profile.py [top_n] > gpurun_out/shape_profile.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib, ops, models, configs  # noqa: E402
from easevoice_trainer_b200.train import s2_step              # noqa: E402

lib.init()
dev = torch.device("cuda", 0)
hps = configs.load_s2_config()
torch.manual_seed(1234)
net_g = models.SynthesizerTrn(hps["data"]["filter_length"] // 2 + 1, hps["train"]["segment_size"] // hps["data"]["hop_length"],
                              n_speakers=hps["data"]["n_speakers"], **hps["model"]).to(dev).train()
net_d = models.MultiPeriodDiscriminator(hps["model"]["use_spectral_norm"]).to(dev).train()
st = s2_step.S2Step(net_g, net_d, hps["train"], hps["data"])
batch = s2_step.to_device_batch(s2_step.synthetic_batch(16, 346, 120, dev, seed=1234), dev, st.bank)
st.step(batch)
st.step(batch)
torch.cuda.synchronize()
models.SIDE_STREAMS = False
ops.PROF_SHAPES = True
torch.cuda._sleep(int(0.04 * 1.9e9))
ops.profile_begin()
st.step(batch)
prof = ops.profile_end()
tot = sum(v["ms"] for v in prof.values())
print(f"library calls of one step: {sum(v['calls'] for v in prof.values())}, {tot:.2f} ms of event time")
top = int(sys.argv[1]) if len(sys.argv) > 1 else 45
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:top]:
    tf = f"{v['flops'] / (v['ms'] * 1e-3) / 1e12:7.1f} TF/s" if v["flops"] else "            "
    print(f"{v['ms']:7.3f} ms {100 * v['ms'] / tot:5.1f}% {v['calls']:4d}x {1e3 * v['ms'] / v['calls']:7.1f} us  {tf}  {k}")
