#!/usr/bin/env python
"""ncu driver: one eager stage-2 step at the bench workload (B = 16 x 10 s @ 22.05 kHz label).
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches.csv python tests/prof_s2.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops, models, configs      # noqa: E402
from easevoice_trainer_b200.train import s2_step                  # noqa: E402

lib.init()
dev = torch.device("cuda", 0)
hps = configs.load_s2_config()
torch.manual_seed(1234)
net_g = models.SynthesizerTrn(hps["data"]["filter_length"] // 2 + 1, hps["train"]["segment_size"] // hps["data"]["hop_length"],
                              n_speakers=hps["data"]["n_speakers"], **hps["model"]).to(dev).train()
net_d = models.MultiPeriodDiscriminator(hps["model"]["use_spectral_norm"]).to(dev).train()
st = s2_step.S2Step(net_g, net_d, hps["train"], hps["data"])
host = s2_step.synthetic_batch(16, 346, 120, dev, seed=1234)
batch = s2_step.to_device_batch(host, dev, st.bank)
st.step(batch)                      # first step: records the packing plans (per-layer path)
torch.cuda.synchronize()
torch.cuda.profiler.start()         # ncu --profile-from-start off: only the steady-state step below is profiled
st.step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
