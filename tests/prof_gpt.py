#!/usr/bin/env python
"""ncu driver: one eager stage-1 AR-GPT training_step (micro-batch fwd+bwd + ScaledAdam update) at BASELINE configs[1] size.
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gpt_launches.csv python tests/prof_gpt.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops                               # noqa: E402
from easevoice_trainer_b200.models_gpt import Text2SemanticDecoder        # noqa: E402
from easevoice_trainer_b200.train import gpt_step                         # noqa: E402
from easevoice_trainer_b200.configs import GPT_MODEL                      # noqa: E402

lib.init()
dev = torch.device("cuda", 0)
layers = int(os.environ.get("GPT_LAYERS", "24"))
net = Text2SemanticDecoder({"model": dict(GPT_MODEL, n_layer=layers)}, seed=1234).to(dev).train()
st = gpt_step.GptStep(net)
b = gpt_step.synthetic_batch(16, 256, 1024, seed=1, device=dev)
b["bert_feature"] = ops.to_channels_last(b["bert_feature"]); b["bert_channels_last"] = True
st.batch_idx = 1
st.step(b)                          # first micro-batch: records the packing plan
torch.cuda.synchronize()
st.batch_idx = 4                    # steady state micro-batch with the optimizer update
torch.cuda.profiler.start()         # ncu --profile-from-start off
st.step(b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", float(st.last[0]))
