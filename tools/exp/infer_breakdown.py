#!/usr/bin/env python
"""Where does a token of infer_panel go?  Wall-clock of (a) the replayed 24-layer token-step graph alone, (b) the sampling code alone
(logits_to_probs + exponential race + the two EOS tests with their host syncs), (c) the whole loop."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from easevoice_trainer_b200 import lib  # noqa: E402
from easevoice_trainer_b200.models_gpt import Text2SemanticDecoder  # noqa: E402
from oracle import gpt_oracle  # noqa: E402

lib.init()
dev = torch.device("cuda", 0)
m = dict(gpt_oracle.GPT_MODEL)
P = gpt_oracle.init_params(gpt_oracle.gpt_param_spec(m), 35)
net = Text2SemanticDecoder({"model": m})
net.load_state_dict(P)
net = net.to(dev).eval()
X, Yp = 120, 150
g = torch.Generator().manual_seed(2)
x = torch.randint(0, m["phoneme_vocab_size"], (1, X), generator=g).to(dev)
bert = torch.randn(1, 1024, X, generator=g).to(dev)
prompts = torch.randint(0, 1024, (1, Yp), generator=g).to(dev)
xl = torch.tensor([X], device=dev)
for n in (8, 300):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y, idx = net.infer_panel(x, xl, prompts, bert, top_k=15, top_p=1, early_stop_num=n, temperature=1.0)
    torch.cuda.synchronize(); print(f"infer_panel early_stop {n}: {time.perf_counter() - t0:.4f} s, {y.shape[1] - Yp} tokens")
st = net.__dict__["_infer_st"]
with torch.no_grad():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300):
        st["graph"].replay()
    torch.cuda.synchronize(); print(f"graph replay alone: {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms per token")
    st["n"].fill_(X + Yp)
    logits0 = torch.randn(1, 1025, device=dev)
    yy = prompts.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(300):
        logits = logits0.clone()
        probs = net.logits_to_probs(logits, yy, temperature=1.0, top_k=15, top_p=1, repetition_penalty=1.35)
        q = torch.empty_like(probs).exponential_(1)
        samples = torch.argmax(probs / q, dim=-1, keepdim=True).to(torch.int64)
        yy = torch.cat([yy, samples], dim=1)
        stop = int(torch.argmax(logits, dim=-1)[0]) == 1024 or int(samples[0, 0]) == 1024
    torch.cuda.synchronize(); print(f"sampling alone: {(time.perf_counter() - t0) / 300 * 1e3:.3f} ms per token")
if "--ncu" in sys.argv:                      # ncu --profile-from-start off --metrics gpu__time_duration.sum: two eager token steps
    import easevoice_trainer_b200.models_gpt as mg
    caches = st["caches"]
    last = torch.randn(1, 1, 512, device=dev)
    with torch.no_grad():
        net._active, net._memo_pack = net.packed_for_inference(), True
        try:
            n = X + Yp + 300
            for rep in range(3):
                if rep == 1:
                    torch.cuda.synchronize(); torch.cuda.profiler.start()
                h = last
                for i in range(net.num_layers):
                    h = net._infer_layer(i, h, caches[i], n)
                n += 1
            torch.cuda.synchronize(); torch.cuda.profiler.stop()
        finally:
            net._active, net._memo_pack = None, False
