import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops
from easevoice_trainer_b200.models_gpt import Text2SemanticDecoder
from easevoice_trainer_b200.train import gpt_step
from easevoice_trainer_b200.configs import GPT_MODEL
lib.init()
dev = torch.device("cuda", 0)
net = Text2SemanticDecoder({"model": dict(GPT_MODEL, n_layer=1)}, seed=1).to(dev).train()
b = gpt_step.synthetic_batch(4, 32, 64, seed=1, device=dev)
named = dict(net.named_parameters())
def fb(names):
    loss, acc = net.forward_old(b["phoneme_ids"], b["phoneme_ids_len"], b["semantic_ids"], b["semantic_ids_len"], b["bert_feature"])
    if names:
        return torch.autograd.grad(loss, [named[n] for n in names], allow_unused=True)
    return loss
def trial(tag, names):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fb(names); fb(names)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            out = fb(names)
        g.replay(); torch.cuda.synchronize()
        print("OK  ", tag, flush=True)
    except Exception as e:
        print("FAIL", tag, str(e).split("\n")[0], flush=True)
        torch.cuda.synchronize()
trial("forward only", [])
for n in named:
    trial(n, [n])
