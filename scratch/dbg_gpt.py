import json, os, sys, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpt_oracle
from easevoice_trainer_b200 import lib, ops
from easevoice_trainer_b200.models_gpt import Text2SemanticDecoder
L = lib.init()
L.evk_set_precise(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
tag = sys.argv[1] if len(sys.argv) > 1 else "ragged"
gold = json.load(open(f"tests/golden/gpt_{tag}.json"))
m = gold["model"]
spec = gpt_oracle.gpt_param_spec(m)
P = gpt_oracle.init_params(spec, gold["param_seed"])
P["ar_text_position.alpha"].fill_(gold["alpha_text"]); P["ar_audio_position.alpha"].fill_(gold["alpha_audio"])
net = Text2SemanticDecoder({"model": m}, layer_dropout=0.0)
net.load_state_dict(P); net = net.cuda()
x, xl, y, yl, bert = gpt_oracle.synthetic_gpt_batch(gold["B"], gold["X"], gold["Y"], gold["batch_seed"], gold["ragged"])
Pq = {k: v.clone().requires_grad_(True) for k, v in P.items()}
loss_o = gpt_oracle.forward_old(Pq, x, xl, y, yl, bert, m)[0]; loss_o.backward()
loss, acc = net.forward_old(x.cuda(), xl.cuda(), y.cuda(), yl.cuda(), bert.cuda())
names = [n for n, _ in net.named_parameters()]
grads = torch.autograd.grad(loss, [p for _, p in net.named_parameters()])
rows = []
for n, g in zip(names, grads):
    r = Pq[n].grad
    rows.append((float((g.cpu() - r).norm() / (r.norm() + 1e-20)), n, float(g.norm()), float(r.norm())))
for e, n, a, b in sorted(rows, reverse=True)[:12]:
    print(f"{e:.3e} {n:50s} |g|={a:.4e} |ref|={b:.4e}")
print(xl, yl)
