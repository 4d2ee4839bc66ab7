import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from easevoice_trainer_b200 import lib, ops
from easevoice_trainer_b200.models_gpt import Text2SemanticDecoder
from easevoice_trainer_b200.train import gpt_step
from easevoice_trainer_b200.configs import GPT_MODEL
lib.init()
dev = torch.device("cuda", 0)
for eager, cl in ((False, False), (True, False), (False, True)):
    net = Text2SemanticDecoder({"model": dict(GPT_MODEL, n_layer=2)}, seed=1).to(dev).train()
    st = gpt_step.GptStep(net)
    b = gpt_step.synthetic_batch(4, 32, 64, seed=1, device=dev)
    if cl:
        b["bert_feature"] = ops.to_channels_last(b["bert_feature"]); b["bert_channels_last"] = True
    try:
        if eager:
            st.batch_idx = 4; st.step(b); st.batch_idx = 1
        st.graph_step(b)
        torch.cuda.synchronize()
        for i in range(6):
            out = st.graph_step(b)
        print("OK", eager, cl, float(out[0]), float(out[1]), st.opt.step_count, flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
        print("FAIL", eager, cl, str(e).split("\n")[0], flush=True)
        try: torch.cuda.synchronize()
        except Exception as e2: print("sync fail", e2)
