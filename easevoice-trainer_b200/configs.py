"""Stage-2 hyper-parameter file: same schema and path convention as the reference (configs/s2.json, resolved by
src/utils/config/__init__.py:32).  Note the reference runs ``fp16_run: true``; this implementation computes in
fp32 storage / TF32 tensor-core math and ignores the flag (no GradScaler needed)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
S2_CONFIG_PATH = os.path.join(ROOT, "configs", "s2.json")


def load_s2_config(path=None):
    with open(path or S2_CONFIG_PATH) as f:
        return json.load(f)


# configs/gpt.yaml `model:` section of the reference (vocab_size 1025, phoneme_vocab_size 732 in GPT-SoVITS v2; EOS = 1024)
GPT_MODEL = dict(vocab_size=1025, phoneme_vocab_size=732, embedding_dim=512, hidden_dim=512, head=16, linear_units=2048,
                 n_layer=24, dropout=0, EOS=1024, random_bert=0)
