"""Output-directory scheme of the reference (src/train/helper.py:10-34)."""
import datetime
import os
from dataclasses import dataclass

train_logs_path = "logs"


@dataclass
class TrainOutput:
    model_path: str


def generate_random_name():
    return datetime.datetime.now().strftime("%Y%m%d-%H%M%S")


def get_sovits_train_dir(project_dir, name):
    if not name:
        name = "sovits_" + generate_random_name()
    return os.path.join(project_dir, "models", "sovits_train", name)


def get_gpt_train_dir(project_dir, name):
    """src/train/helper.py:21-24."""
    if not name:
        name = "gpt_" + generate_random_name()
    return os.path.join(project_dir, "models", "gpt_train", name)
