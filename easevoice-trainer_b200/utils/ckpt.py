"""Checkpoint I/O with the reference's on-disk layouts (src/utils/path/ckpt.py:13-93, src/train/sovits.py:179-196).

  resumable: {"model": state_dict, "iteration": epoch, "optimizer": optimizer.state_dict(), "learning_rate": lr}
             at logs/{G,D}_latest.pth or {G,D}_{step}.pth
  export:    OrderedDict(weight={k: v.half() for k without "enc_q"}, config=<hps dict>, info="<epoch>epoch_<step>iteration")
"""
import glob
import os
import shutil
import time
from collections import OrderedDict

import torch


def save_with_torch(obj, path):
    tmp = os.path.join(os.path.dirname(os.path.abspath(path)), f".{time.time()}.pth.tmp")   # next to the target, not in the CWD
    torch.save(obj, tmp)
    shutil.move(tmp, path)


def save_checkpoint(model, optimizer, learning_rate, iteration, checkpoint_path):
    state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    save_with_torch({"model": state, "iteration": iteration, "optimizer": optimizer.state_dict(),
                     "learning_rate": learning_rate}, checkpoint_path)


def load_checkpoint(checkpoint_path, model, optimizer=None, skip_optimizer=False):
    assert os.path.isfile(checkpoint_path)
    ck = torch.load(checkpoint_path, map_location="cpu")
    saved = ck["model"]
    cur = model.state_dict()
    new = {}
    for k, v in cur.items():                      # tolerant of missing / mis-shaped keys, like the reference
        new[k] = saved[k] if (k in saved and tuple(saved[k].shape) == tuple(v.shape)) else v
    with torch.no_grad():
        for k, v in cur.items():
            v.copy_(new[k])                       # in place: parameters may be views into a flat optimizer arena
    if optimizer is not None and not skip_optimizer and ck.get("optimizer") is not None:
        optimizer.load_state_dict(ck["optimizer"])
    return model, optimizer, ck["learning_rate"], ck["iteration"]


def latest_checkpoint_path(dir_path, regex="G_*.pth"):
    f_list = glob.glob(os.path.join(dir_path, regex))
    latest = [x for x in f_list if "latest" in x]
    if latest:
        return latest[0]
    f_list.sort(key=lambda f: int("".join(filter(str.isdigit, f))))
    return f_list[-1]


def export_weights(state_dict, hps, name, epoch, steps, save_dir):
    """sovits.py:179-196: half-precision inference export without the posterior encoder."""
    opt = OrderedDict()
    opt["weight"] = {k: v.detach().cpu().half() for k, v in state_dict.items() if "enc_q" not in k}
    opt["config"] = hps
    opt["info"] = "%sepoch_%siteration" % (epoch, steps)
    path = os.path.join(save_dir, f"{name}.pth")
    save_with_torch(opt, path)
    return path
