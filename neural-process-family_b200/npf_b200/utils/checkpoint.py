"""Checkpoint / resume in upstream's on-disk layout (the directory skorch's ``Checkpoint`` callback writes through
utils/train.py:203-222 and ``_eval_save_load`` 308-330, e.g. results/pretrained/RBF_Kernel/CNP/run_0/):

    params.pt          model.state_dict()  -- key-for-key upstream's, so either code base loads the other's file
    optimizer.pt       torch.optim.Adam state dict (``FlatAdam.torch_state_dict``; upstream's own files load too)
    history.json       list of per-epoch records (free-form dicts)
    eval.csv           test log-likelihood per task, one value per line (``numpy.savetxt`` format)
    model_summary.txt  str(model)
"""
import json
import os

import numpy as np
import torch

__all__ = ["save_checkpoint", "load_checkpoint"]


def save_checkpoint(dirname, model, optimizer=None, history=None, eval_loglik=None):
    os.makedirs(dirname, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(dirname, "params.pt"))
    if optimizer is not None:
        sd = optimizer.torch_state_dict() if hasattr(optimizer, "torch_state_dict") else optimizer.state_dict()
        torch.save(sd, os.path.join(dirname, "optimizer.pt"))
    if history is not None:
        with open(os.path.join(dirname, "history.json"), "w") as f:
            json.dump(history, f)
    if eval_loglik is not None:
        np.savetxt(os.path.join(dirname, "eval.csv"), np.asarray(torch.as_tensor(eval_loglik).detach().cpu(), dtype=np.float64))
    with open(os.path.join(dirname, "model_summary.txt"), "w") as f:
        f.write(str(model))


def load_checkpoint(dirname, model, optimizer=None, strict=True, trust=False):
    """Load ``params.pt`` into ``model`` (and ``optimizer.pt`` into ``optimizer`` when given and present); returns the
    stored history (``[]`` if none).  Works on upstream's own checkpoint directories.
    Both files are read with the restricted unpickler (``weights_only=True``: tensors, dicts, lists, numbers -- all a
    state dict holds), so a downloaded checkpoint directory cannot execute code; ``trust=True`` falls back to the full
    unpickler for files that need it."""
    sd = torch.load(os.path.join(dirname, "params.pt"), map_location="cpu", weights_only=not trust)
    model.load_state_dict(sd, strict=strict)
    opt_file = os.path.join(dirname, "optimizer.pt")
    if optimizer is not None and os.path.exists(opt_file):
        osd = torch.load(opt_file, map_location="cpu", weights_only=not trust)
        if hasattr(optimizer, "load_torch_state_dict"):
            optimizer.load_torch_state_dict(osd)
        else:
            optimizer.load_state_dict(osd)
    hist_file = os.path.join(dirname, "history.json")
    if os.path.exists(hist_file):
        with open(hist_file) as f:
            return json.load(f)
    return []
