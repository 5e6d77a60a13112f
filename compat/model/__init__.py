"""`model` package replacement (reference: dgl/model/).  Importing it also restores the checkpoint-loading behaviour the
reference's scripts were written for: they pickle the WHOLE module (`torch.save({'model': model, ...})`, dgl/training.py:345-355)
and read it back with a bare `torch.load(path)` (dgl/training.py:291, dgl/eval.py:251).  Since torch 2.6 a bare `torch.load`
defaults to weights_only=True and refuses pickled modules, so the unchanged scripts could not resume or evaluate.  The scripts do
`from model import models` before they load anything, so this is the place to put the default back: a `torch.load` call that does
not pass `weights_only` itself gets weights_only=False (the scripts' own trust model: they only load checkpoints they wrote).
Callers that pass weights_only explicitly are untouched; ROBOT_GNN_STRICT_LOAD=1 disables the shim."""
import functools
import os

import torch

_ORIG_LOAD = getattr(torch.load, "__wrapped_by_robot_gnn__", None) or torch.load


def enable_legacy_checkpoint_load():
    if getattr(torch.load, "__wrapped_by_robot_gnn__", None) is not None or os.environ.get("ROBOT_GNN_STRICT_LOAD") == "1":
        return

    @functools.wraps(_ORIG_LOAD)
    def load(*args, **kwargs):
        kwargs.setdefault("weights_only", False)
        return _ORIG_LOAD(*args, **kwargs)

    load.__wrapped_by_robot_gnn__ = _ORIG_LOAD
    torch.load = load


def disable_legacy_checkpoint_load():
    orig = getattr(torch.load, "__wrapped_by_robot_gnn__", None)
    if orig is not None:
        torch.load = orig


enable_legacy_checkpoint_load()
