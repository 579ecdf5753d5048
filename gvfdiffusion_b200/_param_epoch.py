"""A process-wide counter of optimiser steps, part of the engines' "have the parameters changed?" signature.

The engines (vae_engine / vae_train / vae_encode / sparse.transformer) keep fp16 copies and transposes of the fp32 master
Parameters and refresh them when a parameter's (data_ptr, _version) pair has moved.  torch's FUSED optimisers
(`AdamW(fused=True)`, what a training loop on this hardware uses) update parameters WITHOUT bumping `_version` -- measured:
the counter stays put across `opt.step()` and the engines kept stepping on the initial weights -- so every optimiser step of
any optimiser also bumps this counter, through torch's global step post-hook."""
import torch

_epoch = [0]


def _bump(optimizer, args, kwargs):
    _epoch[0] += 1


try:
    import importlib
    importlib.import_module("torch.optim.optimizer").register_optimizer_step_post_hook(_bump)
    HOOKED = True
except (AttributeError, ImportError):       # very old torch: the training Functions then refresh on every forward
    HOOKED = False


def epoch():
    return _epoch[0]
